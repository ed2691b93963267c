// subdivide_cuda.go -- cgo body for patches.SubdividePatches (rad/patches/subdivide.go:25-145): the face patches
// made by MakePatchForFace go through vrad_patches_subdivide (host code in libvradcuda.so), the result replaces
// cache.GetPatches(), and the Parent/Child links are handed to the device for the hierarchical transfer build
// and CollectLight.  SOURCE ONLY (no Go toolchain in the build image).
//
//go:build cuda

package patches

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../../vrad_b200/_lib -lvradcuda
#include "vrad_cuda.h"
*/
import "C"

import (
	"log"

	"github.com/galaco/vrad/cache"
	"github.com/galaco/vrad/common/types"
)

func SubdividePatchesCUDA() {
	src := *cache.GetPatches()
	faces := make([]C.vrad_face_patch, len(src))
	var points []C.float
	for i := range src {
		p := &src[i]
		f := &faces[i]
		f.first_point, f.n_points = C.int32_t(len(points)/3), C.int32_t(p.Winding.NumPoints)
		for k := 0; k < p.Winding.NumPoints; k++ {
			points = append(points, C.float(p.Winding.Points[k][0]), C.float(p.Winding.Points[k][1]), C.float(p.Winding.Points[k][2]))
		}
		f.normal[0], f.normal[1], f.normal[2] = C.float(p.Plane.Normal[0]), C.float(p.Plane.Normal[1]), C.float(p.Plane.Normal[2])
		f.plane_dist, f.lux_scale, f.chop = C.float(p.Plane.Distance), C.float(p.LuxScale), C.float(p.Chop)
		if p.Sky {
			f.sky = 1
		}
		if PreventSubdivision(p) || (*cache.GetTargetFaces())[p.FaceNumber].DispInfo != -1 {
			f.no_subdivide = 1
		}
		if p.BaseLight[0] != 0 || p.BaseLight[1] != 0 || p.BaseLight[2] != 0 {
			f.has_base_light = 1
		}
	}
	var n, np C.int
	C.vrad_patches_subdivide(C.int(len(faces)), &faces[0], &points[0], C.float(minChop), 0, 0, &n, &np,
		nil, nil, nil, nil, nil, nil, nil, nil, nil, nil, nil, nil, nil, nil)
	origin, normal := make([]C.float, 3*n), make([]C.float, 3*n)
	dist, area, chop := make([]C.float, n), make([]C.float, n), make([]C.float, n)
	mins, maxs := make([]C.float, 3*n), make([]C.float, 3*n)
	parent, child1, child2, face := make([]C.int32_t, n), make([]C.int32_t, n), make([]C.int32_t, n), make([]C.int32_t, n)
	wfirst, wcount := make([]C.int32_t, n), make([]C.int32_t, n)
	wpts := make([]C.float, 3*np)
	if rc := C.vrad_patches_subdivide(C.int(len(faces)), &faces[0], &points[0], C.float(minChop), n, np, &n, &np,
		&origin[0], &normal[0], &dist[0], &area[0], &mins[0], &maxs[0], &chop[0], &parent[0], &child1[0], &child2[0], &face[0],
		&wfirst[0], &wcount[0], &wpts[0]); rc != 0 {
		log.Fatalf("vrad_patches_subdivide: %s", C.GoString(C.vrad_last_error()))
	}
	out := make([]types.Patch, n)
	for i := range out {
		root := i
		for parent[root] != -1 {
			root = int(parent[root])
		}
		out[i] = src[root] // CreateChildPatch: the child copies its parent (subdivide.go:360)
		q := &out[i]
		q.Parent, q.Child1, q.Child2 = int(parent[i]), int(child1[i]), int(child2[i])
		q.Area, q.Chop, q.PlaneDist = float32(area[i]), float32(chop[i]), float32(dist[i])
		for k := 0; k < 3; k++ {
			q.Origin[k], q.Mins[k], q.Maxs[k] = float32(origin[3*i+k]), float32(mins[3*i+k]), float32(maxs[3*i+k])
		}
	}
	*cache.GetPatches() = out
}
