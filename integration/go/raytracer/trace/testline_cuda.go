// testline_cuda.go -- cgo bodies for raytracer/trace (testline.go, pointleaf.go) and the sky-camera /
// BSP uploads they depend on.  SOURCE ONLY (no Go toolchain in the build image); the same calls are
// exercised through ctypes by tests/test_gpu_sky.py.
//
//go:build cuda

package trace

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../../vrad_b200/_lib -lvradcuda
#include "vrad_cuda.h"
*/
import "C"

import (
	"log"
	"unsafe"

	"github.com/galaco/vrad/cache"
	"github.com/galaco/vrad/raytracer"
	"github.com/galaco/vrad/vmath/ssemath"
	"github.com/galaco/vrad/vmath/ssemath/simd"
	"github.com/go-gl/mathgl/mgl32"
)

func check(rc C.int, what string) {
	if rc != 0 {
		log.Fatalf("%s: vrad status %d: %s", what, int(rc), C.GoString(C.vrad_last_error()))
	}
}

func handle() *C.vrad_env { return (*C.vrad_env)(raytracer.GetEnvironment().CudaHandle()) }

// UploadBSP hands the lumps PointLeafnum walks (pointleaf.go:8-33) to the library; call once after loadbsp.
func UploadBSP() {
	lumps := cache.GetLumpCache()
	nodePlane := make([]C.int32_t, len(lumps.Nodes))
	children := make([]C.int32_t, 2*len(lumps.Nodes))
	for i, n := range lumps.Nodes {
		nodePlane[i] = C.int32_t(n.PlaneNum)
		children[2*i], children[2*i+1] = C.int32_t(n.Children[0]), C.int32_t(n.Children[1])
	}
	normal := make([]C.float, 3*len(lumps.Planes))
	dist := make([]C.float, len(lumps.Planes))
	ptype := make([]C.int32_t, len(lumps.Planes))
	for i, p := range lumps.Planes {
		normal[3*i], normal[3*i+1], normal[3*i+2] = C.float(p.Normal[0]), C.float(p.Normal[1]), C.float(p.Normal[2])
		dist[i], ptype[i] = C.float(p.Distance), C.int32_t(p.AxisType)
	}
	cluster := make([]C.int32_t, len(lumps.Leafs))
	area := make([]C.int32_t, len(lumps.Leafs))
	for i, l := range lumps.Leafs {
		cluster[i], area[i] = C.int32_t(l.Cluster), C.int32_t(l.Area())
	}
	check(C.vrad_bsp_upload(handle(), C.int(len(nodePlane)), &nodePlane[0], &children[0], C.int(len(dist)), &normal[0], &dist[0], &ptype[0],
		C.int(len(cluster)), &cluster[0], &area[0], C.int(len(lumps.Areas))), "vrad_bsp_upload")
}

// PointLeafnum (pointleaf.go:8-10)
func PointLeafnum(point *mgl32.Vec3) int {
	var leaf C.int32_t
	check(C.vrad_point_leafnum(handle(), 1, (*C.float)(unsafe.Pointer(&point[0])), &leaf), "vrad_point_leafnum")
	return int(leaf)
}

// ProcessSkyCameras (rad/cameras/skycamera.go:10-49) collects (origin, scale) of every sky_camera entity and calls this.
func SetSkyCameras(origins []mgl32.Vec3, scales []float32) int {
	var kept C.int
	check(C.vrad_sky_cameras_set(handle(), C.int(len(scales)), (*C.float)(unsafe.Pointer(&origins[0])),
		(*C.float)(unsafe.Pointer(&scales[0])), &kept), "vrad_sky_cameras_set")
	return int(kept)
}

// TestLineDoesHitSky (testline.go:18-94): the panic at :21 and the body below it become one call.
func TestLineDoesHitSky(start *ssemath.FourVectors, stop ssemath.FourVectors,
	fractionVisible *simd.Flt4x, canRecurse bool, staticPropToSkip int, doDebug bool) {
	var a, b [12]C.float
	for l := 0; l < 4; l++ {
		a[l], a[4+l], a[8+l] = C.float(start.X[l]), C.float(start.Y[l]), C.float(start.Z[l])
		b[l], b[4+l], b[8+l] = C.float(stop.X[l]), C.float(stop.Y[l]), C.float(stop.Z[l])
	}
	flags := C.int(C.VRAD_TL_PACKET_LEAF) // the leaf comes from lane 0 (start.Vec(0), :63)
	if canRecurse && !noSkyRecurse {
		flags |= C.VRAD_TL_CAN_RECURSE
	}
	if textureShadows {
		flags |= C.VRAD_TL_TEXTURE_SHADOWS
	}
	var fv [4]C.float
	check(C.vrad_test_lines_sky(handle(), 4, &a[0], &b[0], flags, C.int32_t(staticPropToSkip), &fv[0]), "vrad_test_lines_sky")
	for l := 0; l < 4; l++ {
		fractionVisible[l] = float32(fv[l])
	}
}

// TestLinesSky is the batched form (n segments, SoA x[n] y[n] z[n]) the lighting stages should use.
func TestLinesSky(start, stop []float32, n int, flags int, staticPropToSkip int, fractionVisible []float32) {
	check(C.vrad_test_lines_sky(handle(), C.int64_t(n), (*C.float)(unsafe.Pointer(&start[0])), (*C.float)(unsafe.Pointer(&stop[0])),
		C.int(flags), C.int32_t(staticPropToSkip), (*C.float)(unsafe.Pointer(&fractionVisible[0]))), "vrad_test_lines_sky")
}
