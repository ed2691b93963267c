// pairedges_cuda.go -- cgo bodies for lightmap.PairEdges, SaveVertexNormals, GetPhongNormal and
// BuildVisForLightEnvironment (rad/lightmap/lightmap.go:37-265, 284-397; normallist.go:52-143): host code in
// libvradcuda.so (vrad_b200/csrc/bsp_light.cpp, bsp_input.cpp); only the LEAF_FLAGS_RADIAL branch of the sky vis
// (CanLeafTraceToSky, lightmap.go:373-381) reaches the GPU.  SOURCE ONLY (no Go toolchain in the build image).
//
//go:build cuda

package lightmap

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../../vrad_b200/_lib -lvradcuda
#include "vrad_bsp.h"
*/
import "C"

import (
	"log"
	"unsafe"

	"github.com/galaco/bsp/primitives/vertnormal"
	"github.com/galaco/vrad/cache"
	"github.com/galaco/vrad/raytracer"
	"github.com/go-gl/mathgl/mgl32"
)

var cudaVertexNormals []C.float // per face-vertex, face order: what faceNeighbour[i].Normal holds in the reference
var cudaNormalFirst []int       // offset of each face's block

func fatal(what string, rc C.int) {
	if rc != 0 {
		log.Fatalf("%s: %s", what, C.GoString(C.vrad_last_error()))
	}
}

func PairEdgesCUDA(lumps *cache.CLumps) {
	faces := *cache.GetTargetFaces()
	total := 0
	cudaNormalFirst = make([]int, len(faces)+1)
	for i := range faces {
		cudaNormalFirst[i] = total
		total += int(faces[i].NumEdges)
	}
	cudaNormalFirst[len(faces)] = total
	cudaVertexNormals = make([]C.float, 3*total+3)
	first := make([]C.int32_t, len(faces)+1)
	nb := make([]C.int32_t, 64*len(faces)+1)
	fatal("vrad_bsp_pair_edges", C.vrad_bsp_pair_edges(&lumps.L, C.float(smoothingThreshold), &cudaVertexNormals[0], &first[0], &nb[0], C.int(len(nb))))
	for i := range faces { // faceNeighbour, as the rest of the Go code reads it
		fn := &faceNeighbour[i]
		fn.FaceNormal = cache.GetLumpCache().Planes[faces[i].Planenum].Normal
		fn.Normal = make([]mgl32.Vec3, faces[i].NumEdges)
		for j := range fn.Normal {
			k := 3 * (cudaNormalFirst[i] + j)
			fn.Normal[j] = mgl32.Vec3{float32(cudaVertexNormals[k]), float32(cudaVertexNormals[k+1]), float32(cudaVertexNormals[k+2])}
		}
		fn.NumNeighbours = int(first[i+1] - first[i])
		fn.Neighbour = make([]int, fn.NumNeighbours)
		for m := range fn.Neighbour {
			fn.Neighbour[m] = int(nb[int(first[i])+m])
		}
	}
}

func SaveVertexNormalsCUDA() {
	total := cudaNormalFirst[len(cudaNormalFirst)-1]
	normals := make([]C.float, 3*total+3)
	indices := make([]C.uint16_t, total+1)
	var n C.int
	fatal("vrad_bsp_save_vertex_normals", C.vrad_bsp_save_vertex_normals(C.int(total), &cudaVertexNormals[0], C.int(total), &normals[0], &indices[0], &n))
	lc := cache.GetLumpCache()
	lc.VertNormals = make([]vertnormal.VertNormal, int(n))
	for i := range lc.VertNormals {
		lc.VertNormals[i].Pos = mgl32.Vec3{float32(normals[3*i]), float32(normals[3*i+1]), float32(normals[3*i+2])}
	}
	computedNumVertNormalIndices = total
	for i := 0; i < total; i++ {
		lc.VertNormalIndices[i] = uint16(indices[i])
	}
}

// GetPhongNormalsCUDA is the batched GetPhongNormal: one call for all child patches of a subdivision pass.
func GetPhongNormalsCUDA(lumps *cache.CLumps, faceNum []int32, spots []mgl32.Vec3) []mgl32.Vec3 {
	n := len(faceNum)
	out := make([]mgl32.Vec3, n)
	if n == 0 {
		return out
	}
	cent := make([]C.float, 3*len(*cache.GetTargetFaces()))
	for i, c := range cache.GetFaceCentroids()[:len(*cache.GetTargetFaces())] {
		cent[3*i], cent[3*i+1], cent[3*i+2] = C.float(c[0]), C.float(c[1]), C.float(c[2])
	}
	fatal("vrad_bsp_phong_normals", C.vrad_bsp_phong_normals(&lumps.L, C.float(smoothingThreshold), &cudaVertexNormals[0], &cent[0],
		C.int64_t(n), (*C.int32_t)(unsafe.Pointer(&faceNum[0])), (*C.float)(unsafe.Pointer(&spots[0])), (*C.float)(unsafe.Pointer(&out[0]))))
	return out
}

func BuildVisForLightEnvironmentCUDA(lumps *cache.CLumps) {
	lc := cache.GetLumpCache()
	flags := make([]C.uint8_t, len(lc.Leafs)+1)
	row := (int(lc.Visibility.NumClusters) + 7) / 8
	pvs := make([]C.uint8_t, row+1)
	var has C.int
	env := (*C.vrad_env)(unsafe.Pointer(raytracer.GetEnvironment().CudaHandle()))
	fatal("vrad_bsp_vis_for_light_environment", C.vrad_bsp_vis_for_light_environment(env, &lumps.L, &flags[0], &pvs[0], &has))
	for i := range lc.Leafs {
		lc.Leafs[i].SetFlags(int(flags[i]))
	}
	if has != 0 { // MergeDLightVis on both sky lights (lightmap.go:305-306)
		merged := C.GoBytes(unsafe.Pointer(&pvs[0]), C.int(row))
		if globalSkyLight != nil {
			globalSkyLight.PVS = append([]byte(nil), merged...)
		}
		if globalAmbient != nil {
			globalAmbient.PVS = append([]byte(nil), merged...)
		}
	}
}
