// radworld_cuda.go -- what `RadWorld_Go()` (commented out at cmd/tasks/computerad/main.go:7) becomes
// when the radiosity stages run in libvradcuda.  SOURCE ONLY (no Go toolchain in the build image).
//
//go:build cuda

package rad

/*
#include "vrad_cuda.h"
*/
import "C"

import (
	"log"
	"unsafe"

	"github.com/galaco/vrad/cache"
	"github.com/galaco/vrad/common/types"
)

func fatal(rc C.int, what string) {
	if rc != 0 {
		log.Fatalf("%s: vrad status %d: %s", what, int(rc), C.GoString(C.vrad_last_error()))
	}
}

// RadWorldCUDA: BuildFacelights (K3) -> transfers (K2) -> BounceLight (K4) on the leaf patches that
// rad.Start (rad/start.go:21-100) left in cache.GetPatches().
func RadWorldCUDA(env *C.vrad_env, luxelPos, luxelNormal []float32, lights []C.vrad_light, numBounce int) (lightmap, bounced []float32) {
	patches := *cache.GetPatches()
	n := 0
	var origin, normal, refl, planeDist, area []float32
	var cluster []int32
	var flags []uint8
	for i := range patches {
		p := &patches[i]
		if p.Child1 != -1 { // leaf patches only (common/types/patch.go:49-51)
			continue
		}
		origin = append(origin, p.Origin[0], p.Origin[1], p.Origin[2])
		normal = append(normal, p.Normal[0], p.Normal[1], p.Normal[2])
		refl = append(refl, p.Reflectivity[0], p.Reflectivity[1], p.Reflectivity[2])
		planeDist = append(planeDist, p.PlaneDist)
		area = append(area, p.Area)
		cluster = append(cluster, int32(p.ClusterNumber))
		f := uint8(0)
		if p.Sky {
			f = 1
		}
		flags = append(flags, f)
		n++
	}
	fatal(C.vrad_patches_upload(env, C.int(n), (*C.float)(&origin[0]), (*C.float)(&normal[0]), (*C.float)(&planeDist[0]),
		(*C.float)(&area[0]), (*C.float)(&refl[0]), (*C.int32_t)(&cluster[0]), (*C.uint8_t)(&flags[0])), "vrad_patches_upload")

	// K3: direct light per luxel
	nLux := len(luxelPos) / 3
	lightmap = make([]float32, 3*nLux)
	fatal(C.vrad_direct_light(env, C.int64_t(nLux), (*C.float)(&luxelPos[0]), (*C.float)(&luxelNormal[0]),
		C.int(len(lights)), &lights[0], (*C.float)(&lightmap[0])), "vrad_direct_light")

	// K2: PVS bytes come from rad/lightmap/vis.go (DecompressVis) expanded to one byte per cluster pair
	var nnz C.int64_t
	fatal(C.vrad_build_transfers(env, 0, nil, &nnz), "vrad_build_transfers")

	// K4: emit0 = direct light averaged onto patches (Patch.DirectLight); result -> Patch.TotalLight
	emit0 := make([]float32, 3*n)
	bounced = make([]float32, 3*n)
	var added [3]C.float
	var done C.int
	fatal(C.vrad_bounce(env, (*C.float)(&emit0[0]), C.int(numBounce), 1, (*C.float)(&bounced[0]), &added[0], &done), "vrad_bounce")
	log.Printf("%d bounces, last added RGB(%.0f, %.0f, %.0f), %d transfers", int(done), float32(added[0]), float32(added[1]), float32(added[2]), int64(nnz))
	return
}

var _ = types.Transfer{}
var _ = unsafe.Pointer(nil)
