// radworld_cuda.go -- what `RadWorld_Go()` (commented out at cmd/tasks/computerad/main.go:7) becomes
// when the radiosity stages run in libvradcuda.  SOURCE ONLY (no Go toolchain in the build image); the same
// sequence runs through ctypes in tests/test_gpu_pipeline.py.
//
// Order (rad/start.go:21-100 has done: MakePatches, cameras.ProcessSkyCameras, patches.SubdividePatchesCUDA):
//   lights  : CreateDirectLights          -> vrad_lights_from_patches + vrad_lights_from_entities
//   PVS     : lightmap.GetVisCache        -> vrad_pvs_from_vis_lump
//   patches : cache.GetPatches()          -> vrad_patches_upload + vrad_patches_set_hierarchy (+ vrad_patches_set_bump)
//   K3      : BuildFacelights             -> vrad_direct_light (light rays through the complete TestLineDoesHitSky)
//   K2      : BuildVisMatrix/MakeScales   -> vrad_build_transfers (hierarchical)
//   K4      : BounceLight                 -> vrad_bounce (+ vrad_bounce_bump_totals) -> Patch.TotalLight
//
//go:build cuda

package rad

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../vrad_b200/_lib -lvradcuda
#include "vrad_cuda.h"
*/
import "C"

import (
	"log"

	"github.com/galaco/vrad/cache"
	"github.com/galaco/vrad/raytracer"
)

func fatal(rc C.int, what string) {
	if rc != 0 {
		log.Fatalf("%s: vrad status %d: %s", what, int(rc), C.GoString(C.vrad_last_error()))
	}
}

// LightEntities is filled by the caller from cache.GetAllEntities() with Entity.FloatForKey / VectorForKey /
// LightForKey (C.vrad_light_for_string), one record per "light*" entity (lights.go:90-113).
// PolygonFormFactor switches MakeTransfer to the polygon-to-differential form factor for emitters that are large for their
// distance (upstream's rule; off = the differential form everywhere).
var PolygonFormFactor = false

func RadWorldCUDA(luxelPos, luxelNormal []float32, lightEntities []C.vrad_light_entity, numBounce int) (lightmap []float32) {
	env := (*C.vrad_env)(raytracer.GetEnvironment().CudaHandle())
	patches := *cache.GetPatches()
	n := len(patches)

	// ---- patches, with the Parent/Child links SubdividePatches left (common/types/patch.go:33,49-51) ----
	origin, normal, refl := make([]C.float, 3*n), make([]C.float, 3*n), make([]C.float, 3*n)
	planeDist, area := make([]C.float, n), make([]C.float, n)
	cluster, parent, child1, child2, face := make([]C.int32_t, n), make([]C.int32_t, n), make([]C.int32_t, n), make([]C.int32_t, n), make([]C.int32_t, n)
	flags, needsBump := make([]C.uint8_t, n), make([]C.uint8_t, n)
	baseLight, scale2, baseArea := make([]C.float, 3*n), make([]C.float, 2*n), make([]C.float, n)
	bumpNormals := make([]C.float, 9*n)
	anyBump := false
	for i := range patches {
		p := &patches[i]
		for k := 0; k < 3; k++ {
			origin[3*i+k], normal[3*i+k], refl[3*i+k] = C.float(p.Origin[k]), C.float(p.Normal[k]), C.float(p.Reflectivity[k])
			baseLight[3*i+k] = C.float(p.BaseLight[k])
		}
		planeDist[i], area[i], baseArea[i] = C.float(p.PlaneDist), C.float(p.Area), C.float(p.BaseArea)
		scale2[2*i], scale2[2*i+1] = C.float(p.Scale[0]), C.float(p.Scale[1])
		cluster[i] = C.int32_t(p.ClusterNumber)
		parent[i], child1[i], child2[i], face[i] = C.int32_t(p.Parent), C.int32_t(p.Child1), C.int32_t(p.Child2), C.int32_t(p.FaceNumber)
		if p.Sky {
			flags[i] = 1
		}
		if p.NeedsBumpMap { // face.go:66-70 (SURF_BUMPLIGHT)
			needsBump[i] = 1
			anyBump = true
			tx := &cache.GetLumpCache().TexInfo[(*cache.GetTargetFaces())[p.FaceNumber].TexInfo]
			var s, t, flat [3]C.float
			for k := 0; k < 3; k++ {
				s[k], t[k] = C.float(tx.TextureVecsTexelsPerWorldUnits[0][k]), C.float(tx.TextureVecsTexelsPerWorldUnits[1][k])
				flat[k] = C.float(p.Plane.Normal[k])
			}
			fatal(C.vrad_bump_normals(&s[0], &t[0], &flat[0], &normal[3*i], &bumpNormals[9*i]), "vrad_bump_normals")
		}
	}
	fatal(C.vrad_patches_upload(env, C.int(n), &origin[0], &normal[0], &planeDist[0], &area[0], &refl[0], &cluster[0], &flags[0]), "vrad_patches_upload")
	fatal(C.vrad_patches_set_hierarchy(env, C.int(n), &parent[0], &child1[0], &child2[0], &face[0]), "vrad_patches_set_hierarchy")
	if anyBump {
		fatal(C.vrad_patches_set_bump(env, C.int(n), &needsBump[0], &bumpNormals[0]), "vrad_patches_set_bump")
	}
	if PolygonFormFactor { // Patch.Winding (common/types/patch.go:10): MakeTransfer integrates over the emitter's polygon for near pairs
		windFirst, windCount := make([]C.int32_t, n), make([]C.int32_t, n)
		var windPoints []C.float
		for i := range patches {
			windFirst[i] = C.int32_t(len(windPoints) / 3)
			if w := patches[i].Winding; w != nil {
				windCount[i] = C.int32_t(w.NumPoints)
				for q := 0; q < w.NumPoints; q++ {
					windPoints = append(windPoints, C.float(w.Points[q][0]), C.float(w.Points[q][1]), C.float(w.Points[q][2]))
				}
			}
		}
		if len(windPoints) > 0 {
			fatal(C.vrad_patches_set_windings(env, C.int(n), &windFirst[0], &windCount[0], C.int(len(windPoints)/3), &windPoints[0]), "vrad_patches_set_windings")
		}
	}

	// ---- lights: surface lights from the emitting leaf patches, then the light entities (lights.go:49-113) ----
	lights := make([]C.vrad_light, n+2*len(lightEntities)+1)
	var nSurf, nEnt C.int
	fatal(C.vrad_lights_from_patches(C.int(n), &origin[0], &normal[0], &baseLight[0], &area[0], &scale2[0], &baseArea[0], &child1[0],
		0.1, C.int(n), &lights[0], &nSurf), "vrad_lights_from_patches")
	if len(lightEntities) > 0 {
		fatal(C.vrad_lights_from_entities(C.int(len(lightEntities)), &lightEntities[0], C.int(len(lights))-nSurf, &lights[nSurf], &nEnt), "vrad_lights_from_entities")
	}

	// ---- PVS: the visibility lump, run-length decoded (rad/lightmap/vis.go:9-94) ----
	vis := cache.GetLumpCache().Visibility
	nc := int(vis.NumClusters)
	ofs := make([]C.int32_t, 2*nc)
	for c := 0; c < nc; c++ {
		ofs[2*c], ofs[2*c+1] = C.int32_t(vis.ByteOffset[c][0]), C.int32_t(vis.ByteOffset[c][1])
	}
	raw := cache.GetLumpCache().VisDataRaw
	pvs := make([]C.uint8_t, nc*nc)
	fatal(C.vrad_pvs_from_vis_lump(C.int(nc), &ofs[0], (*C.uint8_t)(&raw[0]), C.int64_t(len(raw)), &pvs[0]), "vrad_pvs_from_vis_lump")

	// ---- K3: direct light per luxel; sky lights clip into the 3D sky boxes like CanLeafTraceToSky's calls (lightmap.go:444) ----
	fatal(C.vrad_set_light_trace_flags(env, C.VRAD_TL_CAN_RECURSE), "vrad_set_light_trace_flags")
	nLux := len(luxelPos) / 3
	lightmap = make([]float32, 3*nLux)
	fatal(C.vrad_direct_light(env, C.int64_t(nLux), (*C.float)(&luxelPos[0]), (*C.float)(&luxelNormal[0]),
		nSurf+nEnt, &lights[0], (*C.float)(&lightmap[0])), "vrad_direct_light")

	// ---- K2 + K4: transfers, then BounceLight; emit0 = Patch.DirectLight (the direct light averaged onto the patches) ----
	var nnz C.int64_t
	fatal(C.vrad_build_transfers(env, C.int(nc), &pvs[0], &nnz), "vrad_build_transfers")
	emit0, bounced := make([]C.float, 3*n), make([]C.float, 3*n)
	for i := range patches {
		for k := 0; k < 3; k++ {
			emit0[3*i+k] = C.float(patches[i].DirectLight[k])
		}
	}
	var added [3]C.float
	var done C.int
	fatal(C.vrad_bounce(env, &emit0[0], C.int(numBounce), 1, &bounced[0], &added[0], &done), "vrad_bounce")
	bump := make([]C.float, 9*n)
	if anyBump {
		fatal(C.vrad_bounce_bump_totals(env, &bump[0]), "vrad_bounce_bump_totals")
	}
	for i := range patches { // Patch.TotalLight = BumpLights{Light[0] flat, Light[1..3] bump} (common/types/bumpLights.go:8-10)
		for k := 0; k < 3; k++ {
			patches[i].TotalLight.Light[0][k] = float32(bounced[3*i+k])
			for b := 0; b < 3; b++ {
				patches[i].TotalLight.Light[b+1][k] = float32(bump[9*i+3*b+k])
			}
		}
	}
	log.Printf("%d bounces, last added RGB(%.0f, %.0f, %.0f), %d transfers, %d lights", int(done), float32(added[0]), float32(added[1]),
		float32(added[2]), int64(nnz), int(nSurf+nEnt))
	return
}
