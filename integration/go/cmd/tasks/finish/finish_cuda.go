// finish_cuda.go -- cgo body for the finish task (cmd/tasks/finish/main.go:8-40), whose "Writing %s" step is commented out in
// the reference: lay the lighting lump out, pack the final light per luxel on the GPU (K5), and write the .bsp with
// LUMP_LIGHTING, the face lump (lightofs / styles / extents) and the vertex-normal lumps replaced.
// SOURCE ONLY (no Go toolchain in the build image).
//
//go:build cuda

package finish

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../../../vrad_b200/_lib -lvradcuda
#include <stdlib.h>
#include "vrad_bsp.h"
*/
import "C"

import (
	"log"
	"unsafe"

	"github.com/galaco/vrad/cache"
	"github.com/galaco/vrad/raytracer"
)

func fatal(what string, rc C.int) {
	if rc != 0 {
		log.Fatalf("%s: %s", what, C.GoString(C.vrad_last_error()))
	}
}

// RadialIndirectCUDA: the bounced light of every luxel by the radial filter (vrad_bsp_radial_entries on the host, vrad_luxel_radial_light on the
// GPU).  patchFace / child1 / origin / windFirst / windCount / windPoints describe cache.GetPatches() as rad/radworld_cuda.go uploads them;
// nbFirst / nb are lightmap.PairEdgesCUDA's neighbour lists; mins / size come from vrad_bsp_face_extents.
func RadialIndirectCUDA(lumps *cache.CLumps, litFaces []C.vrad_dface, mins, size []C.int32_t, faceOrigins []C.float, luxelFace []C.int32_t, luxelFirst []C.int64_t,
	patchFace, child1 []C.int32_t, origin []C.float, windFirst, windCount []C.int32_t, windPoints []C.float, nbFirst, nb []C.int32_t,
	patchTotal []C.float, patchBump []C.float) []C.float {
	env := (*C.vrad_env)(unsafe.Pointer(raytracer.GetEnvironment().CudaHandle()))
	view := lumps.L
	view.faces = &litFaces[0]
	nFaces, nPatches, n := len(litFaces), len(patchFace), int(luxelFirst[len(luxelFirst)-1])
	entryFirst := make([]C.int64_t, nFaces+1)
	var nEntries C.int64_t
	var pnb *C.int32_t
	if len(nb) > 0 {
		pnb = &nb[0]
	}
	fatal("vrad_bsp_radial_entries", C.vrad_bsp_radial_entries(&view, &mins[0], &faceOrigins[0], C.int(nPatches), &patchFace[0], &child1[0], &origin[0],
		&windFirst[0], &windCount[0], &windPoints[0], &nbFirst[0], pnb, 0, &entryFirst[0], nil, &nEntries))
	entries := make([]C.vrad_radial_entry, int(nEntries)+1)
	fatal("vrad_bsp_radial_entries", C.vrad_bsp_radial_entries(&view, &mins[0], &faceOrigins[0], C.int(nPatches), &patchFace[0], &child1[0], &origin[0],
		&windFirst[0], &windCount[0], &windPoints[0], &nbFirst[0], pnb, nEntries, &entryFirst[0], &entries[0], &nEntries))
	indirect := make([]C.float, 3*n+3)
	var bump *C.float
	if len(patchBump) > 0 {
		bump = &patchBump[0]
	}
	fatal("vrad_luxel_radial_light", C.vrad_luxel_radial_light(env, C.int64_t(n), &luxelFace[0], C.int(nFaces), &luxelFirst[0], &size[0], &entryFirst[0], &entries[0],
		C.int(nPatches), &patchTotal[0], bump, &indirect[0]))
	return indirect
}

// WriteLightingCUDA: direct = per-luxel RGB from vrad_direct_light (luxel order of vrad_bsp_face_luxels), indirect = RadialIndirectCUDA's result (or nil:
// luxelPatch from vrad_luxel_nearest_patch + patchTotal = vrad_bounce's totals, the piecewise-constant form).  inPath is re-read so that every lump the
// program never touched goes back byte for byte.
func WriteLightingCUDA(inPath, outPath string, lumps *cache.CLumps, luxelFirst []C.int64_t, litFaces []C.vrad_dface,
	direct []C.float, indirect []C.float, luxelPatch []C.int32_t, patchTotal []C.float) {
	env := (*C.vrad_env)(unsafe.Pointer(raytracer.GetEnvironment().CudaHandle()))
	n := int(luxelFirst[len(luxelFirst)-1])
	colors := make([]C.vrad_color_rgbexp32, n+1)
	if indirect != nil {
		fatal("vrad_lightmap_finalize", C.vrad_lightmap_finalize(env, C.int64_t(n), &direct[0], &indirect[0], &colors[0]))
	} else {
		fatal("vrad_lightmap_finalize_patches", C.vrad_lightmap_finalize_patches(env, C.int64_t(n), &direct[0], &luxelPatch[0],
			C.int(len(patchTotal)/3), &patchTotal[0], &colors[0]))
	}

	view := lumps.L // the face records laid out by vrad_bsp_layout_lighting carry the offsets
	view.faces = &litFaces[0]
	var lumpBytes C.int64_t
	for i := range litFaces {
		if end := C.int64_t(litFaces[i].lightofs) + 4*(luxelFirst[i+1]-luxelFirst[i]); litFaces[i].lightofs >= 0 && end > lumpBytes {
			lumpBytes = end
		}
	}
	lump := C.malloc(C.size_t(lumpBytes) + 1)
	defer C.free(lump)
	fatal("vrad_bsp_pack_lighting", C.vrad_bsp_pack_lighting(&view, &luxelFirst[0], &colors[0], (*C.uint8_t)(lump), lumpBytes))

	var f *C.vrad_bspfile
	cin, cout := C.CString(inPath), C.CString(outPath)
	defer C.free(unsafe.Pointer(cin))
	defer C.free(unsafe.Pointer(cout))
	fatal("vrad_bspfile_open", C.vrad_bspfile_open(cin, &f))
	defer C.vrad_bspfile_close(f)
	fatal("set lighting", C.vrad_bspfile_set_lump(f, C.VRAD_LUMP_LIGHTING, lump, lumpBytes, 1))
	fatal("set faces", C.vrad_bspfile_set_lump(f, C.VRAD_LUMP_FACES, unsafe.Pointer(&litFaces[0]), C.int64_t(len(litFaces))*C.sizeof_vrad_dface, 1))
	lc := cache.GetLumpCache()
	if len(lc.VertNormals) > 0 { // SaveVertexNormals' output, "for use in the engine" (rad/start.go:82-85)
		vn := make([]C.float, 3*len(lc.VertNormals))
		for i := range lc.VertNormals {
			vn[3*i], vn[3*i+1], vn[3*i+2] = C.float(lc.VertNormals[i].Pos[0]), C.float(lc.VertNormals[i].Pos[1]), C.float(lc.VertNormals[i].Pos[2])
		}
		fatal("set vertnormals", C.vrad_bspfile_set_lump(f, C.VRAD_LUMP_VERTNORMALS, unsafe.Pointer(&vn[0]), C.int64_t(4*len(vn)), 0))
	}
	log.Printf("Writing %s\n", outPath) // finish/main.go:15
	fatal("vrad_bspfile_save", C.vrad_bspfile_save(f, cout))
}
