// raytrace_cuda.go -- cgo body for the "Setup ray tracer" block of loadbsp.Main (cmd/tasks/loadbsp/main.go:97-150):
// ExtractBrushEntityShadowCasters + addBrushesForRayTrace + SetupAccelerationStructure become two library calls.
// The windings (vmath/polygon/winding.go), GetBrushRecursive (brush/brush.go) and the matrix (vmath/matrix/mat4.go) run as
// host code inside libvradcuda.so (vrad_b200/csrc/bsp_input.cpp), the SAH build and upload in vrad_env_build.
// SOURCE ONLY (no Go toolchain in the build image).
//
//go:build cuda

package loadbsp

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../../../vrad_b200/_lib -lvradcuda
#include "vrad_bsp.h"
*/
import "C"

import (
	"log"
	"strconv"
	"strings"
	"unsafe"

	"github.com/galaco/vmf"
	"github.com/galaco/vrad/cache"
	"github.com/galaco/vrad/raytracer"
)

func vec3(s string) [3]C.float {
	var out [3]C.float
	for i, f := range strings.Fields(s) {
		if i < 3 {
			v, _ := strconv.ParseFloat(f, 32)
			out[i] = C.float(v)
		}
	}
	return out
}

// SetupRayTracerCUDA replaces main.go:97 (ExtractBrushEntityShadowCasters), :133 (addBrushesForRayTrace) and :148
// (SetupAccelerationStructure).  lumps = cache.BuildCLumps(), kept for rad.Start and the finish task.
func SetupRayTracerCUDA(entities *vmf.Node, lumps *cache.CLumps) {
	var model []C.int32_t
	var origin, angles []C.float
	for _, iEntity := range *entities.GetAllValues() {
		entity := iEntity.(vmf.Node)
		if !entity.HasProperty("vrad_brush_cast_shadows") {
			continue
		}
		name := entity.GetProperty("model") // "*N"
		if len(name) < 2 || name[0] != '*' {
			continue
		}
		n, _ := strconv.Atoi(name[1:])
		o, a := vec3(entity.GetProperty("origin")), vec3(entity.GetProperty("angles"))
		model = append(model, C.int32_t(n))
		origin = append(origin, o[0], o[1], o[2])
		angles = append(angles, a[0], a[1], a[2])
	}
	var pm *C.int32_t
	var po, pa *C.float
	if len(model) > 0 {
		pm, po, pa = &model[0], &origin[0], &angles[0]
	}
	env := (*C.vrad_env)(unsafe.Pointer(raytracer.GetEnvironment().CudaHandle()))
	var added C.int
	if rc := C.vrad_env_add_bsp(env, &lumps.L, C.int(len(model)), pm, po, pa, &added); rc != 0 {
		log.Fatalf("vrad_env_add_bsp: %s", C.GoString(C.vrad_last_error()))
	}
	log.Printf("%d triangles in the ray-trace environment\n", int(added))
	if rc := C.vrad_env_build(env); rc != 0 {
		log.Fatalf("vrad_env_build: %s", C.GoString(C.vrad_last_error()))
	}
}
