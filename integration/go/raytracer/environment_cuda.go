// environment_cuda.go -- cgo binding a VRADiant maintainer would add next to raytracer/environment.go.
//
// SOURCE ONLY: the build image has no Go toolchain, so this file has not been compiled.  It shows
// the exact call-for-call mapping between raytracer.Environment's methods and libvradcuda's C-ABI
// (include/vrad_cuda.h).  The same mapping, executed through ctypes, is what tests/ exercise.
//
// Build: CGO_CFLAGS="-I${VRAD_B200}/include" CGO_LDFLAGS="-L${VRAD_B200}/vrad_b200/_lib -lvradcuda" go build -tags cuda
//
//go:build cuda

package raytracer

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../vrad_b200/_lib -lvradcuda
#include <stdlib.h>
#include "vrad_cuda.h"
*/
import "C"

import (
	"log"
	"os"
	"strconv"
	"strings"
	"unsafe"

	"github.com/galaco/vrad/raytracer/types"
	"github.com/galaco/vrad/vmath/ssemath/simd"
	"github.com/go-gl/mathgl/mgl32"
)

// cudaEnv is stored inside Environment (new field `cuda *cudaEnv`); GetEnvironment() keeps its singleton.
type cudaEnv struct {
	h *C.vrad_env
}

func check(rc C.int, what string) {
	if rc != 0 {
		// the reference has no error returns: failures are fatal (cf. raytracer/trace/testline.go:21)
		log.Fatalf("%s: vrad status %d: %s", what, int(rc), C.GoString(C.vrad_last_error()))
	}
}

// CudaHandle exposes the handle to the sibling packages' shims (raytracer/trace, rad/patches, rad).
func (environment *Environment) CudaHandle() unsafe.Pointer { return unsafe.Pointer(environment.cuda.h) }

// newCudaEnv makes ONE handle over every GPU in `devices` (vrad_env_create_multi): the driver is a single goroutine behind a
// package-level singleton (environment.go:17-25; common/constants/constants.go:43), so the devices live inside the handle --
// batches are split and the patch rows sharded there, and every call on it takes Go (host) memory and returns when it is done.
// One device: the plain single-GPU handle.  VRAD_DEVICES="0,1,2,3" selects the devices; the default is device 0.
func newCudaEnv(devices []int) *cudaEnv {
	var h *C.vrad_env
	if len(devices) <= 1 {
		d := 0
		if len(devices) == 1 {
			d = devices[0]
		}
		cfg := C.vrad_config{device: C.int(d), rank: 0, world: 1, flags: 0}
		check(C.vrad_env_create(&cfg, &h), "vrad_env_create")
		return &cudaEnv{h: h}
	}
	var cfg C.vrad_multi_config
	cfg.n_devices = C.int(len(devices))
	for i, d := range devices {
		if i < 8 {
			cfg.devices[i] = C.int(d)
		}
	}
	check(C.vrad_env_create_multi(&cfg, &h), "vrad_env_create_multi")
	return &cudaEnv{h: h}
}

// devicesFromEnv parses VRAD_DEVICES ("0,1,2,3"); empty or unset = device 0.
func devicesFromEnv() []int {
	var out []int
	for _, f := range strings.Split(os.Getenv("VRAD_DEVICES"), ",") {
		if d, err := strconv.Atoi(strings.TrimSpace(f)); err == nil {
			out = append(out, d)
		}
	}
	if len(out) == 0 {
		out = []int{0}
	}
	return out
}

// AddTriangleWithMaterial (environment.go:45-69) keeps appending to OptimizedTriangleList on the Go
// side; the triangles are handed to the library in one batch when the tree is built.
// SetupAccelerationStructureFastCUDA is SetupAccelerationStructure with RTE_FLAGS_FAST_TREE_GENERATION (raytracer/constants.go:5) set in
// environment.Flags: the binned-SAH tree is built on the GPU from the triangles already handed over with vrad_env_add_triangles /
// vrad_env_add_bsp (vrad_env_build_fast, include/vrad_cuda.h).
func (environment *Environment) SetupAccelerationStructureFastCUDA() {
	check(C.vrad_env_build_fast(environment.cuda.h, C.VRAD_BUILD_AUTO), "vrad_env_build_fast")
}

func (environment *Environment) SetupAccelerationStructureCUDA() {
	n := len(environment.OptimizedTriangleList)
	ids := make([]C.int32_t, n)
	verts := make([]C.float, 9*n)
	flags := make([]C.uint8_t, n)
	for i := range environment.OptimizedTriangleList {
		g := &environment.OptimizedTriangleList[i].TriGeometryData
		ids[i] = C.int32_t(g.NTriangleID)
		for k := 0; k < 9; k++ {
			verts[9*i+k] = C.float(g.VertexCoordData[k])
		}
		flags[i] = C.uint8_t(g.NFlags)
	}
	environment.cuda = newCudaEnv(devicesFromEnv())
	check(C.vrad_env_add_triangles(environment.cuda.h, C.int(n), &ids[0], &verts[0], &flags[0]), "vrad_env_add_triangles")
	// replaces RefineNode/CalculateCostsOfSplit/ChangeIntoIntersectionFormat (environment.go:119-138)
	check(C.vrad_env_build(environment.cuda.h), "vrad_env_build")
	// optional: mirror the built arrays back so GetTriangle()/OptimizedKDTree keep working on the Go side
	var nNodes, nIdx, nTris C.int
	check(C.vrad_env_stats(environment.cuda.h, &nNodes, &nIdx, &nTris, nil, nil, nil, nil), "vrad_env_stats")
	children := make([]C.int32_t, nNodes)
	split := make([]C.float, nNodes)
	triIndex := make([]C.int32_t, nIdx)
	tris := make([]C.vrad_tri48, nTris)
	check(C.vrad_env_download_tree(environment.cuda.h, &children[0], &split[0], &triIndex[0], &tris[0]), "vrad_env_download_tree")
	environment.TriangleIndexList = make([]int32, nIdx)
	for i := range triIndex {
		environment.TriangleIndexList[i] = int32(triIndex[i])
	}
	for i := range tris {
		d := &environment.OptimizedTriangleList[i].TriIntersectData
		d.FlNx, d.FlNy, d.FlNz, d.FlD = float32(tris[i].nx), float32(tris[i].ny), float32(tris[i].nz), float32(tris[i].d)
		d.NTriangleID = int32(tris[i].id)
		for k := 0; k < 6; k++ {
			d.ProjectedEdgeEquations[k] = float32(tris[i].e[k])
		}
		d.NCoordSelect0, d.NCoordSelect1, d.NFlags = uint8(tris[i].sel0), uint8(tris[i].sel1), uint8(tris[i].flags)
	}
}

// Trace4Rays (environment.go:140-145) -- the stub body becomes one call.  FourVectors is x[4] y[4] z[4],
// which is exactly the origin_xyz4 / dir_xyz4 layout.
func (environment *Environment) Trace4Rays(rays *types.FourRays, TMin simd.Flt4x, TMax simd.Flt4x,
	resultOut *types.RayTracingResult, skipId int, callback *types.ITransparentTriangleCallback) {
	var origin, dir, normal [12]C.float
	for l := 0; l < 4; l++ {
		origin[l], origin[4+l], origin[8+l] = C.float(rays.Origin.X[l]), C.float(rays.Origin.Y[l]), C.float(rays.Origin.Z[l])
		dir[l], dir[4+l], dir[8+l] = C.float(rays.Direction.X[l]), C.float(rays.Direction.Y[l]), C.float(rays.Direction.Z[l])
	}
	var tmin, tmax, dist [4]C.float
	var ids [4]C.int32_t
	for l := 0; l < 4; l++ {
		tmin[l], tmax[l] = C.float(TMin[l]), C.float(TMax[l])
	}
	check(C.vrad_trace4(environment.cuda.h, &origin[0], &dir[0], &tmin[0], &tmax[0], C.int32_t(skipId), &ids[0], &dist[0], &normal[0]), "vrad_trace4")
	for l := 0; l < 4; l++ {
		resultOut.HitIds[l] = int32(ids[l])
		resultOut.HitDistance[l] = float32(dist[l])
		resultOut.SurfaceNormal.X[l], resultOut.SurfaceNormal.Y[l], resultOut.SurfaceNormal.Z[l] =
			float32(normal[l]), float32(normal[4+l]), float32(normal[8+l])
	}
}

// TestLines is the batched form the radiosity stages should call instead of looping over
// TestLineDoesHitSky packets: n segments as SoA blocks x[n] y[n] z[n].
func (environment *Environment) TestLines(start, stop []float32, n int, skyMode int, visBits []uint32) {
	check(C.vrad_test_lines(environment.cuda.h, C.int64_t(n), (*C.float)(unsafe.Pointer(&start[0])),
		(*C.float)(unsafe.Pointer(&stop[0])), C.int(skyMode), (*C.uint32_t)(unsafe.Pointer(&visBits[0]))), "vrad_test_lines")
}

var _ = mgl32.Vec3{}
