// vrad_environment.hpp -- host-side C++ mirror of the reference's ray-tracing call surface on top of
// the C-ABI (include/vrad_cuda.h).  The reference is compiled code (Go) whose toolchain is absent from
// the build image, so the host side above the C-ABI is written in C++ with the reference's own names,
// argument meaning and error behaviour (fatal on failure, like log.Fatalf / log.Panicln):
//
//   raytracer::Environment            raytracer/environment.go:28-39
//     AddTriangle                     :41-43
//     AddTriangleWithMaterial         :45-69
//     AddQuad                         :71-75
//     AddAxisAlignedRectangularSolid  :77-117
//     SetupAccelerationStructure      :119-138
//     Trace4Rays                      :140-145   (FourRays raytracer/types/fourrays.go:8-11,
//                                                 RayTracingResult raytracer/types/result.go:8-12)
//     GetTriangle / GetTriangleColor  :422-424, :430-432
//   trace::TestLineDoesHitSky         raytracer/trace/testline.go:18-94   (textureShadows / noSkyRecurse :13-16)
//   trace::PointLeafnum               raytracer/trace/pointleaf.go:8-10
//   cameras::ProcessSkyCameras        rad/cameras/skycamera.go:10-49
//   patches::SubdividePatches         rad/patches/subdivide.go:25-145 (+ MakePatchForFace, rad/patches/face.go:29-197)
//   loadbsp::LoadBSP / AddBrushesForRayTrace   cmd/tasks/loadbsp/main.go:163-170, 186-340
//   patches::MakePatches              rad/patches/build.go:21-65
//   lightmap::PairEdges               rad/lightmap/lightmap.go:37-216
#pragma once
#include <array>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "vrad_cuda.h"
#include "vrad_bsp.h"

namespace raytracer {

constexpr int32_t TRACE_ID_SKY = VRAD_TRACE_ID_SKY;                  // raytracer/constants.go:9-11
constexpr int32_t TRACE_ID_OPAQUE = VRAD_TRACE_ID_OPAQUE;
constexpr int32_t TRACE_ID_STATICPROP = VRAD_TRACE_ID_STATICPROP;

using Vec3 = std::array<float, 3>;
using Flt4x = std::array<float, 4>;                                   // vmath/ssemath/simd/simd.go:9
struct FourVectors { Flt4x X, Y, Z; };                                // vmath/ssemath/fourvectors.go:10-12
struct FourRays { FourVectors Origin, Direction; };
struct RayTracingResult { FourVectors SurfaceNormal; std::array<int32_t, 4> HitIds; Flt4x HitDistance; };

inline void fatal_on(int rc, const char* what) {
    if (rc != 0) { std::fprintf(stderr, "%s: vrad status %d: %s\n", what, rc, vrad_last_error()); std::abort(); }
}

class Environment {
public:
    explicit Environment(int device = 0) {
        vrad_config cfg{device, 0, 1, 0};
        fatal_on(vrad_env_create(&cfg, &h_), "vrad_env_create");
    }
    // one Environment over several GPUs of this process (vrad_env_create_multi): the singleton the reference's single-goroutine
    // driver holds (raytracer/environment.go:17-25) reaches every device through it -- batches are split and transfer rows sharded inside
    explicit Environment(const std::vector<int>& devices) {
        vrad_multi_config cfg{};
        cfg.n_devices = static_cast<int>(devices.size());
        for (int i = 0; i < cfg.n_devices && i < 8; i++) cfg.devices[i] = devices[i];
        fatal_on(vrad_env_create_multi(&cfg, &h_), "vrad_env_create_multi");
    }
    ~Environment() { vrad_env_destroy(h_); }
    Environment(const Environment&) = delete;
    Environment& operator=(const Environment&) = delete;

    void AddTriangle(int32_t id, const Vec3& v1, const Vec3& v2, const Vec3& v3, const Vec3& colour = {1, 0, 0}) {
        AddTriangleWithMaterial(id, v1, v2, v3, colour, 0, 0);
    }
    void AddTriangleWithMaterial(int32_t id, const Vec3& v1, const Vec3& v2, const Vec3& v3, const Vec3& colour, uint16_t flags, int) {
        ids_.push_back(id);
        for (const Vec3* v : {&v1, &v2, &v3}) for (float c : *v) verts_.push_back(c);
        flags_.push_back(static_cast<uint8_t>(flags));
        for (float c : colour) colours_.push_back(c);                  // TriangleColors (:61-63)
    }
    Vec3 GetTriangleColor(int index) const { return {colours_[3 * index], colours_[3 * index + 1], colours_[3 * index + 2]}; }   // :430-432
    void AddQuad(int32_t id, const Vec3& v1, const Vec3& v2, const Vec3& v3, const Vec3& v4, const Vec3& colour = {1, 0, 0}) {
        AddTriangle(id, v1, v2, v3, colour);
        AddTriangle(id + 1, v1, v3, v4, colour);
    }
    void AddAxisAlignedRectangularSolid(int32_t id, const Vec3& mn, const Vec3& mx, const Vec3& colour = {1, 0, 0}) {
        AddQuad(id, {mn[0], mx[1], mx[2]}, {mx[0], mx[1], mx[2]}, {mx[0], mn[1], mx[2]}, {mn[0], mn[1], mx[2]}, colour);
        AddQuad(id, {mn[0], mx[1], mn[2]}, {mx[0], mx[1], mn[2]}, {mx[0], mn[1], mn[2]}, {mn[0], mn[1], mn[2]}, colour);
        AddQuad(id, {mn[0], mx[1], mx[2]}, {mn[0], mx[1], mn[2]}, {mn[0], mn[1], mn[2]}, {mn[0], mn[1], mx[2]}, colour);
        AddQuad(id, {mx[0], mx[1], mx[2]}, {mx[0], mx[1], mn[2]}, {mx[0], mn[1], mn[2]}, {mx[0], mn[1], mx[2]}, colour);
        AddQuad(id, {mn[0], mx[1], mx[2]}, {mx[0], mx[1], mx[2]}, {mx[0], mx[1], mn[2]}, {mn[0], mx[1], mn[2]}, colour);
        AddQuad(id, {mn[0], mn[1], mx[2]}, {mx[0], mn[1], mx[2]}, {mx[0], mn[1], mn[2]}, {mn[0], mn[1], mn[2]}, colour);
    }
    void SetupAccelerationStructure() {
        fatal_on(vrad_env_add_triangles(h_, static_cast<int>(ids_.size()), ids_.data(), verts_.data(), flags_.data()), "vrad_env_add_triangles");
        fatal_on(vrad_env_build(h_), "vrad_env_build");
        fatal_on(vrad_env_set_triangle_colors(h_, static_cast<int>(ids_.size()), colours_.data()), "vrad_env_set_triangle_colors");
    }
    void Trace4Rays(const FourRays& rays, const Flt4x& TMin, const Flt4x& TMax, RayTracingResult* resultOut, int skipId = -1) const {
        float o[12], d[12], n[12];
        for (int l = 0; l < 4; l++) {
            o[l] = rays.Origin.X[l]; o[4 + l] = rays.Origin.Y[l]; o[8 + l] = rays.Origin.Z[l];
            d[l] = rays.Direction.X[l]; d[4 + l] = rays.Direction.Y[l]; d[8 + l] = rays.Direction.Z[l];
        }
        fatal_on(vrad_trace4(h_, o, d, TMin.data(), TMax.data(), skipId, resultOut->HitIds.data(), resultOut->HitDistance.data(), n), "vrad_trace4");
        for (int l = 0; l < 4; l++) { resultOut->SurfaceNormal.X[l] = n[l]; resultOut->SurfaceNormal.Y[l] = n[4 + l]; resultOut->SurfaceNormal.Z[l] = n[8 + l]; }
    }
    vrad_tri48 GetTriangle(int index) const {
        int nn, ni, nt;
        fatal_on(vrad_env_stats(h_, &nn, &ni, &nt, nullptr, nullptr, nullptr, nullptr), "vrad_env_stats");
        std::vector<vrad_tri48> tris(nt);
        fatal_on(vrad_env_download_tree(h_, nullptr, nullptr, nullptr, tris.data()), "vrad_env_download_tree");
        return tris.at(index);
    }
    vrad_env* handle() const { return h_; }

private:
    vrad_env* h_ = nullptr;
    std::vector<int32_t> ids_;
    std::vector<float> verts_;
    std::vector<uint8_t> flags_;
    std::vector<float> colours_;
};

} // namespace raytracer

namespace trace {

// package variables of raytracer/trace/testline.go:13-16
inline bool textureShadows = false;
inline bool noSkyRecurse = false;

// TestLineDoesHitSky (raytracer/trace/testline.go:18-94) for one FourVectors pair, complete: static-prop skip
// id (:36), sky-id rule (:42-51), coverage (:52-55), 3D-skybox recursion with the leaf of lane 0 (:57-89).
inline void TestLineDoesHitSky(const raytracer::Environment& env, const raytracer::FourVectors& start, const raytracer::FourVectors& stop,
                               raytracer::Flt4x* fractionVisible, bool canRecurse = true, int staticPropToSkip = -1, bool /*doDebug*/ = false) {
    float a[12], b[12];
    for (int l = 0; l < 4; l++) {
        a[l] = start.X[l]; a[4 + l] = start.Y[l]; a[8 + l] = start.Z[l];
        b[l] = stop.X[l]; b[4 + l] = stop.Y[l]; b[8 + l] = stop.Z[l];
    }
    const int flags = VRAD_TL_PACKET_LEAF | ((canRecurse && !noSkyRecurse) ? VRAD_TL_CAN_RECURSE : 0) | (textureShadows ? VRAD_TL_TEXTURE_SHADOWS : 0);
    raytracer::fatal_on(vrad_test_lines_sky(env.handle(), 4, a, b, flags, staticPropToSkip, fractionVisible->data()), "vrad_test_lines_sky");
}

// PointLeafnum (raytracer/trace/pointleaf.go:8-10); needs the BSP lumps (vrad_bsp_upload)
inline int PointLeafnum(const raytracer::Environment& env, const raytracer::Vec3& point) {
    int32_t leaf = -1;
    raytracer::fatal_on(vrad_point_leafnum(env.handle(), 1, point.data(), &leaf), "vrad_point_leafnum");
    return leaf;
}

} // namespace trace

namespace cameras {

// ProcessSkyCameras (rad/cameras/skycamera.go:10-49): origin and scale of every sky_camera entity; returns the cameras kept
inline int ProcessSkyCameras(const raytracer::Environment& env, const std::vector<raytracer::Vec3>& origins, const std::vector<float>& scales) {
    int kept = 0;
    raytracer::fatal_on(vrad_sky_cameras_set(env.handle(), static_cast<int>(scales.size()), origins.empty() ? nullptr : origins[0].data(),
                                             scales.data(), &kept), "vrad_sky_cameras_set");
    return kept;
}

} // namespace cameras

namespace patches {

struct PatchTree {              // the fields of common/types/patch.go:9-64 this stage produces, one entry per patch
    std::vector<float> origin, normal, plane_dist, area, mins, maxs, chop, wind_points;
    std::vector<int32_t> parent, child1, child2, face, wind_first, wind_count;
    int size() const { return static_cast<int>(area.size()); }
};

// MakePatchForFace for every face + SubdividePatches (rad/patches/face.go:29-197, subdivide.go:25-437)
inline PatchTree SubdividePatches(const std::vector<vrad_face_patch>& faces, const std::vector<float>& points3, float minChop = 4.0f) {
    int n = 0, np = 0;
    vrad_patches_subdivide(static_cast<int>(faces.size()), faces.data(), points3.data(), minChop, 0, 0, &n, &np,
                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    PatchTree t;
    t.origin.resize(3 * n); t.normal.resize(3 * n); t.mins.resize(3 * n); t.maxs.resize(3 * n); t.wind_points.resize(3 * np);
    t.plane_dist.resize(n); t.area.resize(n); t.chop.resize(n);
    t.parent.resize(n); t.child1.resize(n); t.child2.resize(n); t.face.resize(n); t.wind_first.resize(n); t.wind_count.resize(n);
    raytracer::fatal_on(vrad_patches_subdivide(static_cast<int>(faces.size()), faces.data(), points3.data(), minChop, n, np, &n, &np,
                                               t.origin.data(), t.normal.data(), t.plane_dist.data(), t.area.data(), t.mins.data(), t.maxs.data(),
                                               t.chop.data(), t.parent.data(), t.child1.data(), t.child2.data(), t.face.data(),
                                               t.wind_first.data(), t.wind_count.data(), t.wind_points.data()), "vrad_patches_subdivide");
    return t;
}

} // namespace patches

// ---- the BSP side (include/vrad_bsp.h) ------------------------------------------------------------------------------------------
namespace loadbsp {

// loadBSP + cache.BuildLumpCache (cmd/tasks/loadbsp/main.go:163-170, cache/bsp.go:51-91): the file and typed views of its lumps
struct Bsp {
    vrad_bspfile* file = nullptr;
    vrad_bsp_lumps lumps{};
    int faceLump = VRAD_LUMP_FACES, lightingLump = VRAD_LUMP_LIGHTING;       // cache.SetTargetFaces (main.go:79-89)
    explicit Bsp(const char* path, bool hdr = false) {
        raytracer::fatal_on(vrad_bspfile_open(path, &file), "vrad_bspfile_open");
        raytracer::fatal_on(vrad_bspfile_set_target_faces(file, hdr, &faceLump, &lightingLump), "vrad_bspfile_set_target_faces");
        raytracer::fatal_on(vrad_bspfile_lumps(file, &lumps), "vrad_bspfile_lumps");
    }
    ~Bsp() { vrad_bspfile_close(file); }
    Bsp(const Bsp&) = delete;
    Bsp& operator=(const Bsp&) = delete;
};

// ExtractBrushEntityShadowCasters + addBrushesForRayTrace (main.go:186-340): ids and vertices of the triangles, in the reference's order
struct RayTraceTriangles { std::vector<int32_t> ids; std::vector<float> verts9; };
inline RayTraceTriangles BrushesForRayTrace(const vrad_bsp_lumps& L, const std::vector<int32_t>& casterModel = {}, const std::vector<float>& casterOrigin3 = {},
                                            const std::vector<float>& casterAngles3 = {}) {
    const int nc = static_cast<int>(casterModel.size());
    int n = 0;
    raytracer::fatal_on(vrad_bsp_raytrace_triangles(&L, nc, casterModel.data(), casterOrigin3.data(), casterAngles3.data(), 0, nullptr, nullptr, &n), "vrad_bsp_raytrace_triangles");
    RayTraceTriangles t;
    t.ids.resize(n); t.verts9.resize(9 * static_cast<size_t>(n));
    raytracer::fatal_on(vrad_bsp_raytrace_triangles(&L, nc, casterModel.data(), casterOrigin3.data(), casterAngles3.data(), n, t.ids.data(), t.verts9.data(), &n), "vrad_bsp_raytrace_triangles");
    return t;
}
// the same, straight into the environment ("Setup ray tracer", main.go:132-150); SetupAccelerationStructure is the caller's next step
inline int AddBrushesForRayTrace(raytracer::Environment& env, const vrad_bsp_lumps& L) {
    int n = 0;
    raytracer::fatal_on(vrad_env_add_bsp(env.handle(), &L, 0, nullptr, nullptr, nullptr, &n), "vrad_env_add_bsp");
    return n;
}

}  // namespace loadbsp

namespace patches {

// MakePatches (rad/patches/build.go:21-65): one record per non-displacement face, ready for SubdividePatches
struct FacePatches {
    std::vector<vrad_face_patch> faces; std::vector<float> points3; std::vector<int32_t> faceNumber;
    std::vector<float> reflectivity3, baseArea, scale2; std::vector<uint8_t> needsBump;
};
inline FacePatches MakePatches(const vrad_bsp_lumps& L, const std::vector<float>& modelOrigins3 = {}, float maxChop = 4.0f) {
    const float* mo = modelOrigins3.empty() ? nullptr : modelOrigins3.data();
    int nf = 0, np = 0;
    raytracer::fatal_on(vrad_bsp_face_patches(&L, mo, maxChop, 0, 0, &nf, &np, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), "vrad_bsp_face_patches");
    FacePatches f;
    f.faces.resize(nf); f.points3.resize(3 * static_cast<size_t>(np)); f.faceNumber.resize(nf); f.reflectivity3.resize(3 * static_cast<size_t>(nf));
    f.baseArea.resize(nf); f.scale2.resize(2 * static_cast<size_t>(nf)); f.needsBump.resize(nf);
    raytracer::fatal_on(vrad_bsp_face_patches(&L, mo, maxChop, nf, np, &nf, &np, f.faces.data(), f.points3.data(), f.faceNumber.data(), f.reflectivity3.data(),
                                              f.baseArea.data(), f.needsBump.data(), f.scale2.data()), "vrad_bsp_face_patches");
    return f;
}

}  // namespace patches

namespace lightmap {

// PairEdges (rad/lightmap/lightmap.go:37-216): smoothed normals per face-vertex + neighbour lists
struct FaceNeighbours { std::vector<float> vertexNormals3; std::vector<int32_t> first, neighbours; };
inline FaceNeighbours PairEdges(const vrad_bsp_lumps& L, float smoothingThreshold = 0.7071067f) {
    size_t nfv = 0;
    for (int i = 0; i < L.n_faces; i++) nfv += static_cast<size_t>(L.faces[i].numedges);
    FaceNeighbours fn;
    fn.vertexNormals3.resize(3 * nfv + 3); fn.first.resize(static_cast<size_t>(L.n_faces) + 1); fn.neighbours.resize(64 * static_cast<size_t>(L.n_faces) + 1);
    raytracer::fatal_on(vrad_bsp_pair_edges(&L, smoothingThreshold, fn.vertexNormals3.data(), fn.first.data(), fn.neighbours.data(), static_cast<int>(fn.neighbours.size())), "vrad_bsp_pair_edges");
    fn.neighbours.resize(static_cast<size_t>(fn.first[L.n_faces]));
    return fn;
}

}  // namespace lightmap
