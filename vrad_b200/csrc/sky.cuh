// sky.cuh -- device functions of the complete trace.TestLineDoesHitSky (raytracer/trace/testline.go:18-94), shared by
// the batched sky kernels (k1_sky.cu) and the direct-light kernels (k3_direct.cu).  Arithmetic order follows
// oracle/skytrace.cpp statement for statement.
#pragma once
#include "env_internal.cuh"

namespace vrad {

constexpr float kMaxTraceLength = (float)(1.732050807569 * 32768.0);   // common/constants/constants.go:15-19
constexpr float kTestEpsilon = 0.03125f;                               // vmath/constants.go:7
constexpr int   kPilStack = 64;

// raytracer/trace/pointleaf.go:8-33
__device__ __forceinline__ int point_leafnum(const DevBsp& B, float px, float py, float pz) {
    if (B.n_nodes == 0) return 0;
    int node = 0;
    while (node >= 0) {
        const int4 nd = __ldg(&B.nodes[node]);
        const float4 pl = __ldg(&B.planes[nd.x]);
        float dist;
        if (nd.w < 3) dist = pick3(nd.w, px, py, pz) - pl.w;
        else dist = (((pl.x * px) + (pl.y * py)) + (pl.z * pz)) - pl.w;
        node = (dist < 0.0f) ? nd.z : nd.y;
    }
    return -1 - node;
}

// rad/clustertable/point.go:14-38, the recursion unrolled into a stack of pending back children:
// the front branch wins unless it ends in a cluster -1 leaf
__device__ __forceinline__ int point_in_leaf(const DevBsp& B, float px, float py, float pz) {
    if (B.n_nodes == 0) return 0;
    int pending[kPilStack];
    int sp = 0;
    int node = 0;
    for (;;) {
        while (node >= 0) {
            const int4 nd = __ldg(&B.nodes[node]);
            const float4 pl = __ldg(&B.planes[nd.x]);
            const float dist = (((px * pl.x) + (py * pl.y)) + (pz * pl.z)) - pl.w;
            if (dist > kTestEpsilon) node = nd.y;
            else if (dist < -kTestEpsilon) node = nd.z;
            else { if (sp < kPilStack) pending[sp++] = nd.z; node = nd.y; }
        }
        const int leaf = -1 - node;
        if (sp == 0 || __ldg(&B.leaf_cluster[leaf]) != -1) return leaf;
        node = pending[--sp];
    }
}

// testline.go:22-55 for one lane: occlusion before the recursion.  Warp-synchronous.
// SKY_RULE = false is trace.TestLine: any hit before the segment end occludes, sky triangles included (App. B.1).
template <bool COVER, bool SKY_RULE = true>
__device__ __forceinline__ float primary_occlusion(const DevScene& S, bool valid, float ax, float ay, float az,
                                                   float bx, float by, float bz, int skip_id, bool& degenerate) {
    Ray r; float len = 0.0f;
    r.ox = r.oy = r.oz = 0.0f; r.dx = r.dy = r.dz = 1.0f;
    const bool ok = segment_to_ray(ax, ay, az, bx, by, bz, r, len);
    degenerate = !ok;
    int tri; float t; float cov = 0.0f;
    if (COVER) trace_ray_cover(S, r, valid && ok, 0.0f, len, skip_id, tri, t, cov);
    else trace_ray<false>(S, r, valid && ok, 0.0f, len, skip_id, 0.0f, tri, t);
    float occ = 0.0f;
    if (valid && ok) {
        if (tri != -1 && t < len && (!SKY_RULE || (__float_as_int(__ldg(&S.q2[tri]).z) & 0x01000000) == 0)) occ = 1.0f;
        if (COVER) occ = max_sel(occ, cov);
    }
    return occ;
}

__device__ __forceinline__ float finish_fraction(float occ) {     // testline.go:91-93
    occ = max_sel(occ, 0.0f);
    occ = min_sel(occ, 1.0f);
    return 1.0f - occ;
}

// The whole of testline.go:22-93 for one lane (warp-synchronous: all 32 lanes call it together).  (lx,ly,lz) is the
// point whose leaf/area decides the recursion (the segment start, or lane 0's start for the FourVectors form).
template <bool COVER>
__device__ __forceinline__ float sky_fraction(const DevScene& S, const DevBsp& B, bool valid, float ax, float ay, float az,
                                              float bx, float by, float bz, bool can_recurse, float lx, float ly, float lz, int skip_id) {
    bool degenerate;
    float occ = primary_occlusion<COVER>(S, valid, ax, ay, az, bx, by, bz, skip_id, degenerate);
    bool recurse = valid && !degenerate && can_recurse && B.n_cams > 0 && occ < 1.0f;
    if (recurse) {
        const int leaf = point_leafnum(B, lx, ly, lz);
        recurse = false;
        if (leaf >= 0 && leaf < B.n_leafs) {
            const int area = __ldg(&B.leaf_area[leaf]);
            if (area >= 0 && area < B.n_areas) recurse = __ldg(&B.area_camera[area]) < 0;
        }
    }
    if (__any_sync(0xffffffffu, recurse)) {
        float dx = bx - ax, dy = by - ay, dz = bz - az;
        float magsq = dx * dx;
        magsq = (dy * dy) + magsq;
        magsq = (dz * dz) + magsq;
        const float rs = (float)(1.0 / sqrt((double)magsq));
        dx = dx * rs; dy = dy * rs; dz = dz * rs;
        for (int c = 0; c < B.n_cams; c++) {
            const float4 cam = __ldg(&B.cams[c]);
            const float sx = cam.x + (ax * cam.w), sy = cam.y + (ay * cam.w), sz = cam.z + (az * cam.w);
            const float ex = (dx * kMaxTraceLength) + sx, ey = (dy * kMaxTraceLength) + sy, ez = (dz * kMaxTraceLength) + sz;
            bool deg2;
            const float occ2 = primary_occlusion<COVER>(S, recurse, sx, sy, sz, ex, ey, ez, skip_id, deg2);
            if (recurse) {
                const float fv2 = deg2 ? 1.0f : finish_fraction(occ2);
                occ = occ + 1.0f;
                occ = occ - fv2;
            }
        }
    }
    return finish_fraction(occ);
}

} // namespace vrad
