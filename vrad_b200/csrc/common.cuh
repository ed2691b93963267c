// common.cuh -- device-side scene layout and the single-ray kd traversal shared by K1/K2/K3.
//
// Arithmetic contract (must match oracle/trace.cpp bit for bit): IEEE fp32, one rounding per
// operation (this translation unit is compiled with -fmad=false), exact division and sqrt
// (nvcc defaults -prec-div=true -prec-sqrt=true, no --use_fast_math), denormals kept (no -ftz),
// min/max as compare+select returning the second operand on NaN (SSE minps/maxps behaviour).
//
// Reference map: node format raytracer/cache/optimisedkdnode.go:15-54; triangle record
// raytracer/cache/triangle/triintersectdata.go:3-22; Trace4Rays contract
// raytracer/environment.go:140-145; traversal + leaf test per SURVEY.md App. B.1 (upstream
// semantics; the reference body is a stub).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrad {

constexpr int   kStackMax = 32;          // uploaded trees are validated to depth <= kStackMax - 1
constexpr float kHitInit = 1.0e23f;      // RayTracingResult.HitDistance initial value
constexpr float kDdotNEps = 1.1920929e-7f;
constexpr float kFltEpsilon = 1.1920929e-7f;

// HBM layout (SoA, every record loaded as one 64- or 128-bit vector):
//   nodes   int2   [n_nodes]   .x = children word, .y = split bits (leaf: triangle count as float)
//   tri_idx int32  [n_idx]     leaf triangle lists
//   q0      float4 [n_tris]    nx ny nz d          -- plane, read for every tested triangle
//   q1      float4 [n_tris]    e0 e1 e2 e3         -- read only when the plane test passes
//   q2      float4 [n_tris]    e4 e5 id sel|flags  -- id read first for the skip test
struct DevScene {
    const int2*    nodes;
    const int32_t* tri_index;
    const float4*  q0;
    const float4*  q1;
    const float4*  q2;
    float bmin[3], bmax[3];
    int n_nodes, n_idx, n_tris;
};

__device__ __forceinline__ float min_sel(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float max_sel(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float pick3(int k, float x, float y, float z) { return k == 0 ? x : (k == 1 ? y : z); }

struct Ray {
    float ox, oy, oz, dx, dy, dz;
};

// One triangle test.  Returns true when (t, ti) improves on (best_t, best_tri).
__device__ __forceinline__ bool test_triangle(const DevScene& S, int ti, const Ray& r, int skip_id,
                                              float best_t, int best_tri, float& t_out) {
    const float4 a = __ldg(&S.q0[ti]);
    const float ddotn = ((r.dx * a.x) + (r.dy * a.y)) + (r.dz * a.z);
    if (!(ddotn > kDdotNEps || ddotn < -kDdotNEps)) return false;
    const float odotn = ((r.ox * a.x) + (r.oy * a.y)) + (r.oz * a.z);
    const float t = (a.w - odotn) / ddotn;
    if (!(t > 0.0f)) return false;
    if (!(t < best_t || (t == best_t && ti < best_tri))) return false;
    const float4 c = __ldg(&S.q2[ti]);
    if (__float_as_int(c.z) == skip_id) return false;
    const int sel = __float_as_int(c.w);
    const int s0 = sel & 3, s1 = (sel >> 8) & 3;
    const float c0 = pick3(s0, r.ox, r.oy, r.oz) + (t * pick3(s0, r.dx, r.dy, r.dz));
    const float c1 = pick3(s1, r.ox, r.oy, r.oz) + (t * pick3(s1, r.dx, r.dy, r.dz));
    const float4 b = __ldg(&S.q1[ti]);
    const float b0 = ((b.x * c0) + (b.y * c1)) + b.z;
    if (!(b0 >= 0.0f)) return false;
    const float b1 = ((b.w * c0) + (c.x * c1)) + c.y;
    if (!(b1 >= 0.0f)) return false;
    if (!((b0 + b1) <= 1.0f)) return false;
    t_out = t;
    return true;
}

// Single-ray traversal, executed warp-synchronously.
//
// Every lane of the warp must call this together (callers pad their index space to whole warps
// and pass valid=false for the padding lanes).  Each lane traverses its own ray with its own
// stack -- results are exactly those of the scalar algorithm -- but the warp moves in phases:
// all lanes descend to their next leaf, reconverge, test their leaf's triangles, reconverge,
// pop.  Without the explicit phase barriers independent thread scheduling lets the lanes drift
// apart for good and the warp executes ~2 threads per instruction (ncu, profiles/k1_r01_*).
// The node step is branch-free (selects + one predicated push).
//
// ANY_HIT: stop at the first hit with t < any_len (visibility only); the occlusion bit equals that
// of the closest-hit traversal (same leaf sequence until the first leaf that holds such a hit).
// Closest hit: ties resolve to the lower triangle index.
template <bool ANY_HIT>
__device__ __forceinline__ void trace_ray(const DevScene& S, const Ray& r, bool valid, float tmin, float tmax,
                                          int skip_id, float any_len, int& hit_tri, float& hit_t) {
    constexpr unsigned kFull = 0xffffffffu;
    hit_tri = -1; hit_t = kHitInit;
    const float ix = 1.0f / (r.dx == 0.0f ? kFltEpsilon : r.dx);
    const float iy = 1.0f / (r.dy == 0.0f ? kFltEpsilon : r.dy);
    const float iz = 1.0f / (r.dz == 0.0f ? kFltEpsilon : r.dz);
    {
        float t0 = (S.bmin[0] - r.ox) * ix, t1 = (S.bmax[0] - r.ox) * ix;
        tmin = max_sel(tmin, min_sel(t0, t1)); tmax = min_sel(tmax, max_sel(t0, t1));
        t0 = (S.bmin[1] - r.oy) * iy; t1 = (S.bmax[1] - r.oy) * iy;
        tmin = max_sel(tmin, min_sel(t0, t1)); tmax = min_sel(tmax, max_sel(t0, t1));
        t0 = (S.bmin[2] - r.oz) * iz; t1 = (S.bmax[2] - r.oz) * iz;
        tmin = max_sel(tmin, min_sel(t0, t1)); tmax = min_sel(tmax, max_sel(t0, t1));
    }
    bool active = valid && (tmin <= tmax);
    const int negx = r.dx < 0.0f, negy = r.dy < 0.0f, negz = r.dz < 0.0f;

    int   st_node[kStackMax];
    float st_tmin[kStackMax], st_tmax[kStackMax];
    int sp = 0;
    int node = 0;
    while (__any_sync(kFull, active)) {
        int2 nd = make_int2(3, 0);
        // ---- phase 1: descend to the next leaf
        if (active) {
            nd = __ldg(&S.nodes[node]);
            while ((nd.x & 3) != 3) {
                const int axis = nd.x & 3;
                const int left = nd.x >> 2;
                const int neg = axis == 0 ? negx : (axis == 1 ? negy : negz);
                const float o = pick3(axis, r.ox, r.oy, r.oz);
                const float inv = pick3(axis, ix, iy, iz);
                const float t = (__int_as_float(nd.y) - o) * inv;
                const int front = left + neg, back = left + (neg ^ 1);
                const bool back_only = !(t >= tmin);
                const bool both = !back_only && (t <= tmax);
                const float tmin_far = max_sel(tmin, t);
                if (both) { st_node[sp] = back; st_tmin[sp] = tmin_far; st_tmax[sp] = tmax; sp++; }
                node = back_only ? back : front;
                tmax = back_only ? tmax : min_sel(tmax, t);
                tmin = back_only ? tmin_far : tmin;
                nd = __ldg(&S.nodes[node]);
            }
        }
        __syncwarp(kFull);
        // ---- phase 2: the leaf's triangles, then terminate or pop
        if (active) {
            const int start = nd.x >> 2;
            const int cnt = (int)__int_as_float(nd.y);
            for (int k = 0; k < cnt; k++) {
                const int ti = __ldg(&S.tri_index[start + k]);
                float t;
                if (ANY_HIT) {
                    if (test_triangle(S, ti, r, skip_id, any_len, -1, t)) { hit_tri = ti; hit_t = t; active = false; break; }
                } else {
                    if (test_triangle(S, ti, r, skip_id, hit_t, hit_tri, t)) { hit_tri = ti; hit_t = t; }
                }
            }
            if (active) {
                if (!(tmax <= hit_t) || sp == 0) active = false;
                else { sp--; node = st_node[sp]; tmin = st_tmin[sp]; tmax = st_tmax[sp]; }
            }
        }
        __syncwarp(kFull);
    }
}

// trace.TestLine front end (raytracer/trace/testline.go:22-27): segment -> normalised ray.
// Returns false for a zero-length segment (visible by definition).
__device__ __forceinline__ bool segment_to_ray(float ax, float ay, float az, float bx, float by, float bz,
                                               Ray& r, float& len) {
    float dx = bx - ax, dy = by - ay, dz = bz - az;
    const float len2 = ((dx * dx) + (dy * dy)) + (dz * dz);
    if (len2 == 0.0f) return false;
    len = sqrtf(len2);
    const float inv = 1.0f / len;
    r.ox = ax; r.oy = ay; r.oz = az;
    r.dx = dx * inv; r.dy = dy * inv; r.dz = dz * inv;
    return true;
}

// Occlusion rule of testline.go:42-51.  Returns 1 when the segment is visible.  Warp-synchronous:
// all 32 lanes call it; lanes with valid=false return 1 without tracing.
__device__ __forceinline__ int segment_visible(const DevScene& S, bool valid, float ax, float ay, float az,
                                               float bx, float by, float bz, int sky_mode) {
    Ray r; float len = 0.0f;
    r.ox = r.oy = r.oz = 0.0f; r.dx = r.dy = r.dz = 1.0f;
    const bool trace = valid && segment_to_ray(ax, ay, az, bx, by, bz, r, len);
    int tri; float t;
    if (!sky_mode) {      // sky_mode is warp-uniform (kernel argument)
        trace_ray<true>(S, r, trace, 0.0f, len, -1, len, tri, t);
        return tri == -1;
    }
    trace_ray<false>(S, r, trace, 0.0f, len, -1, 0.0f, tri, t);
    if (tri != -1 && t < len) {
        const int id = __float_as_int(__ldg(&S.q2[tri]).z);
        if ((id & 0x01000000) == 0) return 0;
    }
    return 1;
}

} // namespace vrad
