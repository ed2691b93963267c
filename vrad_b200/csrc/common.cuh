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
//   tri_cov float  [n_tris]    colour.X of the triangle = its coverage (texture-shadow traces only; may be null)
struct DevScene {
    const int2*    nodes;
    const int32_t* tri_index;
    const float4*  q0;
    const float4*  q1;
    const float4*  q2;
    const float*   tri_cov;
    float bmin[3], bmax[3];
    int n_nodes, n_idx, n_tris;
    // optional: the top of the tree in breadth-first order, copied to shared memory by the kernels that stage it
    // (north_star: "stage the top tree levels in shared memory").  A node reference with kTopRef set indexes this
    // array instead of `nodes`; inside it, the children word of a node whose children are staged too carries the
    // flag already shifted into place, so `word >> 2` yields a flagged reference with no extra instruction.
    const int2*    top;
    int n_top;
};
constexpr int kTopRef = 1 << 28;

__device__ __forceinline__ float min_sel(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float max_sel(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float pick3(int k, float x, float y, float z) { return k == 0 ? x : (k == 1 ? y : z); }

// Axis selects of the node step as predicated selects.  Written in PTX because nvcc turned the ternary
// chains into a three-way divergent branch inside the traversal loop (ncu r01 v3: BSSY/BSYNC pair run
// 45 M times at 4-7 active lanes).
__device__ __forceinline__ void pick_axis(int axis, float ox, float oy, float oz, float ix, float iy, float iz,
                                          float& o, float& inv) {
    asm("{\n\t"
        ".reg .pred p1, p2;\n\t"
        "setp.eq.s32 p1, %8, 1;\n\t"
        "setp.eq.s32 p2, %8, 2;\n\t"
        "selp.f32 %0, %3, %2, p1;\n\t"
        "selp.f32 %0, %4, %0, p2;\n\t"
        "selp.f32 %1, %6, %5, p1;\n\t"
        "selp.f32 %1, %7, %1, p2;\n\t"
        "}"
        : "=&f"(o), "=&f"(inv)
        : "f"(ox), "f"(oy), "f"(oz), "f"(ix), "f"(iy), "f"(iz), "r"(axis));
}

struct Ray {
    float ox, oy, oz, dx, dy, dz;
};

// Per-lane traversal state.  All member functions are WARP-SYNCHRONOUS: every lane of the warp calls
// them together (idle lanes carry active=false).  Each lane traverses its own ray with its own stack
// -- results are exactly those of the scalar algorithm in oracle/trace.cpp -- but the warp moves in
// phases: descend() takes every active lane to its next leaf, leaf() tests that leaf's triangles and
// then pops or retires the lane.  Without explicit phase barriers independent thread scheduling lets
// the lanes drift apart for good (ncu r01 v1: 2.5 threads per instruction).
//
// Closest hit: ties resolve to the lower triangle index.  ANY_HIT: the lane retires at the first
// hit with t < any_len; the occlusion bit equals that of the closest-hit traversal (same leaf
// sequence until the first leaf that holds such a hit).
//
// Stack entries are {far child, tmin of the far interval} -- 8 bytes, one 64-bit local store per push.  The far
// interval's tmax is not stored: at any moment the current tmax equals the tmin of the top entry (or the ray's
// clipped end when the stack is empty) -- a push leaves tmax = t = the pushed tmin, front-only / back-only
// steps leave tmax alone -- so a pop restores it from the entry below, bit for bit the value the 12-byte form
// stored.  A third less local-memory traffic than {node, tmin, tmax} (r01: 1.2 GB of stack write-backs per
// 4 M segments on the 1 M-triangle map).
struct TraversalStack {        // lives in local memory (dynamically indexed); kept apart from the
    float2 ent[kStackMax];     // scalar state so that the scalars stay in registers
};

// CoverageCount state of one ray (raytracer/types/coverageCount.go:16-48): transparent triangles
// (flag bit 0) add their colour.X to the coverage; the hit is dropped while coverage < 1 and kept once
// it reaches 1.  A triangle counts once per ray however many leaves hold it (exact list of the counted
// triangles, 16 entries; a ray that crosses more distinct ones is treated as fully covered) and only
// when its hit lies before the segment end -- the same rules as oracle/skytrace.cpp.
// Only the texture-shadow kernels instantiate this.
constexpr int kCoverList = 16;
struct Coverage {
    float cov;
    int   n;
    int   counted[kCoverList];
    __device__ __forceinline__ void reset() { cov = 0.0f; n = 0; }
    // returns true when the hit is to be kept
    __device__ __forceinline__ bool visit(const float* __restrict__ tri_cov, int ti) {
        for (int k = 0; k < n; k++) if (counted[k] == ti) return false;
        if (n == kCoverList) cov = 1.0f;
        else {
            counted[n++] = ti;
            const float c = tri_cov ? __ldg(&tri_cov[ti]) : 0.0f;
            cov = min_sel(cov + c, 1.0f);
        }
        return cov == 1.0f;
    }
};

struct Traversal {
    Ray   r;
    float ix, iy, iz;          // 1/d with zero components replaced by FLT_EPSILON first
    float tmin, tmax;          // current interval
    float tend;                // the ray's end clipped to the scene box (tmax of an empty stack)
    float hit_t;
    int   hit_tri;
    int   node, sp;
    int   neg;                 // bit a set when d[a] < 0 (front child = right)
    bool  active;

    __device__ __forceinline__ void idle() { active = false; hit_tri = -1; hit_t = kHitInit; }

    // start a new ray on this lane (may be called by a subset of lanes: no warp collectives inside)
    __device__ __forceinline__ void begin(const DevScene& S, const Ray& ray, bool valid, float t0, float t1, int root = 0) {
        r = ray;
        hit_tri = -1; hit_t = kHitInit;
        ix = 1.0f / (r.dx == 0.0f ? kFltEpsilon : r.dx);
        iy = 1.0f / (r.dy == 0.0f ? kFltEpsilon : r.dy);
        iz = 1.0f / (r.dz == 0.0f ? kFltEpsilon : r.dz);
        float a0 = (S.bmin[0] - r.ox) * ix, a1 = (S.bmax[0] - r.ox) * ix;
        t0 = max_sel(t0, min_sel(a0, a1)); t1 = min_sel(t1, max_sel(a0, a1));
        a0 = (S.bmin[1] - r.oy) * iy; a1 = (S.bmax[1] - r.oy) * iy;
        t0 = max_sel(t0, min_sel(a0, a1)); t1 = min_sel(t1, max_sel(a0, a1));
        a0 = (S.bmin[2] - r.oz) * iz; a1 = (S.bmax[2] - r.oz) * iz;
        t0 = max_sel(t0, min_sel(a0, a1)); t1 = min_sel(t1, max_sel(a0, a1));
        tmin = t0; tmax = t1; tend = t1;
        active = valid && (t0 <= t1);
        neg = (r.dx < 0.0f ? 1 : 0) | (r.dy < 0.0f ? 2 : 0) | (r.dz < 0.0f ? 4 : 0);
        node = root; sp = 0;
    }

    // phase 1: branch-free node steps down to the next leaf; returns that leaf's node word
    template <bool TOP = false>
    __device__ __forceinline__ int2 fetch_node(const DevScene& S, const int2* top_s) const {
        if (TOP && (node & kTopRef)) return top_s[node & (kTopRef - 1)];
        return __ldg(&S.nodes[node]);
    }

    template <bool TOP = false>
    __device__ __forceinline__ int2 descend(const DevScene& S, TraversalStack& st, const int2* top_s = nullptr) {
        int2 nd = make_int2(3, 0);
        if (active) {
            nd = fetch_node<TOP>(S, top_s);
            while ((nd.x & 3) != 3) {
                const int axis = nd.x & 3;
                const int left = nd.x >> 2;
                const int ng = (neg >> axis) & 1;
                float o, inv;
                pick_axis(axis, r.ox, r.oy, r.oz, ix, iy, iz, o, inv);
                const float t = (__int_as_float(nd.y) - o) * inv;
                const int front = left + ng, back = left + (ng ^ 1);
                const bool back_only = !(t >= tmin);
                const bool both = !back_only && (t <= tmax);
                const float tmin_far = max_sel(tmin, t);
                if (both) { st.ent[sp] = make_float2(__int_as_float(back), tmin_far); sp++; }
                node = back_only ? back : front;
                tmax = back_only ? tmax : min_sel(tmax, t);
                tmin = back_only ? tmin_far : tmin;
                nd = fetch_node<TOP>(S, top_s);
            }
        }
        __syncwarp(0xffffffffu);
        return nd;
    }

    // phase 2: the leaf's triangles in two converged sub-phases per round -- (A) every lane scans
    // forward to its next triangle whose plane hit is in range (cheap, one 16-B load each), (B) the
    // lanes that found one run the projected edge tests together -- then terminate or pop.
    template <bool ANY_HIT, bool COVER = false>
    __device__ __forceinline__ void leaf(const DevScene& S, TraversalStack& st, int2 nd, int skip_id, float any_len,
                                         Coverage* cv = nullptr) {
        const int start = nd.x >> 2;
        int cnt = active ? (int)__int_as_float(nd.y) : 0;
        int k = 0;
        for (;;) {
            int cand = -1; float tc = 0.0f;
            while (k < cnt) {
                const int ti = __ldg(&S.tri_index[start + k]);
                k++;
                const float4 a = __ldg(&S.q0[ti]);
                const float ddotn = ((r.dx * a.x) + (r.dy * a.y)) + (r.dz * a.z);
                const float odotn = ((r.ox * a.x) + (r.oy * a.y)) + (r.oz * a.z);
                const float num = a.w - odotn;
                // sign pre-filter (exact: t > 0 needs num and ddotn of one sign, both non-zero) keeps
                // back-facing and parallel planes off the IEEE divider and its slow path
                if ((ddotn > kDdotNEps && num > 0.0f) || (ddotn < -kDdotNEps && num < 0.0f)) {
                    const float t = num / ddotn;
                    const bool better = ANY_HIT ? (t < any_len) : (t < hit_t || (t == hit_t && ti < hit_tri));
                    if (t > 0.0f && better) { cand = ti; tc = t; break; }
                }
            }
            if (!__any_sync(0xffffffffu, cand >= 0)) break;
            if (cand >= 0) {
                const float4 c = __ldg(&S.q2[cand]);
                const float4 b = __ldg(&S.q1[cand]);
                const int sel = __float_as_int(c.w);
                const int s0 = sel & 3, s1 = (sel >> 8) & 3;
                const float c0 = pick3(s0, r.ox, r.oy, r.oz) + (tc * pick3(s0, r.dx, r.dy, r.dz));
                const float c1 = pick3(s1, r.ox, r.oy, r.oz) + (tc * pick3(s1, r.dx, r.dy, r.dz));
                const float b0 = ((b.x * c0) + (b.y * c1)) + b.z;
                const float b1 = ((b.w * c0) + (c.x * c1)) + c.y;
                const bool inside = (b0 >= 0.0f) && (b1 >= 0.0f) && ((b0 + b1) <= 1.0f);
                if (inside && __float_as_int(c.z) != skip_id) {
                    bool keep = true;
                    if (COVER) { if ((sel >> 16) & 1) keep = (tc < any_len) && cv->visit(S.tri_cov, cand); }   // any_len = the ray's own tmax
                    if (keep) {
                        hit_tri = cand; hit_t = tc;
                        if (ANY_HIT) { active = false; cnt = 0; }
                    }
                }
            }
        }
        if (active) {
            if (!(tmax <= hit_t) || sp == 0) active = false;
            else {
                sp--;
                const float2 en = st.ent[sp];
                tmax = sp > 0 ? st.ent[sp - 1].y : tend;
                node = __float_as_int(en.x); tmin = en.y;
            }
        }
        __syncwarp(0xffffffffu);
    }
};

// One ray per lane, to completion (used where each lane has exactly one ray: K2, K3, trace4).
template <bool ANY_HIT>
__device__ __forceinline__ void trace_ray(const DevScene& S, const Ray& r, bool valid, float tmin, float tmax,
                                          int skip_id, float any_len, int& hit_tri, float& hit_t) {
    Traversal T;
    TraversalStack st;
    T.begin(S, r, valid, tmin, tmax);
    while (__any_sync(0xffffffffu, T.active)) {
        const int2 nd = T.descend(S, st);
        T.leaf<ANY_HIT>(S, st, nd, skip_id, any_len);
    }
    hit_tri = T.hit_tri; hit_t = T.hit_t;
}

// closest hit with the transparent-triangle coverage rule; cov_out = accumulated coverage of the ray
__device__ __forceinline__ void trace_ray_cover(const DevScene& S, const Ray& r, bool valid, float tmin, float tmax,
                                                int skip_id, int& hit_tri, float& hit_t, float& cov_out) {
    Traversal T;
    TraversalStack st;
    Coverage cv;
    cv.reset();
    T.begin(S, r, valid, tmin, tmax);
    while (__any_sync(0xffffffffu, T.active)) {
        const int2 nd = T.descend(S, st);
        T.leaf<false, true>(S, st, nd, skip_id, tmax, &cv);
    }
    hit_tri = T.hit_tri; hit_t = T.hit_t; cov_out = cv.cov;
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

// Streaming traversal with lane refill: a warp owns rays [base, base + kRaysPerWarp) and keeps its 32
// lanes busy -- whenever lanes retire (their ray finished) they take the next rays of the chunk before
// the next descend/leaf round, so the warp does not idle on the longest ray of a fixed batch of 32
// (ncu r01 v2: 6.2 threads per instruction with fixed batches; rays average 2.9 leaf visits but a
// batch needs 9.3 rounds).  Results are per ray, so the order in which lanes pick rays is irrelevant.
// `more` (warp-uniform) is asked for another range of rays when the current one is used up and lanes are idle: a
// warp whose supply comes in many small ranges (the sorted order hands them out through a global counter) never
// drains -- its lanes go from the last rays of one range straight to the first rays of the next.
struct NoMoreRays { __device__ __forceinline__ bool operator()(int64_t&, int64_t&) const { return false; } };

template <typename Fetch, typename Retire, bool ANY_HIT, bool TOP = false, typename More = NoMoreRays>
__device__ __forceinline__ void stream_rays(const DevScene& S, int64_t base, int64_t end, int skip_id, Fetch fetch, Retire retire, const int2* top_s = nullptr,
                                            More more = More()) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Traversal T;
    TraversalStack st;
    T.idle();
    int64_t my_ray = -1;
    float my_len = 0.0f;
    int64_t next = base;
    bool exhausted = false;
    for (;;) {
        if (my_ray >= 0 && !T.active) { retire(my_ray, T.hit_tri, T.hit_t, my_len); my_ray = -1; }
        const unsigned idle = __ballot_sync(0xffffffffu, my_ray < 0);
        if (idle && next >= end && !exhausted) exhausted = !more(next, end);
        if (idle && next < end) {
            const int64_t idx = next + __popc(idle & lt_mask);
            if (my_ray < 0 && idx < end) {
                Ray r; float t0, t1;
                my_ray = idx;
                const bool ok = fetch(my_ray, r, t0, t1, my_len);   // false: nothing to trace (zero-length segment); may rename the ray (sorted order)
                T.begin(S, r, ok, t0, t1, TOP ? kTopRef : 0);
            }
            next += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, T.active)) {
            if (__all_sync(0xffffffffu, my_ray < 0) && next >= end && exhausted) break;
            continue;                                               // only retirements / refills pending
        }
        const int2 nd = T.template descend<TOP>(S, st, top_s);
        T.leaf<ANY_HIT>(S, st, nd, skip_id, ANY_HIT ? my_len : 0.0f);
    }
}

} // namespace vrad
