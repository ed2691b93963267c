// k2_transfers.cu -- K2: patch-to-patch visibility + form factor -> resident CSR transfer lists.
//
// Absent from the reference: only the element type exists (common/types/transfer.go:3-6, counters
// Patch.NumTransfers/Transfers at common/types/patch.go:60-61); the patch fields read here are
// those of common/types/patch.go:9-64.  Semantics follow SURVEY.md App. B.3 (vismat.cpp):
// for receiver i and every candidate j in the clusters visible from cluster(i):
//   keep (j, area_j * scale)  iff  j not sky, area_j > 0, origin_j in front of i's plane (+0.01),
//   scale = -(D.n_i)(D.n_j) / (pi * len^2) > 0, area_j*scale > 1e-7, and the segment between
//   origin+normal of the two patches (always traced from the lower to the higher index) is clear.
// MakeScales: rows whose weights sum above 1 are rescaled to sum 1.
//
// With a patch hierarchy (vrad_patches_set_hierarchy; Patch.Parent/Child1/Child2, common/types/patch.go:33,49-51)
// only leaf patches gather, and the emitters of receiver i are those the top-down walk of vismat.cpp
// TestPatchToPatch stops at: from every face root visible from cluster(i) (faces other than i's own), descend
// into the children while |origin_i - origin_j|^2 / 16 < area_j.  The walk is evaluated per candidate instead:
// j is an emitter for i iff every ancestor of j descends and j itself does not (or is a leaf) -- the same set,
// but each (row, candidate) thread decides alone by walking j's parent chain (<= tree depth steps, early out at
// the first ancestor that is small for its distance), so the candidate list stays a flat sorted array, the
// bit matrix / count / fill passes are unchanged and the rows come out in ascending patch order.
//
// Rows are independent -> sharded by rank with no collective.  Three device passes:
//   A  thread per (row, candidate): cheap tests, then the shadow ray -- one ray per unordered pair where both
//      rows are local -- into a bit matrix (rays are generated on the device: no per-pair HBM input at all);
//   B  warp per row: popcount -> row lengths; exclusive scan (cub) over 4-entry padded lengths;
//   C  warp per row: expand set bits in order into (col, w), sequential row sum, MakeScales.
// Algorithmic HBM bytes: 8*nnz + 64*N + 4*(N+1).
#include "env_internal.cuh"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <thread>

namespace vrad {

int comm_allreduce_i32(vrad_env* e, int32_t* d, size_t n);
int build_gather_plan(vrad_env* e, const int32_t* rowlen, int64_t nloc);

constexpr float kPlaneTestEpsilon = 0.01f;
constexpr float kTransEpsilon = 1.0e-7f;
constexpr float kPiF = 3.14159265358979323846f;

struct PatchView {
    const float4* origin_area;
    const float4* normal_dist;
    const float4* refl;        // .w = sky flag
    const int2*   wind;        // Patch.Winding {first point, count}; nullptr = differential form factor everywhere
    const float4* wind_pts;
};

// Form factor from the emitter's polygon to the differential receiver (upstream vismat.cpp uses it for near pairs; SURVEY App. B.3
// "optional"): the contour integral of Hottel / Baum -- for every edge of the polygon, the angle it subtends at the receiver times
// the unit normal of the plane through the receiver and the edge, dotted with the receiver's normal; the sum over the edges is
// 2*pi times the fraction of the receiver's hemisphere the polygon covers.  Returned per unit emitter area and without the 1/pi
// (MakeTransfer divides by pi and multiplies by the area, as for the differential form).  Windings are clockwise seen from the
// emitter's front, which makes the sum positive for a receiver in front of it.  The angle comes from asin of the cross product's
// length, as upstream (exact up to 90 degrees per edge).  PARITY UNPINNED (not in the reference).
__device__ __forceinline__ float poly_to_diff_form_factor(const PatchView& P, int j, float area_j, const float4 oi, const float4 ni) {
    const int2 w = __ldg(&P.wind[j]);
    if (w.y < 3) return __int_as_float(0x7fc00000);          // no polygon for this patch: the caller keeps the differential form
    float ff = 0.0f;
    for (int k = 0; k < w.y; k++) {
        const float4 p1 = __ldg(&P.wind_pts[w.x + k]);
        const float4 p2 = __ldg(&P.wind_pts[w.x + (k + 1 < w.y ? k + 1 : 0)]);
        float ax = p1.x - oi.x, ay = p1.y - oi.y, az = p1.z - oi.z;
        float bx = p2.x - oi.x, by = p2.y - oi.y, bz = p2.z - oi.z;
        const float la = sqrtf(((ax * ax) + (ay * ay)) + (az * az)), lb = sqrtf(((bx * bx) + (by * by)) + (bz * bz));
        if (la > 0.0f) { const float r = 1.0f / la; ax = ax * r; ay = ay * r; az = az * r; }
        if (lb > 0.0f) { const float r = 1.0f / lb; bx = bx * r; by = by * r; bz = bz * r; }
        float gx = (ay * bz) - (az * by), gy = (az * bx) - (ax * bz), gz = (ax * by) - (ay * bx);
        const float sin_alpha = sqrtf(((gx * gx) + (gy * gy)) + (gz * gz));
        if (sin_alpha > 1.0f) return 0.0f;
        if (sin_alpha > 0.0f) {
            const float m = asinf(sin_alpha) * (1.0f / sin_alpha);
            gx = gx * m; gy = gy * m; gz = gz * m;
        }
        ff = ff + (((gx * ni.x) + (gy * ni.y)) + (gz * ni.z));
    }
    return ff * (0.5f / area_j);
}

// MakeTransfer weight (0 = nothing transfers).  Same operation order as the CPU formulation.
__device__ __forceinline__ float transfer_weight(const PatchView& P, const float4 oi, const float4 ni, int j, const float4 oj, const float4 nj, float sky_j) {
    if (sky_j != 0.0f || !(oj.w > 0.0f)) return 0.0f;
    const float side = ((oj.x * ni.x) + (oj.y * ni.y)) + (oj.z * ni.z);
    if (!(side > ni.w + kPlaneTestEpsilon)) return 0.0f;
    float dx = oi.x - oj.x, dy = oi.y - oj.y, dz = oi.z - oj.z;
    const float len2 = ((dx * dx) + (dy * dy)) + (dz * dz);
    const float len = sqrtf(len2);
    if (!(len > 0.0f)) return 0.0f;
    const float r = 1.0f / len;
    dx = dx * r; dy = dy * r; dz = dz * r;
    const float d1 = ((dx * ni.x) + (dy * ni.y)) + (dz * ni.z);
    const float d2 = ((dx * nj.x) + (dy * nj.y)) + (dz * nj.z);
    float scale = -(d1 * d2) / ((len * len) * kPiF);
    if (!(scale > 0.0f)) return 0.0f;
    if (P.wind != nullptr && ((len * len) * kPiF) * 0.04f < oj.w) {          // emitter large for its distance: integrate over its polygon
        const float ff = poly_to_diff_form_factor(P, j, oj.w, oi, ni);
        if (ff == ff) scale = ff / kPiF;                                     // NaN: keep the differential form
        if (!(scale > 0.0f)) return 0.0f;
    }
    const float trans = oj.w * scale;
    if (!(trans > kTransEpsilon)) return 0.0f;
    return trans;
}

// TestPatchToPatch seen from the emitter (see the header): does the top-down walk for a receiver at `o_recv`
// stop exactly at patch j?  tree[] = {parent, child1, face, root cluster}.
__device__ __forceinline__ bool emitter_accepted(const PatchView& P, const int4* __restrict__ tree, const float4 o_recv,
                                                 const float4 oj, const int4 tj) {
    float dx = o_recv.x - oj.x, dy = o_recv.y - oj.y, dz = o_recv.z - oj.z;
    if (tj.y != -1 && (((dx * dx) + (dy * dy)) + (dz * dz)) * 0.0625f < oj.w) return false;     // the walk goes on into j's children
    int a = tj.x;
    while (a != -1) {
        const float4 oa = __ldg(&P.origin_area[a]);
        dx = o_recv.x - oa.x; dy = o_recv.y - oa.y; dz = o_recv.z - oa.z;
        if (!((((dx * dx) + (dy * dy)) + (dz * dz)) * 0.0625f < oa.w)) return false;             // the walk stopped above j
        a = __ldg(&tree[a]).x;
    }
    return true;
}

// Pass A.  One block per local row; threads sweep that row's candidate list.
//
// The shadow segment of a pair always runs from the lower to the higher patch index, so visibility is
// symmetric and ONE ray serves both directions: thread (i, j) with i < j also runs the cheap tests from
// j's side and, if the segment is clear, sets bit (j, i) in row j as well (position of i in j's candidate
// list by binary search); thread (j, i) is then skipped.  This halves the ray count.  It applies when row
// j is built by this rank and lists i (PVS entry [cluster j][cluster i]); otherwise each side traces
// for itself.  All bit writes are atomicOr (the own-row word is warp-aggregated), so the bit matrix does
// not depend on scheduling.
template <bool HIER>
__global__ void __launch_bounds__(256)
k2_visibility(DevScene S, PatchView P, int nloc, int64_t row0, const int32_t* __restrict__ cluster,
              const int64_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_idx,
              const int64_t* __restrict__ bit_ptr, uint32_t* __restrict__ bits,
              const uint8_t* __restrict__ pvs, int n_clusters, const int4* __restrict__ tree) {
    for (int row = blockIdx.x; row < nloc; row += gridDim.x) {
        const int i = (int)(row0 + row);
        const float4 oi = __ldg(&P.origin_area[i]), ni = __ldg(&P.normal_dist[i]);
        const float sky_i = __ldg(&P.refl[i]).w;
        const int ci = __ldg(&cluster[i]);
        int4 ti = make_int4(-1, -1, -1, ci);
        if (HIER) ti = __ldg(&tree[i]);
        const int64_t c0 = __ldg(&cand_ptr[ci]);
        const int K = (int)(__ldg(&cand_ptr[ci + 1]) - c0);
        const int Kpad = (K + 31) & ~31;
        uint32_t* out = bits + bit_ptr[row];
        for (int p = threadIdx.x; p < Kpad; p += blockDim.x) {
            bool need_ray = false, pass_ij = false, pass_ji = false;
            int j = i, cj = ci;
            float4 a = oi, an = ni, b = oi, bn = ni;
            if (p < K) {
                j = __ldg(&cand_idx[c0 + p]);
                if (j != i) {
                    cj = __ldg(&cluster[j]);
                    // row j lists i when the cluster of i's face root is visible from j's cluster (ti.w == ci when flat)
                    // (a face root in no cluster, ti.w == -1, is in nobody's list; cj is a real cluster: j came out of a candidate list)
                    const bool mirror = j >= row0 && j < row0 + nloc && (pvs == nullptr || (ti.w >= 0 && ti.w < n_clusters && __ldg(&pvs[(size_t)cj * n_clusters + ti.w]) != 0));
                    if (!(mirror && j < i)) {                       // otherwise thread (j, i) covers this pair
                        const float4 oj = __ldg(&P.origin_area[j]), nj = __ldg(&P.normal_dist[j]);
                        const float sky_j = __ldg(&P.refl[j]).w;
                        pass_ij = sky_i == 0.0f && transfer_weight(P, oi, ni, j, oj, nj, sky_j) != 0.0f;
                        pass_ji = mirror && sky_j == 0.0f && transfer_weight(P, oj, nj, i, oi, ni, sky_i) != 0.0f;
                        if (HIER && (pass_ij || pass_ji)) {           // cheap tests first: the parent-chain walk runs only for pairs that could transfer
                            const int4 tj = __ldg(&tree[j]);
                            const bool other_face = !(ti.z >= 0 && ti.z == tj.z);       // "don't check patches on the same face"
                            pass_ij = pass_ij && ti.y == -1 && other_face && emitter_accepted(P, tree, oi, oj, tj);
                            pass_ji = pass_ji && tj.y == -1 && other_face && emitter_accepted(P, tree, oj, oi, ti);
                        }
                        need_ray = pass_ij || pass_ji;
                        if (i < j) { b = oj; bn = nj; } else { a = oj; an = nj; }
                    }
                }
            }
            // warp-synchronous shadow test: every lane calls, lanes without a ray idle inside
            const int vis = segment_visible(S, need_ray, a.x + an.x, a.y + an.y, a.z + an.z,
                                            b.x + bn.x, b.y + bn.y, b.z + bn.z, 0) && need_ray;
            const uint32_t m = __ballot_sync(0xffffffffu, vis && pass_ij);
            if ((threadIdx.x & 31) == 0 && m) atomicOr(&out[p >> 5], m);
            if (vis && pass_ji) {
                // position of i in row j's candidate list (sorted patch indices of the clusters j sees)
                const int64_t j0 = __ldg(&cand_ptr[cj]);
                int lo = 0, hi = (int)(__ldg(&cand_ptr[cj + 1]) - j0);
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&cand_idx[j0 + mid]) < i) lo = mid + 1; else hi = mid; }
                atomicOr(&bits[bit_ptr[j - row0] + (lo >> 5)], 1u << (lo & 31));
            }
        }
    }
}

// Pass A, streaming form (round 2 experiment, k2_stream=1; NOT the default).  The idea: decouple the cheap tests from the rays per warp --
// a warp draws rows from a global counter and sweeps the row's candidate list 32 at a time, pairs that need a ray are compacted by
// ballot into the warp's queue in shared memory, and the traversal takes its rays from that queue through stream_rays (common.cuh)
// with lane refill, testing further candidates (of this row or the next) whenever the queue runs dry.  Same pairs, same segments,
// same bits as above (checked: tools/k2_stream_probe.py) -- but slower: 114 ms against 86 ms on the C4 build, 99e9 against 75e9 warp
// instructions at 16 instead of 22 threads per instruction (profiles/r02_k2_stream_ncu.csv).  The per-candidate kernel was never
// short of rays: 2.1e9 (row, candidate) pairs produce ~5e8 rays (most pairs between rooms face each other and are blocked by a wall),
// 8 per 32-pair batch, all towards one end point and mostly ended by the first wall they meet -- ~150 warp instructions per ray,
// the rate K1 reaches on this scene; queueing adds its loop overhead and takes the rays out of candidate order.
constexpr int kVsBlock = 128, kVsWarps = kVsBlock / 32, kVsQueue = 64, kVsBlocksPerSM = 8;
template <bool HIER>
__global__ void __launch_bounds__(kVsBlock, kVsBlocksPerSM)
k2_visibility_stream(DevScene S, PatchView P, int nloc, int64_t row0, const int32_t* __restrict__ cluster,
                     const int64_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_idx,
                     const int64_t* __restrict__ bit_ptr, uint32_t* __restrict__ bits,
                     const uint8_t* __restrict__ pvs, int n_clusters, const int4* __restrict__ tree, unsigned int* __restrict__ row_counter) {
    __shared__ int4 queue[kVsWarps][kVsQueue];          // {row, candidate position | pass_ij << 30 | pass_ji << 31, j, -}
    const unsigned lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
    int4* q = queue[threadIdx.x >> 5];
    int cur_row = 0, cur_p = 0, cur_K = 0;              // warp-uniform: the row being swept and the next candidate position
    int my_row = 0, my_pf = 0, my_j = 0;                // the pair whose ray this lane holds
    auto more = [&](int64_t& next, int64_t& end) {
        int count = 0;
        while (count <= kVsQueue - 32) {
            if (cur_p >= cur_K) {                       // next row
                unsigned r = 0;
                if (lane == 0) r = atomicAdd(row_counter, 1u);
                r = __shfl_sync(0xffffffffu, r, 0);
                if (r >= (unsigned)nloc) { cur_p = cur_K = 0; break; }
                const int c = __ldg(&cluster[row0 + r]);
                cur_row = (int)r; cur_p = 0; cur_K = (int)(__ldg(&cand_ptr[c + 1]) - __ldg(&cand_ptr[c]));
                continue;
            }
            const int i = (int)(row0 + cur_row);
            const int p = cur_p + (int)lane;
            bool pass_ij = false, pass_ji = false;
            int j = i;
            if (p < cur_K) {
                const float4 oi = __ldg(&P.origin_area[i]), ni = __ldg(&P.normal_dist[i]);
                const float sky_i = __ldg(&P.refl[i]).w;
                const int ci = __ldg(&cluster[i]);
                int4 ti = make_int4(-1, -1, -1, ci);
                if (HIER) ti = __ldg(&tree[i]);
                j = __ldg(&cand_idx[__ldg(&cand_ptr[ci]) + p]);
                if (j != i) {
                    const int cj = __ldg(&cluster[j]);
                    const bool mirror = j >= row0 && j < row0 + nloc && (pvs == nullptr || (ti.w >= 0 && ti.w < n_clusters && __ldg(&pvs[(size_t)cj * n_clusters + ti.w]) != 0));
                    if (!(mirror && j < i)) {           // otherwise pair (j, i) of row j covers this one
                        const float4 oj = __ldg(&P.origin_area[j]), nj = __ldg(&P.normal_dist[j]);
                        const float sky_j = __ldg(&P.refl[j]).w;
                        pass_ij = sky_i == 0.0f && transfer_weight(P, oi, ni, j, oj, nj, sky_j) != 0.0f;
                        pass_ji = mirror && sky_j == 0.0f && transfer_weight(P, oj, nj, i, oi, ni, sky_i) != 0.0f;
                        if (HIER && (pass_ij || pass_ji)) {
                            const int4 tj = __ldg(&tree[j]);
                            const bool other_face = !(ti.z >= 0 && ti.z == tj.z);
                            pass_ij = pass_ij && ti.y == -1 && other_face && emitter_accepted(P, tree, oi, oj, tj);
                            pass_ji = pass_ji && tj.y == -1 && other_face && emitter_accepted(P, tree, oj, oi, ti);
                        }
                    }
                }
            }
            const bool need_ray = pass_ij || pass_ji;
            const unsigned m = __ballot_sync(0xffffffffu, need_ray);
            if (need_ray) q[count + __popc(m & lt_mask)] = make_int4(cur_row, p | (pass_ij ? 1 << 30 : 0) | (pass_ji ? (int)(1u << 31) : 0), j, 0);
            count += __popc(m);
            cur_p += 32;
        }
        __syncwarp(0xffffffffu);
        next = 0; end = count;
        return count > 0;
    };
    auto fetch = [&](int64_t& qi, Ray& r, float& t0, float& t1, float& len) {
        const int4 en = q[(int)qi];
        my_row = en.x; my_pf = en.y; my_j = en.z;
        const int i = (int)(row0 + en.x);
        const bool lower = i < en.z;                    // the segment runs from the lower to the higher patch index
        const int ia = lower ? i : en.z, ib = lower ? en.z : i;
        const float4 a = __ldg(&P.origin_area[ia]), an = __ldg(&P.normal_dist[ia]);
        const float4 b = __ldg(&P.origin_area[ib]), bn = __ldg(&P.normal_dist[ib]);
        t0 = 0.0f; len = 0.0f;
        r = Ray{0.f, 0.f, 0.f, 1.f, 1.f, 1.f};
        const bool ok = segment_to_ray(a.x + an.x, a.y + an.y, a.z + an.z, b.x + bn.x, b.y + bn.y, b.z + bn.z, r, len);
        t1 = len;
        return ok;
    };
    auto retire = [&](int64_t, int tri, float, float) {
        if (tri != -1) return;                          // any-hit with the segment length as the limit: a recorded hit occludes
        const int p = my_pf & 0x3fffffff;
        if (my_pf & (1 << 30)) atomicOr(&bits[__ldg(&bit_ptr[my_row]) + (p >> 5)], 1u << (p & 31));
        if (my_pf < 0) {
            const int i = (int)(row0 + my_row);
            const int cj = __ldg(&cluster[my_j]);
            const int64_t j0 = __ldg(&cand_ptr[cj]);
            int lo = 0, hi = (int)(__ldg(&cand_ptr[cj + 1]) - j0);
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&cand_idx[j0 + mid]) < i) lo = mid + 1; else hi = mid; }
            atomicOr(&bits[__ldg(&bit_ptr[my_j - row0]) + (lo >> 5)], 1u << (lo & 31));
        }
    };
    stream_rays<decltype(fetch), decltype(retire), true, false, decltype(more)>(S, 0, 0, -1, fetch, retire, nullptr, more);
}

// Pass A, hierarchical form, top down (the walk as upstream runs it).  The per-candidate kernel above evaluates every
// patch of every visible face tree for every row (5.8e9 (row, candidate) pairs on the full S2 map for 23.5 M kept
// transfers); here a block expands receiver i's emitter set level by level from the face roots its cluster sees:
// a frontier in shared memory, each thread takes a node, pushes its two children when the node is large for its
// distance, otherwise appends it to the accepted list; then the accepted emitters run the cheap tests and the shadow
// ray warp-synchronously.  A kept emitter sets the bit at its position in the row's sorted candidate list (binary
// search), so the bit matrix, the count/scan/fill passes and the row order are the same as before -- and so is the
// result, bit for bit.  Receivers are the local leaf rows; no pair sharing (rows of ~150 emitters make rays cheap).
// The accepted list is flushed (tested and traced) whenever the next level could overflow it; if a frontier
// overflows its shared-memory capacity the host reruns the row block with the per-candidate kernel.
constexpr int kTdThreads = 128;
constexpr int kTdFrontier = 1024;       // a level's frontier is ~50 nodes per nearby face plane whatever the level (|d|^2 < 16 area)
constexpr int kTdAccepted = 2048;       // flushed (tested + traced) whenever the next level could overflow it

__global__ void __launch_bounds__(kTdThreads)
k2_visibility_topdown(DevScene S, PatchView P, int n_rows, const int32_t* __restrict__ rows, int64_t row0,
                      const int32_t* __restrict__ cluster, const int64_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_idx,
                      const int64_t* __restrict__ root_ptr, const int32_t* __restrict__ root_idx,
                      const int64_t* __restrict__ bit_ptr, uint32_t* __restrict__ bits,
                      const int4* __restrict__ tree, const int32_t* __restrict__ child2, int* __restrict__ overflow) {
    __shared__ int fr[2][kTdFrontier];
    __shared__ int acc[kTdAccepted];
    __shared__ int n_fr[2], n_acc;
    const int tid = threadIdx.x;
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int row = __ldg(&rows[r]);
        const int i = (int)(row0 + row);
        const float4 oi = __ldg(&P.origin_area[i]), ni = __ldg(&P.normal_dist[i]);
        if (__ldg(&P.refl[i]).w != 0.0f) continue;                      // sky patches receive nothing (block-uniform)
        const int ci = __ldg(&cluster[i]);
        const int face_i = __ldg(&tree[i]).z;
        const int64_t c0 = __ldg(&cand_ptr[ci]);
        const int K = (int)(__ldg(&cand_ptr[ci + 1]) - c0);
        uint32_t* out = bits + bit_ptr[row];
        // the accepted emitters so far: cheap tests, shadow ray (warp-synchronous), bit at the emitter's list position
        auto flush = [&]() {
            const int na = min(n_acc, kTdAccepted);
            for (int p = tid; p < ((na + 31) & ~31); p += kTdThreads) {
                bool need = false;
                int j = i;
                float4 a = oi, an = ni, b = oi, bn = ni;
                if (p < na) {
                    j = acc[p];
                    if (j != i) {
                        const float4 oj = __ldg(&P.origin_area[j]), nj = __ldg(&P.normal_dist[j]);
                        need = transfer_weight(P, oi, ni, j, oj, nj, __ldg(&P.refl[j]).w) != 0.0f;
                        if (i < j) { b = oj; bn = nj; } else { a = oj; an = nj; }
                    }
                }
                const int vis = segment_visible(S, need, a.x + an.x, a.y + an.y, a.z + an.z, b.x + bn.x, b.y + bn.y, b.z + bn.z, 0) && need;
                if (vis) {
                    int lo = 0, hi = K;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&cand_idx[c0 + mid]) < j) lo = mid + 1; else hi = mid; }
                    atomicOr(&out[lo >> 5], 1u << (lo & 31));
                }
            }
            __syncthreads();
            if (tid == 0) n_acc = 0;
            __syncthreads();
        };
        if (tid == 0) { n_fr[0] = 0; n_fr[1] = 0; n_acc = 0; }
        __syncthreads();
        for (int64_t k = __ldg(&root_ptr[ci]) + tid; k < __ldg(&root_ptr[ci + 1]); k += kTdThreads) {
            const int rt = __ldg(&root_idx[k]);
            if (face_i >= 0 && __ldg(&tree[rt]).z == face_i) continue;  // "don't check patches on the same face"
            const int pos = atomicAdd(&n_fr[0], 1);
            if (pos < kTdFrontier) fr[0][pos] = rt; else *overflow = 1;
        }
        __syncthreads();
        int cur = 0;
        for (;;) {
            const int n = min(n_fr[cur], kTdFrontier);
            if (n == 0) break;
            if (n_acc + n > kTdAccepted) flush();                       // block-uniform: both counts are shared
            for (int k = tid; k < n; k += kTdThreads) {
                const int j = fr[cur][k];
                const int c1 = __ldg(&tree[j]).y;
                const float4 oj = __ldg(&P.origin_area[j]);
                const float dx = oi.x - oj.x, dy = oi.y - oj.y, dz = oi.z - oj.z;
                if (c1 != -1 && (((dx * dx) + (dy * dy)) + (dz * dz)) * 0.0625f < oj.w) {
                    const int pos = atomicAdd(&n_fr[cur ^ 1], 2);
                    if (pos + 1 < kTdFrontier) { fr[cur ^ 1][pos] = c1; fr[cur ^ 1][pos + 1] = __ldg(&child2[j]); } else *overflow = 1;
                } else {
                    acc[atomicAdd(&n_acc, 1)] = j;                      // room was made above
                }
            }
            __syncthreads();
            if (tid == 0) n_fr[cur] = 0;
            cur ^= 1;
            __syncthreads();
        }
        flush();
    }
}

// Load estimate for the multi-GPU row partition: per row, the number of transfers it will hold, estimated by
// running the full pair test (cheap tests + shadow ray) on every 16th candidate (staggered by row).  The
// per-bounce gather streams exactly these entries, so blocks balanced on this estimate keep the bounce loop --
// which waits for the fullest rank -- even; costs ~1/16 of pass A.
constexpr int kEstimateStride = 16;

template <bool HIER>
__global__ void __launch_bounds__(256)
k2_estimate(DevScene S, PatchView P, int nloc, int64_t row0, const int32_t* __restrict__ cluster,
            const int64_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_idx, int32_t* __restrict__ counts,
            const int4* __restrict__ tree) {
    __shared__ int total;
    for (int row = blockIdx.x; row < nloc; row += gridDim.x) {
        const int i = (int)(row0 + row);
        if (threadIdx.x == 0) total = 0;
        __syncthreads();
        const float4 oi = __ldg(&P.origin_area[i]), ni = __ldg(&P.normal_dist[i]);
        int4 ti = make_int4(-1, -1, -1, 0);
        if (HIER) ti = __ldg(&tree[i]);
        const bool sky_i = __ldg(&P.refl[i]).w != 0.0f || ti.y != -1;      // interior patches gather nothing
        const int64_t c0 = __ldg(&cand_ptr[__ldg(&cluster[i])]);
        const int K = (int)(__ldg(&cand_ptr[__ldg(&cluster[i]) + 1]) - c0);
        const int n_samples = (K + kEstimateStride - 1) / kEstimateStride;
        int mine = 0;
        for (int q = threadIdx.x; q < ((n_samples + 31) & ~31); q += blockDim.x) {
            const int p = q * kEstimateStride + (i & (kEstimateStride - 1));
            bool need = false;
            float4 a = oi, an = ni, b = oi, bn = ni;
            if (q < n_samples && p < K && !sky_i) {
                const int j = __ldg(&cand_idx[c0 + p]);
                if (j != i) {
                    const float4 oj = __ldg(&P.origin_area[j]), nj = __ldg(&P.normal_dist[j]);
                    bool ok = transfer_weight(P, oi, ni, j, oj, nj, __ldg(&P.refl[j]).w) != 0.0f;
                    if (HIER && ok) {
                        const int4 tj = __ldg(&tree[j]);
                        ok = !(ti.z >= 0 && ti.z == tj.z) && emitter_accepted(P, tree, oi, oj, tj);
                    }
                    if (ok) {
                        need = true;
                        if (i < j) { b = oj; bn = nj; } else { a = oj; an = nj; }
                    }
                }
            }
            mine += segment_visible(S, need, a.x + an.x, a.y + an.y, a.z + an.z, b.x + bn.x, b.y + bn.y, b.z + bn.z, 0) && need;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&total, mine);
        __syncthreads();
        if (threadIdx.x == 0) counts[i] = total * kEstimateStride;
        __syncthreads();
    }
}

// Pass B.  Warp per row: number of kept transfers, and the 4-entry padded length for the scan.
__global__ void k2_count(int nloc, const int64_t* __restrict__ bit_ptr, const uint32_t* __restrict__ bits,
                         int32_t* __restrict__ rowlen, int64_t* __restrict__ padded_len) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= nloc) return;
    const int64_t w0 = bit_ptr[row], w1 = bit_ptr[row + 1];
    int cnt = 0;
    for (int64_t w = w0 + lane; w < w1; w += 32) cnt += __popc(bits[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) { rowlen[row] = cnt; padded_len[row] = (cnt + 3) & ~3; }
}

// Pass C.  Warp per row: expand bits in candidate order, then MakeScales.
__global__ void k2_fill(PatchView P, int nloc, int64_t row0, const int32_t* __restrict__ cluster,
                        const int64_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_idx,
                        const int64_t* __restrict__ bit_ptr, const uint32_t* __restrict__ bits,
                        const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rowlen,
                        int2* __restrict__ tr) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= nloc) return;
    const int i = (int)(row0 + row);
    const float4 oi = P.origin_area[i], ni = P.normal_dist[i];
    const int64_t c0 = cand_ptr[cluster[i]];
    const int64_t w0 = bit_ptr[row], w1 = bit_ptr[row + 1];
    const int64_t base = rowptr[row];
    const int len = rowlen[row];
    int pos = 0;
    for (int64_t wd = w0; wd < w1; wd++) {
        const uint32_t m = bits[wd];
        if (m == 0) continue;
        if ((m >> lane) & 1u) {
            const int j = cand_idx[c0 + (wd - w0) * 32 + lane];
            const int k = pos + __popc(m & ((1u << lane) - 1u));
            tr[base + k] = make_int2(j, __float_as_int(transfer_weight(P, oi, ni, j, P.origin_area[j], P.normal_dist[j], P.refl[j].w)));
        }
        pos += __popc(m);
    }
    const int lenp = (len + 3) & ~3;
    for (int k = len + lane; k < lenp; k += 32) tr[base + k] = make_int2(0, 0);               // padding entries (w = 0)
    __syncwarp();
    float total = 0.0f;
    if (lane == 0) for (int k = 0; k < len; k++) total = total + __int_as_float(tr[base + k].y);   // CSR-order fp32 sum
    total = __shfl_sync(0xffffffffu, total, 0);
    if (total > 1.0f) {
        const float s = 1.0f / total;
        for (int k = lane; k < len; k += 32) tr[base + k].y = __float_as_int(__int_as_float(tr[base + k].y) * s);
    }
}

} // namespace vrad
using namespace vrad;

extern "C" {

int vrad_build_transfers(vrad_env* e, int n_clusters, const uint8_t* pvs, int64_t* nnz_out) {
    VRAD_MULTI(e, group_build_transfers(e, n_clusters, pvs, nnz_out));
    if (!e) return VRAD_E_INVALID;
    if (!e->built) { set_error("vrad_build_transfers: acceleration structure not built"); return VRAD_E_STATE; }
    PatchesDev& P = e->patches;
    if (P.n == 0) { set_error("vrad_build_transfers: upload patches first"); return VRAD_E_STATE; }
    if (pvs && n_clusters <= 0) { set_error("vrad_build_transfers: pvs given but n_clusters <= 0"); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const int N = P.n;
    const int C = pvs ? n_clusters : 1;
    static const bool phase_timing = getenv("VRAD_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t_begin = now();
    // A patch with cluster -1 (origin and winding points all in solid space, rad/patches/subdivide.go:100-116) is in no cluster's
    // child list upstream: it neither gathers nor is gathered from.  Here it goes to an extra cluster, index C, that sees nothing
    // and that nothing sees -- its row comes out empty and it is in no candidate list -- so the kernels need no special case.
    const int Cx = C + 1;
    // host: per-cluster patch lists, then per-cluster sorted candidate lists (PVS-visible clusters)
    std::vector<int32_t> clus(N, 0);
    if (pvs) {
        for (int i = 0; i < N; i++) {
            int c = P.h_cluster[i];
            if (c < -1 || c >= C) { set_error("vrad_build_transfers: patch %d has cluster %d outside [-1,%d)", i, c, C); return VRAD_E_INVALID; }
            clus[i] = c < 0 ? C : c;
        }
    }
    // candidates of a cluster = every patch (with a hierarchy: of any tree level) whose face root lies in a visible cluster
    const bool hier = P.hier;
    const int4* d_tree = hier ? P.tree.p : nullptr;
    std::vector<int32_t> rclus(clus);
    if (hier && pvs) {
        for (int i = 0; i < N; i++) {
            int c = P.h_root_cluster[i];
            if (c < -1 || c >= C) { set_error("vrad_build_transfers: the face root of patch %d has cluster %d outside [-1,%d)", i, c, C); return VRAD_E_INVALID; }
            rclus[i] = c < 0 ? C : c;
        }
    }
    std::vector<std::vector<int32_t>> members(Cx);
    for (int i = 0; i < N; i++) members[rclus[i]].push_back(i);
    // list sizes first (a C x C pass over the PVS), then every cluster's list filled and sorted on its own -- on as many host
    // threads as this rank's share of the cores (the lists are the same on every rank; with one thread they were 0.37 s of a
    // 1.4 s build on the C5 map at 8 ranks, r02)
    std::vector<int64_t> cand_ptr(Cx + 1, 0);
    for (int c = 0; c < C; c++) {
        int64_t cnt = 0;
        for (int c2 = 0; c2 < C; c2++) if (!pvs || pvs[(size_t)c * C + c2]) cnt += (int64_t)members[c2].size();
        cand_ptr[c + 1] = cand_ptr[c] + cnt;
    }
    cand_ptr[Cx] = cand_ptr[C];                                                // the extra cluster's list is empty
    std::vector<int32_t> cand_idx((size_t)cand_ptr[C]);
    const int host_threads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, e->cfg.world)));
#pragma omp parallel for schedule(dynamic, 4) num_threads(host_threads)
    for (int c = 0; c < C; c++) {
        int32_t* dst = cand_idx.data() + cand_ptr[c];
        int32_t* p = dst;
        for (int c2 = 0; c2 < C; c2++)
            if (!pvs || pvs[(size_t)c * C + c2]) { std::copy(members[c2].begin(), members[c2].end(), p); p += members[c2].size(); }
        std::sort(dst, p);
    }
    // hierarchical top-down form: per cluster, the face roots (patches without a parent) of the clusters it sees
    std::vector<int64_t> root_ptr(Cx + 1, 0);
    std::vector<int32_t> root_idx;
    if (hier) {
        std::vector<std::vector<int32_t>> roots(Cx);
        for (int i = 0; i < N; i++) if (P.h_parent[i] == -1) roots[rclus[i]].push_back(i);
        for (int c = 0; c < C; c++) {
            for (int c2 = 0; c2 < C; c2++)
                if (!pvs || pvs[(size_t)c * C + c2]) root_idx.insert(root_idx.end(), roots[c2].begin(), roots[c2].end());
            root_ptr[c + 1] = (int64_t)root_idx.size();
        }
        root_ptr[Cx] = root_ptr[C];
    }
    const auto t_lists = now();
    const int world = e->cfg.world;
    // row blocks start on multiples of 4: the gather's block-row streams group 4 consecutive GLOBAL rows, so that any number of ranks sums a row in the same order
    const int64_t rpr = ((((int64_t)N + world - 1) / world) + 3) & ~(int64_t)3;
    int64_t row0 = std::min<int64_t>(N, e->cfg.rank * rpr), row1 = std::min<int64_t>(N, (e->cfg.rank + 1) * rpr);
    PatchView pv{P.origin_area.p, P.normal_dist.p, P.refl.p, P.has_windings ? P.wind.p : nullptr, P.has_windings ? P.wind_pts.p : nullptr};
    static const bool no_balance = [] { const char* v = getenv("VRAD_K2_BALANCE"); return v && v[0] == '0'; }();
    if (world > 1 && has_comm(e) && !no_balance) {
        // Balance the contiguous row blocks by estimated transfers instead of by row count (collective): every
        // rank estimates the row lengths of its equal block (k2_estimate), the estimates are summed over ranks, and
        // block r starts at the first row whose prefix reaches r/world of the total -- identical on every rank.
        DevBuf<int32_t> d_cl, d_ci, d_counts; DevBuf<int64_t> d_cp;
        auto drop = [&]() { d_cl.release(); d_ci.release(); d_counts.release(); d_cp.release(); };
        if (d_cl.alloc(N) || d_ci.alloc(cand_idx.size() + 1) || d_cp.alloc(Cx + 1) || d_counts.alloc(N)) { drop(); set_error("out of device memory (row balance)"); return VRAD_E_NOMEM; }
        std::vector<int32_t> counts(N);
        cudaError_t ce = cudaMemcpyAsync(d_cl.p, clus.data(), (size_t)N * 4, cudaMemcpyHostToDevice, e->stream);
        if (ce == cudaSuccess && !cand_idx.empty()) ce = cudaMemcpyAsync(d_ci.p, cand_idx.data(), cand_idx.size() * 4, cudaMemcpyHostToDevice, e->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_cp.p, cand_ptr.data(), (Cx + 1) * 8, cudaMemcpyHostToDevice, e->stream);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(d_counts.p, 0, (size_t)N * 4, e->stream);
        if (ce != cudaSuccess) { drop(); set_error("row balance staging failed: %s", cudaGetErrorString(ce)); return VRAD_E_CUDA; }
        const int nl0 = (int)(row1 - row0);
        if (nl0 > 0) {
            if (hier) k2_estimate<true><<<std::min(nl0, e->sm_count * 32), 256, 0, e->stream>>>(e->scene, pv, nl0, row0, d_cl.p, d_cp.p, d_ci.p, d_counts.p, d_tree);
            else k2_estimate<false><<<std::min(nl0, e->sm_count * 32), 256, 0, e->stream>>>(e->scene, pv, nl0, row0, d_cl.p, d_cp.p, d_ci.p, d_counts.p, nullptr);
        }
        int rcb = comm_allreduce_i32(e, d_counts.p, (size_t)N);
        if (rcb) { drop(); return rcb; }
        ce = cudaMemcpyAsync(counts.data(), d_counts.p, (size_t)N * 4, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        drop();
        if (ce != cudaSuccess) { set_error("row balance failed: %s", cudaGetErrorString(ce)); return VRAD_E_CUDA; }
        int64_t total = 0;
        for (int i = 0; i < N; i++) total += counts[i] + 1;                 // +1: every row costs something
        int64_t bounds[kMaxWorld + 1];
        bounds[0] = 0; bounds[world] = N;
        int64_t acc = 0; int r = 1;
        for (int i = 0; i < N && r < world; i++) {
            acc += counts[i] + 1;
            while (r < world && acc >= (total * r) / world) bounds[r++] = std::min<int64_t>(N, ((int64_t)i + 1 + 3) & ~(int64_t)3);
        }
        while (r < world) bounds[r++] = N;
        row0 = bounds[e->cfg.rank]; row1 = bounds[e->cfg.rank + 1];
        if (getenv("VRAD_VERBOSE") && e->cfg.rank == 0) {
            fprintf(stderr, "[vrad] balanced row blocks:");
            for (int q = 0; q <= world; q++) fprintf(stderr, " %lld", (long long)bounds[q]);
            fprintf(stderr, " (equal blocks would be %lld rows)\n", (long long)rpr);
        }
    }
    const auto t_balance = now();
    const int nloc = (int)(row1 - row0);
    std::vector<int64_t> bit_ptr(nloc + 1, 0);
    for (int r = 0; r < nloc; r++) {
        int c = clus[row0 + r];
        bit_ptr[r + 1] = bit_ptr[r] + ((cand_ptr[c + 1] - cand_ptr[c] + 31) >> 5);
    }
    const int64_t nwords = bit_ptr[nloc];

    DevBuf<int32_t> d_clus, d_cand_idx; DevBuf<int64_t> d_cand_ptr, d_bit_ptr, d_padlen; DevBuf<uint32_t> d_bits; DevBuf<unsigned char> d_tmp, d_pvs;
    TransfersDev& T = e->transfers;
    T.ready = false;
    auto cleanup = [&]() { d_clus.release(); d_cand_idx.release(); d_cand_ptr.release(); d_bit_ptr.release(); d_padlen.release(); d_bits.release(); d_tmp.release(); d_pvs.release(); };
    if (d_clus.alloc(N) || d_cand_idx.alloc(cand_idx.size() + 1) || d_cand_ptr.alloc(Cx + 1) || d_bit_ptr.alloc(nloc + 1) ||
        d_padlen.alloc(nloc + 1) || d_bits.alloc(nwords + 1) || T.rowptr.alloc(nloc + 1) || T.rowlen.alloc(nloc + 1) ||
        (pvs && d_pvs.alloc((size_t)C * C))) {
        cleanup(); set_error("out of device memory for transfer build (%lld visibility words)", (long long)nwords); return VRAD_E_NOMEM;
    }
#define K2_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); return VRAD_E_CUDA; } } while (0)
    K2_CHECK(cudaMemcpyAsync(d_clus.p, clus.data(), (size_t)N * 4, cudaMemcpyHostToDevice, e->stream));
    if (!cand_idx.empty()) K2_CHECK(cudaMemcpyAsync(d_cand_idx.p, cand_idx.data(), cand_idx.size() * 4, cudaMemcpyHostToDevice, e->stream));
    K2_CHECK(cudaMemcpyAsync(d_cand_ptr.p, cand_ptr.data(), (Cx + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    K2_CHECK(cudaMemcpyAsync(d_bit_ptr.p, bit_ptr.data(), (nloc + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    K2_CHECK(cudaMemsetAsync(d_padlen.p, 0, (nloc + 1) * 8, e->stream));
    K2_CHECK(cudaMemsetAsync(d_bits.p, 0, (size_t)(nwords + 1) * 4, e->stream));
    if (pvs) K2_CHECK(cudaMemcpyAsync(d_pvs.p, pvs, (size_t)C * C, cudaMemcpyHostToDevice, e->stream));

    const auto t_staged = now();
    static const bool verbose = getenv("VRAD_TIMING") != nullptr;
    cudaEvent_t tv0 = nullptr, tv1 = nullptr;
    if (verbose) { cudaEventCreate(&tv0); cudaEventCreate(&tv1); }
    timing_begin(e);
    int launches = 0;
    if (nloc > 0) {
        if (verbose) cudaEventRecord(tv0, e->stream);
        static const bool topdown_off = [] { const char* v = getenv("VRAD_K2_TOPDOWN"); return v && v[0] == '0'; }();
        bool done = false;
        if (hier && !topdown_off) {
            std::vector<int32_t> lrows;
            for (int64_t i = row0; i < row1; i++) if (P.h_child1[i] == -1) lrows.push_back((int32_t)(i - row0));
            DevBuf<int32_t> d_rows, d_root_idx; DevBuf<int64_t> d_root_ptr; DevBuf<int> d_ovf;
            auto drop = [&]() { d_rows.release(); d_root_idx.release(); d_root_ptr.release(); d_ovf.release(); };
            if (d_rows.alloc(lrows.size() + 1) || d_root_idx.alloc(root_idx.size() + 1) || d_root_ptr.alloc(Cx + 1) || d_ovf.alloc(1)) {
                drop(); cleanup(); set_error("out of device memory (top-down transfer build)"); return VRAD_E_NOMEM;
            }
            int ovf = 0;
            cudaError_t ce = cudaSuccess;
            if (!lrows.empty()) ce = cudaMemcpyAsync(d_rows.p, lrows.data(), lrows.size() * 4, cudaMemcpyHostToDevice, e->stream);
            if (ce == cudaSuccess && !root_idx.empty()) ce = cudaMemcpyAsync(d_root_idx.p, root_idx.data(), root_idx.size() * 4, cudaMemcpyHostToDevice, e->stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_root_ptr.p, root_ptr.data(), (Cx + 1) * 8, cudaMemcpyHostToDevice, e->stream);
            if (ce == cudaSuccess) ce = cudaMemsetAsync(d_ovf.p, 0, sizeof(int), e->stream);
            if (ce == cudaSuccess && !lrows.empty())
                k2_visibility_topdown<<<std::min((int)lrows.size(), e->sm_count * 16), kTdThreads, 0, e->stream>>>(
                    e->scene, pv, (int)lrows.size(), d_rows.p, row0, d_clus.p, d_cand_ptr.p, d_cand_idx.p, d_root_ptr.p, d_root_idx.p,
                    d_bit_ptr.p, d_bits.p, d_tree, P.child2.p, d_ovf.p);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(&ovf, d_ovf.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);     // host buffers above must outlive their copies
            drop();
            if (ce != cudaSuccess) { cleanup(); set_error("top-down transfer build failed: %s", cudaGetErrorString(ce)); return VRAD_E_CUDA; }
            done = ovf == 0;
            if (!done) K2_CHECK(cudaMemsetAsync(d_bits.p, 0, (size_t)(nwords + 1) * 4, e->stream));   // a list overflowed: per-candidate form
        }
        if (done) { }
        else if (e->opt.k2_stream) {
            DevBuf<unsigned int> d_ctr;
            if (d_ctr.alloc(1)) { cleanup(); set_error("out of device memory (row counter)"); return VRAD_E_NOMEM; }
            cudaError_t ce = cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned int), e->stream);
            const int grid = std::max(1, std::min((nloc + kVsWarps - 1) / kVsWarps, e->sm_count * kVsBlocksPerSM));
            if (ce == cudaSuccess) {
                if (hier) k2_visibility_stream<true><<<grid, kVsBlock, 0, e->stream>>>(e->scene, pv, nloc, row0, d_clus.p, d_cand_ptr.p, d_cand_idx.p, d_bit_ptr.p, d_bits.p,
                                                                                       pvs ? d_pvs.p : nullptr, C, d_tree, d_ctr.p);
                else k2_visibility_stream<false><<<grid, kVsBlock, 0, e->stream>>>(e->scene, pv, nloc, row0, d_clus.p, d_cand_ptr.p, d_cand_idx.p, d_bit_ptr.p, d_bits.p,
                                                                                   pvs ? d_pvs.p : nullptr, C, nullptr, d_ctr.p);
                ce = cudaStreamSynchronize(e->stream);      // the counter is released below
            }
            d_ctr.release();
            if (ce != cudaSuccess) { cleanup(); set_error("transfer build (visibility pass) failed: %s", cudaGetErrorString(ce)); return VRAD_E_CUDA; }
        }
        else if (hier) k2_visibility<true><<<std::min(nloc, e->sm_count * 32), 256, 0, e->stream>>>(e->scene, pv, nloc, row0, d_clus.p, d_cand_ptr.p, d_cand_idx.p, d_bit_ptr.p, d_bits.p,
                                                                                             pvs ? d_pvs.p : nullptr, C, d_tree);
        else k2_visibility<false><<<std::min(nloc, e->sm_count * 32), 256, 0, e->stream>>>(e->scene, pv, nloc, row0, d_clus.p, d_cand_ptr.p, d_cand_idx.p, d_bit_ptr.p, d_bits.p,
                                                                                           pvs ? d_pvs.p : nullptr, C, nullptr);
        if (verbose) cudaEventRecord(tv1, e->stream);
        launches++;
        const int wblocks = (nloc * 32 + 255) / 256;
        k2_count<<<wblocks, 256, 0, e->stream>>>(nloc, d_bit_ptr.p, d_bits.p, T.rowlen.p, d_padlen.p);
        launches++;
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_padlen.p, T.rowptr.p, nloc + 1, e->stream);
    if (d_tmp.alloc(tmp_bytes + 16)) { cleanup(); set_error("out of device memory (scan scratch)"); return VRAD_E_NOMEM; }
    K2_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_padlen.p, T.rowptr.p, nloc + 1, e->stream));
    launches++;
    int64_t np = 0;
    K2_CHECK(cudaMemcpyAsync(&np, T.rowptr.p + nloc, 8, cudaMemcpyDeviceToHost, e->stream));
    K2_CHECK(cudaStreamSynchronize(e->stream));
    if (T.tr.alloc(np + 4)) { cleanup(); set_error("out of device memory for %lld transfers", (long long)np); return VRAD_E_NOMEM; }
    if (nloc > 0) {
        const int wblocks = (nloc * 32 + 255) / 256;
        k2_fill<<<wblocks, 256, 0, e->stream>>>(pv, nloc, row0, d_clus.p, d_cand_ptr.p, d_cand_idx.p, d_bit_ptr.p, d_bits.p,
                                               T.rowptr.p, T.rowlen.p, T.tr.p);
        launches++;
    }
    timing_end(e, launches);
    // logical nnz = sum of row lengths
    std::vector<int32_t> rl(nloc ? nloc : 1);
    if (nloc) K2_CHECK(cudaMemcpyAsync(rl.data(), T.rowlen.p, (size_t)nloc * 4, cudaMemcpyDeviceToHost, e->stream));
    K2_CHECK(cudaStreamSynchronize(e->stream));
    K2_CHECK(cudaGetLastError());
#undef K2_CHECK
    if (verbose && tv0) {
        float ms = 0.f;
        if (nloc > 0) cudaEventElapsedTime(&ms, tv0, tv1);
        fprintf(stderr, "[vrad] k2_visibility %.3f ms (rows %d, visibility words %lld)\n", ms, nloc, (long long)nwords);
        cudaEventDestroy(tv0); cudaEventDestroy(tv1);
    }
    int64_t nnz = 0;
    for (int r = 0; r < nloc; r++) nnz += rl[r];
    T.row0 = row0; T.row1 = row1; T.nnz = nnz; T.nnz_padded = np; T.rows_serial++;
    T.rows_ascending = true;                 // k2_fill expands the bit matrix in candidate order = ascending patch index
    const auto t_kernels = now();
    cleanup();
    int rcp = build_gather_plan(e, rl.data(), nloc);
    if (rcp) return rcp;
    T.ready = true;
    if (phase_timing)
        fprintf(stderr, "[vrad] rank %d build_transfers phases: candidate lists (host) %.3f s, row balance %.3f s, staging %.3f s, kernels + readback %.3f s, "
                        "free + gather plan %.3f s; %zu candidate entries, %lld visibility words\n", e->cfg.rank, secs(t_begin, t_lists), secs(t_lists, t_balance),
                secs(t_balance, t_staged), secs(t_staged, t_kernels), secs(t_kernels, now()), cand_idx.size(), (long long)nwords);
    if (nnz_out) *nnz_out = nnz;
    return VRAD_OK;
}

} // extern "C"
