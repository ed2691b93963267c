// k4_bounce.cu -- patches, resident transfer lists and K4: the iterative bounce gather.
//
// The reference has no bounce code (cmd/tasks/computerad/main.go:5-10 comments the step out);
// the data it would run on is common/types/patch.go:9-64 (Reflectivity, TotalLight, Sky, ...)
// and common/types/transfer.go:3-6 ({Patch, Transfer} = one CSR entry).  Semantics follow
// SURVEY.md App. B.4 (GatherLight / CollectLight / BounceLight, leaf patches only):
//     add[i]   = sum_k w[i,k] * (emit[col[i,k]] * refl[col[i,k]])
//     total[i] += add[i];  emit[i] = add[i];  added += emit[i]        (sky patches: emit = 0)
//
// HBM layout: transfers are a padded CSR of {col:int32, w:float32} pairs (8 B, the reference's
// Transfer struct) -- every row starts on a 4-entry (32-byte sector) boundary; `er` = emit*refl is
// kept as one float4 per patch (16 B gathers that live in L2: 3.2 MB at 200k patches, 32 MB at 2M).  Algorithmic
// bytes per bounce: 8*nnz + 40*N; the kernel is HBM-bound on the 8*nnz stream.
// One warp per row, warp-shuffle reduction, collect step fused into the epilogue.
#include "env_internal.cuh"
#include <algorithm>
#include <cstdlib>
#include <string>

namespace vrad {

int comm_allgather_rows(vrad_env* e, float4* buf, const int64_t* bounds);
int comm_exchange_bounds(vrad_env* e, int64_t row0, int64_t row1, int64_t n, int64_t* bounds_out);
int comm_setup_peers(vrad_env* e, size_t n_pad);
int comm_allreduce3(vrad_env* e, float* d3);

constexpr int kGatherBlock = 256;
constexpr int kGatherWarps = kGatherBlock / 32;

__global__ void k4_init_er(int n_pad, int n, const float* __restrict__ emit0, const float4* __restrict__ refl,
                           float4* __restrict__ er) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
        float4 r = refl[i];
        if (r.w == 0.0f) v = make_float4(emit0[3 * i] * r.x, emit0[3 * i + 1] * r.y, emit0[3 * i + 2] * r.z, 0.f);
    }
    er[i] = v;
}

// One warp per row.  Entries are {col, w} pairs -- the reference's Transfer struct
// (common/types/transfer.go:3-6) -- read as one coalesced 64-bit load per lane with lanes on
// CONSECUTIVE entries, so the 32 er[] gathers of one instruction hit consecutive patches wherever the
// row has a run of adjacent columns (avg run length ~16 on the synthetic maps): few L1 wavefronts
// per gather instead of one per lane (ncu r01: the int4-per-lane mapping was L1TEX-bound at 86%).
// kGatherUnroll entries per lane are in flight and the {col,w} loads of the NEXT step are issued
// before the current step's gathers (software pipeline), so the two dependent memory latencies
// overlap.  The kernel is latency-bound: what matters is bytes in flight per SM = resident warps x
// entries in flight, so the register budget is capped to keep 5 blocks (40 warps) per SM -- the
// same loop at 60 registers (4 blocks) ran at 4.1 TB/s, at 48 registers 6.4 TB/s (tools/exp/k4_exp.cu).
constexpr int kGatherUnroll = 8;

__global__ void __launch_bounds__(kGatherBlock, 5)
k4_gather(int nloc, int64_t row0, const int64_t* __restrict__ rowptr, const int2* __restrict__ tr,
          const float4* __restrict__ er, const float4* __restrict__ refl,
          float4* __restrict__ er_next, float4* __restrict__ total, float* __restrict__ partials) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * kGatherWarps + warp;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (row < nloc) {
        const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];      // padded to 4 entries; padding has w = 0
        const int2 zero = make_int2(0, 0);                          // out-of-row slots: col 0, weight 0
        int2 cur[kGatherUnroll], nxt[kGatherUnroll];
        int64_t k = k0 + lane;
#pragma unroll
        for (int j = 0; j < kGatherUnroll; j++) cur[j] = k + 32 * j < k1 ? __ldcs(&tr[k + 32 * j]) : zero;
        for (; k < k1; k += 32 * kGatherUnroll) {
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++)
                nxt[j] = k + 32 * (kGatherUnroll + j) < k1 ? __ldcs(&tr[k + 32 * (kGatherUnroll + j)]) : zero;
            float4 x[kGatherUnroll];
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) x[j] = __ldg(&er[cur[j].x]);
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) {
                const float w = __int_as_float(cur[j].y);
                s0 += w * x[j].x; s1 += w * x[j].y; s2 += w * x[j].z;
            }
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) cur[j] = nxt[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            const float4 r = refl[row0 + row];
            if (r.w == 0.0f) {                                              // CollectLight, leaf patch
                float4 t = total[row];
                t.x += s0; t.y += s1; t.z += s2;
                total[row] = t;
                er_next[row0 + row] = make_float4(s0 * r.x, s1 * r.y, s2 * r.z, 0.f);
                e0 = s0; e1 = s1; e2 = s2;
            } else {
                er_next[row0 + row] = make_float4(0.f, 0.f, 0.f, 0.f);     // sky: emit = 0
            }
        }
    }
    // deterministic per-block partial of `added`
    __shared__ float sm[kGatherWarps][3];
    if (lane == 0) { sm[warp][0] = e0; sm[warp][1] = e1; sm[warp][2] = e2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kGatherWarps; k++) a += sm[k][threadIdx.x];
        partials[3 * (size_t)blockIdx.x + threadIdx.x] = a;
    }
}

// GatherLight, bump-mapped branch (upstream vrad.cpp; the reference carries the types: Patch.NeedsBumpMap
// common/types/patch.go:23, BumpLights common/types/bumpLights.go:8-10, NUM_BUMP_VECTS common/constants/constants.go:33).
// For a bump-mapped patch the light of every transfer is also projected on the three bump-basis normals:
//   delta = normalize(origin_j - origin_i); v = emit_j * refl_j * transfer / (delta . n_i)   ("remove normal already factored
//   into transfer steradian"); sum_b += v * (delta . normal_b) for delta . normal_b > 0.
// The flat sum (Light[0], which is what the patch re-emits) stays with the gather kernels above; this kernel runs over the
// bump-mapped leaf rows only, one warp per row, and re-reads their {col,w} entries plus the emitters' origins.
__global__ void __launch_bounds__(256)
k4_gather_bump(int n_rows, const int32_t* __restrict__ rows, int64_t row0, const int64_t* __restrict__ rowptr, const int2* __restrict__ tr,
               const float4* __restrict__ er, const float4* __restrict__ origin_area, const float4* __restrict__ normal_dist,
               const float4* __restrict__ bump_normals, float4* __restrict__ tb0, float4* __restrict__ tb1, float4* __restrict__ tb2) {
    const int lane = threadIdx.x & 31;
    const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= n_rows) return;
    const int row = __ldg(&rows[r]);
    const int64_t i = row0 + row;
    const float4 oi = __ldg(&origin_area[i]), ni = __ldg(&normal_dist[i]);
    const float4 n1 = __ldg(&bump_normals[3 * i]), n2 = __ldg(&bump_normals[3 * i + 1]), n3 = __ldg(&bump_normals[3 * i + 2]);
    float s[9];
#pragma unroll
    for (int k = 0; k < 9; k++) s[k] = 0.f;
    for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
        const int2 en = __ldcs(&tr[k]);
        const float w = __int_as_float(en.y);
        if (w == 0.0f) continue;                                   // row padding
        const float4 oj = __ldg(&origin_area[en.x]);
        const float4 v = __ldg(&er[en.x]);
        float dx = oj.x - oi.x, dy = oj.y - oi.y, dz = oj.z - oi.z;
        const float len = sqrtf(((dx * dx) + (dy * dy)) + (dz * dz));
        if (len != 0.0f) { const float rl = 1.0f / len; dx = dx * rl; dy = dy * rl; dz = dz * rl; }
        const float ws = w * (1.0f / (((dx * ni.x) + (dy * ni.y)) + (dz * ni.z)));
        const float vx = v.x * ws, vy = v.y * ws, vz = v.z * ws;
        const float d1 = ((dx * n1.x) + (dy * n1.y)) + (dz * n1.z);
        const float d2 = ((dx * n2.x) + (dy * n2.y)) + (dz * n2.z);
        const float d3 = ((dx * n3.x) + (dy * n3.y)) + (dz * n3.z);
        if (d1 > 0.0f) { s[0] += vx * d1; s[1] += vy * d1; s[2] += vz * d1; }
        if (d2 > 0.0f) { s[3] += vx * d2; s[4] += vy * d2; s[5] += vz * d2; }
        if (d3 > 0.0f) { s[6] += vx * d3; s[7] += vy * d3; s[8] += vz * d3; }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
    if (lane == 0) {                                               // CollectLight: TotalLight.Light[b] += addlight.light[b] (leaf patches)
        float4 t = tb0[i]; t.x += s[0]; t.y += s[1]; t.z += s[2]; tb0[i] = t;
        t = tb1[i]; t.x += s[3]; t.y += s[4]; t.z += s[5]; tb1[i] = t;
        t = tb2[i]; t.x += s[6]; t.y += s[7]; t.z += s[8]; tb2[i] = t;
    }
}

// Short-row form (patch hierarchy: rows average ~150 transfers instead of ~1000+).  With one warp per row such a
// row is a single partial step and the kernel is a chain of dependent latencies (rowptr -> {col,w} -> er[col] ->
// epilogue) with only 40 rows in flight per SM: measured 1.3 TB/s on the hierarchical S2 matrix.  Here 8 lanes
// own a row (4 rows per warp, lanes on consecutive entries so the gathers of a row still coalesce), 4 entries in
// flight per lane and the register budget capped at 32 so that 64 warps = 256 rows stay resident per SM: 61.9 us per
// bounce; with the one-step {col,w} prefetch of the long-row kernel the loop needs 40 registers (48 warps): 64.1 us,
// and squeezed into 32 it spills: 76 us.  `rows` (optional) lists the local rows to process -- with a
// hierarchy only the leaf patches gather, the interior rows are rewritten by k4_collect_parents.
template <int kShortLanes, int kMinBlocks, int kShortUnroll = 4, bool kPrefetch = true>
__global__ void __launch_bounds__(kGatherBlock, kMinBlocks)
k4_gather_short(int nrows, const int32_t* __restrict__ rows, int64_t row0, const int64_t* __restrict__ rowptr,
                const int2* __restrict__ tr, const float4* __restrict__ er, const float4* __restrict__ refl,
                float4* __restrict__ er_next, float4* __restrict__ total, float* __restrict__ partials) {
    constexpr int kShortRowsPerBlock = kGatherBlock / kShortLanes;
    const int sub = threadIdx.x & (kShortLanes - 1), grp = threadIdx.x / kShortLanes;
    const int r = blockIdx.x * kShortRowsPerBlock + grp;
    const bool valid = r < nrows;
    const int row = valid ? (rows ? __ldg(&rows[r]) : r) : 0;
    const int64_t k0 = valid ? rowptr[row] : 0, k1 = valid ? rowptr[row + 1] : 0;
    const int2 zero = make_int2(0, 0);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    int64_t k = k0 + sub;
    if (kPrefetch) {
        int2 cur[kShortUnroll], nxt[kShortUnroll];
#pragma unroll
        for (int j = 0; j < kShortUnroll; j++) cur[j] = k + kShortLanes * j < k1 ? __ldcs(&tr[k + kShortLanes * j]) : zero;
        for (; k < k1; k += kShortLanes * kShortUnroll) {
#pragma unroll
            for (int j = 0; j < kShortUnroll; j++)
                nxt[j] = k + kShortLanes * (kShortUnroll + j) < k1 ? __ldcs(&tr[k + kShortLanes * (kShortUnroll + j)]) : zero;
            float4 x[kShortUnroll];
#pragma unroll
            for (int j = 0; j < kShortUnroll; j++) x[j] = __ldg(&er[cur[j].x]);
#pragma unroll
            for (int j = 0; j < kShortUnroll; j++) {
                const float w = __int_as_float(cur[j].y);
                s0 += w * x[j].x; s1 += w * x[j].y; s2 += w * x[j].z;
            }
#pragma unroll
            for (int j = 0; j < kShortUnroll; j++) cur[j] = nxt[j];
        }
    } else {
        for (; k < k1; k += kShortLanes * kShortUnroll) {
            int2 cur[kShortUnroll]; float4 x[kShortUnroll];
#pragma unroll
            for (int j = 0; j < kShortUnroll; j++) cur[j] = k + kShortLanes * j < k1 ? __ldcs(&tr[k + kShortLanes * j]) : zero;
#pragma unroll
            for (int j = 0; j < kShortUnroll; j++) x[j] = __ldg(&er[cur[j].x]);
#pragma unroll
            for (int j = 0; j < kShortUnroll; j++) {
                const float w = __int_as_float(cur[j].y);
                s0 += w * x[j].x; s1 += w * x[j].y; s2 += w * x[j].z;
            }
        }
    }
#pragma unroll
    for (int o = kShortLanes / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (valid && sub == 0) {
        const float4 rf = refl[row0 + row];
        if (rf.w == 0.0f) {                                              // CollectLight, leaf patch
            float4 t = total[row];
            t.x += s0; t.y += s1; t.z += s2;
            total[row] = t;
            er_next[row0 + row] = make_float4(s0 * rf.x, s1 * rf.y, s2 * rf.z, 0.f);
            e0 = s0; e1 = s1; e2 = s2;
        } else {
            er_next[row0 + row] = make_float4(0.f, 0.f, 0.f, 0.f);      // sky: emit = 0
        }
    }
    // deterministic per-block partial of `added`
    __shared__ float sm[kShortRowsPerBlock][3];
    if (sub == 0) { sm[grp][0] = e0; sm[grp][1] = e1; sm[grp][2] = e2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < kShortRowsPerBlock; q++) a += sm[q][threadIdx.x];
        partials[3 * (size_t)blockIdx.x + threadIdx.x] = a;
    }
}

// out of line on purpose: keeps the peer-table loads out of the gather loop's register allocation
__device__ __noinline__ void store_row_to_peers(const PeerTable* __restrict__ peers, int next_buf, int64_t row, float x, float y, float z) {
    const int lane = threadIdx.x & 31;
    x = __shfl_sync(0xffffffffu, x, 0); y = __shfl_sync(0xffffffffu, y, 0); z = __shfl_sync(0xffffffffu, z, 0);
    if (lane < peers->world) peers->er[next_buf][lane][row] = make_float4(x, y, z, 0.f);
}

__device__ __noinline__ void wait_for_peers(const PeerTable* __restrict__ peers, int wait_world) {
    if ((int)threadIdx.x < wait_world) {
        volatile uint32_t* local = peers->flags[peers->rank];
        const uint32_t epoch = local[kMaxWorld];
        while ((int32_t)(local[threadIdx.x] - epoch) < 0) { }
    }
    __syncthreads();
}

// Multi-GPU form of the same kernel: `peers` (device memory) lists every rank's er_next buffer and this
// rank's flag words; lanes 0..world-1 store the finished row into one rank each (fused exchange), and
// the prologue is the wait half of the inter-bounce barrier.  Kept as a separate kernel so that the
// single-GPU instantiation keeps its register allocation (the loop is occupancy-bound).
template <bool MULTI>
__global__ void __launch_bounds__(kGatherBlock, 5)
k4_gather_multi(int nloc, int64_t row0, const int64_t* __restrict__ rowptr, const int2* __restrict__ tr,
          const float4* __restrict__ er, const float4* __restrict__ refl,
          float4* __restrict__ er_next, float4* __restrict__ total, float* __restrict__ partials,
          const PeerTable* __restrict__ peers, int next_buf, int wait_world) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * kGatherWarps + warp;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (MULTI) {
        // fused-exchange prologue: the radiance this bounce reads is complete once every rank's epoch has
        // arrived in the local flag words (wait half of the barrier; k4_peer_signal is the other half)
        if (wait_world > 1) wait_for_peers(peers, wait_world);
    }
    if (row < nloc) {
        const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];      // padded to 4 entries; padding has w = 0
        const int2 zero = make_int2(0, 0);                          // out-of-row slots: col 0, weight 0
        int2 cur[kGatherUnroll], nxt[kGatherUnroll];
        int64_t k = k0 + lane;
#pragma unroll
        for (int j = 0; j < kGatherUnroll; j++) cur[j] = k + 32 * j < k1 ? __ldcs(&tr[k + 32 * j]) : zero;
        for (; k < k1; k += 32 * kGatherUnroll) {
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++)
                nxt[j] = k + 32 * (kGatherUnroll + j) < k1 ? __ldcs(&tr[k + 32 * (kGatherUnroll + j)]) : zero;
            float4 x[kGatherUnroll];
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) x[j] = __ldg(&er[cur[j].x]);
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) {
                const float w = __int_as_float(cur[j].y);
                s0 += w * x[j].x; s1 += w * x[j].y; s2 += w * x[j].z;
            }
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) cur[j] = nxt[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);                        // sky: emit = 0
        if (lane == 0) {
            const float4 r = refl[row0 + row];
            if (r.w == 0.0f) {                                              // CollectLight, leaf patch
                float4 t = total[row];
                t.x += s0; t.y += s1; t.z += s2;
                total[row] = t;
                nv = make_float4(s0 * r.x, s1 * r.y, s2 * r.z, 0.f);
                e0 = s0; e1 = s1; e2 = s2;
            }
            if (!MULTI) er_next[row0 + row] = nv;
        }
        if (MULTI) {
            // fused exchange: lane p stores the finished row straight into rank p's next-bounce buffer
            // (NVLink peer store; slot `rank` is the local buffer) -- no separate all-gather pass
            store_row_to_peers(peers, next_buf, row0 + row, nv.x, nv.y, nv.z);
        }
    }
    // deterministic per-block partial of `added`
    __shared__ float sm[kGatherWarps][3];
    if (lane == 0) { sm[warp][0] = e0; sm[warp][1] = e1; sm[warp][2] = e2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kGatherWarps; k++) a += sm[k][threadIdx.x];
        partials[3 * (size_t)blockIdx.x + threadIdx.x] = a;
    }
}

// Fused-exchange barrier between bounces, split in two so that no kernel sits spinning between them:
// k4_peer_signal (one tiny block after each gather) bumps this rank's epoch and stores it into its
// slot of every peer's flag words -- after a system-scope fence that orders the peer row stores of
// the gather that just completed -- and the NEXT k4_gather waits in its prologue until all peers'
// epochs have arrived.  k4_peer_wait closes the last bounce of a call.
// wait half as a stand-alone kernel: closes the last bounce of a call (no later gather would wait for it)
__global__ void k4_peer_wait(const PeerTable* __restrict__ peers) {
    if ((int)threadIdx.x < peers->world) {
        volatile uint32_t* local = peers->flags[peers->rank];
        const uint32_t epoch = local[kMaxWorld];
        while ((int32_t)(local[threadIdx.x] - epoch) < 0) { }
        __threadfence_system();
    }
}

// signal half of the fused-exchange barrier (the wait half is the prologue of k4_gather)
__global__ void k4_peer_signal(const PeerTable* __restrict__ peers) {
    __shared__ uint32_t epoch;
    if (threadIdx.x == 0) epoch = ++peers->flags[peers->rank][kMaxWorld];
    __syncthreads();
    if ((int)threadIdx.x < peers->world) {
        __threadfence_system();
        volatile uint32_t* dst = peers->flags[threadIdx.x] + peers->rank;
        *dst = epoch;
    }
}

// single block: fixed-order tree reduction of the per-block partials -> added[3]
__global__ void k4_reduce_added(int nblocks, const float* __restrict__ partials, float* __restrict__ added) {
    __shared__ float sm[3][256];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int i = threadIdx.x; i < nblocks; i += 256) { a0 += partials[3 * i]; a1 += partials[3 * i + 1]; a2 += partials[3 * i + 2]; }
    sm[0][threadIdx.x] = a0; sm[1][threadIdx.x] = a1; sm[2][threadIdx.x] = a2;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) for (int c = 0; c < 3; c++) sm[c][threadIdx.x] += sm[c][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 3) added[threadIdx.x] = sm[threadIdx.x][0];
}

__global__ void k4_unpack_total(int64_t n, const float4* __restrict__ total, float* __restrict__ out3) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 t = total[i];
    out3[3 * i] = t.x; out3[3 * i + 1] = t.y; out3[3 * i + 2] = t.z;
}

static inline int64_t rows_per_rank(const vrad_env* e, int64_t n) { return (n + e->cfg.world - 1) / e->cfg.world; }

// CollectLight for interior patches (vrad.cpp CollectLight): an interior patch holds the
// area-weighted average of its two children.  Flattened over the subtree: one warp per interior patch sums
// w(p, leaf) * buf[leaf] over its leaves (children inherit the face's reflectivity -- CreateChildPatch copies the
// parent, rad/patches/subdivide.go:360 -- so averaging emit*refl equals averaging emit and then reflecting).
// Rows are ordered long first (host side): blocks [0, long_blocks) give a whole warp to each of the n_long rows
// with >= kCollectLong leaves (face roots and their first levels: up to 1000+ entries), the other blocks give 8
// lanes to each remaining row (half of all interior patches have 2 leaves).  Loads are issued kCollectUnroll deep:
// one warp walking a root row entry by entry was a 28 us tail per bounce.
constexpr int kCollectLong = 128;
constexpr int kCollectUnroll = 4;

template <int LANES>
__device__ __forceinline__ void collect_row(int64_t k0, int64_t k1, int sub, const int2* __restrict__ ent, const float4* buf,
                                            float& s0, float& s1, float& s2) {
    const int2 zero = make_int2(0, 0);
    for (int64_t k = k0 + sub; k < k1; k += LANES * kCollectUnroll) {
        int2 en[kCollectUnroll]; float4 v[kCollectUnroll];
#pragma unroll
        for (int j = 0; j < kCollectUnroll; j++) en[j] = k + LANES * j < k1 ? __ldg(&ent[k + LANES * j]) : zero;
#pragma unroll
        for (int j = 0; j < kCollectUnroll; j++) v[j] = buf[en[j].x];
#pragma unroll
        for (int j = 0; j < kCollectUnroll; j++) {
            const float wt = __int_as_float(en[j].y);
            s0 += wt * v[j].x; s1 += wt * v[j].y; s2 += wt * v[j].z;
        }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
}

__global__ void __launch_bounds__(256)
k4_collect_parents(int n_interior, int n_long, int long_blocks, const int32_t* __restrict__ ids, const int64_t* __restrict__ cptr,
                   const int2* __restrict__ ent, float4* __restrict__ buf) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if ((int)blockIdx.x < long_blocks) {
        const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        const bool valid = w < n_long;
        collect_row<32>(valid ? cptr[w] : 0, valid ? cptr[w + 1] : 0, lane, ent, buf, s0, s1, s2);
        if (valid && lane == 0) buf[ids[w]] = make_float4(s0, s1, s2, 0.f);
    } else {
        const int w = n_long + ((int)blockIdx.x - long_blocks) * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
        const bool valid = w < n_interior;
        collect_row<8>(valid ? cptr[w] : 0, valid ? cptr[w + 1] : 0, sub, ent, buf, s0, s1, s2);
        if (valid && sub == 0) buf[ids[w]] = make_float4(s0, s1, s2, 0.f);
    }
}

} // namespace vrad
using namespace vrad;

extern "C" {

int vrad_patches_upload(vrad_env* e, int n, const float* origin3, const float* normal3, const float* plane_dist,
                        const float* area, const float* reflectivity3, const int32_t* cluster, const uint8_t* flags) {
    if (!e || n <= 0 || !origin3 || !normal3 || !plane_dist || !area || !reflectivity3) { set_error("vrad_patches_upload: bad arguments"); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    PatchesDev& P = e->patches;
    if (P.origin_area.alloc(n) || P.normal_dist.alloc(n) || P.refl.alloc(n) || P.cluster.alloc(n)) { set_error("out of device memory for patches"); return VRAD_E_NOMEM; }
    std::vector<float4> oa(n), nd(n), rf(n);
    P.h_cluster.assign(n, 0); P.h_flags.assign(n, 0);
    P.h_area.assign(area, area + n); P.h_refl.assign(reflectivity3, reflectivity3 + 3 * (size_t)n);
    P.hier = false; P.n_interior = 0; P.h_root_cluster.clear();
    P.bump = false; P.h_needs_bump.clear();
    for (int i = 0; i < n; i++) {
        oa[i] = make_float4(origin3[3 * i], origin3[3 * i + 1], origin3[3 * i + 2], area[i]);
        nd[i] = make_float4(normal3[3 * i], normal3[3 * i + 1], normal3[3 * i + 2], plane_dist[i]);
        uint8_t f = flags ? flags[i] : 0;
        rf[i] = make_float4(reflectivity3[3 * i], reflectivity3[3 * i + 1], reflectivity3[3 * i + 2], (f & 1) ? 1.0f : 0.0f);
        P.h_flags[i] = f;
        if (cluster) P.h_cluster[i] = cluster[i];
    }
    VRAD_CUDA_CHECK(cudaMemcpyAsync(P.origin_area.p, oa.data(), (size_t)n * 16, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(P.normal_dist.p, nd.data(), (size_t)n * 16, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(P.refl.p, rf.data(), (size_t)n * 16, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(P.cluster.p, P.h_cluster.data(), (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    P.n = n;
    e->transfers.ready = false;
    return VRAD_OK;
}

int vrad_patches_set_hierarchy(vrad_env* e, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face) {
    if (!e || !parent || !child1 || !child2) { set_error("vrad_patches_set_hierarchy: bad arguments"); return VRAD_E_INVALID; }
    PatchesDev& P = e->patches;
    if (P.n == 0 || n != P.n) { set_error("vrad_patches_set_hierarchy: %d links for %d uploaded patches", n, P.n); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    // the links of a SubdividePatches result: children are appended after their parent (subdivide.go:352-355), in pairs
    for (int i = 0; i < n; i++) {
        const int c1 = child1[i], c2 = child2[i];
        if ((c1 == -1) != (c2 == -1)) { set_error("vrad_patches_set_hierarchy: patch %d has one child", i); return VRAD_E_INVALID; }
        if (c1 != -1) {
            if (c1 <= i || c2 <= i || c1 >= n || c2 >= n || c1 == c2 || parent[c1] != i || parent[c2] != i) { set_error("vrad_patches_set_hierarchy: bad children of patch %d", i); return VRAD_E_INVALID; }
            for (int k = 0; k < 2; k++) {
                const int c = k ? c2 : c1;
                if (memcmp(&P.h_refl[3 * (size_t)c], &P.h_refl[3 * (size_t)i], 12) != 0 || (P.h_flags[c] & 1) != (P.h_flags[i] & 1)) {
                    set_error("vrad_patches_set_hierarchy: child %d does not carry its parent's reflectivity / sky flag (CreateChildPatch copies the parent)", c);
                    return VRAD_E_INVALID;
                }
            }
        }
        if (parent[i] != -1 && (parent[i] < 0 || parent[i] >= i || (child1[parent[i]] != i && child2[parent[i]] != i))) { set_error("vrad_patches_set_hierarchy: bad parent of patch %d", i); return VRAD_E_INVALID; }
    }
    std::vector<int4> tree(n);
    P.h_root_cluster.assign(n, 0);
    for (int i = 0; i < n; i++) {                            // parents precede children: one forward pass
        P.h_root_cluster[i] = parent[i] == -1 ? P.h_cluster[i] : P.h_root_cluster[parent[i]];
        tree[i] = make_int4(parent[i], child1[i], face ? face[i] : -1, P.h_root_cluster[i]);
    }
    // flattened CollectLight rows: weights top-down, s = area_child / (area_child1 + area_child2)
    // subtree leaf counts (children follow their parents: one backward pass), then the interior patches long rows first
    std::vector<int32_t> n_leaves(n, 1);
    for (int i = n - 1; i >= 0; i--) if (child1[i] != -1) n_leaves[i] = n_leaves[child1[i]] + n_leaves[child2[i]];
    std::vector<int32_t> order;
    for (int pass = 0; pass < 2; pass++)
        for (int p = 0; p < n; p++)
            if (child1[p] != -1 && (n_leaves[p] >= kCollectLong) == (pass == 0)) order.push_back(p);
    int n_long = 0;
    for (int p : order) n_long += n_leaves[p] >= kCollectLong;
    std::vector<int32_t> ids; std::vector<int64_t> cptr(1, 0); std::vector<int2> ent;
    std::vector<std::pair<int, float>> work;
    for (int p : order) {
        ids.push_back(p);
        work.assign(1, std::make_pair(p, 1.0f));
        while (!work.empty()) {
            const auto [q, wq] = work.back(); work.pop_back();
            if (child1[q] == -1) { int2 v; v.x = q; memcpy(&v.y, &wq, 4); ent.push_back(v); continue; }
            const float a1 = P.h_area[child1[q]], a2 = P.h_area[child2[q]];
            work.push_back(std::make_pair(child2[q], wq * (a2 / (a1 + a2))));
            work.push_back(std::make_pair(child1[q], wq * (a1 / (a1 + a2))));
        }
        cptr.push_back((int64_t)ent.size());
    }
    if (P.tree.alloc(n) || P.collect_ids.alloc(ids.size() + 1) || P.collect_ptr.alloc(cptr.size()) || P.collect_ent.alloc(ent.size() + 1)) {
        set_error("out of device memory for the patch hierarchy"); return VRAD_E_NOMEM;
    }
    VRAD_CUDA_CHECK(cudaMemcpyAsync(P.tree.p, tree.data(), (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, e->stream));
    if (!ids.empty()) VRAD_CUDA_CHECK(cudaMemcpyAsync(P.collect_ids.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(P.collect_ptr.p, cptr.data(), cptr.size() * 8, cudaMemcpyHostToDevice, e->stream));
    if (!ent.empty()) VRAD_CUDA_CHECK(cudaMemcpyAsync(P.collect_ent.p, ent.data(), ent.size() * 8, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    P.n_interior = (int)ids.size();
    P.n_collect_long = n_long;
    P.h_child1.assign(child1, child1 + n);
    P.h_parent.assign(parent, parent + n);
    if (P.child2.alloc(n)) { set_error("out of device memory for the patch hierarchy"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpy(P.child2.p, child2, (size_t)n * 4, cudaMemcpyHostToDevice));
    P.leaf_rows_row0 = P.leaf_rows_row1 = -1;
    P.hier = true;
    e->transfers.ready = false;
    return VRAD_OK;
}

// upstream GetBumpNormals (host-only): basis around the phong normal from the texture S vector, mirrored for left-handed
// texture axes, then the fixed tangent-space bump basis rotated into world space
int vrad_bump_normals(const float s_vect[3], const float t_vect[3], const float flat_normal[3], const float phong_normal[3], float out9[9]) {
    if (!s_vect || !t_vect || !flat_normal || !phong_normal || !out9) { set_error("vrad_bump_normals: bad arguments"); return VRAD_E_INVALID; }
    const float kBasis[3][3] = {{0.81649661064147949f, 0.0f, 0.57735025882720947f},
                                {-0.40824821591377258f, 0.70710676908493042f, 0.57735025882720947f},
                                {-0.40824821591377258f, -0.70710676908493042f, 0.57735025882720947f}};
    struct V { float x, y, z; };
    auto cross = [](V a, V b) { return V{(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)}; };
    auto dot = [](V a, V b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); };
    auto unit = [&](V a) { const float len = sqrtf(dot(a, a)); if (len != 0.0f) { const float r = 1.0f / len; a = V{a.x * r, a.y * r, a.z * r}; } return a; };
    const V sv{s_vect[0], s_vect[1], s_vect[2]}, tv{t_vect[0], t_vect[1], t_vect[2]};
    const V flat{flat_normal[0], flat_normal[1], flat_normal[2]}, phong{phong_normal[0], phong_normal[1], phong_normal[2]};
    const bool left_handed = dot(flat, cross(sv, tv)) < 0.0f;
    V by = unit(cross(phong, sv));
    const V bx = unit(cross(by, phong));
    if (left_handed) by = V{-by.x, -by.y, -by.z};
    for (int b = 0; b < 3; b++) {
        out9[3 * b] = ((kBasis[b][0] * bx.x) + (kBasis[b][1] * by.x)) + (kBasis[b][2] * phong.x);
        out9[3 * b + 1] = ((kBasis[b][0] * bx.y) + (kBasis[b][1] * by.y)) + (kBasis[b][2] * phong.y);
        out9[3 * b + 2] = ((kBasis[b][0] * bx.z) + (kBasis[b][1] * by.z)) + (kBasis[b][2] * phong.z);
    }
    return VRAD_OK;
}

int vrad_patches_set_bump(vrad_env* e, int n, const uint8_t* needs_bump, const float* bump_normals9) {
    if (!e || !needs_bump || !bump_normals9) { set_error("vrad_patches_set_bump: bad arguments"); return VRAD_E_INVALID; }
    PatchesDev& P = e->patches;
    if (P.n == 0 || n != P.n) { set_error("vrad_patches_set_bump: %d entries for %d uploaded patches", n, P.n); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    std::vector<float4> bn(3 * (size_t)n);
    for (size_t i = 0; i < 3 * (size_t)n; i++) bn[i] = make_float4(bump_normals9[3 * i], bump_normals9[3 * i + 1], bump_normals9[3 * i + 2], 0.f);
    if (P.bump_normals.alloc(3 * (size_t)n)) { set_error("out of device memory for bump normals"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpy(P.bump_normals.p, bn.data(), bn.size() * sizeof(float4), cudaMemcpyHostToDevice));
    P.h_needs_bump.assign(needs_bump, needs_bump + n);
    P.bump_rows_row0 = P.bump_rows_row1 = -1;
    P.bump = true;
    return VRAD_OK;
}

int vrad_bounce_bump_totals(vrad_env* e, float* out9) {
    if (!e || !out9) { set_error("vrad_bounce_bump_totals: bad arguments"); return VRAD_E_INVALID; }
    PatchesDev& P = e->patches;
    if (!P.bump || P.total_bump[0].n < (size_t)P.n) { set_error("vrad_bounce_bump_totals: no bump-mapped bounce has run (vrad_patches_set_bump, vrad_bounce)"); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    std::vector<float4> t((size_t)P.n);
    for (int b = 0; b < 3; b++) {
        VRAD_CUDA_CHECK(cudaMemcpy(t.data(), P.total_bump[b].p, (size_t)P.n * sizeof(float4), cudaMemcpyDeviceToHost));
        for (int i = 0; i < P.n; i++) { out9[9 * (size_t)i + 3 * b] = t[i].x; out9[9 * (size_t)i + 3 * b + 1] = t[i].y; out9[9 * (size_t)i + 3 * b + 2] = t[i].z; }
    }
    return VRAD_OK;
}

int vrad_transfers_upload(vrad_env* e, int64_t row0, int64_t row1, const int64_t* rowptr, const int32_t* col, const float* w) {
    if (!e || !rowptr || row0 < 0 || row1 < row0) { set_error("vrad_transfers_upload: bad arguments"); return VRAD_E_INVALID; }
    const int64_t N = e->patches.n;
    if (N == 0) { set_error("vrad_transfers_upload: upload patches first"); return VRAD_E_STATE; }
    if (row1 > N) { set_error("vrad_transfers_upload: rows [%lld,%lld) exceed %lld patches", (long long)row0, (long long)row1, (long long)N); return VRAD_E_INVALID; }
    // with several ranks the row blocks must tile [0,N) in rank order; vrad_bounce checks that collectively
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const int64_t nloc = row1 - row0;
    const int64_t nnz = rowptr[nloc] - rowptr[0];
    if (nnz > 0 && (!col || !w)) { set_error("vrad_transfers_upload: col/w missing"); return VRAD_E_INVALID; }
    std::vector<int64_t> prow(nloc + 1);
    std::vector<int32_t> rlen(nloc ? nloc : 1);
    prow[0] = 0;
    for (int64_t i = 0; i < nloc; i++) {
        int64_t len = rowptr[i + 1] - rowptr[i];
        if (len < 0) { set_error("vrad_transfers_upload: rowptr not monotone at row %lld", (long long)i); return VRAD_E_INVALID; }
        rlen[i] = (int32_t)len;
        prow[i + 1] = prow[i] + ((len + 3) & ~(int64_t)3);
    }
    const int64_t np = prow[nloc];
    std::vector<int2> ptr(np ? np : 4, make_int2(0, 0));              // {col, w bits}; padding = {0, 0.0f}
    for (int64_t i = 0; i < nloc; i++) {
        const int64_t s = rowptr[i] - rowptr[0];
        for (int64_t k = 0; k < rlen[i]; k++) {
            int32_t c = col[s + k];
            if (c < 0 || c >= N) { set_error("vrad_transfers_upload: column %d out of range at row %lld", c, (long long)(row0 + i)); return VRAD_E_INVALID; }
            int2 v; v.x = c; memcpy(&v.y, &w[s + k], 4);
            ptr[prow[i] + k] = v;
        }
    }
    TransfersDev& T = e->transfers;
    if (T.rowptr.alloc(nloc + 1) || T.rowlen.alloc(nloc ? nloc : 1) || T.tr.alloc(ptr.size())) { set_error("out of device memory for transfers"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpyAsync(T.rowptr.p, prow.data(), (nloc + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(T.rowlen.p, rlen.data(), rlen.size() * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(T.tr.p, ptr.data(), ptr.size() * 8, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    T.row0 = row0; T.row1 = row1; T.nnz = nnz; T.nnz_padded = np; T.ready = true;
    return VRAD_OK;
}

int vrad_transfers_info(vrad_env* e, int64_t* row0, int64_t* row1, int64_t* nnz) {
    if (!e) return VRAD_E_INVALID;
    if (!e->transfers.ready) { set_error("vrad_transfers_info: no transfers resident"); return VRAD_E_STATE; }
    if (row0) *row0 = e->transfers.row0;
    if (row1) *row1 = e->transfers.row1;
    if (nnz) *nnz = e->transfers.nnz;
    return VRAD_OK;
}

int vrad_transfers_download(vrad_env* e, int64_t* rowptr, int32_t* col, float* w) {
    if (!e) return VRAD_E_INVALID;
    TransfersDev& T = e->transfers;
    if (!T.ready) { set_error("vrad_transfers_download: no transfers resident"); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const int64_t nloc = T.row1 - T.row0;
    std::vector<int64_t> prow(nloc + 1);
    std::vector<int32_t> rlen(nloc ? nloc : 1);
    VRAD_CUDA_CHECK(cudaMemcpy(prow.data(), T.rowptr.p, (nloc + 1) * 8, cudaMemcpyDeviceToHost));
    if (nloc) VRAD_CUDA_CHECK(cudaMemcpy(rlen.data(), T.rowlen.p, nloc * 4, cudaMemcpyDeviceToHost));
    std::vector<int2> ptr(T.nnz_padded ? T.nnz_padded : 1);
    if (T.nnz_padded) VRAD_CUDA_CHECK(cudaMemcpy(ptr.data(), T.tr.p, T.nnz_padded * 8, cudaMemcpyDeviceToHost));
    int64_t pos = 0;
    for (int64_t i = 0; i < nloc; i++) {
        if (rowptr) rowptr[i] = pos;
        for (int32_t k = 0; k < rlen[i]; k++) {
            if (col) col[pos + k] = ptr[prow[i] + k].x;
            if (w) memcpy(&w[pos + k], &ptr[prow[i] + k].y, 4);
        }
        pos += rlen[i];
    }
    if (rowptr) rowptr[nloc] = pos;
    return VRAD_OK;
}

int vrad_transfers_download_rows(vrad_env* e, int64_t row_begin, int64_t row_end, int64_t* rowptr, int32_t* col, float* w, int64_t capacity) {
    if (!e || !rowptr) { set_error("vrad_transfers_download_rows: bad arguments"); return VRAD_E_INVALID; }
    TransfersDev& T = e->transfers;
    if (!T.ready) { set_error("vrad_transfers_download_rows: no transfers resident"); return VRAD_E_STATE; }
    if (row_begin < T.row0 || row_end > T.row1 || row_end < row_begin) {
        set_error("vrad_transfers_download_rows: rows [%lld,%lld) outside this rank's [%lld,%lld)", (long long)row_begin, (long long)row_end, (long long)T.row0, (long long)T.row1);
        return VRAD_E_INVALID;
    }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const int64_t a = row_begin - T.row0, n = row_end - row_begin;
    std::vector<int64_t> prow(n + 1);
    std::vector<int32_t> rlen(n ? n : 1);
    VRAD_CUDA_CHECK(cudaMemcpy(prow.data(), T.rowptr.p + a, (n + 1) * 8, cudaMemcpyDeviceToHost));
    if (n) VRAD_CUDA_CHECK(cudaMemcpy(rlen.data(), T.rowlen.p + a, n * 4, cudaMemcpyDeviceToHost));
    int64_t need = 0;
    for (int64_t i = 0; i < n; i++) need += rlen[i];
    if (need > capacity || (need > 0 && (!col || !w))) { set_error("vrad_transfers_download_rows: %lld entries, capacity %lld", (long long)need, (long long)capacity); return VRAD_E_INVALID; }
    const int64_t span = prow[n] - prow[0];
    std::vector<int2> ptr(span ? span : 1);
    if (span) VRAD_CUDA_CHECK(cudaMemcpy(ptr.data(), T.tr.p + prow[0], span * 8, cudaMemcpyDeviceToHost));
    int64_t pos = 0;
    for (int64_t i = 0; i < n; i++) {
        rowptr[i] = pos;
        const int64_t base = prow[i] - prow[0];
        for (int32_t k = 0; k < rlen[i]; k++) { col[pos + k] = ptr[base + k].x; memcpy(&w[pos + k], &ptr[base + k].y, 4); }
        pos += rlen[i];
    }
    rowptr[n] = pos;
    return VRAD_OK;
}

int vrad_bounce(vrad_env* e, const float* emit0_rgb, int n_bounces, int early_out, float* total_rgb_out,
                float added_last[3], int* bounces_done) {
    if (!e || !emit0_rgb || n_bounces < 0) { set_error("vrad_bounce: bad arguments"); return VRAD_E_INVALID; }
    TransfersDev& T = e->transfers;
    if (!T.ready) { set_error("vrad_bounce: no transfers resident (vrad_build_transfers / vrad_transfers_upload first)"); return VRAD_E_STATE; }
    if (e->cfg.world > 1 && !e->nccl_comm) { set_error("vrad_bounce: world=%d but vrad_comm_init was not called", e->cfg.world); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const int64_t N = e->patches.n;
    const int world = e->cfg.world;
    const int64_t rpr = rows_per_rank(e, N);
    const int64_t n_pad = rpr * world;           // >= N; radiance buffers are indexed by global patch number
    int64_t bounds[kMaxWorld + 1] = {0};
    if (world > 1) {
        if (world > kMaxWorld) { set_error("vrad_bounce: world %d > %d", world, kMaxWorld); return VRAD_E_UNSUPPORTED; }
        int rcb = comm_exchange_bounds(e, T.row0, T.row1, N, bounds);
        if (rcb) return rcb;
    }
    const int nloc = (int)(T.row1 - T.row0);
    const int nblocks = std::max(1, (nloc + kGatherWarps - 1) / kGatherWarps);
    // `total` is indexed by global row too, so that the final gather is in place
    if (e->d_er[0].alloc(n_pad) || e->d_er[1].alloc(n_pad) || e->d_total.alloc(n_pad) || e->d_partials.alloc(3 * (size_t)nblocks + 8)) {
        set_error("out of device memory for bounce state"); return VRAD_E_NOMEM;
    }
    int rc; const void* d_emit0; bool h_in, h_out;
    if ((rc = stage_in(e, 0, emit0_rgb, (size_t)N * 12, &d_emit0, &h_in))) return rc;
    void* d_out3;
    if ((rc = stage_out(e, 1, total_rgb_out, (size_t)N * 12, &d_out3, &h_out))) return rc;
    float* d_added = e->d_partials.p + 3 * (size_t)nblocks;       // 3 floats after the partials
    // with a patch hierarchy the interior patches are recomputed from the exchanged leaf rows on every rank, which
    // needs the complete buffer before the next gather starts: the fused peer-store exchange is not used then
    const PatchesDev& PD = e->patches;
    const bool hier = PD.hier && PD.n_interior > 0;
    const int collect_long_blocks = (PD.n_collect_long + 7) / 8;
    const int collect_blocks = collect_long_blocks + (PD.n_interior - PD.n_collect_long + 31) / 32;
    if (world > 1 && !hier && (rc = comm_setup_peers(e, (size_t)n_pad))) return rc;
    const bool p2p = world > 1 && !hier && e->peers.ready;
    // short-row form: chosen by the average row length of the rows that gather (env VRAD_K4_SHORT=0/1 forces it)
    int n_short = nloc;
    const int32_t* d_rows = nullptr;
    if (hier) {     // local leaf rows only; the list is rebuilt when the row block changed
        if (PD.leaf_rows_row0 != T.row0 || PD.leaf_rows_row1 != T.row1) {
            std::vector<int32_t> lr;
            for (int64_t i = T.row0; i < T.row1; i++) if (PD.h_child1[i] == -1) lr.push_back((int32_t)(i - T.row0));
            PatchesDev& PM = e->patches;
            if (PM.leaf_rows.alloc(lr.size() + 1)) { set_error("out of device memory (leaf row list)"); return VRAD_E_NOMEM; }
            if (!lr.empty()) VRAD_CUDA_CHECK(cudaMemcpyAsync(PM.leaf_rows.p, lr.data(), lr.size() * 4, cudaMemcpyHostToDevice, e->stream));
            VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
            PM.n_leaf_rows = (int)lr.size(); PM.leaf_rows_row0 = T.row0; PM.leaf_rows_row1 = T.row1;
        }
        n_short = PD.n_leaf_rows; d_rows = PD.leaf_rows.p;
    }
    // bump-mapped leaf rows of this rank (TotalLight.Light[1..3] accumulate next to the flat gather)
    const bool bump = PD.bump;
    if (bump) {
        if (world > 1) { set_error("vrad_bounce: bump-mapped patches are not supported with world > 1 yet"); return VRAD_E_UNSUPPORTED; }
        PatchesDev& PM = e->patches;
        if (PM.bump_rows_row0 != T.row0 || PM.bump_rows_row1 != T.row1) {
            std::vector<int32_t> br;
            for (int64_t i = T.row0; i < T.row1; i++)
                if (PM.h_needs_bump[i] && !(PM.h_flags[i] & 1) && (!PM.hier || PM.h_child1[i] == -1)) br.push_back((int32_t)(i - T.row0));
            if (PM.bump_rows.alloc(br.size() + 1)) { set_error("out of device memory (bump row list)"); return VRAD_E_NOMEM; }
            if (!br.empty()) VRAD_CUDA_CHECK(cudaMemcpyAsync(PM.bump_rows.p, br.data(), br.size() * 4, cudaMemcpyHostToDevice, e->stream));
            VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
            PM.n_bump_rows = (int)br.size(); PM.bump_rows_row0 = T.row0; PM.bump_rows_row1 = T.row1;
        }
        for (int b = 0; b < 3; b++) {
            if (PM.total_bump[b].alloc((size_t)n_pad)) { set_error("out of device memory for bump totals"); return VRAD_E_NOMEM; }
            VRAD_CUDA_CHECK(cudaMemsetAsync(PM.total_bump[b].p, 0, (size_t)n_pad * 16, e->stream));
        }
    }
    static const int force_short = [] { const char* v = getenv("VRAD_K4_SHORT"); return v ? atoi(v) : -1; }();
    const bool use_short = !p2p && (force_short >= 0 ? force_short != 0 : (T.nnz < (int64_t)400 * std::max(1, n_short)));
    static const int short_cfg = [] { const char* v = getenv("VRAD_K4_SHORT_CFG"); return v ? atoi(v) : 884; }();   // experiments; see the switch below
    const int short_lanes = short_cfg / 10 == 16 ? 16 : (short_cfg / 10 == 4 ? 4 : 8);
    const int short_rpb = kGatherBlock / short_lanes;
    const int short_blocks = std::max(1, (n_short + short_rpb - 1) / short_rpb);
    const PeerTable* d_peers = p2p ? e->peers.d_table.p : nullptr;

    timing_begin(e);
    int launches = 0;
    float4* total_local = e->d_total.p + (world > 1 ? T.row0 : 0);
    VRAD_CUDA_CHECK(cudaMemsetAsync(e->d_total.p, 0, (size_t)n_pad * 16, e->stream));
    VRAD_CUDA_CHECK(cudaMemsetAsync(d_added, 0, 12, e->stream));
    k4_init_er<<<(int)((n_pad + 255) / 256), 256, 0, e->stream>>>((int)n_pad, (int)N, (const float*)d_emit0, e->patches.refl.p, e->d_er[0].p);
    launches++;

    int cur = 0, done = 0;
    float h_added[3] = {0.f, 0.f, 0.f};
    static const bool verbose = getenv("VRAD_TIMING") != nullptr;
    bool pending_wait = false;      // a peer signal was sent that no gather prologue has waited for yet
    constexpr int kProbe = 16;
    cudaEvent_t pe[kProbe][3];
    int n_probe = 0;
    for (int b = 0; b < n_bounces; b++) {
        const bool probe = verbose && b >= 4 && n_probe < kProbe;
        if (probe) { for (int k = 0; k < 3; k++) cudaEventCreate(&pe[n_probe][k]); cudaEventRecord(pe[n_probe][0], e->stream); }
        if (p2p)    // pending_wait false: nothing outstanding, the buffer read was initialised locally
            k4_gather_multi<true><<<nblocks, kGatherBlock, 0, e->stream>>>(nloc, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p, e->patches.refl.p,
                                                                   e->d_er[cur ^ 1].p, total_local, e->d_partials.p, d_peers, cur ^ 1, pending_wait ? world : 0);
        else if (use_short) {
#define VRAD_SHORT(L, B, ...) k4_gather_short<L, B, ##__VA_ARGS__><<<short_blocks, kGatherBlock, 0, e->stream>>>(n_short, d_rows, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p, \
                                                                 e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, e->d_partials.p)
            switch (short_cfg) {                                  // lanes per row, min blocks/SM[, unroll, prefetch]: measured on the hierarchical S2 matrix
                case 86: VRAD_SHORT(8, 6); break;                 // 64.1 us (40 registers, one-step prefetch)
                case 168: VRAD_SHORT(16, 8); break;               // 77 us
                case 48: VRAD_SHORT(4, 8); break;                 // 82 us
                default: VRAD_SHORT(8, 8, 4, false); break;       // 61.9 us: no prefetch, 32 registers, 64 warps per SM
            }
#undef VRAD_SHORT
        }
        else
            k4_gather<<<nblocks, kGatherBlock, 0, e->stream>>>(nloc, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p, e->patches.refl.p,
                                                             e->d_er[cur ^ 1].p, total_local, e->d_partials.p);
        launches++;
        pending_wait = false;
        if (bump && PD.n_bump_rows > 0) {       // same emitters as the gather above (er[cur]), bump-mapped rows only
            k4_gather_bump<<<(PD.n_bump_rows * 32 + 255) / 256, 256, 0, e->stream>>>(PD.n_bump_rows, PD.bump_rows.p, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p,
                                                                                  PD.origin_area.p, PD.normal_dist.p, PD.bump_normals.p,
                                                                                  PD.total_bump[0].p, PD.total_bump[1].p, PD.total_bump[2].p);
            launches++;
        }
        if (probe) cudaEventRecord(pe[n_probe][1], e->stream);
        if (p2p) { k4_peer_signal<<<1, 32, 0, e->stream>>>(d_peers); launches++; pending_wait = true; }
        else if (world > 1 && (rc = comm_allgather_rows(e, e->d_er[cur ^ 1].p, bounds))) return rc;
        if (hier) {     // CollectLight, interior patches: emit of a parent = area-weighted average of its children
            k4_collect_parents<<<collect_blocks, 256, 0, e->stream>>>(PD.n_interior, PD.n_collect_long, collect_long_blocks, PD.collect_ids.p, PD.collect_ptr.p, PD.collect_ent.p, e->d_er[cur ^ 1].p);
            launches++;
        }
        if (probe) { cudaEventRecord(pe[n_probe][2], e->stream); n_probe++; }
        cur ^= 1; done++;
        const bool last = (b + 1 == n_bounces);
        if (early_out || last) {
            k4_reduce_added<<<1, 256, 0, e->stream>>>(use_short ? short_blocks : nblocks, e->d_partials.p, d_added);
            launches++;
            if (world > 1 && (rc = comm_allreduce3(e, d_added))) return rc;
        }
        if (early_out && !last) {
            VRAD_CUDA_CHECK(cudaMemcpyAsync(h_added, d_added, 12, cudaMemcpyDeviceToHost, e->stream));
            VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
            if (h_added[0] < 1.0f && h_added[1] < 1.0f && h_added[2] < 1.0f) break;
        }
    }
    if (pending_wait) { k4_peer_wait<<<1, 32, 0, e->stream>>>(d_peers); launches++; }
    if (world > 1 && (rc = comm_allgather_rows(e, e->d_total.p, bounds))) return rc;
    if (hier) {         // totallight of the interior patches, from the leaves' totals
        k4_collect_parents<<<collect_blocks, 256, 0, e->stream>>>(PD.n_interior, PD.n_collect_long, collect_long_blocks, PD.collect_ids.p, PD.collect_ptr.p, PD.collect_ent.p, e->d_total.p);
        launches++;
    }
    if (d_out3) {
        k4_unpack_total<<<(int)((N + 255) / 256), 256, 0, e->stream>>>(N, e->d_total.p, (float*)d_out3);
        launches++;
    }
    timing_end(e, launches);
    VRAD_CUDA_CHECK(cudaGetLastError());
    if (n_probe) {
        cudaStreamSynchronize(e->stream);
        float g = 0.f, x = 0.f;
        for (int i = 0; i < n_probe; i++) {
            float a = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, pe[i][0], pe[i][1]); cudaEventElapsedTime(&c, pe[i][1], pe[i][2]);
            g += a; x += c;
            for (int k = 0; k < 3; k++) cudaEventDestroy(pe[i][k]);
        }
        fprintf(stderr, "[vrad] rank %d k4_gather %.1f us, exchange (%s) %.1f us per bounce (%d probed)\n", e->cfg.rank,
                1e3f * g / n_probe, p2p ? "peer stores + barrier" : (world > 1 ? "ncclAllGather" : "none"), 1e3f * x / n_probe, n_probe);
    }
    if ((rc = finish_out(e, total_rgb_out, d_out3, (size_t)N * 12, h_out))) return rc;
    const bool need_sync = h_in || h_out || added_last != nullptr;
    if (added_last) VRAD_CUDA_CHECK(cudaMemcpyAsync(added_last, d_added, 12, cudaMemcpyDeviceToHost, e->stream));
    if (bounces_done) *bounces_done = done;
    return sync_if_needed(e, need_sync);
}

} // extern "C"
