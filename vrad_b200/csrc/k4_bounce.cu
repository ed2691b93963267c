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
//
// The pairs are what vrad_build_transfers / vrad_transfers_upload leave and what downloads, the bump-mapped and the short-row kernels
// read.  The gather itself streams a denser form derived from them where the rows allow it (build_gather_plan; TransfersDev):
//   block rows  4 consecutive rows share the union of their columns: u16 column offset + 4 f32 weights per union entry -- 5.25 B per
//               transfer and one er[] gather per four rows on the C4 map (k4_gather_blocked, k4_gather_items_blocked);
//   packed      f32 weight + u16 column offset, 6 B per transfer (k4_gather_packed, k4_gather_items<.., PACKED>) -- rows that do not share
//               their columns, and slices too small for block-row work items (8 ranks on the C4 map);
//   pairs       everything else (rows over many 65,536-column windows, rows uploaded with unsorted columns, the patch hierarchy).
// Weights are the same f32 values in every form; the forms differ in the order a row's products are added (results agree to ~2e-7).
#include "env_internal.cuh"
#include <cub/device/device_scan.cuh>
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

// Single-GPU form: one warp per row, as many 8-row blocks as that takes, balanced by the hardware block scheduler.
// (r02: on one GPU this plain form beats every persistent / work-item variant of the loop tried below -- 279 us per
// bounce on the C4 matrix against 294-331 us -- the scheduler's dynamic balance is worth more than the start-up
// bubbles it leaves; the work-item form pays off where a barrier follows every bounce, i.e. with several GPUs.)
// Entries are {col, w} pairs -- the reference's Transfer struct
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

// ---------------------------------------------------------------------------------------------------------
// Packed transfer streams (TransfersDev::pk_*): 6 bytes per transfer instead of the 8 of the {col, w} pair -- the gather is
// HBM-bound on that stream, so the bytes are the time.  Columns of a row ascend and stay within the patches its cluster sees,
// so a 16-bit offset from a per-segment base covers them: one segment per row on the C4 map, a handful on the C5 map (a row
// there spans several 65,536-patch windows).  Weights stay f32, bit for bit.
// ---------------------------------------------------------------------------------------------------------
// Warp-uniform walk over the segments of one row of tr[]: f(index in row, first entry (relative), entries, column base).
template <typename F>
__device__ __forceinline__ void for_each_segment(const int2* __restrict__ tr, int64_t k0, int len, int max_seg, F f) {
    if (len <= 0) return;
    const int lane = threadIdx.x & 31;
    const int col0 = tr[k0].x;
    int seg_start = 0, n = 0;
    auto emit = [&](int a, int b) {
        const int base = col0 + (((tr[k0 + a].x - col0) >> 16) << 16);
        while (a < b) { const int l = min(b - a, max_seg); f(n++, a, l, base); a += l; }
    };
    for (int c = 0; c < len; c += 32) {
        const int en = c + lane;
        bool brk = false;
        if (en < len && en > 0) brk = ((tr[k0 + en].x - col0) >> 16) != ((tr[k0 + en - 1].x - col0) >> 16);
        unsigned m = __ballot_sync(0xffffffffu, brk);
        while (m) {
            const int pos = c + __ffs(m) - 1;
            m &= m - 1;
            emit(seg_start, pos);
            seg_start = pos;
        }
    }
    emit(seg_start, len);
}

__global__ void __launch_bounds__(256)
k4_pack_count(int nloc, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rowlen, const int2* __restrict__ tr, int max_seg,
              int32_t* __restrict__ n_segs, int64_t* __restrict__ padded) {
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= nloc) return;
    int ns = 0; int64_t pad = 0;
    for_each_segment(tr, rowptr[row], rowlen[row], max_seg, [&](int, int, int l, int) { ns++; pad += (l + 63) & ~63; });
    if ((threadIdx.x & 31) == 0) { n_segs[row] = ns; padded[row] = pad; }
}

__global__ void __launch_bounds__(256)
k4_pack_fill(int nloc, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rowlen, const int2* __restrict__ tr, int max_seg,
             const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ pk_rowptr, int4* __restrict__ segs,
             float* __restrict__ pw, uint16_t* __restrict__ pc) {
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= nloc) return;
    const int lane = threadIdx.x & 31;
    const int64_t k0 = rowptr[row];
    int64_t out = pk_rowptr[row];
    const int s0 = seg_ptr[row];
    for_each_segment(tr, k0, rowlen[row], max_seg, [&](int idx, int a, int l, int base) {
        const int lp = (l + 63) & ~63;
        if (lane == 0) segs[s0 + idx] = make_int4((int)(uint32_t)(out & 0xffffffff), (int)(out >> 32), lp, base);
        for (int t = lane; t < lp; t += 32) {
            // position t of a 64-entry group holds entry (t odd ? 32 : 0) + t/2 of the group: the gather's lane L reads positions
            // 2L, 2L+1 with one vector load and so holds entries L and 32+L -- its er[] gathers then run over CONSECUTIVE entries
            // across the warp, like the pair kernel's (adjacent columns share cache lines; with lanes on entries 2L, 2L+1 every
            // line was touched by two gather instructions: 305 us per bounce on the C4 matrix against 276 us for the pairs)
            const int src = (t & ~63) + ((t & 1) << 5) + ((t & 63) >> 1);
            int2 en = make_int2(base, 0);
            if (src < l) en = tr[k0 + a + src];
            pw[out + t] = __int_as_float(en.y);
            pc[out + t] = (uint16_t)(en.x - base);
        }
        out += lp;
    });
}

// The gather from the packed streams, one warp per row (the single-GPU form of k4_gather above): a lane takes two consecutive
// entries per step -- one 8-byte weight load and one 4-byte column load, both coalesced -- four steps in flight.
struct PackedRows { const int32_t* seg_ptr; const int4* segs; const float* w; const uint16_t* c; };

template <int kPackUnroll>
__device__ __forceinline__ void gather_packed_segments(const PackedRows& P, int sg0, int sg1, const float4* __restrict__ er, int lane,
                                                       float& s0, float& s1, float& s2) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int sg = sg0; sg < sg1; sg++) {
        const int4 d = __ldg(&P.segs[sg]);
        const int64_t k0 = (int64_t)(uint32_t)d.x | ((int64_t)d.y << 32);
        const int npairs = d.z >> 1, base = d.w;
        const float2* __restrict__ w2 = reinterpret_cast<const float2*>(P.w + k0);
        const uint32_t* __restrict__ c2 = reinterpret_cast<const uint32_t*>(P.c + k0);
        float2 cw[kPackUnroll], nw[kPackUnroll];
        uint32_t cc[kPackUnroll], nc[kPackUnroll];
        int p = lane;
#pragma unroll
        for (int j = 0; j < kPackUnroll; j++) {
            const bool in = p + 32 * j < npairs;
            cw[j] = in ? __ldcs(&w2[p + 32 * j]) : make_float2(0.f, 0.f);
            cc[j] = in ? __ldcs(&c2[p + 32 * j]) : 0u;
        }
        for (; p < npairs; p += 32 * kPackUnroll) {
#pragma unroll
            for (int j = 0; j < kPackUnroll; j++) {
                const bool in = p + 32 * (kPackUnroll + j) < npairs;
                nw[j] = in ? __ldcs(&w2[p + 32 * (kPackUnroll + j)]) : make_float2(0.f, 0.f);
                nc[j] = in ? __ldcs(&c2[p + 32 * (kPackUnroll + j)]) : 0u;
            }
            float4 x[2 * kPackUnroll];
#pragma unroll
            for (int j = 0; j < kPackUnroll; j++) {
                x[2 * j] = __ldg(&er[base + (int)(cc[j] & 0xffffu)]);
                x[2 * j + 1] = __ldg(&er[base + (int)(cc[j] >> 16)]);
            }
#pragma unroll
            for (int j = 0; j < kPackUnroll; j++) {
                s0 += cw[j].x * x[2 * j].x; s1 += cw[j].x * x[2 * j].y; s2 += cw[j].x * x[2 * j].z;
                s0 += cw[j].y * x[2 * j + 1].x; s1 += cw[j].y * x[2 * j + 1].y; s2 += cw[j].y * x[2 * j + 1].z;
            }
#pragma unroll
            for (int j = 0; j < kPackUnroll; j++) { cw[j] = nw[j]; cc[j] = nc[j]; }
        }
        if (sg1 - sg0 > 1) {
            // a row of several segments: each segment's sum is reduced on its own and the sums are added in segment order, lane-strided
            // and then by the xor tree -- the arithmetic of combine_parts() below, where the multi-GPU kernel treats the segments of
            // such a row as parts (so that one GPU and several give the same bits)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == ((sg - sg0) & 31)) { a0 += s0; a1 += s1; a2 += s2; }
            s0 = s1 = s2 = 0.f;
        }
    }
    if (sg1 - sg0 > 1) { s0 = a0; s1 = a1; s2 = a2; }
}

template <int kPackUnroll, int kMinBlocks>
__global__ void __launch_bounds__(kGatherBlock, kMinBlocks)
k4_gather_packed(int nloc, int64_t row0, PackedRows P, const float4* __restrict__ er, const float4* __restrict__ refl,
                 float4* __restrict__ er_next, float4* __restrict__ total, float* __restrict__ partials) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * kGatherWarps + warp;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (row < nloc) {
        gather_packed_segments<kPackUnroll>(P, P.seg_ptr[row], P.seg_ptr[row + 1], er, lane, s0, s1, s2);
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

// ---------------------------------------------------------------------------------------------------------
// Block-row streams (TransfersDev::bk_*): kBlockRows consecutive rows share one column list, the union of theirs.
// Built per block of rows by one warp with a bitmap of the current 65,536-column window in shared memory: the rows' entries set their
// column's bit; the union position of a column is the number of set bits below it (per-word popcount prefix); the rows' weights go to
// component r of the float4 at that position.  Two passes (sizes, then contents) around two scans, like the packed streams.
// ---------------------------------------------------------------------------------------------------------
constexpr int kBlockRows = 4;                       // the most rows a block of rows may hold (TransfersDev::bk_rows = 2 or 4)
constexpr int kBkWords = 65536 / 32;
constexpr int kBkBuildWarps = 2;

struct BkRowSet { int64_t k0[kBlockRows]; int len[kBlockRows]; };

// FILL = false: calls seg(n_entries) per window piece and returns; FILL = true: also writes columns and weights.
template <bool FILL>
__device__ __forceinline__ void block_row_build(const int2* __restrict__ tr, const BkRowSet& R, int max_seg, uint32_t* bm, uint16_t* pre,
                                                int& n_segs, int64_t& n_padded, int4* __restrict__ segs, int64_t out,
                                                float* __restrict__ bw, uint16_t* __restrict__ bc, int rows) {
    const int lane = threadIdx.x & 31;
    int col_min = 0x7fffffff, col_max = -1;
#pragma unroll
    for (int r = 0; r < kBlockRows; r++)
        if (R.len[r] > 0) { col_min = min(col_min, tr[R.k0[r]].x); col_max = max(col_max, tr[R.k0[r] + R.len[r] - 1].x); }
    n_segs = 0; n_padded = 0;
    if (col_max < 0) return;
    int cur[kBlockRows];
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) cur[r] = 0;
    for (int base = col_min; base <= col_max; base += 65536) {
        for (int q = lane; q < kBkWords; q += 32) bm[q] = 0u;
        __syncwarp();
        int first[kBlockRows];
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            first[r] = cur[r];
            for (;;) {
                const int en = cur[r] + lane;
                int c = 0;
                const bool in = en < R.len[r] && (c = tr[R.k0[r] + en].x - base) < 65536;
                if (in) atomicOr(&bm[c >> 5], 1u << (c & 31));
                const int took = __popc(__ballot_sync(0xffffffffu, in));
                cur[r] += took;
                if (took < 32) break;
            }
        }
        __syncwarp();
        // per-word exclusive popcount prefix: lane owns 64 consecutive words
        int mine = 0;
        for (int q = 0; q < kBkWords / 32; q++) mine += __popc(bm[lane * (kBkWords / 32) + q]);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int count = __shfl_sync(0xffffffffu, incl, 31);
        if (count == 0) continue;
        if (FILL) {
            int run = incl - mine;
            for (int q = 0; q < kBkWords / 32; q++) { pre[lane * (kBkWords / 32) + q] = (uint16_t)run; run += __popc(bm[lane * (kBkWords / 32) + q]); }
            __syncwarp();
        }
        // this window's union, cut into pieces of at most max_seg entries
        const int n_pieces = (count + max_seg - 1) / max_seg;
        const int64_t win_out = out + n_padded;                 // pieces are contiguous: piece p starts at win_out + p * max_seg (max_seg is a multiple of 32)
        for (int p = 0; p < n_pieces; p++) {
            const int l = min(max_seg, count - p * max_seg), lp = (l + 31) & ~31;
            if (FILL && lane == 0) {
                const int64_t st = win_out + (int64_t)p * max_seg;
                segs[n_segs] = make_int4((int)(uint32_t)(st & 0xffffffff), (int)(st >> 32), lp, base);
            }
            n_segs++; n_padded += lp;
        }
        if (FILL) {
            // columns: every set bit writes its offset at its union position
            for (int q = lane; q < kBkWords; q += 32) {
                uint32_t m = bm[q];
                int pos = pre[q];
                while (m) { const int b = __ffs(m) - 1; m &= m - 1; bc[win_out + pos] = (uint16_t)(q * 32 + b); pos++; }
            }
            // weights: component r of the entry at the column's union position
#pragma unroll
            for (int r = 0; r < kBlockRows; r++) {
                for (int en = first[r] + lane; en < cur[r]; en += 32) {
                    const int2 t = tr[R.k0[r] + en];
                    const int c = t.x - base;
                    const int pos = pre[c >> 5] + __popc(bm[c >> 5] & ((1u << (c & 31)) - 1u));
                    bw[(win_out + pos) * rows + r] = __int_as_float(t.y);
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ BkRowSet block_rows_of(int blk, int rows, int nloc, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rowlen) {
    BkRowSet R;
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) {
        const int row = blk * rows + r;
        const bool in = r < rows && row < nloc;
        R.k0[r] = in ? rowptr[row] : 0;
        R.len[r] = in ? rowlen[row] : 0;
    }
    return R;
}

__global__ void __launch_bounds__(kBkBuildWarps * 32)
k4_block_count(int n_blocks, int rows, int nloc, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rowlen, const int2* __restrict__ tr, int max_seg,
               int32_t* __restrict__ n_segs, int64_t* __restrict__ padded) {
    __shared__ uint32_t bm_s[kBkBuildWarps][kBkWords];
    const int warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * kBkBuildWarps + warp;
    if (blk >= n_blocks) return;
    int ns; int64_t np;
    block_row_build<false>(tr, block_rows_of(blk, rows, nloc, rowptr, rowlen), max_seg, bm_s[warp], nullptr, ns, np, nullptr, 0, nullptr, nullptr, rows);
    if ((threadIdx.x & 31) == 0) { n_segs[blk] = ns; padded[blk] = np; }
}

__global__ void __launch_bounds__(kBkBuildWarps * 32)
k4_block_fill(int n_blocks, int rows, int nloc, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rowlen, const int2* __restrict__ tr, int max_seg,
              const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ bk_start, int4* __restrict__ segs, float* __restrict__ bw, uint16_t* __restrict__ bc) {
    __shared__ uint32_t bm_s[kBkBuildWarps][kBkWords];
    __shared__ uint16_t pre_s[kBkBuildWarps][kBkWords];
    const int warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * kBkBuildWarps + warp;
    if (blk >= n_blocks) return;
    int ns; int64_t np;
    block_row_build<true>(tr, block_rows_of(blk, rows, nloc, rowptr, rowlen), max_seg, bm_s[warp], pre_s[warp], ns, np, segs + seg_ptr[blk], bk_start[blk], bw, bc, rows);
}

// The gather from the block-row streams: one warp per block of kBlockRows rows; a lane takes one entry per step -- a 2-byte column, a
// 16-byte weight vector, ONE 16-byte er[] gather -- and keeps kBlockRows x 3 sums.
struct BlockedRows { const int32_t* seg_ptr; const int4* segs; const float* w; const uint16_t* c; };

// R weights of one entry: one 8- or 16-byte load
template <int R> __device__ __forceinline__ void load_weights(const float* __restrict__ w, int64_t entry, float (&out)[R]) {
    if constexpr (R == 4) { const float4 v = __ldcs(reinterpret_cast<const float4*>(w) + entry); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
    else { const float2 v = __ldcs(reinterpret_cast<const float2*>(w) + entry); out[0] = v.x; out[1] = v.y; }
}

template <int R, int U>
__device__ __forceinline__ void gather_block_segments(const BlockedRows& P, int sg0, int sg1, const float4* __restrict__ er, int lane, float (&acc)[R][3]) {
    for (int sg = sg0; sg < sg1; sg++) {
        const int4 d = __ldg(&P.segs[sg]);
        const int64_t k0 = (int64_t)(uint32_t)d.x | ((int64_t)d.y << 32);
        const int len = d.z, base = d.w;
        const uint16_t* __restrict__ c1 = P.c + k0;
        float cw[U][R], nw[U][R];
        int cc[U], nc[U];
        int p = lane;
#pragma unroll
        for (int j = 0; j < U; j++) {
#pragma unroll
            for (int r = 0; r < R; r++) cw[j][r] = 0.f;
            cc[j] = 0;
            if (p + 32 * j < len) { load_weights<R>(P.w, k0 + p + 32 * j, cw[j]); cc[j] = (int)__ldcs(&c1[p + 32 * j]); }
        }
        for (; p < len; p += 32 * U) {
#pragma unroll
            for (int j = 0; j < U; j++) {
#pragma unroll
                for (int r = 0; r < R; r++) nw[j][r] = 0.f;
                nc[j] = 0;
                if (p + 32 * (U + j) < len) { load_weights<R>(P.w, k0 + p + 32 * (U + j), nw[j]); nc[j] = (int)__ldcs(&c1[p + 32 * (U + j)]); }
            }
            float4 x[U];
#pragma unroll
            for (int j = 0; j < U; j++) x[j] = __ldg(&er[base + cc[j]]);
#pragma unroll
            for (int j = 0; j < U; j++)
#pragma unroll
                for (int r = 0; r < R; r++) { acc[r][0] += cw[j][r] * x[j].x; acc[r][1] += cw[j][r] * x[j].y; acc[r][2] += cw[j][r] * x[j].z; }
#pragma unroll
            for (int j = 0; j < U; j++) {
#pragma unroll
                for (int r = 0; r < R; r++) cw[j][r] = nw[j][r];
                cc[j] = nc[j];
            }
        }
    }
}

template <int R, int U, int kMinBlocks>
__global__ void __launch_bounds__(kGatherBlock, kMinBlocks)
k4_gather_blocked(int n_blocks, int nloc, int64_t row0, BlockedRows P, const float4* __restrict__ er, const float4* __restrict__ refl,
                  float4* __restrict__ er_next, float4* __restrict__ total, float* __restrict__ partials) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * kGatherWarps + warp;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (blk < n_blocks) {
        float acc[R][3];
#pragma unroll
        for (int r = 0; r < R; r++) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; }
        gather_block_segments<R, U>(P, P.seg_ptr[blk], P.seg_ptr[blk + 1], er, lane, acc);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], o);
        // lane r finishes row r of the block (CollectLight); the block's share of `added` is summed in row order
        float s0 = acc[0][0], s1 = acc[0][1], s2 = acc[0][2];
#pragma unroll
        for (int r = 1; r < R; r++) if (lane == r) { s0 = acc[r][0]; s1 = acc[r][1]; s2 = acc[r][2]; }
        const int row = blk * R + lane;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (lane < R && row < nloc) {
            const float4 rf = refl[row0 + row];
            if (rf.w == 0.0f) {
                float4 t = total[row];
                t.x += s0; t.y += s1; t.z += s2;
                total[row] = t;
                er_next[row0 + row] = make_float4(s0 * rf.x, s1 * rf.y, s2 * rf.z, 0.f);
                a0 = s0; a1 = s1; a2 = s2;
            } else {
                er_next[row0 + row] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            e0 += __shfl_sync(0xffffffffu, a0, r); e1 += __shfl_sync(0xffffffffu, a1, r); e2 += __shfl_sync(0xffffffffu, a2, r);
        }
    }
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

// ---------------------------------------------------------------------------------------------------------
// The gather.  One warp per WORK ITEM (TransfersDev::items): a whole row, or -- for a row longer than `seg`
// entries -- one `seg`-entry part of it.  Entries are {col, w} pairs -- the reference's Transfer struct
// (common/types/transfer.go:3-6) -- read as one coalesced 64-bit load per lane with lanes on CONSECUTIVE
// entries, so the 32 er[] gathers of one instruction hit consecutive patches wherever the row has a run of
// adjacent columns (avg run length ~16 on the synthetic maps): few L1 wavefronts per gather instead of one per
// lane (ncu r01: the int4-per-lane mapping was L1TEX-bound at 86%).  kGatherUnroll entries per lane are in
// flight and the {col,w} loads of the NEXT step are issued before the current step's gathers (software
// pipeline), so the two dependent memory latencies overlap.  The kernel is latency-bound: what matters is bytes
// in flight per SM = resident warps x entries in flight, so the register budget is capped to keep 5 blocks (40
// warps) per SM -- the same loop at 60 registers (4 blocks) ran at 4.1 TB/s, at 48 registers 6.4 TB/s
// (tools/exp/k4_exp.cu).
//
// Why items: with one warp per row a 4000-entry row is a ~30 us chain of 16 dependent steps; started late it
// is the whole tail of a kernel that should take 36 us on an 8-rank slice of the C4 matrix (r01 verdict: 54 us
// per bounce against 36 us of streaming).  Parts bound the longest chain; their sums meet in part_sum[] and the
// part that arrives last adds them in part order (so the row sum does not depend on scheduling) and runs the
// row's epilogue.
//
// Multi-GPU form (MULTI): the epilogue stores the finished row into every rank's next-bounce buffer (lane p ->
// rank p; NVLink peer stores), and the inter-bounce barrier lives inside the kernel --
//   signal: the block that finishes last (ticket) publishes this bounce's epoch in every rank's arrival words;
//   wait:   each block, after it has issued its first {col,w} loads (which do not depend on the peers), spins
//           until all ranks' arrival words (its own included) have reached the previous bounce's epoch.
// Consecutive bounces are therefore ordered by the flags alone, and the next bounce's kernel is launched with
// programmatic stream serialisation (PDL): its blocks fill the SM slots that the current bounce's tail frees
// and have their first loads in flight when the last signal arrives.  One launch per bounce, no barrier kernel.
// Memory model: row stores -> bar.sync -> fence.acq_rel.gpu + ticket atomic (release, gpu scope) -> last block:
// fence.acq_rel.sys -> st.relaxed.sys arrival words (release, sys scope) -> waiter: ld.acquire.sys -> bar.sync
// -> plain (ld.global, not ld.global.nc) loads of er[].


struct GatherAux {                 // rarely used pointers, kept in the parameter bank (no registers until touched)
    float4* add;                   // light added per local row (written when non-null: last bounce / early-out)
    float4* part_sum;
    int32_t* row_ctr;
    const PeerTable* peers;        // MULTI only
    uint32_t* flags;               // this rank's flag words (MULTI only)
    int next_buf;
    uint32_t wait_rel, signal_rel; // epochs relative to flags[kFlagBase]; wait_rel 0 = nothing to wait for
    int wait_world;
    int pool_begin, n_items;       // items [pool_begin, n_items) belong to no block: whoever runs dry takes them one by one (flags[kFlagPool])
    const float* pk_w;             // PACKED only: the weight and column streams (TransfersDev::pk_w / pk_c)
    const uint16_t* pk_c;
    const float* bk_w;             // block-row kernel only (TransfersDev::bk_w / bk_c)
    const uint16_t* bk_c;
    int nloc;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// wait half of the inter-bounce barrier; block-wide.  Gives up after ~2^30 cycles (about half a second: a peer died) and
// sets the error word, which makes every later wait return at once -- vrad_bounce then reports VRAD_E_COMM instead of
// hanging the device.
__device__ __noinline__ void wait_for_peers(const uint32_t* flags, int wait_world, uint32_t wait_rel) {
    if ((int)threadIdx.x < wait_world) {
        // poll with relaxed loads -- an acquire per poll would drop this SM's L1 (CCTL.IVALL) under the blocks of the
        // previous bounce that still gather on it (r02: the PDL chain was 4 us per bounce SLOWER with acquire polls)
        const uint32_t epoch = ld_relaxed_sys(flags + kFlagBase) + wait_rel;
        const long long t0 = clock64();
        while ((int32_t)(ld_relaxed_sys(flags + threadIdx.x) - epoch) < 0) {
            if (ld_relaxed_sys(flags + kFlagError)) break;
            if (clock64() - t0 > (1LL << 30)) { atomicExch((unsigned int*)flags + kFlagError, 1u); break; }
        }
        (void)ld_acquire_sys(flags + threadIdx.x);       // the acquire that orders the radiance loads after the arrival
    }
    __syncthreads();
}

// signal half: called by one thread per block after the block's bar.sync
// An acquire fence makes the SM drop its L1 (CCTL.IVALL) -- paid per caller it wrecks the er[] gathers' L1 hit rate (r02:
// with a full fence per split-row part, 1024-entry items ran at 445 us per bounce against 280 us).  So arrivals only
// RELEASE (atom.release: prior writes ordered, nothing invalidated) and the single thread that finds itself last
// acquires.
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ uint32_t atom_add_release_gpu(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__device__ __noinline__ void signal_if_last(const PeerTable* __restrict__ peers, uint32_t* flags, uint32_t signal_rel) {
    const unsigned int t = atom_add_release_gpu(flags + kFlagTicket, 1u);
    if (t == gridDim.x - 1) {
        flags[kFlagTicket] = 0u;
        flags[kFlagPool] = 0u;
        const uint32_t epoch = flags[kFlagBase] + signal_rel;
        fence_acq_rel_sys();
        const int world = peers->world, rank = peers->rank;
        for (int p = 0; p < world; p++) st_relaxed_sys(peers->flags[p] + rank, epoch);
    }
}

// out of line on purpose: keeps the peer-table loads out of the gather loop's register allocation
__device__ __noinline__ void store_row_to_peers(const PeerTable* __restrict__ peers, int next_buf, int64_t row, float x, float y, float z) {
    const int lane = threadIdx.x & 31;
    x = __shfl_sync(0xffffffffu, x, 0); y = __shfl_sync(0xffffffffu, y, 0); z = __shfl_sync(0xffffffffu, z, 0);
    if (lane < peers->world) peers->er[next_buf][lane][row] = make_float4(x, y, z, 0.f);
}

// a part of a split row has its sum: publish it, and if it is the last part to arrive add all parts in part order.
// Warp-uniform; returns true on the finishing warp with the row sums in s0..s2.
__device__ __noinline__ bool combine_parts(float4* part_sum, int32_t* row_ctr, int row, int slot0, int idx, int n_parts,
                                           float& s0, float& s1, float& s2) {
    const int lane = threadIdx.x & 31;
    int last = 0;
    if (lane == 0) {
        __stcg(&part_sum[slot0 + idx], make_float4(s0, s1, s2, 0.f));
        last = atom_add_release_gpu((uint32_t*)&row_ctr[row], 1u) == (uint32_t)(n_parts - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return false;
    fence_acq_rel_gpu();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int p = lane; p < n_parts; p += 32) {          // fixed order: lane-strided, then the xor tree
        const float4 v = __ldcg(&part_sum[slot0 + p]);
        a0 += v.x; a1 += v.y; a2 += v.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) row_ctr[row] = 0;
    s0 = a0; s1 = a1; s2 = a2;
    return true;
}

template <bool MULTI>
__device__ __forceinline__ float4 load_er(const float4* er, int col) {
    if (MULTI) return er[col];          // peers store into this buffer between bounces: no read-only (.nc) path
    return __ldg(&er[col]);
}

// Blocks: block b owns the items [block_ptr[b], block_ptr[b+1]) and its warps take them one at a time through a
// shared-memory counter.  Two plans (build_gather_plan): "dispatched" -- 8 consecutive items per block, as many blocks
// as that takes, balanced by the hardware block scheduler -- and "persistent" -- exactly one block per resident slot
// (5 per SM), the item list cut into contiguous ranges of equal work and each range ordered longest row first, so that
// every slot runs the whole bounce, finishes within one short row of the others, and -- multi-GPU -- pays the
// release + ticket once per bounce instead of once per 8 rows.
//
// A warp never idles between two items: an item descriptor is self-contained ({row, entries | parts | part index,
// first entry as int64}: no row-pointer hop), the NEXT item is claimed while the current one streams and its
// descriptor travels global -> shared memory by cp.async (no registers held across the loop), and the next item's
// first {col,w} loads are issued BEFORE the current item's reduction and epilogue.  r02 ncu, single-GPU stand-in for
// the 8-rank slice: an isolated bounce took 53 us for 195 MB where the steady rate of the same loop (5.5 TB/s) needs
// 35 us -- the difference is start-up chains (claim -> item -> row pointer -> entries -> radiance, once per row per
// warp) that nothing overlaps at the start and the end of a kernel that short.
constexpr int kNoItem = 0x7fffffff;
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// WARPS x blocks per SM: 8 x 5 = 40 resident warps at 48 registers, or 6 x 6 = 36 resident warps at 56 registers
// CS: the {col,w} stream is read with ld.global.cs (evict first); without it the loads leave the replacement decision to the L2
// access-policy window (k4_l2_mb), which keeps the head of the stream resident from bounce to bounce
template <bool CS> __device__ __forceinline__ int2 load_tr(const int2* p) { return CS ? __ldcs(p) : *p; }

// PACKED: an item is one SEGMENT of the packed streams -- {row, padded entries | parts | part index, first entry / 64, column base} --
// and a lane holds kGatherUnroll / 2 pairs of entries (8-byte weight load + 4-byte column load) instead of kGatherUnroll {col,w} pairs;
// a row of several segments is a row of several parts.  Same lane-to-entry mapping as k4_gather_packed: the two kernels agree bit for bit.
template <bool CS> __device__ __forceinline__ float2 load_w2(const float2* p) { return CS ? __ldcs(p) : *p; }
template <bool CS> __device__ __forceinline__ uint32_t load_c2(const uint32_t* p) { return CS ? __ldcs(p) : *p; }

template <bool MULTI, int WARPS, bool CS = true, bool PACKED = false>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? 5 : 6)
k4_gather_items(const int32_t* __restrict__ block_ptr, const int4* __restrict__ items, const int32_t* __restrict__ item_slot, int64_t row0,
                const int2* __restrict__ tr, const float4* er, const float4* __restrict__ refl,
                float4* er_next, float4* __restrict__ total, GatherAux A) {
    __shared__ int next_item;
    __shared__ int4 desc_s[WARPS], hold_s[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MULTI) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: the next bounce may be scheduled behind this one
    if (threadIdx.x == 0) next_item = __ldg(&block_ptr[blockIdx.x]);
    __syncthreads();
    // next item of this block's range; when the range is used up, of the common pool (global counter) -- equal work is
    // not equal time (cache hit rates and DRAM conflicts differ from range to range), and with a barrier behind every
    // bounce the slots that finish early would otherwise idle until the slowest range is done
    // the common pool's counter belongs to the bounce that is running: a block of the NEXT bounce (launched behind it, PDL) may take from the
    // pool only after its barrier wait, i.e. after the running bounce's last block has reset the counter
    bool pool_ok = !MULTI || A.wait_rel == 0;
    auto claim = [&]() {
        int c = kNoItem;
        if (lane == 0) {
            c = atomicAdd(&next_item, 1);
            if (c >= __ldg(&block_ptr[blockIdx.x + 1])) {
                c = kNoItem;
                if (MULTI && pool_ok && A.pool_begin < A.n_items) {
                    const int q = A.pool_begin + (int)atomicAdd((unsigned int*)A.flags + kFlagPool, 1u);
                    if (q < A.n_items) c = q;
                }
            }
        }
        return __shfl_sync(0xffffffffu, c, 0);
    };
    const int2 zero = make_int2(0, 0);                              // out-of-item slots: col 0, weight 0
    int w = claim();
    int4 it = make_int4(0, 0, 0, 0);
    if (w != kNoItem) it = __ldg(&items[w]);
    constexpr int kHalf = kGatherUnroll / 2;
    int2 cur[PACKED ? 1 : kGatherUnroll], nxt[PACKED ? 1 : kGatherUnroll];
    float2 cw[PACKED ? kHalf : 1], nw[PACKED ? kHalf : 1];
    uint32_t cc[PACKED ? kHalf : 1], nc[PACKED ? kHalf : 1];
    if (PACKED) {
        const int64_t k0 = (int64_t)(uint32_t)it.z << 6;
        const float2* w0 = reinterpret_cast<const float2*>(A.pk_w + k0);
        const uint32_t* c0 = reinterpret_cast<const uint32_t*>(A.pk_c + k0);
        const int l0 = (it.y & 0xffff) >> 1;
#pragma unroll
        for (int j = 0; j < kHalf; j++) {
            const bool in = lane + 32 * j < l0;
            cw[j] = in ? load_w2<CS>(&w0[lane + 32 * j]) : make_float2(0.f, 0.f);
            cc[j] = in ? load_c2<CS>(&c0[lane + 32 * j]) : 0u;
        }
    } else {
        const int2* p0 = tr + (((int64_t)it.w << 32) | (uint32_t)it.z);
        const int l0 = it.y & 0xffff;
#pragma unroll
        for (int j = 0; j < kGatherUnroll; j++) cur[j] = lane + 32 * j < l0 ? load_tr<CS>(&p0[lane + 32 * j]) : zero;
    }
    int wn = w != kNoItem ? claim() : kNoItem;
    if (wn != kNoItem && lane == 0) cp_async16(&desc_s[warp], &items[wn]);
    if (MULTI) {
        // the radiance this bounce reads is complete once every rank's previous-bounce epoch has arrived; every warp
        // passes here exactly once, with its first {col,w} loads (which do not depend on the peers) in flight
        if (A.wait_rel) wait_for_peers(A.flags, A.wait_world, A.wait_rel);
        pool_ok = true;
    }
    bool has = w != kNoItem;
    while (has) {
        // warp-uniform bookkeeping sits in shared memory while the item streams: the loop below runs at the register
        // limit that keeps 5 blocks per SM resident, and anything live across it is paid for in spills inside it
        if (lane == 0) hold_s[warp] = make_int4(it.x, it.y, w, wn);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (PACKED) {
            const int64_t k0 = (int64_t)(uint32_t)it.z << 6;
            const float2* w2 = reinterpret_cast<const float2*>(A.pk_w + k0);
            const uint32_t* c2 = reinterpret_cast<const uint32_t*>(A.pk_c + k0);
            const int npairs = (it.y & 0xffff) >> 1, base = it.w;
            for (int off = lane; off < npairs; off += 32 * kHalf) {
#pragma unroll
                for (int j = 0; j < kHalf; j++) {
                    const bool in = off + 32 * (kHalf + j) < npairs;
                    nw[j] = in ? load_w2<CS>(&w2[off + 32 * (kHalf + j)]) : make_float2(0.f, 0.f);
                    nc[j] = in ? load_c2<CS>(&c2[off + 32 * (kHalf + j)]) : 0u;
                }
                float4 x[kGatherUnroll];
#pragma unroll
                for (int j = 0; j < kHalf; j++) {
                    x[2 * j] = load_er<MULTI>(er, base + (int)(cc[j] & 0xffffu));
                    x[2 * j + 1] = load_er<MULTI>(er, base + (int)(cc[j] >> 16));
                }
#pragma unroll
                for (int j = 0; j < kHalf; j++) {
                    s0 += cw[j].x * x[2 * j].x; s1 += cw[j].x * x[2 * j].y; s2 += cw[j].x * x[2 * j].z;
                    s0 += cw[j].y * x[2 * j + 1].x; s1 += cw[j].y * x[2 * j + 1].y; s2 += cw[j].y * x[2 * j + 1].z;
                }
#pragma unroll
                for (int j = 0; j < kHalf; j++) { cw[j] = nw[j]; cc[j] = nc[j]; }
            }
        } else {
        const int2* p = tr + (((int64_t)it.w << 32) | (uint32_t)it.z);
        const int len = it.y & 0xffff;
        int off = lane;
        for (; off < len; off += 32 * kGatherUnroll) {
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++)
                nxt[j] = off + 32 * (kGatherUnroll + j) < len ? load_tr<CS>(&p[off + 32 * (kGatherUnroll + j)]) : zero;
            float4 x[kGatherUnroll];
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) x[j] = load_er<MULTI>(er, cur[j].x);
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) {
                const float wt = __int_as_float(cur[j].y);
                s0 += wt * x[j].x; s1 += wt * x[j].y; s2 += wt * x[j].z;
            }
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) cur[j] = nxt[j];
        }
        }
        // hand over to the next item before finishing this one: its first loads run under the reduction and epilogue
        __syncwarp();
        const int4 hold = hold_s[warp];
        const int row = hold.x, meta = hold.y, w_done = hold.z;
        wn = hold.w;
        const bool has_next = wn != kNoItem;
        cp_async_wait_all();
        __syncwarp();
        it = desc_s[warp];                                          // stale when !has_next: no entries are read from it then
        __syncwarp();                                               // every lane has its copy before the slots are refilled
        if (PACKED) {
            const int64_t k0 = (int64_t)(uint32_t)it.z << 6;
            const float2* wn2 = reinterpret_cast<const float2*>(A.pk_w + k0);
            const uint32_t* cn2 = reinterpret_cast<const uint32_t*>(A.pk_c + k0);
            const int ln = has_next ? ((it.y & 0xffff) >> 1) : 0;
#pragma unroll
            for (int j = 0; j < kHalf; j++) {
                const bool in = lane + 32 * j < ln;
                cw[j] = in ? load_w2<CS>(&wn2[lane + 32 * j]) : make_float2(0.f, 0.f);
                cc[j] = in ? load_c2<CS>(&cn2[lane + 32 * j]) : 0u;
            }
        } else {
            const int2* pn = tr + (((int64_t)it.w << 32) | (uint32_t)it.z);
            const int ln = has_next ? (it.y & 0xffff) : 0;
#pragma unroll
            for (int j = 0; j < kGatherUnroll; j++) cur[j] = lane + 32 * j < ln ? load_tr<CS>(&pn[lane + 32 * j]) : zero;
        }
        w = wn;
        wn = has_next ? claim() : kNoItem;
        if (wn != kNoItem && lane == 0) cp_async16(&desc_s[warp], &items[wn]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const int n_parts = (meta >> 16) & 0xff;
        bool fin = true;
        if (n_parts > 1) fin = combine_parts(A.part_sum, A.row_ctr, row, __ldg(&item_slot[w_done]), (meta >> 24) & 0xff, n_parts, s0, s1, s2);
        if (fin) {
            float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);                    // sky: emit = 0
            if (lane == 0) {
                const float4 r = refl[row0 + row];
                float4 a = nv;
                if (r.w == 0.0f) {                                          // CollectLight, leaf patch
                    float4 t = total[row];
                    t.x += s0; t.y += s1; t.z += s2;
                    total[row] = t;
                    nv = make_float4(s0 * r.x, s1 * r.y, s2 * r.z, 0.f);
                    a = make_float4(s0, s1, s2, 0.f);
                }
                if (A.add) A.add[row] = a;
                if (!MULTI) er_next[row0 + row] = nv;
            }
            // fused exchange: lane p stores the finished row straight into rank p's next-bounce buffer
            // (NVLink peer store; slot `rank` is the local buffer) -- no separate all-gather pass
            if (MULTI) store_row_to_peers(A.peers, A.next_buf, row0 + row, nv.x, nv.y, nv.z);
        }
        has = has_next;
    }
    if (MULTI) {
        __syncthreads();
        if (threadIdx.x == 0) signal_if_last(A.peers, A.flags, A.signal_rel);
    }
}

// The multi-GPU gather from the block-row streams: the same persistent blocks, claims, pool, barrier and launch chain as k4_gather_items,
// with a block of kBlockRows rows as the work item {block of rows, padded entries | 1 << 16, first entry / 32, column base} (a block
// whose union spans several segments keeps the matrix on the packed streams: build_gather_plan) and kBlockRows epilogues per item.
// Lane-to-entry mapping and accumulation order are those of k4_gather_blocked: one GPU and several give the same bits.
constexpr int kBkUnroll = 4;
template <bool MULTI, int WARPS, int R>
__global__ void __launch_bounds__(WARPS * 32, R == 4 ? 4 : 5)
k4_gather_items_blocked(const int32_t* __restrict__ block_ptr, const int4* __restrict__ items, int64_t row0, const float4* er, const float4* __restrict__ refl,
                        float4* er_next, float4* __restrict__ total, GatherAux A) {
    __shared__ int next_item;
    __shared__ int4 desc_s[WARPS], hold_s[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MULTI) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) next_item = __ldg(&block_ptr[blockIdx.x]);
    __syncthreads();
    // the common pool's counter belongs to the bounce that is running: a block of the NEXT bounce (launched behind it, PDL) may take from the
    // pool only after its barrier wait, i.e. after the running bounce's last block has reset the counter
    bool pool_ok = !MULTI || A.wait_rel == 0;
    auto claim = [&]() {
        int c = kNoItem;
        if (lane == 0) {
            c = atomicAdd(&next_item, 1);
            if (c >= __ldg(&block_ptr[blockIdx.x + 1])) {
                c = kNoItem;
                if (MULTI && pool_ok && A.pool_begin < A.n_items) {
                    const int q = A.pool_begin + (int)atomicAdd((unsigned int*)A.flags + kFlagPool, 1u);
                    if (q < A.n_items) c = q;
                }
            }
        }
        return __shfl_sync(0xffffffffu, c, 0);
    };
    int w = claim();
    int4 it = make_int4(0, 0, 0, 0);
    if (w != kNoItem) it = __ldg(&items[w]);
    float cw[kBkUnroll][R], nw[kBkUnroll][R];
    int cc[kBkUnroll], nc[kBkUnroll];
    auto first_loads = [&](const int4& d, int l0) {
        const int64_t k0 = (int64_t)(uint32_t)d.z << 5;
#pragma unroll
        for (int j = 0; j < kBkUnroll; j++) {
#pragma unroll
            for (int r = 0; r < R; r++) cw[j][r] = 0.f;
            cc[j] = 0;
            if (lane + 32 * j < l0) { load_weights<R>(A.bk_w, k0 + lane + 32 * j, cw[j]); cc[j] = (int)__ldcs(&A.bk_c[k0 + lane + 32 * j]); }
        }
    };
    first_loads(it, it.y & 0xffff);
    int wn = w != kNoItem ? claim() : kNoItem;
    if (wn != kNoItem && lane == 0) cp_async16(&desc_s[warp], &items[wn]);
    if (MULTI) {
        if (A.wait_rel) wait_for_peers(A.flags, A.wait_world, A.wait_rel);
        pool_ok = true;
    }
    bool has = w != kNoItem;
    while (has) {
        if (lane == 0) hold_s[warp] = make_int4(it.x, it.y, w, wn);
        float acc[R][3];
#pragma unroll
        for (int r = 0; r < R; r++) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; }
        {
            const int64_t k0 = (int64_t)(uint32_t)it.z << 5;
            const uint16_t* c1 = A.bk_c + k0;
            const int len = it.y & 0xffff, base = it.w;
            for (int off = lane; off < len; off += 32 * kBkUnroll) {
#pragma unroll
                for (int j = 0; j < kBkUnroll; j++) {
#pragma unroll
                    for (int r = 0; r < R; r++) nw[j][r] = 0.f;
                    nc[j] = 0;
                    if (off + 32 * (kBkUnroll + j) < len) { load_weights<R>(A.bk_w, k0 + off + 32 * (kBkUnroll + j), nw[j]); nc[j] = (int)__ldcs(&c1[off + 32 * (kBkUnroll + j)]); }
                }
                float4 x[kBkUnroll];
#pragma unroll
                for (int j = 0; j < kBkUnroll; j++) x[j] = load_er<MULTI>(er, base + cc[j]);
#pragma unroll
                for (int j = 0; j < kBkUnroll; j++)
#pragma unroll
                    for (int r = 0; r < R; r++) { acc[r][0] += cw[j][r] * x[j].x; acc[r][1] += cw[j][r] * x[j].y; acc[r][2] += cw[j][r] * x[j].z; }
#pragma unroll
                for (int j = 0; j < kBkUnroll; j++) {
#pragma unroll
                    for (int r = 0; r < R; r++) cw[j][r] = nw[j][r];
                    cc[j] = nc[j];
                }
            }
        }
        __syncwarp();
        const int4 hold = hold_s[warp];
        const int blk = hold.x;
        wn = hold.w;
        const bool has_next = wn != kNoItem;
        cp_async_wait_all();
        __syncwarp();
        it = desc_s[warp];
        __syncwarp();
        first_loads(it, has_next ? (it.y & 0xffff) : 0);
        w = wn;
        wn = has_next ? claim() : kNoItem;
        if (wn != kNoItem && lane == 0) cp_async16(&desc_s[warp], &items[wn]);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], o);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int row = blk * R + r;
            if (row < A.nloc) {                                             // warp-uniform
                float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);                // sky: emit = 0
                if (lane == 0) {
                    const float4 rf = refl[row0 + row];
                    float4 a = nv;
                    if (rf.w == 0.0f) {                                     // CollectLight, leaf patch
                        float4 t = total[row];
                        t.x += acc[r][0]; t.y += acc[r][1]; t.z += acc[r][2];
                        total[row] = t;
                        nv = make_float4(acc[r][0] * rf.x, acc[r][1] * rf.y, acc[r][2] * rf.z, 0.f);
                        a = make_float4(acc[r][0], acc[r][1], acc[r][2], 0.f);
                    }
                    if (A.add) A.add[row] = a;
                    if (!MULTI) er_next[row0 + row] = nv;
                }
                if (MULTI) store_row_to_peers(A.peers, A.next_buf, row0 + row, nv.x, nv.y, nv.z);
            }
        }
        has = has_next;
    }
    if (MULTI) {
        __syncthreads();
        if (threadIdx.x == 0) signal_if_last(A.peers, A.flags, A.signal_rel);
    }
}

// closes the last bounce of a call: every rank's final stores into this rank's buffers have landed (and this
// rank's final epoch is out) before anything else touches the radiance buffers
__global__ void k4_peer_wait(uint32_t* flags, int world, uint32_t wait_rel) {
    wait_for_peers(flags, world, wait_rel);
    if (threadIdx.x == 0) { flags[kFlagBase] += wait_rel; __threadfence_system(); }
}

// deterministic `added`: fixed 1024-row blocks -> per-block partial (stage 1), then k4_reduce_added (stage 2)
__global__ void __launch_bounds__(256) k4_sum_added_rows(int nloc, const float4* __restrict__ add, float* __restrict__ partials) {
    __shared__ float sm[3][256];
    const int base = blockIdx.x * 1024;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int r = base + j * 256 + threadIdx.x;
        if (r < nloc) { const float4 v = add[r]; a0 += v.x; a1 += v.y; a2 += v.z; }
    }
    sm[0][threadIdx.x] = a0; sm[1][threadIdx.x] = a1; sm[2][threadIdx.x] = a2;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) for (int c = 0; c < 3; c++) sm[c][threadIdx.x] += sm[c][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 3) partials[3 * (size_t)blockIdx.x + threadIdx.x] = sm[threadIdx.x][0];
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
    // 4 entries per lane in flight: entry loads first, then the two gathers of each (emitter origin, emitter radiance), then the
    // arithmetic -- entry by entry the loop was a chain of three dependent latencies (r02 ncu: 1.19 ms per bounce on the C4 matrix,
    // 16 % of DRAM peak)
    constexpr int kBumpUnroll = 4;
    const int64_t k_end = rowptr[row + 1];
    for (int64_t k = rowptr[row] + lane; k < k_end; k += 32 * kBumpUnroll) {
        int2 en[kBumpUnroll]; float4 ojv[kBumpUnroll], vv[kBumpUnroll];
#pragma unroll
        for (int j = 0; j < kBumpUnroll; j++) en[j] = k + 32 * j < k_end ? __ldcs(&tr[k + 32 * j]) : make_int2(0, 0);
#pragma unroll
        for (int j = 0; j < kBumpUnroll; j++) { ojv[j] = __ldg(&origin_area[en[j].x]); vv[j] = __ldg(&er[en[j].x]); }
#pragma unroll
        for (int j = 0; j < kBumpUnroll; j++) {
            const float w = __int_as_float(en[j].y);
            if (w == 0.0f) continue;                                   // row padding / past the row
            const float4 oj = ojv[j], v = vv[j];
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
// MULTI (several GPUs, patch hierarchy): the finished leaf row goes to every rank's next-bounce buffer (sub-lane p -> rank p) and
// the block that finishes last publishes the bounce epoch, as in k4_gather_items; the wait half of the barrier is the prologue of
// k4_collect_parents, which needs every rank's leaf rows before it averages them into the interior patches.
struct ShortPeers { const PeerTable* peers; uint32_t* flags; int next_buf; uint32_t signal_rel; };

template <int kShortLanes, int kMinBlocks, int kShortUnroll = 4, bool kPrefetch = true, bool MULTI = false>
__global__ void __launch_bounds__(kGatherBlock, kMinBlocks)
k4_gather_short(int nrows, const int32_t* __restrict__ rows, int64_t row0, const int64_t* __restrict__ rowptr,
                const int2* __restrict__ tr, const float4* er, const float4* __restrict__ refl,
                float4* er_next, float4* __restrict__ total, float* __restrict__ partials, ShortPeers SP = ShortPeers{}) {
    static_assert(!MULTI || kShortLanes == 8, "the fused exchange maps the 8 lanes of a row onto the (at most 8) ranks");
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
            for (int j = 0; j < kShortUnroll; j++) x[j] = load_er<MULTI>(er, cur[j].x);
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
            for (int j = 0; j < kShortUnroll; j++) x[j] = load_er<MULTI>(er, cur[j].x);
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
    float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);                        // sky: emit = 0
    if (valid && sub == 0) {
        const float4 rf = refl[row0 + row];
        if (rf.w == 0.0f) {                                              // CollectLight, leaf patch
            float4 t = total[row];
            t.x += s0; t.y += s1; t.z += s2;
            total[row] = t;
            nv = make_float4(s0 * rf.x, s1 * rf.y, s2 * rf.z, 0.f);
            e0 = s0; e1 = s1; e2 = s2;
        }
        if (!MULTI) er_next[row0 + row] = nv;
    }
    if (MULTI) {        // sub-lane p of the row's 8 lanes stores the row into rank p's buffer
        const int src = (threadIdx.x & 31) & ~(kShortLanes - 1);
        nv.x = __shfl_sync(0xffffffffu, nv.x, src); nv.y = __shfl_sync(0xffffffffu, nv.y, src); nv.z = __shfl_sync(0xffffffffu, nv.z, src);
        if (valid && sub < SP.peers->world) SP.peers->er[SP.next_buf][sub][row0 + row] = make_float4(nv.x, nv.y, nv.z, 0.f);
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
    if (MULTI) {
        __syncthreads();
        if (threadIdx.x == 0) signal_if_last(SP.peers, SP.flags, SP.signal_rel);
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
                   const int2* __restrict__ ent, float4* buf, const uint32_t* flags = nullptr, int wait_world = 0, uint32_t wait_rel = 0) {
    // several GPUs with the fused exchange: the leaf rows of this bounce are complete once every rank's epoch has arrived
    if (wait_rel) wait_for_peers(flags, wait_world, wait_rel);
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

// Work items of the gather for the resident rows (rowlen = logical row lengths, host copy).  Called by
// vrad_build_transfers and vrad_transfers_upload once the rows are in place.
// the packed streams from tr[] (after vrad_build_transfers / vrad_transfers_upload); device passes, two scans
static int build_packed_streams(vrad_env* e, int64_t nloc) {
    TransfersDev& T = e->transfers;
    T.packed = false;
    if (!e->opt.k4_pack || e->patches.hier || nloc <= 0 || !T.rows_ascending) return 0;
    int max_seg = 1 << 8;                                  // the longest segment = the longest part of the pair plan (k4_seg)
    while ((max_seg << 1) <= e->opt.k4_seg && max_seg < (1 << 15)) max_seg <<= 1;
    DevBuf<int32_t> d_ns; DevBuf<int64_t> d_pad, d_rowptr; DevBuf<unsigned char> d_tmp;
    auto drop = [&]() { d_ns.release(); d_pad.release(); d_rowptr.release(); d_tmp.release(); };
    if (d_ns.alloc(nloc + 1) || d_pad.alloc(nloc + 1) || d_rowptr.alloc(nloc + 1) || T.pk_seg_ptr.alloc(nloc + 1)) { drop(); set_error("out of device memory (packed transfer streams)"); return VRAD_E_NOMEM; }
#define PK_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { drop(); set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); return VRAD_E_CUDA; } } while (0)
    PK_CHECK(cudaMemsetAsync(d_ns.p, 0, ((size_t)nloc + 1) * 4, e->stream));
    PK_CHECK(cudaMemsetAsync(d_pad.p, 0, ((size_t)nloc + 1) * 8, e->stream));
    const int wblocks = (int)((nloc * 32 + 255) / 256);
    k4_pack_count<<<wblocks, 256, 0, e->stream>>>((int)nloc, T.rowptr.p, T.rowlen.p, T.tr.p, max_seg, d_ns.p, d_pad.p);
    size_t b1 = 0, b2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b1, d_ns.p, T.pk_seg_ptr.p, (int)nloc + 1, e->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, d_pad.p, d_rowptr.p, (int)nloc + 1, e->stream);
    if (d_tmp.alloc(std::max(b1, b2) + 16)) { drop(); set_error("out of device memory (scan scratch)"); return VRAD_E_NOMEM; }
    PK_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, b1, d_ns.p, T.pk_seg_ptr.p, (int)nloc + 1, e->stream));
    PK_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, b2, d_pad.p, d_rowptr.p, (int)nloc + 1, e->stream));
    int32_t n_segs = 0; int64_t n_entries = 0;
    PK_CHECK(cudaMemcpyAsync(&n_segs, T.pk_seg_ptr.p + nloc, 4, cudaMemcpyDeviceToHost, e->stream));
    PK_CHECK(cudaMemcpyAsync(&n_entries, d_rowptr.p + nloc, 8, cudaMemcpyDeviceToHost, e->stream));
    PK_CHECK(cudaStreamSynchronize(e->stream));
    // Rows that span several column windows become several segments; the multi-GPU kernel runs them as parts of a split row (cross-warp
    // combine: an atomic, a fence and a second pass per row), which costs more than the narrower stream saves once most rows are
    // like that (C5 map, 3.0 segments per row: 1224 us per bounce on the 8-rank slice against 926 us from the pairs).  Such a matrix
    // keeps the {col,w} pairs.  k4_pack = 9 packs regardless (measurements).
    if (e->opt.k4_pack != 9 && (int64_t)n_segs > nloc + nloc / 20) {
        if (getenv("VRAD_VERBOSE")) fprintf(stderr, "[vrad] rank %d: %d segments for %lld rows -- transfers stay {col,w} pairs\n", e->cfg.rank, n_segs, (long long)nloc);
        drop();
        return 0;
    }
    if (T.pk_segs.alloc((size_t)n_segs + 1) || T.pk_w.alloc((size_t)n_entries + 8) || T.pk_c.alloc((size_t)n_entries + 8)) {
        // not enough memory for a second copy of the matrix: gather from the pairs
        T.pk_segs.release(); T.pk_w.release(); T.pk_c.release(); drop();
        return 0;
    }
    k4_pack_fill<<<wblocks, 256, 0, e->stream>>>((int)nloc, T.rowptr.p, T.rowlen.p, T.tr.p, max_seg, T.pk_seg_ptr.p, d_rowptr.p, T.pk_segs.p, T.pk_w.p, T.pk_c.p);
    PK_CHECK(cudaGetLastError());
    PK_CHECK(cudaStreamSynchronize(e->stream));
#undef PK_CHECK
    drop();
    T.pk_entries = n_entries; T.pk_n_segs = n_segs; T.packed = true;
    if (getenv("VRAD_VERBOSE"))
        fprintf(stderr, "[vrad] rank %d packed transfer streams: %d segments for %lld rows, %lld entries for %lld transfers (%.1f %% padding), 6 B each\n", e->cfg.rank, n_segs,
                (long long)nloc, (long long)n_entries, (long long)T.nnz, T.nnz > 0 ? 100.0 * (double)(n_entries - T.nnz) / (double)T.nnz : 0.0);
    return 0;
}

// the block-row streams from tr[] (k4_pack = 2); device passes, two scans
static int build_block_streams(vrad_env* e, int64_t nloc) {
    TransfersDev& T = e->transfers;
    T.blocked = false;
    if (!T.rows_ascending) return 0;
    const int rows = e->opt.k4_bk_rows == 2 ? 2 : 4;
    if ((e->opt.k4_pack != 2 && e->opt.k4_pack != 3) || e->patches.hier || nloc <= 0 || (T.row0 % kBlockRows) != 0) return 0;
    int max_seg = 1 << 8;
    while ((max_seg << 1) <= e->opt.k4_seg && max_seg < (1 << 15)) max_seg <<= 1;
    const int nb = (int)((nloc + rows - 1) / rows);
    // With several ranks a block of rows is ONE work item of the persistent kernel (4 blocks x 8 warps per SM).  Below ~2 items per resident
    // warp the warps' ranges no longer balance -- C4 slices on one GPU, us per bounce, block rows / packed streams: 2 ranks 88 / 127,
    // 4 ranks (2.5 items per warp) 56 / 69, 8 ranks (1.2 per warp) 50 / 40 -- and the slice stays on the packed streams, whose items are
    // single rows.  (2 rows per block instead of 4 doubles the items for the same bytes but also the gathers: 43 at 8 ranks, 100 at 2; a
    // thread block per item with the sums combined in shared memory: 43 at 8 ranks, 153 at 2.  profiles/r02_k4_block_rows_sim.json)
    // k4_pack = 3 keeps the block rows whatever the count.
    const bool work_items = e->cfg.world > 1 || e->opt.k4_items;
    if (work_items && e->opt.k4_pack != 3 && (int64_t)nb < (int64_t)e->opt.k4_bk_min_items * e->sm_count * 8 / 10) return 0;
    DevBuf<int32_t> d_ns; DevBuf<int64_t> d_pad, d_start; DevBuf<unsigned char> d_tmp;
    auto drop = [&]() { d_ns.release(); d_pad.release(); d_start.release(); d_tmp.release(); };
    if (d_ns.alloc(nb + 1) || d_pad.alloc(nb + 1) || d_start.alloc(nb + 1) || T.bk_seg_ptr.alloc(nb + 1)) { drop(); set_error("out of device memory (block-row transfer streams)"); return VRAD_E_NOMEM; }
#define BK_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { drop(); set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); return VRAD_E_CUDA; } } while (0)
    BK_CHECK(cudaMemsetAsync(d_ns.p, 0, ((size_t)nb + 1) * 4, e->stream));
    BK_CHECK(cudaMemsetAsync(d_pad.p, 0, ((size_t)nb + 1) * 8, e->stream));
    const int grid = (nb + kBkBuildWarps - 1) / kBkBuildWarps;
    k4_block_count<<<grid, kBkBuildWarps * 32, 0, e->stream>>>(nb, rows, (int)nloc, T.rowptr.p, T.rowlen.p, T.tr.p, max_seg, d_ns.p, d_pad.p);
    size_t b1 = 0, b2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b1, d_ns.p, T.bk_seg_ptr.p, nb + 1, e->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, d_pad.p, d_start.p, nb + 1, e->stream);
    if (d_tmp.alloc(std::max(b1, b2) + 16)) { drop(); set_error("out of device memory (scan scratch)"); return VRAD_E_NOMEM; }
    BK_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, b1, d_ns.p, T.bk_seg_ptr.p, nb + 1, e->stream));
    BK_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, b2, d_pad.p, d_start.p, nb + 1, e->stream));
    int32_t n_segs = 0; int64_t n_entries = 0;
    BK_CHECK(cudaMemcpyAsync(&n_segs, T.bk_seg_ptr.p + nb, 4, cudaMemcpyDeviceToHost, e->stream));
    BK_CHECK(cudaMemcpyAsync(&n_entries, d_start.p + nb, 8, cudaMemcpyDeviceToHost, e->stream));
    BK_CHECK(cudaStreamSynchronize(e->stream));
    if (getenv("VRAD_VERBOSE"))
        fprintf(stderr, "[vrad] rank %d block-row streams: %d blocks of %d rows, %d segments, %lld union entries for %lld transfers (%.2f x one row per block, %.2f B per transfer)\n",
                e->cfg.rank, nb, rows, n_segs, (long long)n_entries, (long long)T.nnz, T.nnz > 0 ? (double)n_entries * rows / (double)T.nnz : 0.0,
                T.nnz > 0 ? (double)n_entries * (2.0 + 4.0 * rows) / (double)T.nnz : 0.0);
    // worth it only where neighbouring rows share their columns (18 B per union entry against 4 rows x 6 B per packed entry), and the work-item
    // kernel takes a block of rows as ONE item: every block must be a single segment (its union inside one 65,536-column window, at most k4_seg entries)
    bool single = (int64_t)n_segs <= nb;
    if (single) {
        std::vector<int32_t> hns((size_t)nb);
        BK_CHECK(cudaMemcpy(hns.data(), d_ns.p, (size_t)nb * 4, cudaMemcpyDeviceToHost));
        for (int b = 0; b < nb && single; b++) single = hns[b] <= 1;
    }
    if (!single || (double)n_entries * (2.0 + 4.0 * rows) > (double)T.nnz * 6.0 * 1.1) { drop(); return 0; }
    if (T.bk_segs.alloc((size_t)n_segs + 1) || T.bk_w.alloc(((size_t)n_entries + 32) * rows) || T.bk_c.alloc((size_t)n_entries + 32)) {
        T.bk_segs.release(); T.bk_w.release(); T.bk_c.release(); drop();
        return 0;
    }
    BK_CHECK(cudaMemsetAsync(T.bk_w.p, 0, ((size_t)n_entries + 32) * rows * sizeof(float), e->stream));
    BK_CHECK(cudaMemsetAsync(T.bk_c.p, 0, ((size_t)n_entries + 32) * sizeof(uint16_t), e->stream));
    k4_block_fill<<<grid, kBkBuildWarps * 32, 0, e->stream>>>(nb, rows, (int)nloc, T.rowptr.p, T.rowlen.p, T.tr.p, max_seg, T.bk_seg_ptr.p, d_start.p, T.bk_segs.p, T.bk_w.p, T.bk_c.p);
    BK_CHECK(cudaGetLastError());
    BK_CHECK(cudaStreamSynchronize(e->stream));
#undef BK_CHECK
    drop();
    T.bk_entries = n_entries; T.bk_n_segs = n_segs; T.bk_n_blocks = nb; T.bk_rows = rows; T.blocked = true;
    return 0;
}

// blocks of the single-GPU gather grid (one warp per row, or per block of rows)
static inline int plain_gather_blocks(const TransfersDev& T, int nloc) {
    if (T.blocked) return std::max(1, (T.bk_n_blocks + kGatherWarps - 1) / kGatherWarps);
    return std::max(1, (nloc + kGatherWarps - 1) / kGatherWarps);
}

int build_gather_plan(vrad_env* e, const int32_t* rowlen, int64_t nloc) {
    TransfersDev& T = e->transfers;
    { const int rcb = build_block_streams(e, nloc); if (rcb) return rcb; }
    T.packed = false;
    if (!T.blocked) { const int rcp = build_packed_streams(e, nloc); if (rcp) return rcp; }
    const bool long_first = e->opt.k4_long_first != 0;
    int max_len = 0;
    for (int64_t r = 0; r < nloc; r++) max_len = std::max(max_len, rowlen[r]);
    int shift = 8;
    while ((1 << (shift + 1)) <= e->opt.k4_seg && shift < 15) shift++;
    while (shift < 15 && ((int64_t)max_len + (1 << shift) - 1) >> shift > 255) shift++;      // at most 255 parts per row
    const int seg = 1 << shift;
    if (((int64_t)max_len + seg - 1) / seg > 255) { set_error("a transfer row of %d entries needs more than 255 parts of %d", max_len, seg); return VRAD_E_UNSUPPORTED; }
    // item = {local row, entries | n_parts << 16 | part index << 24, first entry (padded CSR position) lo, hi}; slot = first part slot of its row
    std::vector<int4> items;
    std::vector<int32_t> slots;
    items.reserve((size_t)nloc + 1024); slots.reserve((size_t)nloc + 1024);
    int n_slots = 0;
    int64_t pos = 0;                                       // rows start on 4-entry boundaries (vrad_transfers_upload / k2_fill)
    T.plan_packed = false; T.plan_blocked = false;
    if (T.blocked) {
        // block-row plan: one item per block of rows {block, padded entries | 1 << 16, first entry / 32, column base}; every block must be one segment
        std::vector<int4> segs((size_t)T.bk_n_segs + 1);
        std::vector<int32_t> sp((size_t)T.bk_n_blocks + 1);
        VRAD_CUDA_CHECK(cudaMemcpy(segs.data(), T.bk_segs.p, (size_t)T.bk_n_segs * sizeof(int4), cudaMemcpyDeviceToHost));
        VRAD_CUDA_CHECK(cudaMemcpy(sp.data(), T.bk_seg_ptr.p, ((size_t)T.bk_n_blocks + 1) * 4, cudaMemcpyDeviceToHost));
        for (int b = 0; b < T.bk_n_blocks; b++) {
            if (sp[b + 1] == sp[b]) { items.push_back(make_int4(b, 1 << 16, 0, 0)); slots.push_back(-1); continue; }      // empty rows: epilogues only
            const int4 sg = segs[sp[b]];
            const int64_t start = (int64_t)(uint32_t)sg.x | ((int64_t)sg.y << 32);
            items.push_back(make_int4(b, sg.z | (1 << 16), (int)(uint32_t)(start >> 5), sg.w));
            slots.push_back(-1);
        }
        T.plan_blocked = true;
    }
    if (T.packed && !T.plan_blocked) {
        // packed plan: one item per segment {row, padded entries | n_parts << 16 | part index << 24, first entry / 64, column base}
        std::vector<int4> segs((size_t)T.pk_n_segs + 1);
        std::vector<int32_t> sp((size_t)nloc + 1);
        VRAD_CUDA_CHECK(cudaMemcpy(segs.data(), T.pk_segs.p, (size_t)T.pk_n_segs * sizeof(int4), cudaMemcpyDeviceToHost));
        VRAD_CUDA_CHECK(cudaMemcpy(sp.data(), T.pk_seg_ptr.p, ((size_t)nloc + 1) * 4, cudaMemcpyDeviceToHost));
        bool fits = true;
        for (int64_t r = 0; r < nloc && fits; r++) fits = sp[r + 1] - sp[r] <= 255;
        if (fits) {
            for (int64_t r = 0; r < nloc; r++) {
                const int n_parts = std::max(1, sp[r + 1] - sp[r]);
                if (sp[r + 1] == sp[r]) { items.push_back(make_int4((int)r, 1 << 16, 0, 0)); slots.push_back(-1); continue; }     // empty row: epilogue only
                for (int q = 0; q < n_parts; q++) {
                    const int4 sg = segs[sp[r] + q];
                    const int64_t start = (int64_t)(uint32_t)sg.x | ((int64_t)sg.y << 32);
                    items.push_back(make_int4((int)r, sg.z | (n_parts << 16) | (q << 24), (int)(uint32_t)(start >> 6), sg.w));
                    slots.push_back(n_parts > 1 ? n_slots : -1);
                }
                if (n_parts > 1) n_slots += n_parts;
            }
            T.plan_packed = true;
        }
    }
    for (int64_t r = 0; r < nloc && !T.plan_packed && !T.plan_blocked; r++) {
        const int len = rowlen[r];
        const int n_parts = len > seg ? (len + seg - 1) / seg : 1;
        for (int p = 0; p < n_parts; p++) {
            const int off = p * seg, l = n_parts == 1 ? len : std::min(seg, len - off);
            const int64_t start = pos + off;
            items.push_back(make_int4((int)r, l | (n_parts << 16) | (p << 24), (int)(uint32_t)(start & 0xffffffff), (int)(start >> 32)));
            slots.push_back(n_parts > 1 ? n_slots : -1);
        }
        if (n_parts > 1) n_slots += n_parts;
        pos += ((int64_t)len + 3) & ~(int64_t)3;
    }
    auto sort_desc = [&](int a, int b) {                   // stable, by entries, items and their slots together
        std::vector<int> ord(b - a);
        for (int i = 0; i < b - a; i++) ord[i] = a + i;
        std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return (items[x].y & 0xffff) > (items[y].y & 0xffff); });
        std::vector<int4> ti(b - a); std::vector<int32_t> ts(b - a);
        for (int i = 0; i < b - a; i++) { ti[i] = items[ord[i]]; ts[i] = slots[ord[i]]; }
        std::copy(ti.begin(), ti.end(), items.begin() + a); std::copy(ts.begin(), ts.end(), slots.begin() + a);
    };
    if (long_first) sort_desc(0, (int)items.size());
    // blocks of the gather grid
    std::vector<int32_t> bp;
    const int n_it = (int)items.size();
    int pool_begin = n_it;
    // 0 = automatic: 8 warps x 5 blocks per SM at 48 registers for the packed streams, 6 x 6 at 56 registers for the pairs (measured at 2, 4 and 8 GPUs)
    const int plan_warps = T.plan_blocked ? 8 : e->opt.k4_block ? (e->opt.k4_block == 192 ? 6 : 8) : (T.plan_packed ? 8 : 6);
    const int n_resident = e->sm_count * (T.plan_blocked ? (T.bk_rows == 4 ? 4 : 5) : plan_warps == 6 ? 6 : 5);      // resident blocks of k4_gather_items<.., WARPS> / k4_gather_items_blocked      // resident blocks of k4_gather_items<.., WARPS> / k4_gather_items_blocked
    if (e->opt.k4_persist && n_it > n_resident * plan_warps) {
        // persistent plan: `slots` contiguous ranges of equal work (entries + a per-item constant), longest item first inside each
        constexpr int64_t kItemCost = 96;                  // a row's fixed work (index loads, reduction, epilogue) in entry equivalents
        int64_t all = 0;
        for (const int4& it : items) all += (it.y & 0xffff) + kItemCost;
        // the tail of the list is the common pool: items no range owns
        const int pool_pct = std::min(std::max(e->opt.k4_pool, 0), 50);
        int64_t total = 0;
        for (pool_begin = 0; pool_begin < n_it && total * 100 < all * (100 - pool_pct); pool_begin++) total += (items[pool_begin].y & 0xffff) + kItemCost;
        bp.assign(1, 0);
        int64_t acc = 0;
        for (int i = 0; i < pool_begin; i++) {
            acc += (items[i].y & 0xffff) + kItemCost;
            while ((int)bp.size() < n_resident && acc * n_resident >= total * (int64_t)bp.size()) bp.push_back(i + 1);
        }
        while ((int)bp.size() < n_resident) bp.push_back(pool_begin);
        bp.push_back(pool_begin);
        for (int b = 0; b < n_resident; b++) sort_desc(bp[b], bp[b + 1]);
        sort_desc(pool_begin, n_it);
    } else {
        for (int i = 0; i < n_it; i += plan_warps) bp.push_back(i);
        if (bp.empty()) bp.push_back(0);
        bp.push_back(n_it);
    }
    if (T.items.alloc(items.size() + 1) || T.item_slot.alloc(slots.size() + 1) || T.block_ptr.alloc(bp.size()) || T.part_sum.alloc((size_t)n_slots + 1) || T.row_ctr.alloc((size_t)nloc + 1)) {
        set_error("out of device memory for the gather plan"); return VRAD_E_NOMEM;
    }
    if (!items.empty()) VRAD_CUDA_CHECK(cudaMemcpyAsync(T.items.p, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice, e->stream));
    if (!slots.empty()) VRAD_CUDA_CHECK(cudaMemcpyAsync(T.item_slot.p, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(T.block_ptr.p, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemsetAsync(T.row_ctr.p, 0, ((size_t)nloc + 1) * 4, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    T.n_blocks = (int)bp.size() - 1;
    T.n_items = (int)items.size(); T.n_slots = n_slots; T.seg_shift = shift; T.plan_warps = plan_warps; T.pool_begin = pool_begin; T.plan_serial++;
    return 0;
}

} // namespace vrad
using namespace vrad;

extern "C" {

int vrad_patches_upload(vrad_env* e, int n, const float* origin3, const float* normal3, const float* plane_dist,
                        const float* area, const float* reflectivity3, const int32_t* cluster, const uint8_t* flags) {
    VRAD_MULTI(e, group_patches_upload(e, n, origin3, normal3, plane_dist, area, reflectivity3, cluster, flags));
    if (!e || n <= 0 || !origin3 || !normal3 || !plane_dist || !area || !reflectivity3) { set_error("vrad_patches_upload: bad arguments"); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    PatchesDev& P = e->patches;
    if (P.origin_area.alloc(n) || P.normal_dist.alloc(n) || P.refl.alloc(n) || P.cluster.alloc(n)) { set_error("out of device memory for patches"); return VRAD_E_NOMEM; }
    std::vector<float4> oa(n), nd(n), rf(n);
    P.h_cluster.assign(n, 0); P.h_flags.assign(n, 0);
    P.h_area.assign(area, area + n); P.h_refl.assign(reflectivity3, reflectivity3 + 3 * (size_t)n);
    P.hier = false; P.n_interior = 0; P.h_root_cluster.clear();
    P.bump = false; P.h_needs_bump.clear();
    P.has_windings = false;
    P.h_normal.assign(normal3, normal3 + 3 * (size_t)n);
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

int vrad_patches_set_windings(vrad_env* e, int n, const int32_t* first, const int32_t* count, int n_points, const float* points3) {
    VRAD_MULTI(e, group_set_windings(e, n, first, count, n_points, points3));
    if (!e || n < 0 || n_points < 0) { set_error("vrad_patches_set_windings: bad arguments"); return VRAD_E_INVALID; }
    PatchesDev& P = e->patches;
    if (n == 0) {                                             // back to the differential form factor everywhere
        P.has_windings = false; P.wind.release(); P.wind_pts.release();
        e->transfers.ready = false;
        return VRAD_OK;
    }
    if (!first || !count || (n_points > 0 && !points3)) { set_error("vrad_patches_set_windings: bad arguments"); return VRAD_E_INVALID; }
    if (P.n == 0 || n != P.n) { set_error("vrad_patches_set_windings: %d windings for %d uploaded patches", n, P.n); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    std::vector<int2> w(n);
    std::vector<float4> pts((size_t)n_points + 1);
    for (int i = 0; i < n_points; i++) pts[i] = make_float4(points3[3 * (size_t)i], points3[3 * (size_t)i + 1], points3[3 * (size_t)i + 2], 0.f);
    std::vector<uint8_t> owned((size_t)n_points + 1, 0);
    for (int i = 0; i < n; i++) {
        const int f = first[i], c = count[i];
        if (c < 0 || f < 0 || (int64_t)f + c > n_points) { set_error("vrad_patches_set_windings: winding of patch %d ([%d, %d + %d)) is outside the %d points", i, f, f, c, n_points); return VRAD_E_INVALID; }
        w[i] = make_int2(f, c < 3 ? 0 : c);
        if (c < 3) continue;
        // the contour integral assumes the points run clockwise seen from the patch's front (the BSP face convention,
        // kept by ClipWindingEpsilon); a winding the other way round is reversed here -- unless it shares its points with another patch
        double nx = 0, ny = 0, nz = 0;                       // Newell normal: along the patch normal for a counter-clockwise winding
        for (int k = 0; k < c; k++) {
            const float4 a = pts[f + k], b = pts[f + (k + 1 < c ? k + 1 : 0)];
            nx += ((double)a.y - b.y) * ((double)a.z + b.z); ny += ((double)a.z - b.z) * ((double)a.x + b.x); nz += ((double)a.x - b.x) * ((double)a.y + b.y);
        }
        const double o = nx * P.h_normal[3 * (size_t)i] + ny * P.h_normal[3 * (size_t)i + 1] + nz * P.h_normal[3 * (size_t)i + 2];
        bool shared = false;
        for (int k = 0; k < c; k++) shared |= owned[f + k] != 0;
        if (o > 0.0) {
            if (shared) { set_error("vrad_patches_set_windings: patch %d is wound counter-clockwise and shares its points with another patch", i); return VRAD_E_INVALID; }
            std::reverse(pts.begin() + f, pts.begin() + f + c);
        }
        for (int k = 0; k < c; k++) owned[f + k] = 1;
    }
    if (P.wind.alloc(n) || P.wind_pts.alloc((size_t)n_points + 1)) { set_error("out of device memory for patch windings"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpy(P.wind.p, w.data(), (size_t)n * sizeof(int2), cudaMemcpyHostToDevice));
    VRAD_CUDA_CHECK(cudaMemcpy(P.wind_pts.p, pts.data(), pts.size() * sizeof(float4), cudaMemcpyHostToDevice));
    P.has_windings = true;
    e->transfers.ready = false;
    return VRAD_OK;
}

int vrad_patches_set_hierarchy(vrad_env* e, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face) {
    VRAD_MULTI(e, group_set_hierarchy(e, n, parent, child1, child2, face));
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
    VRAD_MULTI(e, group_set_bump(e, n, needs_bump, bump_normals9));
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
    VRAD_MULTI_RANK0(e);          // every rank holds the gathered totals
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
    VRAD_MULTI_UNSUPPORTED(e, "vrad_transfers_upload");
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
    bool ascending = true;
    std::vector<int2> ptr(np ? np : 4, make_int2(0, 0));              // {col, w bits}; padding = {0, 0.0f}
    for (int64_t i = 0; i < nloc; i++) {
        const int64_t s = rowptr[i] - rowptr[0];
        for (int64_t k = 0; k < rlen[i]; k++) {
            int32_t c = col[s + k];
            if (c < 0 || c >= N) { set_error("vrad_transfers_upload: column %d out of range at row %lld", c, (long long)(row0 + i)); return VRAD_E_INVALID; }
            if (k > 0 && c <= col[s + k - 1]) ascending = false;
            int2 v; v.x = c; memcpy(&v.y, &w[s + k], 4);
            ptr[prow[i] + k] = v;
        }
    }
    TransfersDev& T = e->transfers;
    // the packed and block-row streams rely on what vrad_build_transfers guarantees -- columns strictly ascending inside a row (MakeScales walks
    // the row in patch order); rows uploaded in another order, or with a column twice, are gathered from the {col,w} pairs as given
    T.rows_ascending = ascending;
    if (T.rowptr.alloc(nloc + 1) || T.rowlen.alloc(nloc ? nloc : 1) || T.tr.alloc(ptr.size())) { set_error("out of device memory for transfers"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpyAsync(T.rowptr.p, prow.data(), (nloc + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(T.rowlen.p, rlen.data(), rlen.size() * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(T.tr.p, ptr.data(), ptr.size() * 8, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    T.row0 = row0; T.row1 = row1; T.nnz = nnz; T.nnz_padded = np; T.rows_serial++;
    int rcp = build_gather_plan(e, rlen.data(), nloc);
    if (rcp) return rcp;
    T.ready = true;
    return VRAD_OK;
}

int vrad_transfers_info(vrad_env* e, int64_t* row0, int64_t* row1, int64_t* nnz) {
    VRAD_MULTI(e, group_transfers_info(e, row0, row1, nnz));
    if (!e) return VRAD_E_INVALID;
    if (!e->transfers.ready) { set_error("vrad_transfers_info: no transfers resident"); return VRAD_E_STATE; }
    if (row0) *row0 = e->transfers.row0;
    if (row1) *row1 = e->transfers.row1;
    if (nnz) *nnz = e->transfers.nnz;
    return VRAD_OK;
}

int vrad_transfers_layout(vrad_env* e, int64_t* pair_entries, int64_t* packed_entries, int64_t* packed_segments, int64_t* block_entries, int64_t* block_rows) {
    if (!e) return VRAD_E_INVALID;
    int64_t a = 0, b = 0, c = 0, d = 0, br = 0;
    if (e->multi) {
        for (vrad_env* r : e->multi->ranks) {
            int64_t x, y, z, u, v;
            const int rc = vrad_transfers_layout(r, &x, &y, &z, &u, &v);
            if (rc) return rc;
            a += x; b += y; c += z; d += u; br = v;
        }
    } else {
        if (!e->transfers.ready) { set_error("vrad_transfers_layout: no transfers resident"); return VRAD_E_STATE; }
        a = e->transfers.nnz_padded;
        if (e->transfers.packed) { b = e->transfers.pk_entries; c = e->transfers.pk_n_segs; }
        if (e->transfers.blocked) { d = e->transfers.bk_entries; br = e->transfers.bk_rows; }
    }
    if (pair_entries) *pair_entries = a;
    if (packed_entries) *packed_entries = b;
    if (packed_segments) *packed_segments = c;
    if (block_entries) *block_entries = d;
    if (block_rows) *block_rows = br;
    return VRAD_OK;
}

int vrad_transfers_download(vrad_env* e, int64_t* rowptr, int32_t* col, float* w) {
    VRAD_MULTI(e, group_transfers_download(e, rowptr, col, w));
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
    VRAD_MULTI_UNSUPPORTED(e, "vrad_transfers_download_rows");
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

// One bounce of the flat (non-hierarchical) gather: a single launch.  p2p: the work-item form with the fused exchange and
// the in-kernel barrier, chained to the previous bounce by PDL when `chained`; otherwise the plain warp-per-row kernel.
static cudaError_t launch_gather(vrad_env* e, bool p2p, bool chained, int cur, bool want_add, uint32_t wait_rel, uint32_t signal_rel, float4* total_local) {
    TransfersDev& T = e->transfers;
    if (!p2p) {
        const int nloc = (int)(T.row1 - T.row0);
        const int nblocks = plain_gather_blocks(T, nloc);
        if (T.blocked) {
            const BlockedRows bk{T.bk_seg_ptr.p, T.bk_segs.p, T.bk_w.p, T.bk_c.p};
            // 4 rows per block: 4 entries in flight per lane, 4 blocks per SM (64 registers); measured alternatives on the C4 matrix (us per bounce, this form 159):
            // 1 entry 256, 2 entries 180 (193 at 5 blocks, 182 at 3), 3 entries 170, 4 entries at 3 blocks per SM 175
            if (T.bk_rows == 4)
                k4_gather_blocked<4, 4, 4><<<nblocks, kGatherBlock, 0, e->stream>>>(T.bk_n_blocks, nloc, T.row0, bk, e->d_er[cur].p, e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, e->d_partials.p);
            else
                k4_gather_blocked<2, 4, 5><<<nblocks, kGatherBlock, 0, e->stream>>>(T.bk_n_blocks, nloc, T.row0, bk, e->d_er[cur].p, e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, e->d_partials.p);
        }
        else if (T.packed) {
            const PackedRows pk{T.pk_seg_ptr.p, T.pk_segs.p, T.pk_w.p, T.pk_c.p};
            // 4 pairs in flight per lane at 48 registers / 5 blocks per SM; measured alternatives on the C4 matrix (us per bounce, this form 250):
            // 6 pairs 277 (spills), 6 pairs / 4 blocks 282, 8 pairs / 4 blocks 417, 3 pairs / 6 blocks 258, 2 pairs / 6 blocks 286, 4 pairs / 4 blocks 303, 4 pairs / 6 blocks 323
            k4_gather_packed<4, 5><<<nblocks, kGatherBlock, 0, e->stream>>>(nloc, T.row0, pk, e->d_er[cur].p, e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, e->d_partials.p);
        }
        else
            k4_gather<<<nblocks, kGatherBlock, 0, e->stream>>>(nloc, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p, e->patches.refl.p,
                                                             e->d_er[cur ^ 1].p, total_local, e->d_partials.p);
        return cudaGetLastError();
    }
    const int nblocks = T.n_blocks;
    GatherAux A{};
    A.add = want_add ? e->d_add.p : nullptr;
    A.part_sum = T.part_sum.p; A.row_ctr = T.row_ctr.p;
    A.pool_begin = T.pool_begin; A.n_items = T.n_items;
    const bool w6 = T.plan_warps == 6;
    A.peers = e->peers.d_table.p; A.flags = e->peers.d_flags.p; A.next_buf = cur ^ 1;
    A.wait_rel = e->opt.k4_sim_peers == 2 ? 0u : wait_rel;      // sim 2: no barrier wait (timing diagnostic)
    A.signal_rel = signal_rel; A.wait_world = e->peers.table_world;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(w6 ? 192 : 256); cfg.dynamicSmemBytes = 0; cfg.stream = e->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (chained) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    if (e->l2_window_bytes > 0) {
        // Keep the head of this rank's transfer stream in L2 from bounce to bounce (the per-rank stream of the C4 matrix at 8 ranks
        // is 192 MB against 126 MB of L2): lines of the window are marked persisting with probability hitRatio (= set-aside / window,
        // so that the set-aside is not thrashed), everything else streams.
        attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[na].val.accessPolicyWindow.base_ptr = (void*)T.tr.p;
        attr[na].val.accessPolicyWindow.num_bytes = e->l2_window_bytes;
        attr[na].val.accessPolicyWindow.hitRatio = e->l2_hit_ratio;
        attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na++;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    const bool cs = e->l2_window_bytes == 0 || e->opt.k4_l2_mb < 0;
    A.pk_w = T.pk_w.p; A.pk_c = T.pk_c.p;
    auto kern = w6 ? (cs ? k4_gather_items<true, 6, true> : k4_gather_items<true, 6, false>) : (cs ? k4_gather_items<true, 8, true> : k4_gather_items<true, 8, false>);
    if (T.plan_packed) kern = w6 ? k4_gather_items<true, 6, true, true> : k4_gather_items<true, 8, true, true>;
    A.bk_w = T.bk_w.p; A.bk_c = T.bk_c.p; A.nloc = (int)(T.row1 - T.row0);
    if (T.plan_blocked)
        return cudaLaunchKernelEx(&cfg, T.bk_rows == 4 ? k4_gather_items_blocked<true, 8, 4> : k4_gather_items_blocked<true, 8, 2>, (const int32_t*)T.block_ptr.p,
                                  (const int4*)T.items.p, T.row0, (const float4*)e->d_er[cur].p, (const float4*)e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, A);
    return cudaLaunchKernelEx(&cfg, kern, (const int32_t*)T.block_ptr.p, (const int4*)T.items.p, (const int32_t*)T.item_slot.p, T.row0,
                              (const int2*)T.tr.p, (const float4*)e->d_er[cur].p, (const float4*)e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, A);
}

// VRAD_K4_SIM_PEERS=1 with world > 1 and no communicator: every "peer" is this device -- rows for the other ranks land in a
// sink buffer and all arrival words are this rank's own -- so that the N-rank slice of a matrix (rows, launches, barrier
// code path) can be timed and profiled on ONE GPU.  The radiance of the other ranks' rows is never refreshed, so the light
// is wrong; only the timing is meaningful.  Never used unless the variable is set.
static int setup_simulated_peers(vrad_env* e, size_t n_pad) {
    PeerLinks& P = e->peers;
    if (P.ready && P.simulated && P.n_pad == n_pad) return 0;
    const int world = e->cfg.world, rank = e->cfg.rank;      // world 1 (k4_items): a one-rank table, every row is this rank's -- results are exact
    if (P.d_flags.alloc(kFlagWords) || P.d_sink.alloc(n_pad) || P.d_table.alloc(1)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemsetAsync(P.d_flags.p, 0, kFlagWords * sizeof(uint32_t), e->stream));
    PeerTable tbl{};
    for (int r = 0; r < world; r++) {
        tbl.er[0][r] = r == rank ? e->d_er[0].p : P.d_sink.p;
        tbl.er[1][r] = r == rank ? e->d_er[1].p : P.d_sink.p;
        tbl.flags[r] = P.d_flags.p + (r - rank);          // flags[r][rank] == own word r: one signal fills every arrival word
    }
    tbl.world = world; tbl.rank = rank;
    VRAD_CUDA_CHECK(cudaMemcpyAsync(P.d_table.p, &tbl, sizeof(tbl), cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    P.ready = true; P.simulated = true; P.n_pad = n_pad; P.table_world = world;
    return 0;
}

int vrad_bounce(vrad_env* e, const float* emit0_rgb, int n_bounces, int early_out, float* total_rgb_out,
                float added_last[3], int* bounces_done) {
    VRAD_MULTI(e, group_bounce(e, emit0_rgb, n_bounces, early_out, total_rgb_out, added_last, bounces_done));
    if (!e || !emit0_rgb || n_bounces < 0) { set_error("vrad_bounce: bad arguments"); return VRAD_E_INVALID; }
    TransfersDev& T = e->transfers;
    if (!T.ready) { set_error("vrad_bounce: no transfers resident (vrad_build_transfers / vrad_transfers_upload first)"); return VRAD_E_STATE; }
    const bool sim = (e->opt.k4_sim_peers && e->cfg.world > 1 && !has_comm(e)) || (e->opt.k4_items && e->cfg.world == 1);
    if (e->cfg.world > 1 && !has_comm(e) && !sim) { set_error("vrad_bounce: world=%d but vrad_comm_init was not called", e->cfg.world); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const int64_t N = e->patches.n;
    const int world = e->cfg.world;
    const int64_t rpr = rows_per_rank(e, N);
    const int64_t n_pad = rpr * world;           // >= N; radiance buffers are indexed by global patch number
    int64_t bounds[kMaxWorld + 1] = {0};
    if (world > 1 && !sim) {
        if (world > kMaxWorld) { set_error("vrad_bounce: world %d > %d", world, kMaxWorld); return VRAD_E_UNSUPPORTED; }
        // the block boundaries change only when the rows do (a collective call on every rank): exchanged and checked once per
        // set of rows, not once per call -- the exchange is a collective plus a host synchronisation in front of every bounce loop
        if (e->bounds_serial != T.rows_serial) {
            int rcb = comm_exchange_bounds(e, T.row0, T.row1, N, e->bounds);
            if (rcb) return rcb;
            e->bounds_serial = T.rows_serial;
        }
        memcpy(bounds, e->bounds, sizeof(bounds));
    } else if (world == 1 && (T.row0 != 0 || T.row1 != N)) {
        set_error("vrad_bounce: world=1 but the resident transfer rows are [%lld,%lld) of %lld", (long long)T.row0, (long long)T.row1, (long long)N);
        return VRAD_E_STATE;
    }
    const int nloc = (int)(T.row1 - T.row0);
    const int nblocks = std::max(1, (nloc + kGatherWarps - 1) / kGatherWarps);
    const int plain_blocks = plain_gather_blocks(T, nloc);          // partials the single-GPU gather writes (fewer with block rows)
    const int add_blocks = std::max(1, (nloc + 1023) / 1024);
    // `total` is indexed by global row too, so that the final gather is in place
    if (e->d_er[0].alloc(n_pad) || e->d_er[1].alloc(n_pad) || e->d_total.alloc(n_pad) || e->d_add.alloc((size_t)nloc + 1) ||
        e->d_partials.alloc(3 * (size_t)std::max(nblocks, add_blocks) + 8)) {
        set_error("out of device memory for bounce state"); return VRAD_E_NOMEM;
    }
    int rc; const void* d_emit0; bool h_in, h_out;
    if ((rc = stage_in(e, 0, emit0_rgb, (size_t)N * 12, &d_emit0, &h_in))) return rc;
    void* d_out3;
    if ((rc = stage_out(e, 1, total_rgb_out, (size_t)N * 12, &d_out3, &h_out))) return rc;
    float* d_added = e->d_partials.p + 3 * (size_t)std::max(nblocks, add_blocks);       // 3 floats after the partials
    // with a patch hierarchy the interior patches are recomputed from the exchanged leaf rows on every rank, which
    // needs the complete buffer before the next gather starts: the fused peer-store exchange is not used then
    const PatchesDev& PD = e->patches;
    const bool hier = PD.hier && PD.n_interior > 0;
    const int collect_long_blocks = (PD.n_collect_long + 7) / 8;
    const int collect_blocks = collect_long_blocks + (PD.n_interior - PD.n_collect_long + 31) / 32;
    if (sim && !hier && !e->patches.bump) { if ((rc = setup_simulated_peers(e, (size_t)n_pad))) return rc; }
    else if (world > 1 && !sim && !e->patches.bump && (rc = comm_setup_peers(e, (size_t)n_pad))) return rc;
    const bool p2p = (world > 1 || sim) && !hier && !e->patches.bump && e->peers.ready;
    // patch hierarchy on several GPUs: the leaf rows travel by peer stores out of the short-row gather, the barrier's wait half is the
    // prologue of k4_collect_parents (two launches per bounce, no all-gather pass)
    const bool p2p_hier = world > 1 && !sim && hier && !e->patches.bump && e->peers.ready && e->opt.k4_hier_p2p;
    if (sim && world > 1 && !p2p) { set_error("vrad_bounce: VRAD_K4_SIM_PEERS does not cover the patch hierarchy"); return VRAD_E_UNSUPPORTED; }
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
    const int force_short = e->opt.k4_short;
    const bool use_short = !p2p && (hier || (force_short >= 0 ? force_short != 0 : (T.nnz < (int64_t)400 * std::max(1, n_short))));
    static const int short_cfg = [] { const char* v = getenv("VRAD_K4_SHORT_CFG"); return v ? atoi(v) : 884; }();   // experiments; see the switch below
    const int short_lanes = short_cfg / 10 == 16 ? 16 : (short_cfg / 10 == 4 ? 4 : 8);
    const int short_rpb = kGatherBlock / short_lanes;
    const int short_blocks = std::max(1, (n_short + short_rpb - 1) / short_rpb);
    const bool use_pdl = e->opt.k4_pdl != 0, use_graph = e->opt.k4_graph != 0;

    // L2 residency of the transfer stream (k4_l2_mb; multi-GPU form)
    {
        const size_t want = p2p && e->opt.k4_l2_mb != 0 ? (size_t)std::abs(e->opt.k4_l2_mb) << 20 : 0;      // negative: window with the .cs loads kept (experiment)
        if (want != e->l2_set_aside_req) {
            cudaDeviceProp prop;
            VRAD_CUDA_CHECK(cudaGetDeviceProperties(&prop, e->cfg.device));
            const size_t set_aside = std::min(want, (size_t)prop.persistingL2CacheMaxSize);
            VRAD_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside));
            e->l2_set_aside_req = want; e->l2_set_aside = set_aside;
            if (getenv("VRAD_VERBOSE")) fprintf(stderr, "[vrad] L2 set-aside %zu MB (max %d MB, window max %d MB)\n", set_aside >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
            e->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
            e->bounce_graph.tag = -1;            // the window is baked into the captured launches
        }
        const size_t stream_bytes = (size_t)T.nnz_padded * 8;
        e->l2_window_bytes = e->l2_set_aside ? std::min(stream_bytes, e->l2_window_max) : 0;
        e->l2_hit_ratio = e->l2_window_bytes ? std::min(1.0f, (float)((double)e->l2_set_aside / (double)e->l2_window_bytes)) : 0.0f;
    }
    timing_begin(e);
    int launches = 0;
    float4* total_local = e->d_total.p + T.row0;
    VRAD_CUDA_CHECK(cudaMemsetAsync(e->d_total.p, 0, (size_t)n_pad * 16, e->stream));
    VRAD_CUDA_CHECK(cudaMemsetAsync(d_added, 0, 12, e->stream));
    k4_init_er<<<(int)((n_pad + 255) / 256), 256, 0, e->stream>>>((int)n_pad, (int)N, (const float*)d_emit0, e->patches.refl.p, e->d_er[0].p);
    launches++;

    int cur = 0, done = 0;
    float h_added[3] = {0.f, 0.f, 0.f};
    static const bool verbose = getenv("VRAD_TIMING") != nullptr;
    constexpr int kProbe = 16;
    cudaEvent_t pe[kProbe][3];
    int n_probe = 0;
    // One bounce of the forms that are nothing but kernel launches whose arguments depend only on the bounce number: the flat gather
    // (plain kernel on one GPU, work-item kernel with the fused exchange on several) and the hierarchical gather (short rows +
    // CollectLight, with the fused exchange and the barrier wait in the collect kernel on several GPUs).
    const bool launch_only = !bump && (world == 1 || p2p || p2p_hier) && (!use_short || short_cfg == 884);
    auto enqueue_bounce = [&](int b, int c, bool want_add) -> cudaError_t {
        if (!use_short) return launch_gather(e, p2p, p2p && use_pdl && b > 0, c, want_add, p2p ? (uint32_t)b : 0u, (uint32_t)b + 1u, total_local);
        if (p2p_hier) {
            const ShortPeers SP{e->peers.d_table.p, e->peers.d_flags.p, c ^ 1, (uint32_t)b + 1u};
            k4_gather_short<8, 8, 4, false, true><<<short_blocks, kGatherBlock, 0, e->stream>>>(n_short, d_rows, T.row0, T.rowptr.p, T.tr.p, e->d_er[c].p,
                                                                                              e->patches.refl.p, e->d_er[c ^ 1].p, total_local, e->d_partials.p, SP);
        } else {
            k4_gather_short<8, 8, 4, false><<<short_blocks, kGatherBlock, 0, e->stream>>>(n_short, d_rows, T.row0, T.rowptr.p, T.tr.p, e->d_er[c].p,
                                                                                        e->patches.refl.p, e->d_er[c ^ 1].p, total_local, e->d_partials.p);
        }
        if (hier)
            k4_collect_parents<<<collect_blocks, 256, 0, e->stream>>>(PD.n_interior, PD.n_collect_long, collect_long_blocks, PD.collect_ids.p, PD.collect_ptr.p, PD.collect_ent.p, e->d_er[c ^ 1].p,
                                                                      p2p_hier ? e->peers.d_flags.p : nullptr, e->peers.table_world, p2p_hier ? (uint32_t)b + 1u : 0u);
        return cudaGetLastError();
    };
    const int launches_per_bounce = use_short && hier ? 2 : 1;
    // Such a loop is captured once per (bounce count, buffers, plan) as a CUDA graph and replayed: the host issues one graph launch
    // per call instead of one or two kernels per bounce.
    const bool graphed = launch_only && use_graph && !early_out && !verbose && n_bounces >= 4;
    if (graphed) {
        GraphCache& G = e->bounce_graph;
        const int64_t graph_tag = ((T.plan_serial * 8 + (use_pdl ? 4 : 0) + (p2p ? 2 : 0) + (p2p_hier ? 1 : 0)) * 4 + (use_short ? 2 : 0) + (hier ? 1 : 0)) * 4 + (T.blocked ? 2 : T.packed ? 1 : 0);
        const void* key_items = use_short ? (const void*)d_rows : (const void*)T.items.p;
        const int key_n = use_short ? n_short : T.n_items;
        const bool hit = G.exec && G.n_bounces == n_bounces && G.items == key_items && G.n_items == key_n && G.er0 == e->d_er[0].p && G.total == total_local &&
                         G.p2p == p2p && G.row0 == T.row0 && G.add == e->d_add.p && G.tr == T.tr.p && G.tag == graph_tag;
        if (!hit) {
            if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
            // capture on the handle's own stream (the caller's stream may be the legacy default stream, which cannot capture)
            cudaStream_t user = e->stream;
            VRAD_CUDA_CHECK(cudaStreamSynchronize(user));
            e->stream = e->own_stream;
            cudaError_t ce = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
            int c = 0;
            for (int b = 0; b < n_bounces && ce == cudaSuccess; b++) {
                ce = enqueue_bounce(b, c, b + 1 == n_bounces);
                c ^= 1;
            }
            cudaGraph_t g = nullptr;
            cudaError_t ce2 = cudaStreamEndCapture(e->stream, &g);
            e->stream = user;
            if (ce == cudaSuccess) ce = ce2;
            if (ce == cudaSuccess) ce = cudaGraphInstantiate(&G.exec, g, 0);
            if (g) cudaGraphDestroy(g);
            if (ce != cudaSuccess) {       // capture not available here (e.g. a driver without programmatic edges): plain launches from now on
                G.exec = nullptr; cudaGetLastError();
                if (getenv("VRAD_VERBOSE")) fprintf(stderr, "[vrad] bounce-loop graph capture failed (%s); using stream launches\n", cudaGetErrorString(ce));
                e->opt.k4_graph = 0;
            }
            G.n_bounces = n_bounces; G.items = key_items; G.n_items = key_n; G.er0 = e->d_er[0].p; G.total = total_local; G.p2p = p2p; G.row0 = T.row0;
            G.add = e->d_add.p; G.tr = T.tr.p; G.tag = graph_tag;
        }
        if (G.exec) {
            VRAD_CUDA_CHECK(cudaGraphLaunch(G.exec, e->stream));
            launches += n_bounces * launches_per_bounce;
            done = n_bounces; cur = n_bounces & 1;
            if (p2p) {
                k4_sum_added_rows<<<add_blocks, 256, 0, e->stream>>>(nloc, e->d_add.p, e->d_partials.p);
                k4_reduce_added<<<1, 256, 0, e->stream>>>(add_blocks, e->d_partials.p, d_added);
                launches += 2;
            } else { k4_reduce_added<<<1, 256, 0, e->stream>>>(use_short ? short_blocks : plain_blocks, e->d_partials.p, d_added); launches++; }
            if (world > 1 && !sim && (rc = comm_allreduce3(e, d_added))) return rc;
        }
    }
    for (int b = 0; b < n_bounces && done < n_bounces; b++) {
        const bool probe = verbose && b >= 4 && n_probe < kProbe;
        const bool last = (b + 1 == n_bounces);
        if (probe) { for (int k = 0; k < 3; k++) cudaEventCreate(&pe[n_probe][k]); cudaEventRecord(pe[n_probe][0], e->stream); }
        if (use_short && p2p_hier) {
            const ShortPeers SP{e->peers.d_table.p, e->peers.d_flags.p, cur ^ 1, (uint32_t)b + 1u};
            k4_gather_short<8, 8, 4, false, true><<<short_blocks, kGatherBlock, 0, e->stream>>>(n_short, d_rows, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p,
                                                                                              e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, e->d_partials.p, SP);
        } else if (use_short) {
#define VRAD_SHORT(L, B, ...) k4_gather_short<L, B, ##__VA_ARGS__><<<short_blocks, kGatherBlock, 0, e->stream>>>(n_short, d_rows, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p, \
                                                                 e->patches.refl.p, e->d_er[cur ^ 1].p, total_local, e->d_partials.p)
            switch (short_cfg) {                                  // lanes per row, min blocks/SM[, unroll, prefetch]: measured on the hierarchical S2 matrix
                case 86: VRAD_SHORT(8, 6); break;                 // 64.1 us (40 registers, one-step prefetch)
                case 168: VRAD_SHORT(16, 8); break;               // 77 us
                case 48: VRAD_SHORT(4, 8); break;                 // 82 us
                default: VRAD_SHORT(8, 8, 4, false); break;       // 61.9 us: no prefetch, 32 registers, 64 warps per SM
            }
#undef VRAD_SHORT
        } else {
            // p2p: bounce b waits for every rank's epoch base+b (bounce 0 reads locally initialised radiance) and signals base+b+1
            VRAD_CUDA_CHECK(launch_gather(e, p2p, p2p && use_pdl && b > 0 && !probe, cur, early_out || last, p2p ? (uint32_t)b : 0u, (uint32_t)b + 1u, total_local));
        }
        launches++;
        if (bump && PD.n_bump_rows > 0) {       // same emitters as the gather above (er[cur]), bump-mapped rows only
            k4_gather_bump<<<(PD.n_bump_rows * 32 + 255) / 256, 256, 0, e->stream>>>(PD.n_bump_rows, PD.bump_rows.p, T.row0, T.rowptr.p, T.tr.p, e->d_er[cur].p,
                                                                                  PD.origin_area.p, PD.normal_dist.p, PD.bump_normals.p,
                                                                                  PD.total_bump[0].p, PD.total_bump[1].p, PD.total_bump[2].p);
            launches++;
        }
        if (probe) cudaEventRecord(pe[n_probe][1], e->stream);
        if (!p2p && !p2p_hier && world > 1 && !sim && (rc = comm_allgather_rows(e, e->d_er[cur ^ 1].p, bounds))) return rc;
        if (hier) {     // CollectLight, interior patches: emit of a parent = area-weighted average of its children
            k4_collect_parents<<<collect_blocks, 256, 0, e->stream>>>(PD.n_interior, PD.n_collect_long, collect_long_blocks, PD.collect_ids.p, PD.collect_ptr.p, PD.collect_ent.p, e->d_er[cur ^ 1].p,
                                                                      p2p_hier ? e->peers.d_flags.p : nullptr, e->peers.table_world, p2p_hier ? (uint32_t)b + 1u : 0u);
            launches++;
        }
        if (probe) { cudaEventRecord(pe[n_probe][2], e->stream); n_probe++; }
        cur ^= 1; done++;
        if (early_out || last) {
            if (use_short || !p2p) { k4_reduce_added<<<1, 256, 0, e->stream>>>(use_short ? short_blocks : plain_blocks, e->d_partials.p, d_added); launches++; }
            else {
                k4_sum_added_rows<<<add_blocks, 256, 0, e->stream>>>(nloc, e->d_add.p, e->d_partials.p);
                k4_reduce_added<<<1, 256, 0, e->stream>>>(add_blocks, e->d_partials.p, d_added);
                launches += 2;
            }
            if (world > 1 && !sim && (rc = comm_allreduce3(e, d_added))) return rc;
        }
        if (early_out && !last) {
            VRAD_CUDA_CHECK(cudaMemcpyAsync(h_added, d_added, 12, cudaMemcpyDeviceToHost, e->stream));
            VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
            if (h_added[0] < 1.0f && h_added[1] < 1.0f && h_added[2] < 1.0f) break;
        }
    }
    if ((p2p || p2p_hier) && done > 0) { k4_peer_wait<<<1, 32, 0, e->stream>>>(e->peers.d_flags.p, e->peers.table_world, (uint32_t)done); launches++; }
    if (world > 1 && !sim && (rc = comm_allgather_rows(e, e->d_total.p, bounds))) return rc;
    if (bump && world > 1 && !sim)          // TotalLight.Light[1..3] of every rank's bump-mapped rows
        for (int bb = 0; bb < 3; bb++) if ((rc = comm_allgather_rows(e, e->patches.total_bump[bb].p, bounds))) return rc;
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
        fprintf(stderr, "[vrad] rank %d gather %.1f us, exchange (%s) %.1f us per bounce (%d probed)\n", e->cfg.rank,
                1e3f * g / n_probe, p2p ? "peer stores + in-kernel barrier" : (world > 1 ? "ncclAllGather" : "none"), 1e3f * x / n_probe, n_probe);
    }
    if ((rc = finish_out(e, total_rgb_out, d_out3, (size_t)N * 12, h_out))) return rc;
    bool need_sync = h_in || h_out || added_last != nullptr;
    if (added_last) VRAD_CUDA_CHECK(cudaMemcpyAsync(added_last, d_added, 12, cudaMemcpyDeviceToHost, e->stream));
    if (bounces_done) *bounces_done = done;
    uint32_t h_err = 0;
    if ((p2p || p2p_hier) && (need_sync || !e->async)) {       // a wait that gave up (dead peer) is an error, not a hang; async callers see it on their next synchronous call
        VRAD_CUDA_CHECK(cudaMemcpyAsync(&h_err, e->peers.d_flags.p + kFlagError, 4, cudaMemcpyDeviceToHost, e->stream));
        need_sync = true;
    }
    if ((rc = sync_if_needed(e, need_sync))) return rc;
    if (h_err) { set_error("vrad_bounce: a rank never signalled the end of a bounce (inter-GPU barrier timed out)"); return VRAD_E_COMM; }
    return VRAD_OK;
}

} // extern "C"
