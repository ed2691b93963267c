// Measured-and-rejected K4 variant (round 1): per-warp shared-memory rings filled by cp.async.bulk (UBLKCP)
// + mbarrier.  406 us/bounce on S2 vs 238-270 us for the register-pipelined warp kernel: the consumer side is
// bound by er[] gather latency, and the ring costs occupancy.  Kept for the record; not compiled into the product.
// ---------------------------------------------------------------------------------------------
// TMA-staged variant (default).  The transfer stream is contiguous in HBM, so instead of pulling it
// through registers each warp owns a ring of shared-memory stages that one elected lane fills with
// 1-D bulk async copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) -- kStages-1 chunks
// of 2 KB are in flight per warp (96 KB per SM) without costing registers or L1 lines, and L1 is
// left to the er[] gathers.  Persistent grid: one CTA per SM, every warp walks a contiguous range
// of rows holding an equal share of the entries (k4_partition), carrying the running row sum in
// registers across chunk boundaries.
constexpr int kTmaWarps = 16;
constexpr int kTmaStages = 3;
constexpr int kTmaChunk = 256;                 // entries per stage (2 KB); 2 CTAs per SM -> 32 warps, 128 KB in flight
constexpr int kTmaThreads = kTmaWarps * 32;
constexpr size_t kTmaSmem = (size_t)kTmaWarps * kTmaStages * kTmaChunk * sizeof(int2) + kTmaWarps * kTmaStages * 8 + 3 * kTmaWarps * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// warp w of the persistent grid handles rows [warp_rows[w], warp_rows[w+1]): equal shares of the padded stream
__global__ void k4_partition(int nloc, int nwarps, const int64_t* __restrict__ rowptr, int32_t* __restrict__ warp_rows) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > nwarps) return;
    if (w == nwarps) { warp_rows[w] = nloc; return; }
    const int64_t total = rowptr[nloc];
    const int64_t target = (total / nwarps) * w + ((total % nwarps) * w) / nwarps;
    int lo = 0, hi = nloc;                      // first row whose start >= target
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rowptr[mid] >= target) hi = mid; else lo = mid + 1; }
    warp_rows[w] = w == 0 ? 0 : lo;
}

__global__ void __launch_bounds__(kTmaThreads, 2)
k4_gather_tma(int64_t row0, const int32_t* __restrict__ warp_rows, const int64_t* __restrict__ rowptr,
              const int2* __restrict__ tr, const float4* __restrict__ er, const float4* __restrict__ refl,
              float4* __restrict__ er_next, float4* __restrict__ total, float* __restrict__ partials) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int2* stage_buf = reinterpret_cast<int2*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kTmaWarps * kTmaStages * kTmaChunk * sizeof(int2));
    float* added_sm = reinterpret_cast<float*>(bars + kTmaWarps * kTmaStages);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kTmaWarps + warp;
    int2* my_buf = stage_buf + (size_t)warp * kTmaStages * kTmaChunk;
    const uint32_t bar0 = smem_u32(bars + warp * kTmaStages);
    if (lane == 0) {
        for (int s = 0; s < kTmaStages; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int ra = warp_rows[gw], rb = warp_rows[gw + 1];
    const int64_t seg0 = rowptr[ra], seg1 = rowptr[rb];
    const int n_chunks = (int)((seg1 - seg0 + kTmaChunk - 1) / kTmaChunk);
    auto issue = [&](int c) {                  // lane 0 only
        const int s = c % kTmaStages;
        const int64_t cb = seg0 + (int64_t)c * kTmaChunk;
        const uint32_t bytes = (uint32_t)(min((int64_t)kTmaChunk, seg1 - cb) * sizeof(int2));
        mbar_expect_tx(bar0 + 8 * s, bytes);
        bulk_g2s(smem_u32(my_buf + s * kTmaChunk), tr + cb, bytes, bar0 + 8 * s);
    };
    if (lane == 0) for (int c = 0; c < kTmaStages - 1 && c < n_chunks; c++) issue(c);

    float e0 = 0.f, e1 = 0.f, e2 = 0.f;        // this warp's share of `added` (lane 0)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;        // running row sum (per lane)
    int row = ra;
    int64_t row_end = row < rb ? rowptr[row + 1] : seg1;
    auto finish_row = [&]() {                   // reduce the row held in (s0,s1,s2) and run CollectLight
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            const float4 r = refl[row0 + row];
            if (r.w == 0.0f) {
                float4 t = total[row];
                t.x += s0; t.y += s1; t.z += s2;
                total[row] = t;
                er_next[row0 + row] = make_float4(s0 * r.x, s1 * r.y, s2 * r.z, 0.f);
                e0 += s0; e1 += s1; e2 += s2;
            } else {
                er_next[row0 + row] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        s0 = s1 = s2 = 0.f;
        row++;
        row_end = row < rb ? rowptr[row + 1] : seg1;
    };
    int64_t p = seg0;
    while (row < rb && row_end == p) finish_row();                     // leading empty rows
    for (int c = 0; c < n_chunks; c++) {
        const int s = c % kTmaStages;
        const uint32_t parity = (uint32_t)((c / kTmaStages) & 1);
        if (lane == 0 && c + kTmaStages - 1 < n_chunks) issue(c + kTmaStages - 1);   // stage freed at the end of step c-1
        while (!mbar_try_wait(bar0 + 8 * s, parity)) {}
        const int2* buf = my_buf + s * kTmaChunk;
        const int64_t cb = seg0 + (int64_t)c * kTmaChunk;
        const int64_t ce = min(cb + kTmaChunk, seg1);
        if (p == cb && ce - cb == kTmaChunk && row_end >= ce) {
            // fast path: the whole chunk belongs to the current row -- 8 entries per lane, all in flight
            float4 x[8]; float wv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { const int2 a = buf[lane + 32 * j]; wv[j] = __int_as_float(a.y); x[j] = __ldg(&er[a.x]); }
            float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                s0 += wv[j] * x[j].x; s1 += wv[j] * x[j].y; s2 += wv[j] * x[j].z;
                t0 += wv[j + 1] * x[j + 1].x; t1 += wv[j + 1] * x[j + 1].y; t2 += wv[j + 1] * x[j + 1].z;
            }
            s0 += t0; s1 += t1; s2 += t2;
            p = ce;
            while (row < rb && row_end == p) finish_row();
        }
        while (p < ce) {
            const int lim = (int)(min(row_end, ce) - cb);
            int k = (int)(p - cb) + lane;
            for (; k < lim; k += 32) {
                const int2 a = buf[k];
                const float4 xa = __ldg(&er[a.x]);
                const float wa = __int_as_float(a.y);
                s0 += wa * xa.x; s1 += wa * xa.y; s2 += wa * xa.z;
            }
            p = cb + lim;
            while (row < rb && row_end == p) finish_row();              // row complete (and any empty rows after it)
        }
        __syncwarp();                                                   // all lanes done with this stage
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the next async write
    }
    while (row < rb) finish_row();                                      // trailing empty rows
    if (lane == 0) { added_sm[3 * warp] = e0; added_sm[3 * warp + 1] = e1; added_sm[3 * warp + 2] = e2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kTmaWarps; k++) a += added_sm[3 * k + threadIdx.x];
        partials[3 * (size_t)blockIdx.x + threadIdx.x] = a;
    }
}

