// Development micro-benchmark for K4 gather variants on a synthetic CSR shaped like S2
// (187k rows x ~1024 entries, columns in runs of ~16 inside a +-2000 window).  Not product code.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U, int MODE, bool PREFETCH>
__global__ void __launch_bounds__(256) gather_warp(int nloc, const int64_t* __restrict__ rowptr, const int2* __restrict__ tr,
                                                   const float4* __restrict__ er, float4* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 8 + warp;
    if (row >= nloc) return;
    const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];
    float s0 = 0, s1 = 0, s2 = 0;
    const int2 zero = make_int2(0, 0);
    int2 cur[U], nxt[U];
    int64_t k = k0 + lane;
#pragma unroll
    for (int j = 0; j < U; j++) cur[j] = k + 32 * j < k1 ? __ldcs(&tr[k + 32 * j]) : zero;
    for (; k < k1; k += 32 * U) {
        if (PREFETCH) {
#pragma unroll
            for (int j = 0; j < U; j++) nxt[j] = k + 32 * (U + j) < k1 ? __ldcs(&tr[k + 32 * (U + j)]) : zero;
        }
        float4 x[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            if (MODE == 0) x[j] = __ldg(&er[cur[j].x]);
            else if (MODE == 1) x[j] = make_float4(1.f, 2.f, 3.f, 0.f);
            else x[j] = __ldg(&er[(cur[j].x & 1023)]);
        }
#pragma unroll
        for (int j = 0; j < U; j++) { float w = __int_as_float(cur[j].y); s0 += w * x[j].x; s1 += w * x[j].y; s2 += w * x[j].z; }
        if (PREFETCH) {
#pragma unroll
            for (int j = 0; j < U; j++) cur[j] = nxt[j];
        } else {
#pragma unroll
            for (int j = 0; j < U; j++) cur[j] = k + 32 * (U + j) < k1 ? __ldcs(&tr[k + 32 * (U + j)]) : zero;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(~0u, s0, o); s1 += __shfl_xor_sync(~0u, s1, o); s2 += __shfl_xor_sync(~0u, s2, o); }
    if (lane == 0) out[row] = make_float4(s0, s1, s2, 0);
}

constexpr int kGatherBlock = 256; constexpr int kGatherWarps = 8;
__global__ void __launch_bounds__(kGatherBlock)
k4_gather(int nloc, int64_t row0, const int64_t* __restrict__ rowptr, const int2* __restrict__ tr,
          const float4* __restrict__ er, const float4* __restrict__ refl,
          float4* __restrict__ er_next, float4* __restrict__ total, float* __restrict__ partials) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * kGatherWarps + warp;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (row < nloc) {
        const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];      // padded to 4 entries; padding has w = 0
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f;
        // software pipeline: the {col,w} stream of step i+1 is in flight while step i's er[] gathers
        // resolve, so the two dependent memory latencies overlap.  Out-of-row slots read as {0, 0.0f}.
        const int2 zero = make_int2(0, 0);
        int64_t k = k0 + lane;
        int2 a = k < k1 ? __ldcs(&tr[k]) : zero, b = k + 32 < k1 ? __ldcs(&tr[k + 32]) : zero;
        int2 c = k + 64 < k1 ? __ldcs(&tr[k + 64]) : zero, d = k + 96 < k1 ? __ldcs(&tr[k + 96]) : zero;
        for (; k < k1; k += 128) {
            const int64_t kn = k + 128;
            const int2 na = kn < k1 ? __ldcs(&tr[kn]) : zero, nb = kn + 32 < k1 ? __ldcs(&tr[kn + 32]) : zero;
            const int2 nc = kn + 64 < k1 ? __ldcs(&tr[kn + 64]) : zero, nd2 = kn + 96 < k1 ? __ldcs(&tr[kn + 96]) : zero;
            const float4 xa = __ldg(&er[a.x]), xb = __ldg(&er[b.x]), xc = __ldg(&er[c.x]), xd = __ldg(&er[d.x]);
            const float wa = __int_as_float(a.y), wb = __int_as_float(b.y), wc = __int_as_float(c.y), wd = __int_as_float(d.y);
            s0 += wa * xa.x; s1 += wa * xa.y; s2 += wa * xa.z;
            t0 += wb * xb.x; t1 += wb * xb.y; t2 += wb * xb.z;
            u0 += wc * xc.x; u1 += wc * xc.y; u2 += wc * xc.z;
            v0 += wd * xd.x; v1 += wd * xd.y; v2 += wd * xd.z;
            a = na; b = nb; c = nc; d = nd2;
        }
        s0 += t0 + u0 + v0; s1 += t1 + u1 + v1; s2 += t2 + u2 + v2;
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


// thread-per-entry flat streaming baseline: pure read bandwidth of the tr stream (sum of weights)
__global__ void stream_only(int64_t n, const int4* __restrict__ tr4, float* __restrict__ out) {
    float s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += (int64_t)gridDim.x * blockDim.x) {
        int4 v = __ldcs(&tr4[i]); s += __int_as_float(v.y) + __int_as_float(v.w);
    }
    if (s == 123.456f) out[0] = s;
}

template <typename F> float timeit(F f, int iters = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < iters; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / iters;
}

int main(int argc, char** argv) {
    int N = 187328; const int L = 1024;
    std::vector<int64_t> rp; std::vector<int2> tr;
    if (argc > 1) {
        FILE* f = fopen(argv[1], "rb"); int64_t n64, nnz; fread(&n64, 8, 1, f); fread(&nnz, 8, 1, f); N = (int)n64;
        std::vector<int64_t> rp0(N + 1); std::vector<int32_t> col(nnz); std::vector<float> w(nnz);
        fread(rp0.data(), 8, N + 1, f); fread(col.data(), 4, nnz, f); fread(w.data(), 4, nnz, f); fclose(f);
        rp.resize(N + 1); rp[0] = 0;
        for (int i = 0; i < N; i++) rp[i + 1] = rp[i] + ((rp0[i + 1] - rp0[i] + 3) & ~3LL);
        tr.assign(rp[N], make_int2(0, 0));
        int64_t mn = 1 << 30, mx = 0;
        for (int i = 0; i < N; i++) { int64_t len = rp0[i + 1] - rp0[i]; mn = std::min(mn, len); mx = std::max(mx, len);
            for (int64_t k = 0; k < len; k++) { int2 v; v.x = col[rp0[i] + k]; v.y = *(int*)&w[rp0[i] + k]; tr[rp[i] + k] = v; } }
        printf("loaded N=%d nnz=%lld rowlen min %lld max %lld\n", N, (long long)nnz, (long long)mn, (long long)mx);
    } else {
    rp.resize(N + 1); tr.resize((size_t)N * L);
    uint64_t st = 12345;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    for (int i = 0; i < N; i++) {
        rp[i] = (int64_t)i * L;
        int c = std::max(0, i - 2000);
        for (int k = 0; k < L;) {
            int run = 1 + rnd() % 31; c += 1 + rnd() % 40; 
            for (int r = 0; r < run && k < L; r++, k++) { float w = 1e-3f; int2 v; v.x = std::min(c++, N - 1); v.y = *(int*)&w; tr[(size_t)i * L + k] = v; }
        }
    }
    rp[N] = (int64_t)N * L;
    }
    int64_t* d_rp; int2* d_tr; float4 *d_er, *d_out; float* d_f;
    CK(cudaMalloc(&d_rp, (N + 1) * 8)); CK(cudaMalloc(&d_tr, tr.size() * 8)); CK(cudaMalloc(&d_er, N * 16)); CK(cudaMalloc(&d_out, N * 16)); CK(cudaMalloc(&d_f, 16));
    CK(cudaMemcpy(d_rp, rp.data(), (N + 1) * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_tr, tr.data(), tr.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_er, 0, N * 16));
    const double gb = tr.size() * 8 / 1e9;
    int blocks = (N + 7) / 8;
    printf("stream bytes %.3f GB\n", gb);
    float ms;
    ms = timeit([&] { stream_only<<<148 * 16, 512>>>((int64_t)tr.size(), (const int4*)d_tr, d_f); }); printf("stream_only int4 flat          : %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
#define RUN(U, MODE, PF, name) ms = timeit([&] { gather_warp<U, MODE, PF><<<blocks, 256>>>(N, d_rp, d_tr, d_er, d_out); }); printf("%-32s: %.1f us  %.0f GB/s\n", name, ms * 1e3, gb / ms * 1e3);
    { float4* d_refl; float4* d_tot; float* d_part; CK(cudaMalloc(&d_refl, N * 16)); CK(cudaMalloc(&d_tot, N * 16)); CK(cudaMalloc(&d_part, blocks * 12 + 64));
      CK(cudaMemset(d_refl, 0, N * 16)); CK(cudaMemset(d_tot, 0, N * 16));
      ms = timeit([&] { k4_gather<<<blocks, 256>>>(N, 0, d_rp, d_tr, d_er, d_refl, d_out, d_tot, d_part); }); printf("%-32s: %.1f us  %.0f GB/s\n", "PRODUCT k4_gather", ms * 1e3, gb / ms * 1e3);
      ms = timeit([&] { k4_gather<<<blocks, 256>>>(N, 0, d_rp, d_tr, d_er, d_refl, d_out, d_tot, d_part); }, 100); printf("%-32s: %.1f us  %.0f GB/s\n", "PRODUCT k4_gather x100", ms * 1e3, gb / ms * 1e3); }
    RUN(4, 1, true, "warp U4 prefetch, no gather")
    RUN(8, 1, true, "warp U8 prefetch, no gather")
    RUN(8, 1, false, "warp U8 noprefetch, no gather")
    RUN(4, 2, true, "warp U4 prefetch, L1 gather")
    RUN(4, 0, true, "warp U4 prefetch, real gather")
    RUN(4, 0, false, "warp U4 noprefetch, real gather")
    RUN(8, 0, true, "warp U8 prefetch, real gather")
    RUN(8, 0, false, "warp U8 noprefetch, real gather")
    RUN(2, 0, true, "warp U2 prefetch, real gather")
    CK(cudaDeviceSynchronize());
    return 0;
}
