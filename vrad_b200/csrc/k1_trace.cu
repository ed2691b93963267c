// k1_trace.cu -- K1: batched kd-tree ray casting (closest hit) and segment visibility.
//
// Replaces Environment.Trace4Rays (raytracer/environment.go:140-145, a stub in the reference) for
// whole batches, and the per-lane post-processing of trace.TestLineDoesHitSky
// (raytracer/trace/testline.go:22-51).  One thread per ray; rays are SoA so every load and
// store of a warp is one coalesced 128-byte transaction; scene records come in through the
// read-only path as 64/128-bit vectors (common.cuh).  Algorithmic HBM bytes: 32 B/ray
// (closest hit: 24 in + 8 out), 24.125 B/segment (visibility).
#include "env_internal.cuh"

namespace vrad {

constexpr int kTraceBlock = 128;
constexpr int kTraceWarps = kTraceBlock / 32;
constexpr int kRaysPerWarp = 512;              // each warp streams a contiguous chunk of the batch

// Streaming traversal with lane refill: a warp owns rays [base, base + kRaysPerWarp) and keeps its 32
// lanes busy -- whenever lanes retire (their ray finished) they take the next rays of the chunk before
// the next descend/leaf round, so the warp does not idle on the longest ray of a fixed batch of 32
// (ncu r01 v2: 6.2 threads per instruction with fixed batches; rays average 2.9 leaf visits but a
// batch needs 9.3 rounds).  Results are per ray, so the order in which lanes pick rays is irrelevant.
template <typename Fetch, typename Retire, bool ANY_HIT>
__device__ __forceinline__ void stream_rays(const DevScene& S, int64_t base, int64_t end, int skip_id, Fetch fetch, Retire retire) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Traversal T;
    TraversalStack st;
    T.idle();
    int64_t my_ray = -1;
    float my_len = 0.0f;
    int64_t next = base;
    for (;;) {
        if (my_ray >= 0 && !T.active) { retire(my_ray, T.hit_tri, T.hit_t, my_len); my_ray = -1; }
        const unsigned idle = __ballot_sync(0xffffffffu, my_ray < 0);
        if (idle && next < end) {
            const int64_t idx = next + __popc(idle & lt_mask);
            if (my_ray < 0 && idx < end) {
                Ray r; float t0, t1;
                my_ray = idx;
                const bool ok = fetch(idx, r, t0, t1, my_len);      // false: nothing to trace (zero-length segment)
                T.begin(S, r, ok, t0, t1);
            }
            next += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, T.active)) {
            if (__all_sync(0xffffffffu, my_ray < 0) && next >= end) break;
            continue;                                               // only retirements / refills pending
        }
        const int2 nd = T.descend(S, st);
        T.leaf<ANY_HIT>(S, st, nd, skip_id, ANY_HIT ? my_len : 0.0f);
    }
}

__global__ void __launch_bounds__(kTraceBlock)
k1_trace_rays(DevScene S, int64_t n, const float* __restrict__ ox, const float* __restrict__ oy,
              const float* __restrict__ oz, const float* __restrict__ dx, const float* __restrict__ dy,
              const float* __restrict__ dz, const float* __restrict__ tmin, const float* __restrict__ tmax,
              int skip_id, int32_t* __restrict__ hit_tri, int32_t* __restrict__ hit_sid,
              float* __restrict__ hit_t, float* __restrict__ normal_soa) {
    const int64_t n_chunks = (n + kRaysPerWarp - 1) / kRaysPerWarp;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += nwarps) {
        const int64_t base = chunk * kRaysPerWarp;
        const int64_t end = base + kRaysPerWarp < n ? base + kRaysPerWarp : n;
        auto fetch = [&](int64_t i, Ray& r, float& t0, float& t1, float& len) {
            r = Ray{__ldcs(&ox[i]), __ldcs(&oy[i]), __ldcs(&oz[i]), __ldcs(&dx[i]), __ldcs(&dy[i]), __ldcs(&dz[i])};   // streamed once
            t0 = tmin ? __ldcs(&tmin[i]) : 0.0f; t1 = __ldcs(&tmax[i]); len = 0.0f;
            return true;
        };
        auto retire = [&](int64_t i, int tri, float t, float) {
            if (hit_tri) hit_tri[i] = tri;
            if (hit_t) hit_t[i] = t;
            if (hit_sid) hit_sid[i] = tri >= 0 ? __float_as_int(__ldg(&S.q2[tri]).z) : -1;
            if (normal_soa) {
                const float4 q = tri >= 0 ? __ldg(&S.q0[tri]) : make_float4(0.f, 0.f, 0.f, 0.f);
                normal_soa[i] = q.x; normal_soa[n + i] = q.y; normal_soa[2 * n + i] = q.z;
            }
        };
        stream_rays<decltype(fetch), decltype(retire), false>(S, base, end, skip_id, fetch, retire);
    }
}

template <bool SKY>
__global__ void __launch_bounds__(kTraceBlock)
k1_test_lines(DevScene S, int64_t n, int64_t stride, const float* __restrict__ a, const float* __restrict__ b,
              uint32_t* __restrict__ bits) {            // a, b: SoA blocks x[stride] y[stride] z[stride]
    __shared__ uint32_t words[kTraceWarps][kRaysPerWarp / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_chunks = (n + kRaysPerWarp - 1) / kRaysPerWarp;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += nwarps) {
        const int64_t base = chunk * kRaysPerWarp;
        const int64_t end = base + kRaysPerWarp < n ? base + kRaysPerWarp : n;
        if (lane < kRaysPerWarp / 32) words[warp][lane] = 0u;
        __syncwarp();
        auto fetch = [&](int64_t i, Ray& r, float& t0, float& t1, float& len) {
            t0 = 0.0f;
            r = Ray{0.f, 0.f, 0.f, 1.f, 1.f, 1.f}; len = 0.0f;
            const bool ok = segment_to_ray(__ldcs(&a[i]), __ldcs(&a[stride + i]), __ldcs(&a[2 * stride + i]),
                                           __ldcs(&b[i]), __ldcs(&b[stride + i]), __ldcs(&b[2 * stride + i]), r, len);   // streamed once
            t1 = len;
            return ok;
        };
        auto retire = [&](int64_t i, int tri, float t, float len) {
            // occlusion rule of raytracer/trace/testline.go:42-51
            bool occluded = tri != -1 && t < len;
            if (SKY && occluded) occluded = (__float_as_int(__ldg(&S.q2[tri]).z) & 0x01000000) == 0;
            if (!occluded) atomicOr(&words[warp][(int)(i - base) >> 5], 1u << ((int)(i - base) & 31));
        };
        stream_rays<decltype(fetch), decltype(retire), !SKY>(S, base, end, -1, fetch, retire);
        __syncwarp();
        const int nw = (int)((end - base + 31) >> 5);
        if (lane < nw) bits[(base >> 5) + lane] = words[warp][lane];
        __syncwarp();
    }
}

// one warp per kRaysPerWarp-ray chunk, capped at 64 resident-block waves per SM (grid-stride beyond that)
static int stream_grid(const vrad_env* e, int64_t n) {
    int64_t blocks = ((n + kRaysPerWarp - 1) / kRaysPerWarp + kTraceWarps - 1) / kTraceWarps;
    const int64_t cap = (int64_t)e->sm_count * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_trace_rays(vrad_env* e, int64_t n, const float* ox, const float* oy, const float* oz, const float* dx,
                      const float* dy, const float* dz, const float* tmin, const float* tmax, int32_t skip_id,
                      int32_t* hit_tri, int32_t* hit_sid, float* hit_t, float* normal_soa) {
    timing_begin(e);
    k1_trace_rays<<<stream_grid(e, n), kTraceBlock, 0, e->stream>>>(
        e->scene, n, ox, oy, oz, dx, dy, dz, tmin, tmax, skip_id, hit_tri, hit_sid, hit_t, normal_soa);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

static void enqueue_test_lines(vrad_env* e, int64_t n, int64_t stride, const float* a, const float* b, int sky_mode, uint32_t* bits) {
    if (sky_mode) k1_test_lines<true><<<stream_grid(e, n), kTraceBlock, 0, e->stream>>>(e->scene, n, stride, a, b, bits);
    else k1_test_lines<false><<<stream_grid(e, n), kTraceBlock, 0, e->stream>>>(e->scene, n, stride, a, b, bits);
}

int launch_test_lines(vrad_env* e, int64_t n, const float* start_soa, const float* stop_soa, int sky_mode, uint32_t* bits) {
    timing_begin(e);
    enqueue_test_lines(e, n, n, start_soa, stop_soa, sky_mode, bits);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Host-buffer path: the segments are copied in chunks on a second stream into two staging buffers
// while the previous chunk is traced, so the call costs max(PCIe, kernel) instead of their sum.
// h_a / h_b are host SoA blocks x[n] y[n] z[n]; d_bits is the device result (n bits).
int launch_test_lines_pipelined(vrad_env* e, int64_t n, const float* h_a, const float* h_b, int sky_mode, uint32_t* d_bits) {
    constexpr int64_t kChunk = (int64_t)1 << 21;       // segments per chunk: 48 MiB of coordinates, multiple of kRaysPerWarp
    for (int s = 0; s < 2; s++)
        if (e->d_stage[s].alloc((size_t)6 * kChunk)) { set_error("out of device memory for staging"); return VRAD_E_NOMEM; }
    timing_begin(e);
    int launches = 0;
    // the copy stream must not overwrite staging that earlier work on the main stream may still read
    VRAD_CUDA_CHECK(cudaEventRecord(e->ev_done[0], e->stream));
    VRAD_CUDA_CHECK(cudaEventRecord(e->ev_done[1], e->stream));
    int c = 0;
    for (int64_t c0 = 0; c0 < n; c0 += kChunk, c++) {
        const int s = c & 1;
        const int64_t m = n - c0 < kChunk ? n - c0 : kChunk;
        float* st = e->d_stage[s].p;
        VRAD_CUDA_CHECK(cudaStreamWaitEvent(e->copy_stream, e->ev_done[s], 0));
        for (int k = 0; k < 3; k++) {
            VRAD_CUDA_CHECK(cudaMemcpyAsync(st + k * kChunk, h_a + k * n + c0, (size_t)m * 4, cudaMemcpyHostToDevice, e->copy_stream));
            VRAD_CUDA_CHECK(cudaMemcpyAsync(st + (3 + k) * kChunk, h_b + k * n + c0, (size_t)m * 4, cudaMemcpyHostToDevice, e->copy_stream));
        }
        VRAD_CUDA_CHECK(cudaEventRecord(e->ev_copied[s], e->copy_stream));
        VRAD_CUDA_CHECK(cudaStreamWaitEvent(e->stream, e->ev_copied[s], 0));
        enqueue_test_lines(e, m, kChunk, st, st + 3 * kChunk, sky_mode, d_bits + (c0 >> 5));
        launches++;
        VRAD_CUDA_CHECK(cudaEventRecord(e->ev_done[s], e->stream));
    }
    timing_end(e, launches);
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace vrad
