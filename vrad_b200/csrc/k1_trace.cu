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

__global__ void __launch_bounds__(kTraceBlock)
k1_trace_rays(DevScene S, int64_t n, const float* __restrict__ ox, const float* __restrict__ oy,
              const float* __restrict__ oz, const float* __restrict__ dx, const float* __restrict__ dy,
              const float* __restrict__ dz, const float* __restrict__ tmin, const float* __restrict__ tmax,
              int skip_id, int32_t* __restrict__ hit_tri, int32_t* __restrict__ hit_sid,
              float* __restrict__ hit_t, float* __restrict__ normal_soa) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + 31) & ~(int64_t)31;          // whole warps enter the traversal together
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        const bool valid = i < n;
        Ray r{0.f, 0.f, 0.f, 1.f, 1.f, 1.f};
        float t0 = 0.0f, t1 = 0.0f;
        if (valid) {
            r = Ray{ox[i], oy[i], oz[i], dx[i], dy[i], dz[i]};
            t0 = tmin ? tmin[i] : 0.0f; t1 = tmax[i];
        }
        int tri; float t;
        trace_ray<false>(S, r, valid, t0, t1, skip_id, 0.0f, tri, t);
        if (!valid) continue;
        if (hit_tri) hit_tri[i] = tri;
        if (hit_t) hit_t[i] = t;
        if (hit_sid) hit_sid[i] = tri >= 0 ? __float_as_int(__ldg(&S.q2[tri]).z) : -1;
        if (normal_soa) {
            float4 q = tri >= 0 ? __ldg(&S.q0[tri]) : make_float4(0.f, 0.f, 0.f, 0.f);
            normal_soa[i] = q.x; normal_soa[n + i] = q.y; normal_soa[2 * n + i] = q.z;
        }
    }
}

__global__ void __launch_bounds__(kTraceBlock)
k1_test_lines(DevScene S, int64_t n, const float* __restrict__ a, const float* __restrict__ b, int sky_mode,
              uint32_t* __restrict__ bits) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        const bool valid = i < n;
        float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        if (valid) { ax = a[i]; ay = a[n + i]; az = a[2 * n + i]; bx = b[i]; by = b[n + i]; bz = b[2 * n + i]; }
        const int vis = segment_visible(S, valid, ax, ay, az, bx, by, bz, sky_mode) && valid;
        const uint32_t m = __ballot_sync(0xffffffffu, vis);
        if ((threadIdx.x & 31) == 0) bits[i >> 5] = m;
    }
}

static int grid_for(const vrad_env* e, int64_t n, int block, int waves_per_sm) {
    int64_t blocks = (n + block - 1) / block;
    int64_t cap = (int64_t)e->sm_count * waves_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_trace_rays(vrad_env* e, int64_t n, const float* ox, const float* oy, const float* oz, const float* dx,
                      const float* dy, const float* dz, const float* tmin, const float* tmax, int32_t skip_id,
                      int32_t* hit_tri, int32_t* hit_sid, float* hit_t, float* normal_soa) {
    timing_begin(e);
    k1_trace_rays<<<grid_for(e, n, kTraceBlock, 64), kTraceBlock, 0, e->stream>>>(
        e->scene, n, ox, oy, oz, dx, dy, dz, tmin, tmax, skip_id, hit_tri, hit_sid, hit_t, normal_soa);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_test_lines(vrad_env* e, int64_t n, const float* start_soa, const float* stop_soa, int sky_mode, uint32_t* bits) {
    timing_begin(e);
    k1_test_lines<<<grid_for(e, n, kTraceBlock, 64), kTraceBlock, 0, e->stream>>>(e->scene, n, start_soa, stop_soa, sky_mode, bits);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace vrad
