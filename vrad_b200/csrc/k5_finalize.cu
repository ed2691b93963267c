// k5_finalize.cu -- K5: final light per luxel -> ColorRGBExp32 ("lightmap finalisation / write-back").
//
// Upstream FinalLightFace (UNCITED; the reference's finish task is a stub, cmd/tasks/finish/main.go:15-18): for every
// luxel the direct light (K3) plus the bounced light sampled for it, negative components clamped to zero, packed by
// VectorToColorRGBExp32 (rgbexp.cuh -- the same inline function the host entry point vrad_color_to_rgbexp32 runs, which
// is what the CPU tests pin).
//
// Roofline: pure streaming, HBM-bound: 12 (+12 with indirect) bytes in and 4 bytes out per luxel, no reuse.  One thread
// per luxel with scalar 4-byte loads would issue three strided requests per array; instead a thread owns FOUR
// consecutive luxels = 48 contiguous bytes per input = three 128-bit loads (the arrays come from cudaMalloc / the
// staging buffers, 256-byte aligned, and 4 luxels x 12 B keeps every thread 16-byte aligned) and one 128-bit store of
// four packed colours.  Grid = ceil(n / 4 / 256) blocks; at C5 (2.0M luxels) that is 1,958 blocks = 13 waves of 148 SMs.
#include "env_internal.cuh"
#include "rgbexp.cuh"
#include "../../include/vrad_bsp.h"

namespace vrad {

__global__ void __launch_bounds__(256)
k5_finalize(int64_t n, const float* __restrict__ direct, const float* __restrict__ indirect, uint32_t* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // group of 4 luxels
    const int64_t first = 4 * q;
    if (first >= n) return;
    float v[12];
    if (first + 4 <= n) {
        const float4* d4 = reinterpret_cast<const float4*>(direct + 3 * first);
        const float4 a = __ldcs(d4), b = __ldcs(d4 + 1), c = __ldcs(d4 + 2);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
        if (indirect) {
            const float4* i4 = reinterpret_cast<const float4*>(indirect + 3 * first);
            const float4 x = __ldcs(i4), y = __ldcs(i4 + 1), z = __ldcs(i4 + 2);
            v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w; v[8] += z.x; v[9] += z.y; v[10] += z.z; v[11] += z.w;
        }
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const RgbExp c4 = pack_rgbexp32(v[3 * k], v[3 * k + 1], v[3 * k + 2]);
            w[k] = (uint32_t)c4.r | ((uint32_t)c4.g << 8) | ((uint32_t)c4.b << 16) | ((uint32_t)(uint8_t)c4.e << 24);
        }
        __stcs(reinterpret_cast<uint4*>(out + first), make_uint4(w[0], w[1], w[2], w[3]));
    } else {                                                              // ragged tail: at most 3 luxels, scalar
        for (int64_t i = first; i < n; i++) {
            float r = direct[3 * i], g = direct[3 * i + 1], b = direct[3 * i + 2];
            if (indirect) { r += indirect[3 * i]; g += indirect[3 * i + 1]; b += indirect[3 * i + 2]; }
            const RgbExp c4 = pack_rgbexp32(r, g, b);
            out[i] = (uint32_t)c4.r | ((uint32_t)c4.g << 8) | ((uint32_t)c4.b << 16) | ((uint32_t)(uint8_t)c4.e << 24);
        }
    }
}

// Same, with the bounced light taken from the patch each luxel belongs to (luxel_patch[i] = index into the N x 3 patch totals
// vrad_bounce returns, -1 = none): 4 index loads as one int4, 4 scattered 12-byte reads that mostly hit L2 (neighbouring
// luxels share a patch; a patch covers (chop/1)^2 = 16+ luxels), the rest as above.
__global__ void __launch_bounds__(256)
k5_finalize_patches(int64_t n, const float* __restrict__ direct, const int32_t* __restrict__ luxel_patch, const float* __restrict__ patch_total,
                    uint32_t* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t first = 4 * q;
    if (first >= n) return;
    const int count = (int)(n - first < 4 ? n - first : 4);
    uint32_t w[4] = {0, 0, 0, 0};
    int32_t idx[4] = {-1, -1, -1, -1};
    float v[12];
    if (count == 4) {
        const int4 i4 = __ldcs(reinterpret_cast<const int4*>(luxel_patch + first));
        idx[0] = i4.x; idx[1] = i4.y; idx[2] = i4.z; idx[3] = i4.w;
        const float4* d4 = reinterpret_cast<const float4*>(direct + 3 * first);
        const float4 a = __ldcs(d4), b = __ldcs(d4 + 1), c = __ldcs(d4 + 2);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
    } else {
        for (int k = 0; k < count; k++) {
            idx[k] = luxel_patch[first + k];
            v[3 * k] = direct[3 * (first + k)]; v[3 * k + 1] = direct[3 * (first + k) + 1]; v[3 * k + 2] = direct[3 * (first + k) + 2];
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k >= count) break;
        float r = v[3 * k], g = v[3 * k + 1], b = v[3 * k + 2];
        if (idx[k] >= 0) {
            const float* t = patch_total + 3 * (int64_t)idx[k];
            r += __ldg(t); g += __ldg(t + 1); b += __ldg(t + 2);
        }
        const RgbExp c4 = pack_rgbexp32(r, g, b);
        w[k] = (uint32_t)c4.r | ((uint32_t)c4.g << 8) | ((uint32_t)c4.b << 16) | ((uint32_t)(uint8_t)c4.e << 24);
    }
    if (count == 4) __stcs(reinterpret_cast<uint4*>(out + first), make_uint4(w[0], w[1], w[2], w[3]));
    else for (int k = 0; k < count; k++) out[first + k] = w[k];
}

}  // namespace vrad

using namespace vrad;

extern "C" int vrad_lightmap_finalize_patches(vrad_env* e, int64_t n, const float* direct3, const int32_t* luxel_patch, int n_patches,
                                              const float* patch_total3, vrad_color_rgbexp32* out) {
    if (!e || n < 0 || n_patches < 0 || (n > 0 && (!direct3 || !luxel_patch || !out)) || (n_patches > 0 && !patch_total3)) {
        set_error("vrad_lightmap_finalize_patches: bad arguments"); return VRAD_E_INVALID;
    }
    if (n == 0) return VRAD_OK;
    if (!is_device_ptr(luxel_patch))                               // indices from the host are checked; device indices are the caller's promise
        for (int64_t i = 0; i < n; i++)
            if (luxel_patch[i] >= n_patches) { set_error("vrad_lightmap_finalize_patches: luxel %lld names patch %d of %d", (long long)i, luxel_patch[i], n_patches); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const void *d_dir, *d_idx, *d_tot = nullptr; void* d_out; bool h0, h1, h2 = false, ho;
    int rc;
    // host buffers land in cudaMalloc'd scratch (256-byte aligned); device pointers must be 16-byte aligned for the 128-bit path
    if ((rc = stage_in(e, 0, direct3, (size_t)n * 12, &d_dir, &h0))) return rc;
    if ((rc = stage_in(e, 1, luxel_patch, (size_t)n * 4, &d_idx, &h1))) return rc;
    if (n_patches && (rc = stage_in(e, 2, patch_total3, (size_t)n_patches * 12, &d_tot, &h2))) return rc;
    if ((rc = stage_out(e, 3, out, (size_t)n * 4, &d_out, &ho))) return rc;
    if ((((uintptr_t)d_dir | (uintptr_t)d_idx | (uintptr_t)d_out) & 15) != 0) { set_error("vrad_lightmap_finalize_patches: device pointers must be 16-byte aligned"); return VRAD_E_INVALID; }
    timing_begin(e);
    const int64_t groups = (n + 3) / 4;
    k5_finalize_patches<<<(unsigned)((groups + 255) / 256), 256, 0, e->stream>>>(n, (const float*)d_dir, (const int32_t*)d_idx, (const float*)d_tot, (uint32_t*)d_out);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    if ((rc = finish_out(e, out, d_out, (size_t)n * 4, ho))) return rc;
    return sync_if_needed(e, h0 | h1 | h2 | ho);
}

extern "C" int vrad_lightmap_finalize(vrad_env* e, int64_t n, const float* direct3, const float* indirect3, vrad_color_rgbexp32* out) {
    if (!e || n < 0 || (n > 0 && (!direct3 || !out))) { set_error("vrad_lightmap_finalize: bad arguments"); return VRAD_E_INVALID; }
    if (n == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const void *d_dir, *d_ind = nullptr; void* d_out; bool h0, h1 = false, ho;
    int rc;
    if ((rc = stage_in(e, 0, direct3, (size_t)n * 12, &d_dir, &h0))) return rc;
    if (indirect3 && (rc = stage_in(e, 1, indirect3, (size_t)n * 12, &d_ind, &h1))) return rc;
    if ((rc = stage_out(e, 2, out, (size_t)n * 4, &d_out, &ho))) return rc;
    // the 128-bit path needs 16-byte aligned bases; a caller's device pointer at an odd offset goes through the tail loop
    const bool aligned = (((uintptr_t)d_dir | (uintptr_t)d_ind | (uintptr_t)d_out) & 15) == 0;
    timing_begin(e);
    if (aligned) {
        const int64_t groups = (n + 3) / 4;
        k5_finalize<<<(unsigned)((groups + 255) / 256), 256, 0, e->stream>>>(n, (const float*)d_dir, (const float*)d_ind, (uint32_t*)d_out);
    } else {
        // unaligned device pointers: stage through aligned scratch copies
        void *s0, *s1 = nullptr, *s2;
        if ((rc = scratch_get(e, 3, (size_t)n * 12, &s0))) return rc;
        VRAD_CUDA_CHECK(cudaMemcpyAsync(s0, d_dir, (size_t)n * 12, cudaMemcpyDeviceToDevice, e->stream));
        if (d_ind) {
            if ((rc = scratch_get(e, 4, (size_t)n * 12, &s1))) return rc;
            VRAD_CUDA_CHECK(cudaMemcpyAsync(s1, d_ind, (size_t)n * 12, cudaMemcpyDeviceToDevice, e->stream));
        }
        if ((rc = scratch_get(e, 5, (size_t)n * 4, &s2))) return rc;
        const int64_t groups = (n + 3) / 4;
        k5_finalize<<<(unsigned)((groups + 255) / 256), 256, 0, e->stream>>>(n, (const float*)s0, (const float*)s1, (uint32_t*)s2);
        VRAD_CUDA_CHECK(cudaMemcpyAsync(d_out, s2, (size_t)n * 4, cudaMemcpyDeviceToDevice, e->stream));
    }
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    if ((rc = finish_out(e, out, d_out, (size_t)n * 4, ho))) return rc;
    return sync_if_needed(e, h0 | h1 | ho);
}
