// k3_direct.cu -- K3: per-luxel direct lighting with shadow rays.
//
// The reference ports only light creation/parsing (rad/lightmap/lights.go:38-116, 210-341; type
// common/types/light.go:10-44); BuildFacelights/GatherSampleLight are absent.  Semantics follow
// SURVEY.md App. B.2; the sky-ambient sampling pattern is the one visible in
// rad/lightmap/lightmap.go:435-447 (162 Anorm directions, TestLineDoesHitSky per direction).
//
// Work decomposition: one thread per (luxel, light) pair computes that pair's scalar
// `falloff * dot * visibility` (rays generated on the device: 36 B of HBM traffic per luxel for
// n_lights rays); sky-ambient lights use one thread per luxel with the warp walking the sample
// directions together (parallel rays per round); a final pass accumulates RGB per luxel in light
// order, which keeps the sum bit-identical to the sequential CPU formulation.
#include "env_internal.cuh"
#include "sky.cuh"

namespace vrad {

constexpr float kEqualEpsilon = 0.001f;                              // vmath/constants.go:9
// (kMaxTraceLength: sky.cuh; common/constants/constants.go:15-19)

// Visibility of a light ray.  MODE 0: the binary TestLine / sky-id TestLine of the base path (any-hit traversal for
// point lights).  MODE 1 / 2: the complete form of raytracer/trace/testline.go:18-94 selected with
// vrad_set_light_trace_flags -- sky lights go through the 3D-skybox recursion; with MODE 2 (texture shadows)
// every light ray accumulates transparent-triangle coverage and the result is a fraction (App. B.2: dot *= fractionVisible).
template <int MODE>
__device__ __forceinline__ float light_ray_fraction(const DevScene& S, const DevBsp& B, bool ok, float px, float py, float pz,
                                                    float sx, float sy, float sz, bool sky_light, bool can_recurse) {
    if (MODE == 0) return (float)(segment_visible(S, ok, px, py, pz, sx, sy, sz, sky_light ? 1 : 0) && ok);
    if (sky_light) {       // warp-uniform (light type)
        const float fv = MODE == 2 ? sky_fraction<true>(S, B, ok, px, py, pz, sx, sy, sz, can_recurse, px, py, pz, VRAD_TRACE_ID_STATICPROP | -1)
                                   : sky_fraction<false>(S, B, ok, px, py, pz, sx, sy, sz, can_recurse, px, py, pz, VRAD_TRACE_ID_STATICPROP | -1);
        return ok ? fv : 0.0f;
    }
    bool degenerate;
    const float occ = MODE == 2 ? primary_occlusion<true, false>(S, ok, px, py, pz, sx, sy, sz, VRAD_TRACE_ID_STATICPROP | -1, degenerate)
                                : primary_occlusion<false, false>(S, ok, px, py, pz, sx, sy, sz, VRAD_TRACE_ID_STATICPROP | -1, degenerate);
    return ok ? finish_fraction(occ) : 0.0f;
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return ((ax * bx) + (ay * by)) + (az * bz);
}

// Pair index space is light-major and padded to whole warps: a warp holds 32 consecutive luxels of
// ONE light, so its rays share an end point (coherent traversal) and the light type / sky mode is
// warp-uniform.  The shadow test is warp-synchronous: every lane calls segment_visible once.
template <int MODE>
__global__ void __launch_bounds__(128)
k3_pair_scale(DevScene S, DevBsp B, int can_recurse, int64_t n_luxels, int n_lights, const float* __restrict__ pos3,
              const float* __restrict__ nrm3, const vrad_light* __restrict__ lights, float* __restrict__ scale_out) {
    const int64_t n_pad = (n_luxels + 31) & ~(int64_t)31;
    const int64_t total = n_pad * n_lights;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int L = (int)(idx / n_pad);
        const int64_t i = idx - (int64_t)L * n_pad;
        const vrad_light& dl = lights[L];
        const int type = dl.type;
        if (!(type == 0 || type == 1 || type == 2 || type == 3)) continue;      // warp-uniform (sky ambient: k3_sky_ambient)
        const bool in_range = i < n_luxels;
        float px = 0.f, py = 0.f, pz = 0.f, nx = 0.f, ny = 0.f, nz = 1.f;
        if (in_range) {
            px = pos3[3 * i]; py = pos3[3 * i + 1]; pz = pos3[3 * i + 2];
            nx = nrm3[3 * i]; ny = nrm3[3 * i + 1]; nz = nrm3[3 * i + 2];
        }
        float value = 0.0f;                 // falloff * dot before visibility
        float sx = px, sy = py, sz = pz;    // far end of the shadow segment
        bool ok = in_range;
        if (type != 3) {
            float dx = dl.origin[0] - px, dy = dl.origin[1] - py, dz = dl.origin[2] - pz;
            const float dist2 = dot3(dx, dy, dz, dx, dy, dz);
            float dist = sqrtf(dist2);
            ok = ok && dist > 0.0f;
            float dot = 0.0f, falloff = 0.0f;
            if (ok) {
                const float r = 1.0f / dist;
                dx = dx * r; dy = dy * r; dz = dz * r;
                dot = dot3(dx, dy, dz, nx, ny, nz);
                ok = dot > 0.0f;
            }
            if (ok && dl.end_fade > dl.start_fade && dist > dl.end_fade) ok = false;
            if (ok) {
                dist = max_sel(dist, 1.0f);
                const float d = min_sel(dist, dl.cap_dist);
                const float denom = (dl.constant_attn + (dl.linear_attn * d)) + ((dl.quadratic_attn * d) * d);
                if (type == 1) {
                    falloff = 1.0f / denom;
                } else if (type == 2) {
                    const float dot2 = -dot3(dx, dy, dz, dl.normal[0], dl.normal[1], dl.normal[2]);
                    if (dot2 <= dl.stopdot2) ok = false;
                    else {
                        falloff = dot2 / denom;
                        if (dot2 <= dl.stopdot) {
                            float m = (dot2 - dl.stopdot2) / (dl.stopdot - dl.stopdot2);
                            m = min_sel(max_sel(m, 0.0f), 1.0f);
                            if (dl.exponent != 0.0f && dl.exponent != 1.0f) m = powf(m, dl.exponent);
                            falloff = falloff * m;
                        }
                    }
                } else {
                    const float dot2 = -dot3(dx, dy, dz, dl.normal[0], dl.normal[1], dl.normal[2]);
                    if (!(dot2 > 0.0f)) ok = false;
                    else { dot = dot * dot2; falloff = 1.0f / (dist * dist); }
                }
            }
            if (ok && dl.end_fade > dl.start_fade && dist > dl.start_fade) {
                float t = (dist - dl.start_fade) / (dl.end_fade - dl.start_fade);
                t = min_sel(max_sel(t, 0.0f), 1.0f);
                const float s = 1.0f - ((t * t) * (3.0f - (2.0f * t)));
                falloff = falloff * s;
            }
            value = falloff * dot;
            sx = dl.origin[0]; sy = dl.origin[1]; sz = dl.origin[2];
        } else {
            const float dot = -dot3(dl.normal[0], dl.normal[1], dl.normal[2], nx, ny, nz);
            ok = ok && dot > 0.0f;
            value = dot;
            sx = px - (dl.normal[0] * kMaxTraceLength); sy = py - (dl.normal[1] * kMaxTraceLength);
            sz = pz - (dl.normal[2] * kMaxTraceLength);
        }
        if (MODE == 0) {
            const int vis = segment_visible(S, ok, px, py, pz, sx, sy, sz, type == 3 ? 1 : 0);
            if (in_range) scale_out[i * n_lights + L] = (ok && vis) ? value : 0.0f;
        } else {
            const float fv = light_ray_fraction<MODE>(S, B, ok, px, py, pz, sx, sy, sz, type == 3, can_recurse != 0);
            if (in_range) scale_out[i * n_lights + L] = (ok && fv > 0.0f) ? value * fv : 0.0f;
        }
    }
}

// Sky ambient: one thread per luxel, the warp walks the sample directions together -- in every round the 32
// lanes trace PARALLEL rays from neighbouring luxels, which traverse the same nodes (a warp per luxel
// with lanes over the 162 directions fans out over the whole hemisphere and ran at 0.8 G rays/s on the
// 1 M-triangle map against 3.2 G rays/s for the parallel sun rays).  Each lane accumulates its own luxel
// in direction order, i.e. exactly the sequential CPU sum.
template <int MODE>
__global__ void __launch_bounds__(128)
k3_sky_ambient(DevScene S, DevBsp B, int can_recurse, int64_t n_luxels, int n_lights, int light_index, int n_dirs, const float* __restrict__ dirs3,
               const float* __restrict__ pos3, const float* __restrict__ nrm3, float* __restrict__ scale_out) {
    const int64_t n_pad = (n_luxels + 31) & ~(int64_t)31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        const bool in_range = i < n_luxels;
        float px = 0.f, py = 0.f, pz = 0.f, nx = 0.f, ny = 0.f, nz = 1.f;
        if (in_range) {
            px = pos3[3 * i]; py = pos3[3 * i + 1]; pz = pos3[3 * i + 2];
            nx = nrm3[3 * i]; ny = nrm3[3 * i + 1]; nz = nrm3[3 * i + 2];
        }
        float sum = 0.0f, possible = 0.0f;
        for (int k = 0; k < n_dirs; k++) {
            const float ax = __ldg(&dirs3[3 * k]), ay = __ldg(&dirs3[3 * k + 1]), az = __ldg(&dirs3[3 * k + 2]);
            const float dot = dot3(ax, ay, az, nx, ny, nz);
            const bool want = in_range && dot > kEqualEpsilon;
            if (!__any_sync(0xffffffffu, want)) continue;
            if (want) possible = possible + dot;
            if (MODE == 0) {
                const int vis = segment_visible(S, want, px, py, pz, px + (ax * kMaxTraceLength), py + (ay * kMaxTraceLength),
                                                pz + (az * kMaxTraceLength), 1);
                if (want && vis) sum = sum + dot;
            } else {
                const float fv = light_ray_fraction<MODE>(S, B, want, px, py, pz, px + (ax * kMaxTraceLength), py + (ay * kMaxTraceLength),
                                                          pz + (az * kMaxTraceLength), true, can_recurse != 0);
                if (want && fv > 0.0f) sum = sum + (dot * fv);
            }
        }
        if (in_range) scale_out[i * n_lights + light_index] = possible > 0.0f ? sum / possible : 0.0f;
    }
}

__global__ void k3_accumulate(int64_t n_luxels, int n_lights, const vrad_light* __restrict__ lights,
                              const float* __restrict__ scale, float* __restrict__ rgb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_luxels) return;
    float r = 0.f, g = 0.f, b = 0.f;
    for (int L = 0; L < n_lights; L++) {
        const float s = scale[i * n_lights + L];
        const int type = lights[L].type;
        const bool known = type == 0 || type == 1 || type == 2 || type == 3 || type == 5;
        if (!known || s == 0.0f) continue;       // adding +0 would not change the sum
        r = r + (lights[L].intensity[0] * s);
        g = g + (lights[L].intensity[1] * s);
        b = b + (lights[L].intensity[2] * s);
    }
    rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
}

} // namespace vrad
using namespace vrad;

extern "C" {

int vrad_set_sky_dirs(vrad_env* e, int n, const float* dirs3) {
    VRAD_MULTI(e, group_set_sky_dirs(e, n, dirs3));
    if (!e || n < 0 || (n > 0 && !dirs3)) { set_error("vrad_set_sky_dirs: bad arguments"); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    if (e->d_sky_dirs.alloc(3 * (size_t)n + 1)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    if (n) VRAD_CUDA_CHECK(cudaMemcpy(e->d_sky_dirs.p, dirs3, 12 * (size_t)n, cudaMemcpyHostToDevice));
    e->n_sky_dirs = n;
    return VRAD_OK;
}

int vrad_set_light_trace_flags(vrad_env* e, int flags) {
    VRAD_MULTI(e, group_set_light_trace_flags(e, flags));
    if (!e || (flags & ~(VRAD_TL_CAN_RECURSE | VRAD_TL_TEXTURE_SHADOWS))) { set_error("vrad_set_light_trace_flags: bad arguments"); return VRAD_E_INVALID; }
    e->light_trace_flags = flags;
    return VRAD_OK;
}

int vrad_direct_light(vrad_env* e, int64_t n_luxels, const float* pos3, const float* normal3, int n_lights,
                      const vrad_light* lights, float* rgb_out) {
    VRAD_MULTI(e, group_direct_light(e, n_luxels, pos3, normal3, n_lights, lights, rgb_out));
    if (!e || n_luxels < 0 || n_lights < 0 || (n_luxels > 0 && (!pos3 || !normal3 || !rgb_out)) || (n_lights > 0 && !lights)) {
        set_error("vrad_direct_light: bad arguments"); return VRAD_E_INVALID;
    }
    if (!e->built) { set_error("vrad_direct_light: acceleration structure not built"); return VRAD_E_STATE; }
    if (n_luxels == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    // the light table is tiny and host-resident in every caller: copy it, and find sky-ambient entries
    std::vector<vrad_light> hl(n_lights ? n_lights : 1);
    if (n_lights) {
        if (is_device_ptr(lights)) VRAD_CUDA_CHECK(cudaMemcpy(hl.data(), lights, sizeof(vrad_light) * n_lights, cudaMemcpyDeviceToHost));
        else memcpy(hl.data(), lights, sizeof(vrad_light) * n_lights);
    }
    for (int L = 0; L < n_lights; L++)
        if (hl[L].type == 5 && e->n_sky_dirs == 0) { set_error("vrad_direct_light: sky-ambient light but vrad_set_sky_dirs was not called"); return VRAD_E_STATE; }
    int rc; const void *d_pos, *d_nrm, *d_l; bool h0, h1, h2, ho;
    if ((rc = stage_in(e, 0, pos3, (size_t)n_luxels * 12, &d_pos, &h0))) return rc;
    if ((rc = stage_in(e, 1, normal3, (size_t)n_luxels * 12, &d_nrm, &h1))) return rc;
    if ((rc = stage_in(e, 2, hl.data(), sizeof(vrad_light) * hl.size(), &d_l, &h2))) return rc;
    void* d_rgb;
    if ((rc = stage_out(e, 3, rgb_out, (size_t)n_luxels * 12, &d_rgb, &ho))) return rc;
    void* d_scale;                       // per-(luxel, light) scalar falloff*dot*visibility
    const size_t nscale = (size_t)n_luxels * (n_lights ? n_lights : 1);
    if ((rc = scratch_get(e, 4, nscale * 4, &d_scale))) return rc;
    timing_begin(e);
    int launches = 0;
    if (n_lights) {
        VRAD_CUDA_CHECK(cudaMemsetAsync(d_scale, 0, nscale * 4, e->stream));
        const int64_t total = ((n_luxels + 31) & ~(int64_t)31) * n_lights;
        int64_t blocks = (total + 127) / 128, cap = (int64_t)e->sm_count * 64;
        // light_trace_flags (vrad_set_light_trace_flags): 0 = base path; recursion and/or texture shadows = complete TestLineDoesHitSky
        const int mode = (e->light_trace_flags & VRAD_TL_TEXTURE_SHADOWS) ? 2 : ((e->light_trace_flags & VRAD_TL_CAN_RECURSE) ? 1 : 0);
        const int rec = (e->light_trace_flags & VRAD_TL_CAN_RECURSE) && e->bsp_ready ? 1 : 0;
        const int grid1 = (int)(blocks < cap ? blocks : cap);
#define K3_ARGS e->scene, e->bsp, rec, n_luxels, n_lights, (const float*)d_pos, (const float*)d_nrm, (const vrad_light*)d_l, (float*)d_scale
        if (mode == 0) k3_pair_scale<0><<<grid1, 128, 0, e->stream>>>(K3_ARGS);
        else if (mode == 1) k3_pair_scale<1><<<grid1, 128, 0, e->stream>>>(K3_ARGS);
        else k3_pair_scale<2><<<grid1, 128, 0, e->stream>>>(K3_ARGS);
#undef K3_ARGS
        launches++;
        for (int L = 0; L < n_lights; L++) {
            if (hl[L].type != 5) continue;
            int64_t b2 = (n_luxels + 127) / 128;
            const int grid2 = (int)(b2 < cap ? b2 : cap);
#define K3_ARGS e->scene, e->bsp, rec, n_luxels, n_lights, L, e->n_sky_dirs, e->d_sky_dirs.p, (const float*)d_pos, (const float*)d_nrm, (float*)d_scale
            if (mode == 0) k3_sky_ambient<0><<<grid2, 128, 0, e->stream>>>(K3_ARGS);
            else if (mode == 1) k3_sky_ambient<1><<<grid2, 128, 0, e->stream>>>(K3_ARGS);
            else k3_sky_ambient<2><<<grid2, 128, 0, e->stream>>>(K3_ARGS);
#undef K3_ARGS
            launches++;
        }
    }
    k3_accumulate<<<(int)((n_luxels + 255) / 256), 256, 0, e->stream>>>(n_luxels, n_lights, (const vrad_light*)d_l, (const float*)d_scale, (float*)d_rgb);
    launches++;
    timing_end(e, launches);
    VRAD_CUDA_CHECK(cudaGetLastError());
    if ((rc = finish_out(e, rgb_out, d_rgb, (size_t)n_luxels * 12, ho))) return rc;
    return sync_if_needed(e, true);   // the host light table above must outlive the H2D copy
}

} // extern "C"
