// radial.cu -- bounced light per luxel: the patch lights of a face and of its smoothing neighbours filtered onto the face's luxel
// grid (lightmap finalisation).  UNCITED: this restates, from recollection of Source SDK 2013
// (utils/vrad/radial.cpp: BuildPatchRadial, PatchLightmapCoordRange, AddBouncedToRadial, SampleRadial), a stage the reference does
// not have (its finish task only logs, cmd/tasks/finish/main.go:15-18); treat it like SURVEY App. B -- design intent, pinned by the
// oracle restatement (oracle/bspside.py) and by properties, not by the reference.  What the reference does supply: the face
// neighbour lists (PairEdges, rad/lightmap/lightmap.go:37-216) and the lightmap vectors / extents (rad/world/face.go:14-90).
//
//   for every leaf patch p of face f or of a neighbour of f:   (cs, ct) = luxel-space position of p's origin on f,
//                                                              (ds, dt) = extent of p's winding in f's luxel space, at least 1
//   luxel (s, t) of f:   r = 2 - ((cs - s)^2 / ds^2 + (ct - t)^2 / dt^2);   r > 0:  light += r * TotalLight(p),  weight += r
//   indirect(s, t) = light / weight   (0 when no patch reaches the luxel)
//
// Upstream scatters (patch -> luxels in reach, into a per-face accumulator); here it is a gather: one thread per luxel walks the
// entries of its face in a fixed order, so there are no atomics and the sum does not depend on scheduling.  A face has tens of leaf
// patches and ~1,000 luxels, all threads of a block read the same 20-byte entries (broadcast from L1) and 12-byte patch totals; per
// luxel the traffic is 12 B out plus a share of the entries -- the kernel is latency / L1-bound, not HBM-bound.  The functor runs
// under two policies (a CUDA kernel, OpenMP on the host) like kd_fast.cu; the CPU tests exercise the host policy.
#include <cmath>
#include <cstring>
#include <vector>
#include "env_internal.cuh"
#include "../../include/vrad_bsp.h"

namespace vrad {
namespace radial {

#if defined(__CUDACC__)
#define RD_HD __host__ __device__ __forceinline__
#else
#define RD_HD inline
#endif

constexpr float kRadialDist2 = 2.0f;        // RADIALDIST2: the filter reaches sqrt(2) patch widths
constexpr float kWeightEps = 0.00001f;      // WEIGHT_EPS

struct Gather {
    const int32_t* luxel_face; const int64_t* luxel_first; const int32_t* size2; const int64_t* entry_first; const vrad_radial_entry* entries;
    const float* patch_total3; const float* patch_bump9;      // bump9 may be null: every block uses the flat totals
    float* out3;
    RD_HD void operator()(int64_t l) const {
        const int f = luxel_face[l];
        float acc[3] = {0.0f, 0.0f, 0.0f}, wsum = 0.0f;
        if (f >= 0) {
            const int w = size2[2 * (int64_t)f] + 1, h = size2[2 * (int64_t)f + 1] + 1;
            const int64_t k = l - luxel_first[f];
            const int64_t per_block = (int64_t)w * h;
            const int block = (int)(k / per_block);               // 0 = flat, 1..3 = bump basis (SURF_BUMPLIGHT faces)
            const int64_t within = k - block * per_block;
            const float s = (float)(within % w), t = (float)(within / w);
            for (int64_t e = entry_first[f]; e < entry_first[f + 1]; e++) {
                const vrad_radial_entry en = entries[e];
                const float ds = (en.s - s) * en.inv_ds, dt = (en.t - t) * en.inv_dt;
                const float r = kRadialDist2 - ((ds * ds) + (dt * dt));
                if (r > 0.0f) {
                    const float* v = (block > 0 && patch_bump9) ? patch_bump9 + 9 * (int64_t)en.patch + 3 * (block - 1) : patch_total3 + 3 * (int64_t)en.patch;
                    acc[0] = acc[0] + (v[0] * r); acc[1] = acc[1] + (v[1] * r); acc[2] = acc[2] + (v[2] * r);
                    wsum = wsum + r;
                }
            }
        }
        if (wsum > kWeightEps) { const float inv = 1.0f / wsum; acc[0] = acc[0] * inv; acc[1] = acc[1] * inv; acc[2] = acc[2] * inv; }
        else { acc[0] = acc[1] = acc[2] = 0.0f; }
        out3[3 * l] = acc[0]; out3[3 * l + 1] = acc[1]; out3[3 * l + 2] = acc[2];
    }
};

template <class F> __global__ void __launch_bounds__(256) k_radial(int64_t n, F f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}

}  // namespace radial
}  // namespace vrad

using namespace vrad;

// ---- host: which patches light which face, in that face's luxel space --------------------------------------------------------
extern "C" int vrad_bsp_radial_entries(const vrad_bsp_lumps* L, const int32_t* mins2, const float* face_origins3,
                                       int n_patches, const int32_t* patch_face, const int32_t* child1, const float* origin3,
                                       const int32_t* wind_first, const int32_t* wind_count, const float* wind_points3,
                                       const int32_t* neighbour_first, const int32_t* neighbours,
                                       int64_t max_entries, int64_t* entry_first, vrad_radial_entry* entries, int64_t* n_entries_out) {
    if (!L || !mins2 || n_patches < 0 || (n_patches && (!patch_face || !origin3 || !wind_first || !wind_count || !wind_points3)) || !entry_first || !n_entries_out ||
        (neighbours && !neighbour_first)) { set_error("vrad_bsp_radial_entries: bad arguments"); return VRAD_E_INVALID; }
    const int nf = L->n_faces;
    // leaf patches per face (ascending patch index)
    std::vector<int64_t> pf_first((size_t)nf + 1, 0);
    for (int p = 0; p < n_patches; p++) {
        if (patch_face[p] < 0 || patch_face[p] >= nf) { set_error("vrad_bsp_radial_entries: patch %d names face %d of %d", p, patch_face[p], nf); return VRAD_E_INVALID; }
        if (!child1 || child1[p] == -1) pf_first[patch_face[p] + 1]++;
    }
    for (int f = 0; f < nf; f++) pf_first[f + 1] += pf_first[f];
    std::vector<int32_t> pf((size_t)pf_first[nf]);
    std::vector<int64_t> fill(pf_first.begin(), pf_first.end() - 1);
    for (int p = 0; p < n_patches; p++) if (!child1 || child1[p] == -1) pf[fill[patch_face[p]]++] = p;

    int64_t n = 0;
    const bool store = entries != nullptr;
    for (int f = 0; f < nf; f++) {
        entry_first[f] = n;
        const vrad_texinfo& tx = L->texinfo[L->faces[f].texinfo];
        if (tx.flags & (VRAD_SURF_SKY | VRAD_SURF_NOLIGHT)) continue;       // no lightmap on this face
        const float* off = face_origins3 ? face_origins3 + 3 * (size_t)f : nullptr;
        auto to_luxel = [&](const float* p, float st[2]) {                  // WorldToLuxelSpace, relative to the face's lightmap mins
            const float x = off ? p[0] - off[0] : p[0], y = off ? p[1] - off[1] : p[1], z = off ? p[2] - off[2] : p[2];
            for (int k = 0; k < 2; k++) {
                const float* lv = tx.lightmap_vecs[k];
                st[k] = ((((x * lv[0]) + (y * lv[1])) + (z * lv[2])) + lv[3]) - (float)mins2[2 * (size_t)f + k];
            }
        };
        auto add_face_patches = [&](int src) {
            for (int64_t q = pf_first[src]; q < pf_first[src + 1]; q++) {
                const int p = pf[q];
                if (store && n < max_entries) {
                    float mn[2] = {1e30f, 1e30f}, mx[2] = {-1e30f, -1e30f}, st[2];
                    for (int k = 0; k < wind_count[p]; k++) {                // PatchLightmapCoordRange
                        to_luxel(wind_points3 + 3 * (size_t)(wind_first[p] + k), st);
                        for (int a = 0; a < 2; a++) { if (st[a] < mn[a]) mn[a] = st[a]; if (st[a] > mx[a]) mx[a] = st[a]; }
                    }
                    to_luxel(origin3 + 3 * (size_t)p, st);
                    float ds = mx[0] - mn[0], dt = mx[1] - mn[1];
                    if (!(ds > 1.0f)) ds = 1.0f;                             // patches smaller than a luxel would be filtered away
                    if (!(dt > 1.0f)) dt = 1.0f;
                    vrad_radial_entry e = {p, st[0], st[1], 1.0f / ds, 1.0f / dt};
                    entries[n] = e;
                }
                n++;
            }
        };
        add_face_patches(f);
        if (neighbours)
            for (int32_t k = neighbour_first[f]; k < neighbour_first[f + 1]; k++) {
                const int o = neighbours[k];
                if (o < 0 || o >= nf) { set_error("vrad_bsp_radial_entries: face %d has neighbour %d of %d", f, o, nf); return VRAD_E_INVALID; }
                add_face_patches(o);
            }
    }
    entry_first[nf] = n;
    *n_entries_out = n;
    if (store && n > max_entries) { set_error("vrad_bsp_radial_entries: %lld entries, room for %lld", (long long)n, (long long)max_entries); return VRAD_E_NOMEM; }
    return VRAD_OK;
}

static int check_radial_args(const char* who, int64_t n, const int32_t* luxel_face, int n_faces, const int64_t* luxel_first, const int32_t* size2,
                             const int64_t* entry_first, const vrad_radial_entry* entries, int n_patches, const float* patch_total3, float* out) {
    if (n < 0 || n_faces < 0 || n_patches < 0 || (n && (!luxel_face || !out)) || (n_faces && (!luxel_first || !size2 || !entry_first)) ||
        (n_patches && !patch_total3)) { set_error("%s: bad arguments", who); return VRAD_E_INVALID; }
    (void)entries;
    return VRAD_OK;
}

// the host policy (CPU tests; also usable when the totals are on the host anyway)
extern "C" int vrad_luxel_radial_light_host(int64_t n, const int32_t* luxel_face, int n_faces, const int64_t* luxel_first, const int32_t* size2,
                                            const int64_t* entry_first, const vrad_radial_entry* entries, int n_patches, const float* patch_total3,
                                            const float* patch_bump9, float* indirect3_out) {
    int rc = check_radial_args("vrad_luxel_radial_light_host", n, luxel_face, n_faces, luxel_first, size2, entry_first, entries, n_patches, patch_total3, indirect3_out);
    if (rc) return rc;
    for (int64_t l = 0; l < n; l++) if (luxel_face[l] >= n_faces) { set_error("vrad_luxel_radial_light_host: luxel %lld names face %d of %d", (long long)l, luxel_face[l], n_faces); return VRAD_E_INVALID; }
    const int64_t ne = n_faces ? entry_first[n_faces] : 0;
    for (int64_t e = 0; e < ne; e++) if (entries[e].patch < 0 || entries[e].patch >= n_patches) { set_error("vrad_luxel_radial_light_host: entry %lld names patch %d of %d", (long long)e, entries[e].patch, n_patches); return VRAD_E_INVALID; }
    radial::Gather g{luxel_face, luxel_first, size2, entry_first, entries, patch_total3, patch_bump9, indirect3_out};
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < n; l++) g(l);
    return VRAD_OK;
}

// the device policy: all arrays host or device memory (host arrays are staged)
extern "C" int vrad_luxel_radial_light(vrad_env* e, int64_t n, const int32_t* luxel_face, int n_faces, const int64_t* luxel_first, const int32_t* size2,
                                       const int64_t* entry_first, const vrad_radial_entry* entries, int n_patches, const float* patch_total3,
                                       const float* patch_bump9, float* indirect3_out) {
    if (!e) { set_error("vrad_luxel_radial_light: bad arguments"); return VRAD_E_INVALID; }
    int rc = check_radial_args("vrad_luxel_radial_light", n, luxel_face, n_faces, luxel_first, size2, entry_first, entries, n_patches, patch_total3, indirect3_out);
    if (rc) return rc;
    if (n == 0) return VRAD_OK;
    if (is_device_ptr(entry_first) || is_device_ptr(luxel_first)) { set_error("vrad_luxel_radial_light: the per-face offset tables must be host memory"); return VRAD_E_INVALID; }
    const int64_t ne = n_faces ? entry_first[n_faces] : 0;
    if (!is_device_ptr(entries)) for (int64_t k = 0; k < ne; k++) if (entries[k].patch < 0 || entries[k].patch >= n_patches) { set_error("vrad_luxel_radial_light: entry %lld names patch %d of %d", (long long)k, entries[k].patch, n_patches); return VRAD_E_INVALID; }
    if (!is_device_ptr(luxel_face)) for (int64_t l = 0; l < n; l++) if (luxel_face[l] >= n_faces) { set_error("vrad_luxel_radial_light: luxel %lld names face %d of %d", (long long)l, luxel_face[l], n_faces); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const void *d_lf, *d_first, *d_size, *d_efirst, *d_ent, *d_tot, *d_bump = nullptr; void* d_out; bool h[7] = {false, false, false, false, false, false, false}, ho;
    if ((rc = stage_in(e, 0, luxel_face, (size_t)n * 4, &d_lf, &h[0]))) return rc;
    if ((rc = stage_in(e, 1, luxel_first, ((size_t)n_faces + 1) * 8, &d_first, &h[1]))) return rc;
    if ((rc = stage_in(e, 2, size2, (size_t)n_faces * 8, &d_size, &h[2]))) return rc;
    if ((rc = stage_in(e, 3, entry_first, ((size_t)n_faces + 1) * 8, &d_efirst, &h[3]))) return rc;
    if ((rc = stage_in(e, 4, entries, (size_t)(ne ? ne : 1) * sizeof(vrad_radial_entry), &d_ent, &h[4]))) return rc;
    if ((rc = stage_in(e, 5, patch_total3, (size_t)(n_patches ? n_patches : 1) * 12, &d_tot, &h[5]))) return rc;
    if (patch_bump9 && (rc = stage_in(e, 6, patch_bump9, (size_t)(n_patches ? n_patches : 1) * 36, &d_bump, &h[6]))) return rc;
    if ((rc = stage_out(e, 7, indirect3_out, (size_t)n * 12, &d_out, &ho))) return rc;
    radial::Gather g{(const int32_t*)d_lf, (const int64_t*)d_first, (const int32_t*)d_size, (const int64_t*)d_efirst, (const vrad_radial_entry*)d_ent,
                     (const float*)d_tot, (const float*)d_bump, (float*)d_out};
    timing_begin(e);
    radial::k_radial<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(n, g);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    if ((rc = finish_out(e, indirect3_out, d_out, (size_t)n * 12, ho))) return rc;
    return sync_if_needed(e, h[0] | h[1] | h[2] | h[3] | h[4] | h[5] | h[6] | ho);
}
