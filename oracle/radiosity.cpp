// radiosity.cpp -- ORACLE (test infrastructure): transfers (K2), direct light (K3), bounce (K4).
//
// None of these stages exists in the reference: `Transfer` is only a type
// (common/types/transfer.go:3-6), `Patch` carries the fields read here (common/types/patch.go:9-64),
// `DirectLight` + worldlight carry the light parameters (common/types/light.go:10-44,
// rad/lightmap/lights.go:216-341) and the radiosity step is commented out
// (cmd/tasks/computerad/main.go:5-10).  The algorithms encode the upstream semantics restated in
// SURVEY.md App. B.2-B.4 (uncited) -- PARITY UNPINNED; analytic known-answer tests are the arbiter.
#include "oracle_impl.hpp"
#include <omp.h>
#include <algorithm>

namespace orc {

static const float PLANE_TEST_EPSILON = 0.01f;
static const float TRANS_EPSILON = 1.0e-7f;
static const float PI_F = 3.14159265358979323846f;           // vmath/constants.go:5
static const float EQUAL_EPSILON = 0.001f;                   // vmath/constants.go:9
static const float MAX_TRACE_LENGTH = 1.732050807569f * 32768.0f;   // common/constants/constants.go:15-19

static inline float dot3(const float* a, const float* b) { return ((a[0] * b[0]) + (a[1] * b[1])) + (a[2] * b[2]); }

// Polygon-to-differential form factor (upstream vismat.cpp, near pairs; SURVEY App. B.3 "optional"; not in the reference): the
// contour integral over the emitter's edges -- angle subtended by the edge at the receiver times the unit normal of the plane
// through the receiver and the edge, dotted with the receiver's normal.  Per unit emitter area, without the 1/pi.  NaN = no polygon.
static inline float poly_to_diff_form_factor(const Patches& P, int j, const float* oi, const float* ni) {
    const int first = P.wind_first[j], cnt = P.wind_count[j];
    if (cnt < 3) return NAN;
    float ff = 0.0f;
    for (int k = 0; k < cnt; k++) {
        const float* p1 = &P.wind_pts[3 * (size_t)(first + k)];
        const float* p2 = &P.wind_pts[3 * (size_t)(first + (k + 1 < cnt ? k + 1 : 0))];
        float a[3] = {p1[0] - oi[0], p1[1] - oi[1], p1[2] - oi[2]};
        float b[3] = {p2[0] - oi[0], p2[1] - oi[1], p2[2] - oi[2]};
        const float la = sqrtf(dot3(a, a)), lb = sqrtf(dot3(b, b));
        if (la > 0.0f) { const float r = 1.0f / la; a[0] = a[0] * r; a[1] = a[1] * r; a[2] = a[2] * r; }
        if (lb > 0.0f) { const float r = 1.0f / lb; b[0] = b[0] * r; b[1] = b[1] * r; b[2] = b[2] * r; }
        float g[3] = {(a[1] * b[2]) - (a[2] * b[1]), (a[2] * b[0]) - (a[0] * b[2]), (a[0] * b[1]) - (a[1] * b[0])};
        const float sin_alpha = sqrtf(dot3(g, g));
        if (sin_alpha > 1.0f) return 0.0f;
        if (sin_alpha > 0.0f) {
            const float m = asinf(sin_alpha) * (1.0f / sin_alpha);
            g[0] = g[0] * m; g[1] = g[1] * m; g[2] = g[2] * m;
        }
        ff = ff + dot3(g, ni);
    }
    return ff * (0.5f / P.area[j]);
}

// App. B.3 MakeTransfer: differential-to-differential form factor * emitter area (polygon-to-differential for an emitter that is
// large for its distance, once windings are set).  Returns 0 when the pair transfers nothing.
static inline float transfer_weight(const Patches& P, int i, int j) {
    if ((P.flags[j] & 1) || !(P.area[j] > 0.0f)) return 0.0f;
    const float* oi = &P.origin[3 * i]; const float* oj = &P.origin[3 * j];
    const float* ni = &P.normal[3 * i]; const float* nj = &P.normal[3 * j];
    if (!(dot3(oj, ni) > P.plane_dist[i] + PLANE_TEST_EPSILON)) return 0.0f;
    float dl[3] = {oi[0] - oj[0], oi[1] - oj[1], oi[2] - oj[2]};
    float len2 = dot3(dl, dl);
    float len = sqrtf(len2);
    if (!(len > 0.0f)) return 0.0f;
    float r = 1.0f / len;
    dl[0] = dl[0] * r; dl[1] = dl[1] * r; dl[2] = dl[2] * r;
    float d1 = dot3(dl, ni), d2 = dot3(dl, nj);
    float scale = -(d1 * d2) / ((len * len) * PI_F);
    if (!(scale > 0.0f)) return 0.0f;
    if (!P.wind_count.empty() && ((len * len) * PI_F) * 0.04f < P.area[j]) {
        const float ff = poly_to_diff_form_factor(P, j, oi, ni);
        if (ff == ff) scale = ff / PI_F;
        if (!(scale > 0.0f)) return 0.0f;
    }
    float trans = P.area[j] * scale;
    if (!(trans > TRANS_EPSILON)) return 0.0f;
    return trans;
}

// shadow test between two patches; the segment always runs from the lower to the higher patch
// index so that visibility is symmetric by construction.
static inline int patches_see(const orc_env* e, int i, int j) {
    const Patches& P = e->patches;
    int lo = i < j ? i : j, hi = i < j ? j : i;
    float a[3], b[3];
    for (int c = 0; c < 3; c++) {
        a[c] = P.origin[3 * lo + c] + P.normal[3 * lo + c];
        b[c] = P.origin[3 * hi + c] + P.normal[3 * hi + c];
    }
    return test_line1(e, a, b, 0, 0);
}

// vismat.cpp TestPatchToPatch: descend into the children of an emitter that is large for its
// distance (|origin_i - origin_j|^2 / 16 < area_j), otherwise the emitter itself is the candidate.
static void test_patch_to_patch(const Patches& P, int i, int j, std::vector<int32_t>& cand) {
    if (P.child1[j] != -1) {
        const float* oi = &P.origin[3 * i]; const float* oj = &P.origin[3 * j];
        float tmp[3] = {oi[0] - oj[0], oi[1] - oj[1], oi[2] - oj[2]};
        if (dot3(tmp, tmp) * 0.0625f < P.area[j]) {
            test_patch_to_patch(P, i, P.child1[j], cand);
            test_patch_to_patch(P, i, P.child2[j], cand);
            return;
        }
    }
    cand.push_back(j);
}

// candidates of receiver i in ascending patch order (MakeScales walks the vis row in patch order)
static void row_candidates(const Patches& P, int i, int n_clusters, const uint8_t* pvs, std::vector<int32_t>& cand) {
    cand.clear();
    // a patch whose cluster is -1 (origin and winding in solid space, rad/patches/subdivide.go:100-116) is in no cluster's
    // child list (clusterChildren): it neither gathers nor is gathered from
    if (pvs && P.cluster[i] < 0) return;
    if (!P.hier()) {
        for (int j = 0; j < P.n; j++) {
            if (j == i) continue;
            if (pvs && (P.cluster[j] < 0 || !pvs[(size_t)P.cluster[i] * n_clusters + P.cluster[j]])) continue;
            cand.push_back(j);
        }
        return;
    }
    if (P.child1[i] != -1) return;                       // only leaf patches gather (BuildVisLeafs walks clusterChildren)
    for (int r = 0; r < P.n; r++) {
        if (P.parent[r] != -1) continue;                 // face root patches (faceParents)
        if (pvs && (P.cluster[r] < 0 || !pvs[(size_t)P.cluster[i] * n_clusters + P.cluster[r]])) continue;
        if (P.face[i] >= 0 && P.face[r] == P.face[i]) continue;     // "don't check patches on the same face"
        test_patch_to_patch(P, i, r, cand);
    }
    std::sort(cand.begin(), cand.end());
    cand.erase(std::remove(cand.begin(), cand.end(), i), cand.end());
}

static void build_row(const orc_env* e, int i, int n_clusters, const uint8_t* pvs, std::vector<int32_t>& cols, std::vector<float>& ws) {
    const Patches& P = e->patches;
    cols.clear(); ws.clear();
    if (P.flags[i] & 1) return;                          // sky patches receive nothing
    std::vector<int32_t> cand;
    row_candidates(P, i, n_clusters, pvs, cand);
    for (int32_t j : cand) {
        float tr = transfer_weight(P, i, j);
        if (tr == 0.0f) continue;
        if (!patches_see(e, i, j)) continue;
        cols.push_back(j); ws.push_back(tr);
    }
    // MakeScales (App. B.3): cap the row sum at 1
    float total = 0.0f;
    for (float v : ws) total = total + v;
    if (total > 1.0f) {
        float s = 1.0f / total;
        for (float& v : ws) v = v * s;
    }
}

} // namespace orc
using namespace orc;

extern "C" {

int orc_patches_set_hierarchy(orc_env* e, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face) {
    if (!e || n != e->patches.n) return -1;
    Patches& P = e->patches;
    P.parent.assign(parent, parent + n); P.child1.assign(child1, child1 + n); P.child2.assign(child2, child2 + n);
    if (face) P.face.assign(face, face + n); else P.face.assign(n, -1);
    e->rowptr.clear(); e->col.clear(); e->w.clear();
    return 0;
}

int orc_patches_set_windings(orc_env* e, int n, const int32_t* first, const int32_t* count, int n_points, const float* points3) {
    if (!e || n < 0 || n_points < 0) return -1;
    Patches& P = e->patches;
    e->rowptr.clear(); e->col.clear(); e->w.clear();
    if (n == 0) { P.wind_first.clear(); P.wind_count.clear(); P.wind_pts.clear(); return 0; }
    if (n != P.n || !first || !count || (n_points > 0 && !points3)) return -1;
    P.wind_first.assign(first, first + n); P.wind_count.assign(count, count + n);
    P.wind_pts.assign(points3, points3 + 3 * (size_t)n_points);
    for (int i = 0; i < n; i++) {
        const int f = first[i], c = count[i];
        if (c < 0 || f < 0 || (int64_t)f + c > n_points) return -1;
        if (c < 3) { P.wind_count[i] = 0; continue; }
        double nrm[3] = {0, 0, 0};                            // Newell normal: along the patch normal for a counter-clockwise winding
        for (int k = 0; k < c; k++) {
            const float* a = &P.wind_pts[3 * (size_t)(f + k)]; const float* b = &P.wind_pts[3 * (size_t)(f + (k + 1 < c ? k + 1 : 0))];
            nrm[0] += ((double)a[1] - b[1]) * ((double)a[2] + b[2]); nrm[1] += ((double)a[2] - b[2]) * ((double)a[0] + b[0]); nrm[2] += ((double)a[0] - b[0]) * ((double)a[1] + b[1]);
        }
        const double o = nrm[0] * P.normal[3 * (size_t)i] + nrm[1] * P.normal[3 * (size_t)i + 1] + nrm[2] * P.normal[3 * (size_t)i + 2];
        if (o > 0.0)
            for (int k = 0; k < c / 2; k++)
                for (int d = 0; d < 3; d++) std::swap(P.wind_pts[3 * (size_t)(f + k) + d], P.wind_pts[3 * (size_t)(f + c - 1 - k) + d]);
    }
    return 0;
}

int orc_patches_set(orc_env* e, int n, const float* origin3, const float* normal3, const float* plane_dist,
                    const float* area, const float* reflectivity3, const int32_t* cluster, const uint8_t* flags) {
    if (!e || n < 0) return -1;
    Patches& P = e->patches;
    P.n = n;
    P.origin.assign(origin3, origin3 + 3 * (size_t)n);
    P.normal.assign(normal3, normal3 + 3 * (size_t)n);
    P.plane_dist.assign(plane_dist, plane_dist + n);
    P.area.assign(area, area + n);
    P.refl.assign(reflectivity3, reflectivity3 + 3 * (size_t)n);
    if (cluster) P.cluster.assign(cluster, cluster + n); else P.cluster.assign(n, 0);
    if (flags) P.flags.assign(flags, flags + n); else P.flags.assign(n, 0);
    P.parent.clear(); P.child1.clear(); P.child2.clear(); P.face.clear();
    e->rowptr.clear(); e->col.clear(); e->w.clear();
    return 0;
}

int orc_build_transfers(orc_env* e, int n_clusters, const uint8_t* pvs, int64_t* nnz_out, int threads) {
    if (!e || !e->built) return -1;
    const Patches& P = e->patches;
    int N = P.n;
    std::vector<std::vector<int32_t>> cols(N);
    std::vector<std::vector<float>> ws(N);
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads > 0 ? threads : 1)
    for (int i = 0; i < N; i++) build_row(e, i, n_clusters, pvs, cols[i], ws[i]);
    e->rowptr.assign(N + 1, 0);
    for (int i = 0; i < N; i++) e->rowptr[i + 1] = e->rowptr[i] + (int64_t)cols[i].size();
    e->col.resize(e->rowptr[N]); e->w.resize(e->rowptr[N]);
    for (int i = 0; i < N; i++) {
        std::copy(cols[i].begin(), cols[i].end(), e->col.begin() + e->rowptr[i]);
        std::copy(ws[i].begin(), ws[i].end(), e->w.begin() + e->rowptr[i]);
    }
    if (nnz_out) *nnz_out = e->rowptr[N];
    return 0;
}

// one row of the transfer matrix (same rule as orc_build_transfers incl. MakeScales), for spot checks on maps whose
// full matrix is too large for the CPU (C5).  Returns the number of entries, or -1 if `cap` is too small.
int64_t orc_transfer_row(orc_env* e, int i, int n_clusters, const uint8_t* pvs, int32_t* col_out, float* w_out, int64_t cap) {
    if (!e || !e->built) return -1;
    const Patches& P = e->patches;
    if (i < 0 || i >= P.n) return -1;
    std::vector<int32_t> cols; std::vector<float> ws;
    build_row(e, i, n_clusters, pvs, cols, ws);
    if ((int64_t)cols.size() > cap) return -1;
    std::copy(cols.begin(), cols.end(), col_out); std::copy(ws.begin(), ws.end(), w_out);
    return (int64_t)cols.size();
}

int orc_transfers_get(orc_env* e, int64_t* rowptr, int32_t* col, float* w) {
    if (!e || e->rowptr.empty()) return -1;
    if (rowptr) memcpy(rowptr, e->rowptr.data(), e->rowptr.size() * sizeof(int64_t));
    if (col) memcpy(col, e->col.data(), e->col.size() * sizeof(int32_t));
    if (w) memcpy(w, e->w.data(), e->w.size() * sizeof(float));
    return 0;
}

// App. B.2.  Anorm directions for sky ambient arrive through lights of type 5 via `sky_dirs`.
static std::vector<float> g_sky_dirs; static int g_n_sky_dirs = 0;
int orc_set_sky_dirs(int n, const float* dirs3) { g_sky_dirs.assign(dirs3, dirs3 + 3 * (size_t)n); g_n_sky_dirs = n; return 0; }

int orc_env_set_light_trace_flags(orc_env* e, int flags) { if (!e) return -1; e->light_trace_flags = flags; return 0; }

int orc_direct_light(orc_env* e, int64_t n_luxels, const float* pos3, const float* normal3, int n_lights,
                     const orc_light* lights, float* rgb_out, int threads) {
    if (!e || !e->built) return -1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < n_luxels; i++) {
        const float* pos = &pos3[3 * i]; const float* n = &normal3[3 * i];
        float rgb[3] = {0, 0, 0};
        for (int L = 0; L < n_lights; L++) {
            const orc_light& dl = lights[L];
            float scale = 0.0f;
            if (dl.type == 0 || dl.type == 1 || dl.type == 2) {        // surface / point / spotlight
                float delta[3] = {dl.origin[0] - pos[0], dl.origin[1] - pos[1], dl.origin[2] - pos[2]};
                float dist2 = dot3(delta, delta);
                float dist = sqrtf(dist2);
                if (!(dist > 0.0f)) continue;
                float r = 1.0f / dist;
                delta[0] = delta[0] * r; delta[1] = delta[1] * r; delta[2] = delta[2] * r;
                float dot = dot3(delta, n);
                if (!(dot > 0.0f)) continue;
                if (dl.end_fade > dl.start_fade && dist > dl.end_fade) continue;
                dist = max_sel(dist, 1.0f);
                float d = min_sel(dist, dl.cap_dist);
                float denom = (dl.constant_attn + (dl.linear_attn * d)) + ((dl.quadratic_attn * d) * d);
                float falloff;
                if (dl.type == 1) {
                    falloff = 1.0f / denom;
                } else if (dl.type == 2) {
                    float dot2 = -dot3(delta, dl.normal);
                    if (dot2 <= dl.stopdot2) continue;
                    falloff = dot2 / denom;
                    if (dot2 <= dl.stopdot) {
                        float m = (dot2 - dl.stopdot2) / (dl.stopdot - dl.stopdot2);
                        m = min_sel(max_sel(m, 0.0f), 1.0f);
                        if (dl.exponent != 0.0f && dl.exponent != 1.0f) m = powf(m, dl.exponent);
                        falloff = falloff * m;
                    }
                } else {
                    float dot2 = -dot3(delta, dl.normal);
                    if (!(dot2 > 0.0f)) continue;
                    dot = dot * dot2;
                    falloff = 1.0f / (dist * dist);
                }
                if (dl.end_fade > dl.start_fade && dist > dl.start_fade) {
                    float t = (dist - dl.start_fade) / (dl.end_fade - dl.start_fade);
                    t = min_sel(max_sel(t, 0.0f), 1.0f);
                    float s = 1.0f - ((t * t) * (3.0f - (2.0f * t)));
                    falloff = falloff * s;
                }
                if (e->light_trace_flags == 0) {
                    if (!test_line1(e, pos, dl.origin, 0, 0)) continue;
                    scale = falloff * dot;
                } else {                                                 // App. B.2: dot *= fractionVisible
                    float fv = test_line_fraction(e, pos, dl.origin, e->light_trace_flags, ORC_TRACE_ID_STATICPROP | -1);
                    if (!(fv > 0.0f)) continue;
                    scale = (falloff * dot) * fv;
                }
            } else if (dl.type == 3) {                                  // skylight (sun)
                float dot = -dot3(dl.normal, n);
                if (!(dot > 0.0f)) continue;
                float stop[3] = {pos[0] - (dl.normal[0] * MAX_TRACE_LENGTH), pos[1] - (dl.normal[1] * MAX_TRACE_LENGTH),
                                 pos[2] - (dl.normal[2] * MAX_TRACE_LENGTH)};
                if (e->light_trace_flags == 0) {
                    if (!test_line1(e, pos, stop, 1, 0)) continue;
                    scale = dot;
                } else {
                    float fv = test_line_sky1(e, pos, stop, pos, e->light_trace_flags, ORC_TRACE_ID_STATICPROP | -1);
                    if (!(fv > 0.0f)) continue;
                    scale = dot * fv;
                }
            } else if (dl.type == 5) {                                  // sky ambient
                float sum = 0.0f, possible = 0.0f;
                for (int k = 0; k < g_n_sky_dirs; k++) {
                    const float* a = &g_sky_dirs[3 * k];
                    float dot = dot3(a, n);
                    if (!(dot > EQUAL_EPSILON)) continue;
                    possible = possible + dot;
                    float stop[3] = {pos[0] + (a[0] * MAX_TRACE_LENGTH), pos[1] + (a[1] * MAX_TRACE_LENGTH),
                                     pos[2] + (a[2] * MAX_TRACE_LENGTH)};
                    if (e->light_trace_flags == 0) {
                        if (test_line1(e, pos, stop, 1, 0)) sum = sum + dot;
                    } else {
                        float fv = test_line_sky1(e, pos, stop, pos, e->light_trace_flags, ORC_TRACE_ID_STATICPROP | -1);
                        if (fv > 0.0f) sum = sum + (dot * fv);
                    }
                }
                if (!(possible > 0.0f)) continue;
                scale = sum / possible;
            } else continue;
            rgb[0] = rgb[0] + (dl.intensity[0] * scale);
            rgb[1] = rgb[1] + (dl.intensity[1] * scale);
            rgb[2] = rgb[2] + (dl.intensity[2] * scale);
        }
        rgb_out[3 * i] = rgb[0]; rgb_out[3 * i + 1] = rgb[1]; rgb_out[3 * i + 2] = rgb[2];
    }
    return 0;
}

// App. B.4 GatherLight for rows [row0,row1): sequential fp32 sums in CSR order.
int orc_gather_rows(int64_t row0, int64_t row1, const int64_t* rowptr, const int32_t* col, const float* w,
                    const float* emit_rgb, const float* refl_rgb, float* out_rgb, int threads) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = row0; i < row1; i++) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) {
            int32_t j = col[k]; float wt = w[k];
            s0 = s0 + (wt * (emit_rgb[3 * j] * refl_rgb[3 * j]));
            s1 = s1 + (wt * (emit_rgb[3 * j + 1] * refl_rgb[3 * j + 1]));
            s2 = s2 + (wt * (emit_rgb[3 * j + 2] * refl_rgb[3 * j + 2]));
        }
        float* o = &out_rgb[3 * (i - row0)];
        o[0] = s0; o[1] = s1; o[2] = s2;
    }
    return 0;
}

// Upstream GetBumpNormals (SURVEY App. B.4 mentions the 4-normal variant; constants NUM_BUMP_VECTS = 3,
// common/constants/constants.go:33): a basis around the phong normal from the texture S vector, mirrored for left-handed
// texture axes, then the fixed tangent-space basis g_localBumpBasis rotated into world space.
int orc_bump_normals(const float s_vect[3], const float t_vect[3], const float flat_normal[3], const float phong_normal[3], float out9[9]) {
    static const float OO_SQRT_2 = 0.70710676908493042f, OO_SQRT_3 = 0.57735025882720947f, OO_SQRT_6 = 0.40824821591377258f,
                       OO_SQRT_2_OVER_3 = 0.81649661064147949f;
    static const float basis[3][3] = {{OO_SQRT_2_OVER_3, 0.0f, OO_SQRT_3}, {-OO_SQRT_6, OO_SQRT_2, OO_SQRT_3}, {-OO_SQRT_6, -OO_SQRT_2, OO_SQRT_3}};
    auto cross = [](const float* a, const float* b, float* o) {
        o[0] = (a[1] * b[2]) - (a[2] * b[1]); o[1] = (a[2] * b[0]) - (a[0] * b[2]); o[2] = (a[0] * b[1]) - (a[1] * b[0]);
    };
    auto normalize = [](float* v) {
        float len = sqrtf(dot3(v, v));
        if (len != 0.0f) { float r = 1.0f / len; v[0] = v[0] * r; v[1] = v[1] * r; v[2] = v[2] * r; }
    };
    float tmp[3];
    cross(s_vect, t_vect, tmp);
    bool left_handed = dot3(flat_normal, tmp) < 0.0f;
    float sb[3][3];
    cross(phong_normal, s_vect, sb[1]); normalize(sb[1]);
    cross(sb[1], phong_normal, sb[0]); normalize(sb[0]);
    for (int k = 0; k < 3; k++) sb[2][k] = phong_normal[k];
    if (left_handed) for (int k = 0; k < 3; k++) sb[1][k] = -sb[1][k];
    for (int i = 0; i < 3; i++)                                   // VectorIRotate: in.x * row0 + in.y * row1 + in.z * row2
        for (int k = 0; k < 3; k++) out9[3 * i + k] = ((basis[i][0] * sb[0][k]) + (basis[i][1] * sb[1][k])) + (basis[i][2] * sb[2][k]);
    return 0;
}

int orc_patches_set_bump(orc_env* e, int n, const uint8_t* needs_bump, const float* bump_normals9) {
    if (!e || n != e->patches.n) return -1;
    e->patches.needs_bump.assign(needs_bump, needs_bump + n);
    e->patches.bump_normals.assign(bump_normals9, bump_normals9 + 9 * (size_t)n);
    return 0;
}

int orc_bounce_bump_totals(orc_env* e, float* out9) {
    if (!e || e->patches.total_bump.empty()) return -1;
    memcpy(out9, e->patches.total_bump.data(), e->patches.total_bump.size() * sizeof(float));
    return 0;
}

// upstream GatherLight, bump-mapped branch, for one patch: the three bump sums (the flat sum stays with orc_gather_rows)
static void gather_bump_row(const Patches& P, int i, const int64_t* rowptr, const int32_t* col, const float* w, const float* emit, float sum9[9]) {
    for (int k = 0; k < 9; k++) sum9[k] = 0.0f;
    const float* oi = &P.origin[3 * i]; const float* ni = &P.normal[3 * i];
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) {
        int32_t j = col[k];
        float delta[3] = {P.origin[3 * j] - oi[0], P.origin[3 * j + 1] - oi[1], P.origin[3 * j + 2] - oi[2]};   // towards the emitter
        float len = sqrtf(dot3(delta, delta));
        if (len != 0.0f) { float r = 1.0f / len; delta[0] = delta[0] * r; delta[1] = delta[1] * r; delta[2] = delta[2] * r; }
        float scale = 1.0f / dot3(delta, ni);                       // "remove normal already factored into transfer steradian"
        float ws = w[k] * scale;
        float v[3] = {(emit[3 * j] * P.refl[3 * j]) * ws, (emit[3 * j + 1] * P.refl[3 * j + 1]) * ws, (emit[3 * j + 2] * P.refl[3 * j + 2]) * ws};
        for (int b = 0; b < 3; b++) {
            float d = dot3(delta, &P.bump_normals[9 * (size_t)i + 3 * b]);
            if (d <= 0.0f) continue;
            sum9[3 * b] = sum9[3 * b] + (v[0] * d); sum9[3 * b + 1] = sum9[3 * b + 1] + (v[1] * d); sum9[3 * b + 2] = sum9[3 * b + 2] + (v[2] * d);
        }
    }
}

int orc_bounce(orc_env* e, const float* emit0_rgb, int n_bounces, int early_out, float* total_rgb_out,
               float added_last[3], int* bounces_done, int threads) {
    if (!e || e->rowptr.empty()) return -1;
    const Patches& P = e->patches;
    int N = P.n;
    std::vector<float> emit(emit0_rgb, emit0_rgb + 3 * (size_t)N), add(3 * (size_t)N), total(3 * (size_t)N, 0.0f);
    float added[3] = {0, 0, 0};
    int done = 0;
    const bool bump = !P.needs_bump.empty();
    Patches& PW = e->patches;
    if (bump) PW.total_bump.assign(9 * (size_t)N, 0.0f); else PW.total_bump.clear();
    for (int b = 0; b < n_bounces; b++) {
        orc_gather_rows(0, N, e->rowptr.data(), e->col.data(), e->w.data(), emit.data(), P.refl.data(), add.data(), threads);
        if (bump) {                                        // GatherLight, bump branch; CollectLight adds it to TotalLight.Light[1..3] of leaf patches
            for (int i = 0; i < N; i++) {
                if (!P.needs_bump[i] || (P.flags[i] & 1) || (P.hier() && P.child1[i] != -1)) continue;
                float s9[9];
                gather_bump_row(P, i, e->rowptr.data(), e->col.data(), e->w.data(), emit.data(), s9);
                for (int k = 0; k < 9; k++) PW.total_bump[9 * (size_t)i + k] = PW.total_bump[9 * (size_t)i + k] + s9[k];
            }
        }
        // CollectLight (vrad.cpp, App. B.4).  Flat patch sets: forward order (every patch is a leaf).  With a hierarchy:
        // reverse index order so that children come before their parents; an interior patch takes the
        // area-weighted average of its two children for both totallight and emitlight.
        added[0] = added[1] = added[2] = 0.0f;
        const bool hier = P.hier();
        for (int n = 0; n < N; n++) {
            const int i = hier ? N - 1 - n : n;
            if (P.flags[i] & 1) { emit[3 * i] = emit[3 * i + 1] = emit[3 * i + 2] = 0.0f; continue; }
            if (!hier || P.child1[i] == -1) {
                for (int c = 0; c < 3; c++) {
                    total[3 * i + c] = total[3 * i + c] + add[3 * i + c];
                    emit[3 * i + c] = add[3 * i + c];
                    added[c] = added[c] + emit[3 * i + c];
                }
            } else {
                const int c1 = P.child1[i], c2 = P.child2[i];
                const float s1 = P.area[c1] / (P.area[c1] + P.area[c2]);
                const float s2 = P.area[c2] / (P.area[c1] + P.area[c2]);
                for (int c = 0; c < 3; c++) {
                    total[3 * i + c] = (total[3 * c1 + c] * s1) + (s2 * total[3 * c2 + c]);      // VectorScale, VectorMA
                    emit[3 * i + c] = (emit[3 * c1 + c] * s1) + (s2 * emit[3 * c2 + c]);
                }
            }
        }
        done++;
        if (early_out && added[0] < 1.0f && added[1] < 1.0f && added[2] < 1.0f) break;
    }
    memcpy(total_rgb_out, total.data(), total.size() * sizeof(float));
    if (added_last) { added_last[0] = added[0]; added_last[1] = added[1]; added_last[2] = added[2]; }
    if (bounces_done) *bounces_done = done;
    return 0;
}

int orc_num_threads(void) { return omp_get_max_threads(); }

} // extern "C"
