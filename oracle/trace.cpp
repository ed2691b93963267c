// trace.cpp -- ORACLE (test infrastructure): ray/kd-tree intersection.
//
// The reference's Trace4Rays is a stub (raytracer/environment.go:140-145); its signature, the
// FourRays/RayTracingResult layouts (raytracer/types/fourrays.go:8-11, result.go:8-12) and the
// consumer in raytracer/trace/testline.go:22-51 fix the contract.  The algorithm is the upstream
// packet traversal as restated in SURVEY.md App. B.1, with these spec decisions:
//   * per-lane arithmetic is IEEE fp32, exact 1/x and sqrt (simd.go:110-117,162-178);
//   * a zero direction component is replaced by FLT_EPSILON before the reciprocal
//     ("reciprocal saturate"); the front/back child order uses d[axis] < 0;
//   * equal-t ties resolve to the lower triangle index (order independent);
//   * a FourRays result is defined as 4 independent single-ray results: trace1() IS THE SPEC,
//     trace4() is the coherent-packet CPU baseline and may differ from it only on fp near-ties.
#include "oracle_impl.hpp"
#include <cfloat>

namespace orc {

static const float HIT_INIT = 1.0e23f;       // RayTracingResult init (App. B.1)
static const float DDOTN_EPS = 1.1920929e-7f;

// One triangle test (App. B.1 "Leaf").  Updates best when (t, index) improves.
static inline void test_triangle(const orc_tri48& T, int32_t ti, const float o[3], const float d[3],
                                 int32_t skip_id, Hit& best) {
    if (T.id == skip_id) return;
    float ddotn = ((d[0] * T.nx) + (d[1] * T.ny)) + (d[2] * T.nz);
    if (!(ddotn > DDOTN_EPS || ddotn < -DDOTN_EPS)) return;
    float odotn = ((o[0] * T.nx) + (o[1] * T.ny)) + (o[2] * T.nz);
    float t = (T.d - odotn) / ddotn;
    if (!(t > 0.0f)) return;
    if (!(t < best.t || (t == best.t && ti < best.tri))) return;
    float c0 = o[T.sel0] + (t * d[T.sel0]);
    float c1 = o[T.sel1] + (t * d[T.sel1]);
    float b0 = ((T.e[0] * c0) + (T.e[1] * c1)) + T.e[2];
    if (!(b0 >= 0.0f)) return;
    float b1 = ((T.e[3] * c0) + (T.e[4] * c1)) + T.e[5];
    if (!(b1 >= 0.0f)) return;
    if (!((b0 + b1) <= 1.0f)) return;
    best.tri = ti; best.t = t;
}

static inline void inv_dir(const float d[3], float inv[3]) {
    for (int a = 0; a < 3; a++) {
        float dd = (d[a] == 0.0f) ? FLT_EPSILON : d[a];
        inv[a] = 1.0f / dd;
    }
}

static inline bool clip_to_bounds(const orc_env* e, const float o[3], const float inv[3], float& tmin, float& tmax) {
    for (int a = 0; a < 3; a++) {
        float t0 = (e->bmin[a] - o[a]) * inv[a];
        float t1 = (e->bmax[a] - o[a]) * inv[a];
        tmin = max_sel(tmin, min_sel(t0, t1));
        tmax = min_sel(tmax, max_sel(t0, t1));
    }
    return tmin <= tmax;
}

Hit trace_brute(const orc_env* e, const float o[3], const float d[3], float tmin, float tmax, int32_t skip_id) {
    (void)tmin; (void)tmax;
    Hit best{-1, HIT_INIT};
    int n = (int)e->tris.size();
    for (int i = 0; i < n; i++) test_triangle(e->tris[i], i, o, d, skip_id, best);
    return best;
}

Hit trace1(const orc_env* e, const float o[3], const float d[3], float tmin, float tmax,
           int32_t skip_id, Counters* ctr) {
    Hit best{-1, HIT_INIT};
    float inv[3];
    inv_dir(d, inv);
    if (!clip_to_bounds(e, o, inv, tmin, tmax)) return best;
    struct Entry { int32_t node; float tmin, tmax; } stack[64];
    int sp = 0;
    int32_t node = 0;
    const KDNode* nodes = e->nodes.data();
    for (;;) {
        KDNode nd = nodes[node];
        while ((nd.children & 3) != ORC_KDNODE_LEAF) {
            if (ctr) ctr->nodes++;
            int axis = nd.children & 3;
            int32_t left = nd.children >> 2;
            bool neg = d[axis] < 0.0f;
            int32_t front = left + (neg ? 1 : 0), back = left + (neg ? 0 : 1);
            float t = (nd.split - o[axis]) * inv[axis];
            if (!(t >= tmin)) {                       // misses the front child
                node = back; tmin = max_sel(tmin, t);
            } else if (!(t <= tmax)) {                // misses the back child
                node = front; tmax = min_sel(tmax, t);
            } else {
                stack[sp].node = back; stack[sp].tmin = max_sel(tmin, t); stack[sp].tmax = tmax; sp++;
                node = front; tmax = min_sel(tmax, t);
            }
            nd = nodes[node];
        }
        int32_t start = nd.children >> 2;
        int cnt = (int)nd.split;
        if (ctr) { ctr->leaves++; ctr->tris += cnt; }
        for (int k = 0; k < cnt; k++) {
            int32_t ti = e->tri_index[start + k];
            test_triangle(e->tris[ti], ti, o, d, skip_id, best);
        }
        if (!(tmax <= best.t)) return best;           // hit lies inside this leaf's interval: done
        if (sp == 0) return best;
        sp--; node = stack[sp].node; tmin = stack[sp].tmin; tmax = stack[sp].tmax;
    }
}

// 4-wide coherent packet (App. B.1).  All four lanes must agree on direction signs.
static void trace4_coherent(const orc_env* e, const float o[3][4], const float d[3][4], const float tmin_in[4],
                            const float tmax_in[4], int32_t skip_id, int32_t hit_tri[4], float hit_t[4]) {
    float inv[3][4], tmin[4], tmax[4];
    bool neg[3];
    for (int l = 0; l < 4; l++) { hit_tri[l] = -1; hit_t[l] = HIT_INIT; tmin[l] = tmin_in[l]; tmax[l] = tmax_in[l]; }
    for (int a = 0; a < 3; a++) {
        neg[a] = d[a][0] < 0.0f;
        for (int l = 0; l < 4; l++) {
            float dd = (d[a][l] == 0.0f) ? FLT_EPSILON : d[a][l];
            inv[a][l] = 1.0f / dd;
        }
    }
    bool any = false;
    for (int l = 0; l < 4; l++) {
        for (int a = 0; a < 3; a++) {
            float t0 = (e->bmin[a] - o[a][l]) * inv[a][l];
            float t1 = (e->bmax[a] - o[a][l]) * inv[a][l];
            tmin[l] = max_sel(tmin[l], min_sel(t0, t1));
            tmax[l] = min_sel(tmax[l], max_sel(t0, t1));
        }
        any |= tmin[l] <= tmax[l];
    }
    if (!any) return;
    struct Entry { int32_t node; float tmin[4], tmax[4]; } stack[64];
    int sp = 0;
    int32_t node = 0;
    const KDNode* nodes = e->nodes.data();
    for (;;) {
        KDNode nd = nodes[node];
        while ((nd.children & 3) != ORC_KDNODE_LEAF) {
            int axis = nd.children & 3;
            int32_t left = nd.children >> 2;
            int32_t front = left + (neg[axis] ? 1 : 0), back = left + (neg[axis] ? 0 : 1);
            float t[4];
            bool hits_front = false, hits_back = false;
            for (int l = 0; l < 4; l++) {
                t[l] = (nd.split - o[axis][l]) * inv[axis][l];
                bool active = tmin[l] <= tmax[l];
                hits_front |= active && (t[l] >= tmin[l]);
                hits_back  |= active && (t[l] <= tmax[l]);
            }
            if (!hits_front) {
                node = back;
                for (int l = 0; l < 4; l++) tmin[l] = max_sel(tmin[l], t[l]);
            } else if (!hits_back) {
                node = front;
                for (int l = 0; l < 4; l++) tmax[l] = min_sel(tmax[l], t[l]);
            } else {
                stack[sp].node = back;
                for (int l = 0; l < 4; l++) { stack[sp].tmin[l] = max_sel(tmin[l], t[l]); stack[sp].tmax[l] = tmax[l]; }
                sp++;
                node = front;
                for (int l = 0; l < 4; l++) tmax[l] = min_sel(tmax[l], t[l]);
            }
            nd = nodes[node];
        }
        int32_t start = nd.children >> 2;
        int cnt = (int)nd.split;
        for (int k = 0; k < cnt; k++) {
            int32_t ti = e->tri_index[start + k];
            const orc_tri48& T = e->tris[ti];
            if (T.id == skip_id) continue;
            for (int l = 0; l < 4; l++) {
                float ol[3] = {o[0][l], o[1][l], o[2][l]}, dl[3] = {d[0][l], d[1][l], d[2][l]};
                Hit h{hit_tri[l], hit_t[l]};
                test_triangle(T, ti, ol, dl, skip_id, h);
                hit_tri[l] = h.tri; hit_t[l] = h.t;
            }
        }
        bool cont = false;
        for (int l = 0; l < 4; l++) cont |= tmax[l] <= hit_t[l];
        if (!cont) return;
        if (sp == 0) return;
        sp--; node = stack[sp].node;
        for (int l = 0; l < 4; l++) { tmin[l] = stack[sp].tmin[l]; tmax[l] = stack[sp].tmax[l]; }
    }
}

// FourRays with arbitrary lanes: if the lanes disagree on a direction sign the packet is
// re-traced lane by lane with the lane replicated (fourrays.go:13-29 comment).
void trace4(const orc_env* e, const float o[3][4], const float d[3][4], const float tmin[4], const float tmax[4],
            int32_t skip_id, int32_t hit_tri[4], float hit_t[4]) {
    bool coherent = true;
    for (int a = 0; a < 3; a++)
        for (int l = 1; l < 4; l++) coherent &= ((d[a][l] < 0.0f) == (d[a][0] < 0.0f));
    if (coherent) { trace4_coherent(e, o, d, tmin, tmax, skip_id, hit_tri, hit_t); return; }
    for (int l = 0; l < 4; l++) {
        float o1[3][4], d1[3][4], tn[4], tx[4]; int32_t ht[4]; float hd[4];
        for (int a = 0; a < 3; a++) for (int k = 0; k < 4; k++) { o1[a][k] = o[a][l]; d1[a][k] = d[a][l]; }
        for (int k = 0; k < 4; k++) { tn[k] = tmin[l]; tx[k] = tmax[l]; }
        trace4_coherent(e, o1, d1, tn, tx, skip_id, ht, hd);
        hit_tri[l] = ht[0]; hit_t[l] = hd[0];
    }
}

// testline.go:22-27 (segment -> normalised ray) and :42-51 (occlusion rule), one lane.
static inline bool segment_to_ray(const float a[3], const float b[3], float d[3], float& len) {
    d[0] = b[0] - a[0]; d[1] = b[1] - a[1]; d[2] = b[2] - a[2];
    float len2 = ((d[0] * d[0]) + (d[1] * d[1])) + (d[2] * d[2]);   // fourvectors.go:71-91
    if (len2 == 0.0f) return false;                                 // zero-length segment: visible by definition
    len = sqrtf(len2);                                              // simd.go:162-178 (exact sqrt)
    float r = 1.0f / len;                                           // simd.go:110-117 (exact reciprocal)
    d[0] = d[0] * r; d[1] = d[1] * r; d[2] = d[2] * r;
    return true;
}

static inline int occlusion_rule(const orc_env* e, Hit h, float len, int sky_mode) {
    if (h.tri != -1 && h.t < len) {
        if (!sky_mode) return 0;
        if ((e->tris[h.tri].id & ORC_TRACE_ID_SKY) == 0) return 0;
    }
    return 1;
}

int test_line1(const orc_env* e, const float a[3], const float b[3], int sky_mode, int mode) {
    float d[3], len;
    if (!segment_to_ray(a, b, d, len)) return 1;
    Hit h = (mode == 2) ? trace_brute(e, a, d, 0.0f, len, -1) : trace1(e, a, d, 0.0f, len, -1, nullptr);
    return occlusion_rule(e, h, len, sky_mode);
}

} // namespace orc

using namespace orc;

extern "C" {

int orc_trace_brute(orc_env* e, int64_t n, const float* ox, const float* oy, const float* oz,
                    const float* dx, const float* dy, const float* dz, const float* tmin, const float* tmax,
                    int32_t skip_id, int32_t* hit_tri, int32_t* hit_sid, float* hit_t, int threads) {
    if (!e || !e->built) return -1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < n; i++) {
        float o[3] = {ox[i], oy[i], oz[i]}, d[3] = {dx[i], dy[i], dz[i]};
        Hit h = trace_brute(e, o, d, tmin ? tmin[i] : 0.0f, tmax[i], skip_id);
        hit_tri[i] = h.tri; hit_t[i] = h.t;
        if (hit_sid) hit_sid[i] = h.tri >= 0 ? e->tris[h.tri].id : -1;
    }
    return 0;
}

int orc_trace1(orc_env* e, int64_t n, const float* ox, const float* oy, const float* oz,
               const float* dx, const float* dy, const float* dz, const float* tmin, const float* tmax,
               int32_t skip_id, int32_t* hit_tri, int32_t* hit_sid, float* hit_t, int threads) {
    if (!e || !e->built) return -1;
    if (threads <= 1) {
        e->counters = Counters();
        for (int64_t i = 0; i < n; i++) {
            float o[3] = {ox[i], oy[i], oz[i]}, d[3] = {dx[i], dy[i], dz[i]};
            Hit h = trace1(e, o, d, tmin ? tmin[i] : 0.0f, tmax[i], skip_id, &e->counters);
            hit_tri[i] = h.tri; hit_t[i] = h.t;
            if (hit_sid) hit_sid[i] = h.tri >= 0 ? e->tris[h.tri].id : -1;
        }
        return 0;
    }
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
    for (int64_t i = 0; i < n; i++) {
        float o[3] = {ox[i], oy[i], oz[i]}, d[3] = {dx[i], dy[i], dz[i]};
        Hit h = trace1(e, o, d, tmin ? tmin[i] : 0.0f, tmax[i], skip_id, nullptr);
        hit_tri[i] = h.tri; hit_t[i] = h.t;
        if (hit_sid) hit_sid[i] = h.tri >= 0 ? e->tris[h.tri].id : -1;
    }
    return 0;
}

int orc_trace4(orc_env* e, int64_t n, const float* ox, const float* oy, const float* oz,
               const float* dx, const float* dy, const float* dz, const float* tmin, const float* tmax,
               int32_t skip_id, int32_t* hit_tri, int32_t* hit_sid, float* hit_t, int threads) {
    if (!e || !e->built) return -1;
    int64_t npk = (n + 3) / 4;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)
    for (int64_t p = 0; p < npk; p++) {
        float o[3][4], d[3][4], tn[4], tx[4]; int32_t ht[4]; float hd[4];
        for (int l = 0; l < 4; l++) {
            int64_t i = p * 4 + l; if (i >= n) i = n - 1;        // tail: replicate the last ray
            o[0][l] = ox[i]; o[1][l] = oy[i]; o[2][l] = oz[i];
            d[0][l] = dx[i]; d[1][l] = dy[i]; d[2][l] = dz[i];
            tn[l] = tmin ? tmin[i] : 0.0f; tx[l] = tmax[i];
        }
        trace4(e, o, d, tn, tx, skip_id, ht, hd);
        for (int l = 0; l < 4; l++) {
            int64_t i = p * 4 + l; if (i >= n) break;
            hit_tri[i] = ht[l]; hit_t[i] = hd[l];
            if (hit_sid) hit_sid[i] = ht[l] >= 0 ? e->tris[ht[l]].id : -1;
        }
    }
    return 0;
}

int orc_trace4_packet(orc_env* e, const float origin_xyz4[12], const float dir_xyz4[12], const float tmin[4],
                      const float tmax[4], int32_t skip_id, int32_t hit_ids[4], float hit_dist[4], float normal_xyz4[12]) {
    if (!e || !e->built) return -1;
    // spec: a FourRays result == 4 independent single-ray results
    for (int l = 0; l < 4; l++) {
        float o[3] = {origin_xyz4[l], origin_xyz4[4 + l], origin_xyz4[8 + l]};
        float d[3] = {dir_xyz4[l], dir_xyz4[4 + l], dir_xyz4[8 + l]};
        Hit h = trace1(e, o, d, tmin[l], tmax[l], skip_id, nullptr);
        hit_ids[l] = h.tri; hit_dist[l] = h.t;
        if (normal_xyz4) {
            normal_xyz4[l]     = h.tri >= 0 ? e->tris[h.tri].nx : 0.0f;
            normal_xyz4[4 + l] = h.tri >= 0 ? e->tris[h.tri].ny : 0.0f;
            normal_xyz4[8 + l] = h.tri >= 0 ? e->tris[h.tri].nz : 0.0f;
        }
    }
    return 0;
}

int orc_test_lines(orc_env* e, int64_t n, const float* start_soa, const float* stop_soa, int sky_mode,
                   uint32_t* vis_bits, int mode, int threads) {
    if (!e || !e->built) return -1;
    int64_t nwords = (n + 31) / 32;
    const float *sx = start_soa, *sy = start_soa + n, *sz = start_soa + 2 * n;
    const float *ex = stop_soa, *ey = stop_soa + n, *ez = stop_soa + 2 * n;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t wd = 0; wd < nwords; wd++) {
        uint32_t bits = 0;
        int64_t base = wd * 32;
        if (mode == 1) {
            for (int p = 0; p < 8; p++) {
                float o[3][4], d[3][4], tn[4], tx[4]; int32_t ht[4]; float hd[4]; bool valid[4];
                bool anyv = false;
                for (int l = 0; l < 4; l++) {
                    int64_t i = base + p * 4 + l;
                    valid[l] = false;
                    if (i < n) {
                        float a[3] = {sx[i], sy[i], sz[i]}, b[3] = {ex[i], ey[i], ez[i]}, dd[3], len;
                        if (segment_to_ray(a, b, dd, len)) {
                            valid[l] = true; anyv = true;
                            for (int c = 0; c < 3; c++) { o[c][l] = a[c]; d[c][l] = dd[c]; }
                            tn[l] = 0.0f; tx[l] = len;
                        }
                    }
                }
                if (!anyv) { for (int l = 0; l < 4; l++) if (base + p * 4 + l < n) bits |= 1u << (p * 4 + l); continue; }
                int first = 0; while (!valid[first]) first++;
                for (int l = 0; l < 4; l++) if (!valid[l]) {
                    for (int c = 0; c < 3; c++) { o[c][l] = o[c][first]; d[c][l] = d[c][first]; }
                    tn[l] = tn[first]; tx[l] = tx[first];
                }
                trace4(e, o, d, tn, tx, -1, ht, hd);
                for (int l = 0; l < 4; l++) {
                    int64_t i = base + p * 4 + l;
                    if (i >= n) continue;
                    int vis = valid[l] ? occlusion_rule(e, Hit{ht[l], hd[l]}, tx[l], sky_mode) : 1;
                    if (vis) bits |= 1u << (p * 4 + l);
                }
            }
        } else {
            for (int k = 0; k < 32; k++) {
                int64_t i = base + k;
                if (i >= n) break;
                float a[3] = {sx[i], sy[i], sz[i]}, b[3] = {ex[i], ey[i], ez[i]};
                if (test_line1(e, a, b, sky_mode, mode)) bits |= 1u << k;
            }
        }
        vis_bits[wd] = bits;
    }
    return 0;
}

int orc_trace_counters(orc_env* e, int64_t* nodes_visited, int64_t* tris_tested, int64_t* leaves_visited) {
    if (!e) return -1;
    if (nodes_visited) *nodes_visited = e->counters.nodes;
    if (tris_tested) *tris_tested = e->counters.tris;
    if (leaves_visited) *leaves_visited = e->counters.leaves;
    return 0;
}

} // extern "C"
