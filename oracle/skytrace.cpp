// skytrace.cpp -- ORACLE (test infrastructure): the full trace.TestLineDoesHitSky call surface.
//
// Follows raytracer/trace/testline.go:18-94 statement by statement, lane by lane:
//   :22-27  segment -> normalised ray (exact sqrt / reciprocal, simd.go:110-117,171-178)
//   :28-36  Trace4Rays with skip id TRACE_ID_STATICPROP | staticPropToSkip and, when textureShadows
//           is on, the coverage callback (raytracer/types/coverageCount.go:16-48)
//   :42-51  occlusion = 1 when the nearest hit is closer than the segment end and is not TRACE_ID_SKY
//   :52-55  occlusion = max(occlusion, coverage)
//   :57-89  3D-skybox recursion: PointLeafnum(start) (raytracer/trace/pointleaf.go:8-33) -> leaf area;
//           if that area has no sky camera, re-trace through every sky camera
//           (skystart = cam.Origin + start*WorldToSky, skystop = skystart + dir*MAX_TRACE_LENGTH,
//            occlusion = occlusion + 1 - fractionVisible(recursive call, canRecurse=false))
//   :91-93  fractionVisible = 1 - clamp(occlusion, 0, 1)
// plus the callers' data: ProcessSkyCameras (rad/cameras/skycamera.go:10-49), CanLeafTraceToSky
// (rad/lightmap/lightmap.go:425-451), clustertable.PointInLeaf (rad/clustertable/point.go:14-38) and
// DecompressVis (rad/lightmap/vis.go:54-94).
//
// Spec decisions (PARITY UNPINNED, see oracle.h; all shared with the CUDA path):
//   * per-lane semantics: the packet's `fullyOccluded` early-out (:57) only skips work, a lane with
//     occlusion >= 1 stays at 1 after :86-87 and the final clamp, so each lane is independent -- except
//     that the packet form takes the leaf from lane 0 (start.Vec(0), :63): flag ORC_TL_PACKET_LEAF
//     reproduces that for every group of 4 segments; without it each segment uses its own start.
//   * cache.CountSkyCameras() returns the array capacity (cache/skycameras.go:30-32); intent = numSkyCameras.
//   * ProcessSkyCameras discards e.VectorForKey("origin") (skycamera.go:26); intent = the entity origin.
//   * CanLeafTraceToSky never copies `center` into center4 (lightmap.go:428-433) and sums int16 bounds in
//     int16 (:430); intent = DuplicateVector(center) and an int sum as in the C original.
//   * transparent triangles (flag bit 0, upstream FCACHETRI_TRANSPARENT) only behave differently when a
//     callback is given (textureShadows); the callback is CoverageCount: coverage += colour.X of the
//     triangle, clamped to 1; the hit is dropped while coverage < 1 ("continue") and kept once it reaches 1.
//     A triangle is counted once per ray however many kd leaves hold it: the counted triangles are kept in
//     an exact per-ray list of 16 entries; a ray that crosses more distinct transparent triangles than that
//     is treated as fully covered.  (Upstream's Trace4Rays avoids retests with a 256-entry direct-mapped
//     mailbox, which double counts after an eviction.)
//     A transparent triangle only counts when its hit lies before the segment end (t < the ray's tmax):
//     upstream tests hits against HitDistance only, which would make the coverage depend on which
//     triangles happen to share the last kd leaf with the segment end.
//     CoverageCountTexture calls staticprops.ComputeCoverageFromTexture, which returns 0 in the reference
//     (rad/staticprops/compute.go:4-12), so the colour form is the one with observable behaviour.
//   * a zero-length segment is visible (fraction 1) and does not recurse.
//   * VectorNormalize (fourvectors.go:85-88) = v * float32(1/sqrt(float64(v.v))) (simd.go:162-169).
#include "oracle_impl.hpp"
#include <cfloat>

namespace orc {

static const float HIT_INIT = 1.0e23f;
static const float DDOTN_EPS = 1.1920929e-7f;
static const float MAX_TRACE_LENGTH = (float)(1.732050807569 * 32768.0);   // common/constants/constants.go:15-19
static const float TEST_EPSILON = 0.03125f;                                // vmath/constants.go:7

struct CovState {
    float cov = 0.0f;
    float seg_len = 0.0f;                                // the ray's own tmax: panes beyond the segment end do not count
    int n = 0;
    int32_t counted[16];
};

static inline void test_triangle_cov(const orc_env* e, const orc_tri48& T, int32_t ti, const float o[3], const float d[3],
                                     int32_t skip_id, Hit& best, CovState& cs) {
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
    if (T.flags & 1) {                                   // FCACHETRI_TRANSPARENT + callback
        if (!(t < cs.seg_len)) return;                   // beyond the segment end: neither counted nor a hit
        for (int k = 0; k < cs.n; k++) if (cs.counted[k] == ti) return;   // already counted for this ray
        if (cs.n == 16) cs.cov = 1.0f;                   // list full: treated as fully covered
        else {
            cs.counted[cs.n++] = ti;
            float c = (size_t)(3 * ti) < e->tri_color.size() ? e->tri_color[3 * (size_t)ti] : 0.0f;   // colour.X, coverageCount.go:29
            cs.cov = min_sel(cs.cov + c, 1.0f);          // :29-30
        }
        if (!(cs.cov == 1.0f)) return;                   // :32-37 "continue": the hit is dropped
    }
    best.tri = ti; best.t = t;
}

Hit trace1_coverage(const orc_env* e, const float o[3], const float d[3], float tmin, float tmax,
                    int32_t skip_id, float* coverage) {
    if (!coverage) return trace1(e, o, d, tmin, tmax, skip_id, nullptr);
    CovState cs;
    cs.seg_len = tmax;
    Hit best{-1, HIT_INIT};
    float inv[3];
    for (int a = 0; a < 3; a++) inv[a] = 1.0f / ((d[a] == 0.0f) ? FLT_EPSILON : d[a]);
    for (int a = 0; a < 3; a++) {
        float t0 = (e->bmin[a] - o[a]) * inv[a];
        float t1 = (e->bmax[a] - o[a]) * inv[a];
        tmin = max_sel(tmin, min_sel(t0, t1));
        tmax = min_sel(tmax, max_sel(t0, t1));
    }
    *coverage = 0.0f;
    if (!(tmin <= tmax)) return best;
    struct Entry { int32_t node; float tmin, tmax; } stack[64];
    int sp = 0;
    int32_t node = 0;
    const KDNode* nodes = e->nodes.data();
    for (;;) {
        KDNode nd = nodes[node];
        while ((nd.children & 3) != ORC_KDNODE_LEAF) {
            int axis = nd.children & 3;
            int32_t left = nd.children >> 2;
            bool neg = d[axis] < 0.0f;
            int32_t front = left + (neg ? 1 : 0), back = left + (neg ? 0 : 1);
            float t = (nd.split - o[axis]) * inv[axis];
            if (!(t >= tmin)) { node = back; tmin = max_sel(tmin, t); }
            else if (!(t <= tmax)) { node = front; tmax = min_sel(tmax, t); }
            else {
                stack[sp].node = back; stack[sp].tmin = max_sel(tmin, t); stack[sp].tmax = tmax; sp++;
                node = front; tmax = min_sel(tmax, t);
            }
            nd = nodes[node];
        }
        int32_t start = nd.children >> 2;
        int cnt = (int)nd.split;
        for (int k = 0; k < cnt; k++) {
            int32_t ti = e->tri_index[start + k];
            test_triangle_cov(e, e->tris[ti], ti, o, d, skip_id, best, cs);
        }
        if (!(tmax <= best.t) || sp == 0) break;
        sp--; node = stack[sp].node; tmin = stack[sp].tmin; tmax = stack[sp].tmax;
    }
    *coverage = cs.cov;
    return best;
}

// raytracer/trace/pointleaf.go:8-33
static int point_leafnum(const Bsp& B, const float p[3]) {
    int node = 0;
    if (B.node_plane.empty()) return 0;                  // a map without nodes is its single leaf
    while (node >= 0) {
        int pl = B.node_plane[node];
        float dist;
        int type = B.plane_type[pl];
        if (type < 3) dist = p[type] - B.plane_dist[pl];
        else dist = (((B.plane_normal[3 * pl] * p[0]) + (B.plane_normal[3 * pl + 1] * p[1])) + (B.plane_normal[3 * pl + 2] * p[2])) - B.plane_dist[pl];
        node = (dist < 0.0f) ? B.node_children[2 * node + 1] : B.node_children[2 * node];
    }
    return -1 - node;
}

// rad/clustertable/point.go:14-38 (recursive; first branch wins unless it ends in a cluster -1 leaf)
static int point_in_leaf(const Bsp& B, int node, const float p[3]) {
    if (node < 0) return -1 - node;
    int pl = B.node_plane[node];
    float dist = (((p[0] * B.plane_normal[3 * pl]) + (p[1] * B.plane_normal[3 * pl + 1])) + (p[2] * B.plane_normal[3 * pl + 2])) - B.plane_dist[pl];
    if (dist > TEST_EPSILON) return point_in_leaf(B, B.node_children[2 * node], p);
    if (dist < -TEST_EPSILON) return point_in_leaf(B, B.node_children[2 * node + 1], p);
    int l = point_in_leaf(B, B.node_children[2 * node], p);
    if (B.leaf_cluster[l] != -1) return l;
    return point_in_leaf(B, B.node_children[2 * node + 1], p);
}

static inline bool segment_to_ray(const float a[3], const float b[3], float d[3], float& len) {
    d[0] = b[0] - a[0]; d[1] = b[1] - a[1]; d[2] = b[2] - a[2];
    float len2 = ((d[0] * d[0]) + (d[1] * d[1])) + (d[2] * d[2]);
    if (len2 == 0.0f) return false;
    len = sqrtf(len2);
    float r = 1.0f / len;
    d[0] = d[0] * r; d[1] = d[1] * r; d[2] = d[2] * r;
    return true;
}

// testline.go:22-55: occlusion of one lane before the recursion
static float primary_occlusion(const orc_env* e, const float a[3], const float b[3], int flags, int32_t skip_id, bool* degenerate,
                               bool sky_rule = true) {
    float d[3], len;
    *degenerate = false;
    if (!segment_to_ray(a, b, d, len)) { *degenerate = true; return 0.0f; }
    float cov = 0.0f;
    Hit h = trace1_coverage(e, a, d, 0.0f, len, skip_id, (flags & ORC_TL_TEXTURE_SHADOWS) ? &cov : nullptr);
    float occ = 0.0f;
    if (h.tri != -1 && h.t < len && (!sky_rule || (e->tris[h.tri].id & ORC_TRACE_ID_SKY) == 0)) occ = 1.0f;
    if (flags & ORC_TL_TEXTURE_SHADOWS) occ = max_sel(occ, cov);
    return occ;
}

static inline float finish(float occ) {                  // testline.go:91-93
    occ = max_sel(occ, 0.0f);
    occ = min_sel(occ, 1.0f);
    return 1.0f - occ;
}

// leaf_point: the point whose leaf/area decides the recursion (own start, or lane 0's with ORC_TL_PACKET_LEAF)
// trace.TestLine as a fraction (App. B.1: no sky rule, no recursion), with coverage when ORC_TL_TEXTURE_SHADOWS is set
float test_line_fraction(const orc_env* e, const float a[3], const float b[3], int flags, int32_t skip_id) {
    bool degenerate;
    float occ = primary_occlusion(e, a, b, flags, skip_id, &degenerate, false);
    return degenerate ? 1.0f : finish(occ);
}

float test_line_sky1(const orc_env* e, const float a[3], const float b[3], const float leaf_point[3], int flags, int32_t skip_id) {
    bool degenerate;
    float occ = primary_occlusion(e, a, b, flags, skip_id, &degenerate);
    if (degenerate) return 1.0f;
    if (occ < 1.0f && (flags & ORC_TL_CAN_RECURSE) && e->bsp.set && !e->cams.area.empty()) {
        float dir[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        float magsq = (dir[0] * dir[0]);                  // fourvectors.go:71-78: mul, madd, madd
        magsq = (dir[1] * dir[1]) + magsq;
        magsq = (dir[2] * dir[2]) + magsq;
        float rs = (float)(1.0 / sqrt((double)magsq));    // simd.go:162-169
        dir[0] *= rs; dir[1] *= rs; dir[2] *= rs;
        int leaf = point_leafnum(e->bsp, leaf_point);
        if (leaf >= 0 && leaf < (int)e->bsp.leaf_area.size()) {
            int area = e->bsp.leaf_area[leaf];
            if (area >= 0 && area < e->bsp.n_areas && e->cams.area_camera[area] < 0) {
                const int ncam = (int)e->cams.area.size();
                for (int c = 0; c < ncam; c++) {
                    float w = e->cams.world_to_sky[c];
                    float s0[3], s1[3];
                    for (int k = 0; k < 3; k++) {
                        s0[k] = e->cams.origin[3 * c + k] + (a[k] * w);          // :73-76
                        s1[k] = (dir[k] * MAX_TRACE_LENGTH) + s0[k];             // :78-80
                    }
                    bool deg2;
                    float occ2 = primary_occlusion(e, s0, s1, flags, skip_id, &deg2);   // :81 canRecurse=false
                    float fv2 = deg2 ? 1.0f : finish(occ2);
                    occ = occ + 1.0f;                                            // :82
                    occ = occ - fv2;                                             // :83
                }
            }
        }
    }
    return finish(occ);
}

} // namespace orc

using namespace orc;

extern "C" {

int orc_env_set_triangle_colors(orc_env* e, int n, const float* rgb3) {
    if (!e || n < 0 || (n > 0 && !rgb3)) return -1;
    e->tri_color.assign(rgb3, rgb3 + 3 * (size_t)n);
    return 0;
}

int orc_bsp_set(orc_env* e, int n_nodes, const int32_t* node_plane, const int32_t* node_children2, int n_planes,
                const float* plane_normal3, const float* plane_dist, const int32_t* plane_type, int n_leafs,
                const int32_t* leaf_cluster, const int32_t* leaf_area, int n_areas) {
    if (!e || n_nodes < 0 || n_planes < 0 || n_leafs < 1) return -1;
    for (int i = 0; i < n_nodes; i++) {
        if (node_plane[i] < 0 || node_plane[i] >= n_planes) return -1;
        for (int c = 0; c < 2; c++) {
            int ch = node_children2[2 * i + c];
            if (ch >= n_nodes || -1 - ch >= n_leafs) return -1;
            if (ch >= 0 && ch <= i) return -1;             // children after parents: no cycles
        }
    }
    Bsp& B = e->bsp;
    B.node_plane.assign(node_plane, node_plane + n_nodes);
    B.node_children.assign(node_children2, node_children2 + 2 * (size_t)n_nodes);
    B.plane_normal.assign(plane_normal3, plane_normal3 + 3 * (size_t)n_planes);
    B.plane_dist.assign(plane_dist, plane_dist + n_planes);
    B.plane_type.assign(plane_type, plane_type + n_planes);
    B.leaf_cluster.assign(leaf_cluster, leaf_cluster + n_leafs);
    B.leaf_area.assign(leaf_area, leaf_area + n_leafs);
    B.n_areas = n_areas;
    B.set = true;
    e->cams = SkyCameras();
    e->cams.area_camera.assign(n_areas > 0 ? n_areas : 0, -1);
    return 0;
}

int orc_point_leafnum(orc_env* e, int64_t n, const float* pts3, int32_t* leaf_out) {
    if (!e || !e->bsp.set) return -1;
    for (int64_t i = 0; i < n; i++) leaf_out[i] = point_leafnum(e->bsp, pts3 + 3 * i);
    return 0;
}

int orc_cluster_from_point(orc_env* e, int64_t n, const float* pts3, int32_t* cluster_out) {
    if (!e || !e->bsp.set) return -1;
    for (int64_t i = 0; i < n; i++) {
        int l = e->bsp.node_plane.empty() ? 0 : point_in_leaf(e->bsp, 0, pts3 + 3 * i);
        cluster_out[i] = e->bsp.leaf_cluster[l];         // ClusterFromPoint, point.go:10-12
    }
    return 0;
}

// rad/cameras/skycamera.go:10-49
int orc_sky_cameras_set(orc_env* e, int n, const float* origin3, const float* scale) {
    if (!e || !e->bsp.set || n < 0) return -1;
    SkyCameras& S = e->cams;
    S = SkyCameras();
    S.area_camera.assign(e->bsp.n_areas > 0 ? e->bsp.n_areas : 0, -1);               // :12-14
    for (int i = 0; i < n; i++) {
        int leaf = point_leafnum(e->bsp, origin3 + 3 * i);                           // :27
        int area = -1;
        if (leaf >= 0 && leaf < (int)e->bsp.leaf_area.size()) area = e->bsp.leaf_area[leaf];   // :30-32
        float sc = scale[i];
        if (sc > 0.0f) {                                                             // :35
            S.origin.insert(S.origin.end(), origin3 + 3 * i, origin3 + 3 * i + 3);
            S.sky_to_world.push_back(sc);
            S.world_to_sky.push_back(1.0f / sc);
            S.area.push_back(area);
            if (area >= 0 && area < e->bsp.n_areas) S.area_camera[area] = (int)S.area.size() - 1;   // :41-43
        }
    }
    return (int)S.area.size();
}

int orc_sky_cameras_get(orc_env* e, int32_t* cam_area, float* world_to_sky, int32_t* area_camera) {
    if (!e) return -1;
    const SkyCameras& S = e->cams;
    for (size_t i = 0; i < S.area.size(); i++) { if (cam_area) cam_area[i] = S.area[i]; if (world_to_sky) world_to_sky[i] = S.world_to_sky[i]; }
    if (area_camera) for (size_t i = 0; i < S.area_camera.size(); i++) area_camera[i] = S.area_camera[i];
    return (int)S.area.size();
}

int orc_test_lines_sky(orc_env* e, int64_t n, const float* start_soa, const float* stop_soa, int flags,
                       int32_t static_prop_to_skip, float* fraction_visible, int threads) {
    if (!e || !e->built) return -1;
    const int32_t skip_id = ORC_TRACE_ID_STATICPROP | static_prop_to_skip;          // testline.go:36
    const float *sx = start_soa, *sy = start_soa + n, *sz = start_soa + 2 * n;
    const float *ex = stop_soa, *ey = stop_soa + n, *ez = stop_soa + 2 * n;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < n; i++) {
        float a[3] = {sx[i], sy[i], sz[i]}, b[3] = {ex[i], ey[i], ez[i]};
        int64_t li = (flags & ORC_TL_PACKET_LEAF) ? (i & ~(int64_t)3) : i;
        float lp[3] = {sx[li], sy[li], sz[li]};
        fraction_visible[i] = test_line_sky1(e, a, b, lp, flags, skip_id);
    }
    return 0;
}

// rad/lightmap/lightmap.go:425-451
int orc_leafs_trace_to_sky(orc_env* e, int n_leafs, const int16_t* mins3, const int16_t* maxs3, int n_dirs,
                           const float* dirs3, uint8_t* can_out, int threads) {
    if (!e || !e->built) return -1;
    const int32_t skip_id = ORC_TRACE_ID_STATICPROP | -1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 0 ? threads : 1)
    for (int l = 0; l < n_leafs; l++) {
        float c[3];
        for (int k = 0; k < 3; k++) c[k] = (float)((int)mins3[3 * l + k] + (int)maxs3[3 * l + k]) * 0.5f;   // :430
        uint8_t can = 0;
        for (int j = 0; j < n_dirs && !can; j++) {
            float b[3];
            for (int k = 0; k < 3; k++) b[k] = (dirs3[3 * j + k] * (-MAX_TRACE_LENGTH)) + c[k];             // :440-441
            float fv = test_line_sky1(e, c, b, c, ORC_TL_CAN_RECURSE, skip_id);                             // :444
            if (fv > 0.0f) can = 1;                                                                          // :445-447
        }
        can_out[l] = can;
    }
    return 0;
}

// rad/lightmap/vis.go:54-94 with the App. A #23 intent: standard Quake/Source PVS run-length code --
// a non-zero byte is copied, a zero byte is followed by a repeat count of zero bytes; stops after
// `row` = (numclusters+7)>>3 output bytes.  Returns the number of input bytes consumed, <0 on error.
int64_t orc_decompress_vis(const uint8_t* in, int64_t in_len, int n_clusters, uint8_t* out_row) {
    const int row = (n_clusters + 7) >> 3;
    int64_t ip = 0;
    int op = 0;
    while (op < row) {
        if (ip >= in_len) return -1;
        if (in[ip]) { out_row[op++] = in[ip++]; continue; }
        if (ip + 1 >= in_len) return -1;
        int c = in[ip + 1];
        if (c == 0) return -2;                            // "DecompressVis: 0 repeat" (:80-82)
        ip += 2;
        if (op + c > row) c = row - op;                   // overrun is clamped with a warning (:85-88)
        while (c-- > 0) out_row[op++] = 0;
    }
    return ip;
}

} // extern "C"
