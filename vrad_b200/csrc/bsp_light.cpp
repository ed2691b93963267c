// bsp_light.cpp -- smoothing normals, luxel sample points and the lighting lump: what sits
// between the BSP faces and K3's luxel arrays on the way in, and between the baked light and the .bsp on the way out.
// Pure host code.
//
// Reference map
//   rad/lightmap/lightmap.go:37-216      PairEdges: faces per vertex, neighbour lists, smoothed vertex normals
//   rad/lightmap/lightmap.go:218-265     SaveVertexNormals
//   rad/lightmap/lightmap.go:267-282     EdgeVertex
//   rad/lightmap/normallist.go:12-49     NormalList.FindOrAddNormal (8x8x8 grid over [-1,1]^3)
//   rad/lightmap/normallist.go:52-143    GetPhongNormal
//   common/types/face.go:5-13            FaceNeighbour
//   vmath/polygon/face.go:5-17           ValidDispFace
// Intent adopted where the literal text is defective: lightmap.go:74 walks len(Edges) instead of the face's edges;
// normallist.go:40 compares pointers and :47 returns the length after the append -- equal means
// squared distance < 1e-5 and the index of the new element is returned; lightmap.go:258-264 `make(..., len)` then append
// doubles the list -- the lump holds each unique normal once.
// UNCITED (absent from the reference): luxel sample points (upstream InitLightinfo / CalcPoints without the
// off-face sample nudging) and the lighting lump layout (upstream PrecompLightmapOffsets / FinalLightFace).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/vrad_bsp.h"
#include "rgbexp.cuh"

namespace vrad { void set_error(const char* fmt, ...); }

namespace {

constexpr uint32_t kSmoothingGroupHardEdge = 0xff000000u;     // lightmap.go:23
constexpr int kMaxNeighbours = 64;                            // lightmap.go:41

struct V3 { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
inline V3 sub(const V3& a, const V3& b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
inline V3 add(const V3& a, const V3& b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
inline V3 scale(const V3& a, float s) { return {{a[0] * s, a[1] * s, a[2] * s}}; }
inline float dot(const V3& a, const V3& b) { return ((a[0] * b[0]) + (a[1] * b[1])) + (a[2] * b[2]); }
inline V3 cross(const V3& a, const V3& b) { return {{(a[1] * b[2]) - (a[2] * b[1]), (a[2] * b[0]) - (a[0] * b[2]), (a[0] * b[1]) - (a[1] * b[0])}}; }
inline float len(const V3& a) { return (float)std::sqrt((double)(((a[0] * a[0]) + (a[1] * a[1])) + (a[2] * a[2]))); }
inline V3 normalize(const V3& a) { const float l = 1.0f / len(a); return {{a[0] * l, a[1] * l, a[2] * l}}; }
inline V3 load3(const float* p) { return {{p[0], p[1], p[2]}}; }

inline int face_vertex(const vrad_bsp_lumps& L, const vrad_dface& f, int k) {    // EdgeVertex, lightmap.go:267-282
    const int n = f.numedges;
    if (k < 0) k += n; else if (k >= n) k %= n;
    const int32_t se = L.surfedges[f.firstedge + k];
    return se < 0 ? L.edges[-(int64_t)se].v[1] : L.edges[se].v[0];
}
inline bool valid_disp_face(const vrad_dface& f) { return f.dispinfo != -1 && f.numedges == 4; }

inline bool face_is_lit(const vrad_bsp_lumps& L, const vrad_dface& f) {
    return !(L.texinfo[f.texinfo].flags & (VRAD_SURF_SKY | VRAD_SURF_NOLIGHT));
}
inline bool face_is_bumped(const vrad_bsp_lumps& L, const vrad_dface& f) { return (L.texinfo[f.texinfo].flags & VRAD_SURF_BUMPLIGHT) != 0; }

}  // namespace

extern "C" int vrad_bsp_pair_edges(const vrad_bsp_lumps* Lp, float smoothing_threshold, float* vertex_normals3,
                                   int32_t* neighbour_first, int32_t* neighbours, int max_neighbours) {
    if (!Lp || !vertex_normals3 || !neighbour_first) { vrad::set_error("vrad_bsp_pair_edges: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    // faces per vertex, each face once, in face order (lightmap.go:48-87) -- as CSR
    std::vector<int32_t> vfirst((size_t)L.n_vertexes + 1, 0);
    std::vector<int32_t> last_face((size_t)L.n_vertexes, -1);
    for (int i = 0; i < L.n_faces; i++)
        for (int j = 0; j < L.faces[i].numedges; j++) { const int v = face_vertex(L, L.faces[i], j); if (last_face[v] != i) { last_face[v] = i; vfirst[v + 1]++; } }
    for (int v = 0; v < L.n_vertexes; v++) vfirst[v + 1] += vfirst[v];
    std::vector<int32_t> vfaces((size_t)vfirst[L.n_vertexes]), fill(vfirst.begin(), vfirst.end() - 1);
    std::fill(last_face.begin(), last_face.end(), -1);
    for (int i = 0; i < L.n_faces; i++)
        for (int j = 0; j < L.faces[i].numedges; j++) { const int v = face_vertex(L, L.faces[i], j); if (last_face[v] != i) { last_face[v] = i; vfaces[fill[v]++] = i; } }

    size_t out = 0;
    int32_t nn_total = 0;
    for (int i = 0; i < L.n_faces; i++) {
        const vrad_dface& f = L.faces[i];
        const V3 face_normal = load3(L.planes[f.planenum].normal);          // :96
        const bool has_disp = valid_disp_face(f);
        int tmp[kMaxNeighbours], nn = 0;
        neighbour_first[i] = nn_total;
        for (int j = 0; j < f.numedges; j++) {
            V3 acc = {{0, 0, 0}};
            const int v = face_vertex(L, f, j);
            for (int32_t k = vfirst[v]; k < vfirst[v + 1]; k++) {
                const int o = vfaces[k];
                if (o == i) continue;                                        // skip self
                const vrad_dface& of = L.faces[o];
                if (!has_disp && valid_disp_face(of)) continue;              // :138-140
                const V3 nb = load3(L.planes[of.planenum].normal);
                const float cos_angle = dot(nb, face_normal);
                if (has_disp) acc = add(acc, nb);                            // always smooth with and against a displacement
                else if (f.smoothing_groups == 0 && of.smoothing_groups == 0) {
                    if (cos_angle >= smoothing_threshold) acc = add(acc, nb); else continue;
                } else {
                    const uint32_t g = f.smoothing_groups & of.smoothing_groups;
                    if (g & kSmoothingGroupHardEdge) continue;
                    if (g != 0) acc = add(acc, nb); else continue;
                }
                int m = 0;
                for (; m < nn; m++) if (tmp[m] == o) break;
                if (m >= nn) {
                    if (nn >= kMaxNeighbours) { vrad::set_error("Stack overflow in neighbors (face %d)", i); return VRAD_E_INVALID; }
                    tmp[nn++] = o;
                }
            }
            acc = normalize(add(acc, face_normal));                          // fixup (:209-213)
            for (int k = 0; k < 3; k++) vertex_normals3[3 * out + k] = acc[k];
            out++;
        }
        if (neighbours) {
            if (nn_total + nn > max_neighbours) { vrad::set_error("vrad_bsp_pair_edges: more than %d neighbour entries", max_neighbours); return VRAD_E_NOMEM; }
            for (int m = 0; m < nn; m++) neighbours[nn_total + m] = tmp[m];
        }
        nn_total += nn;
    }
    neighbour_first[L.n_faces] = nn_total;
    return VRAD_OK;
}

extern "C" int vrad_bsp_save_vertex_normals(int n_face_vertices, const float* vertex_normals3, int max_normals,
                                            float* normals3, uint16_t* indices, int* n_normals_out) {
    if (n_face_vertices < 0 || (n_face_vertices && !vertex_normals3) || !n_normals_out) { vrad::set_error("vrad_bsp_save_vertex_normals: bad arguments"); return VRAD_E_INVALID; }
    constexpr int kSub = 8;                                                  // numSubDivs, normallist.go:11
    std::vector<int32_t> grid[kSub][kSub][kSub];
    std::vector<V3> list;
    for (int i = 0; i < n_face_vertices; i++) {
        const V3 nv = load3(vertex_normals3 + 3 * (size_t)i);
        int gi[3];
        for (int d = 0; d < 3; d++) {
            // (int)(((n + 1) * 0.5) * numSubDivs - 0.000001): Go evaluates the untyped constants in fp32 here
            int g = (int)(((nv[d] + 1.0f) * 0.5f) * (float)kSub - 0.000001f);
            g = g < kSub - 1 ? g : kSub - 1;                                 // math.Min(g, numSubDivs) would index out of range at 8
            g = g > 0 ? g : 0;
            gi[d] = g;
        }
        std::vector<int32_t>& cell = grid[gi[0]][gi[1]][gi[2]];
        int found = -1;
        for (int32_t idx : cell) {
            const V3 d = sub(list[idx], nv);
            if (dot(d, d) < 0.00001f) { found = idx; break; }
        }
        if (found < 0) { found = (int)list.size(); cell.push_back(found); list.push_back(nv); }
        if (found > 0xffff) { vrad::set_error("g_numvertnormals > MAX_MAP_VERTNORMALS"); return VRAD_E_INVALID; }
        if (indices) indices[i] = (uint16_t)found;
    }
    *n_normals_out = (int)list.size();
    if (normals3) {
        if ((int)list.size() > max_normals) { vrad::set_error("vrad_bsp_save_vertex_normals: %zu normals, room for %d", list.size(), max_normals); return VRAD_E_NOMEM; }
        for (size_t i = 0; i < list.size(); i++) for (int k = 0; k < 3; k++) normals3[3 * i + k] = list[i][k];
    }
    return VRAD_OK;
}

extern "C" int vrad_bsp_phong_normals(const vrad_bsp_lumps* Lp, float smoothing_threshold, const float* vertex_normals3, const float* centroids3,
                                      int64_t n, const int32_t* face, const float* points3, float* normals3_out) {
    if (!Lp || !vertex_normals3 || !centroids3 || n < 0 || (n && (!face || !points3 || !normals3_out))) { vrad::set_error("vrad_bsp_phong_normals: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    std::vector<int64_t> first((size_t)L.n_faces + 1, 0);                   // offset of each face's block in vertex_normals3
    for (int i = 0; i < L.n_faces; i++) first[i + 1] = first[i] + L.faces[i].numedges;
    for (int64_t p = 0; p < n; p++) {
        const int fn = face[p];
        if (fn < 0 || fn >= L.n_faces) { vrad::set_error("vrad_bsp_phong_normals: point %lld names face %d of %d", (long long)p, fn, L.n_faces); return VRAD_E_INVALID; }
        const vrad_dface& f = L.faces[fn];
        const V3 face_normal = load3(L.planes[f.planenum].normal);
        V3 result = face_normal;
        if (smoothing_threshold != 1.0f) {
            const V3 centre = load3(centroids3 + 3 * (size_t)fn), spot = load3(points3 + 3 * p);
            const float* vn = vertex_normals3 + 3 * first[fn];
            for (int j = 0; j < f.numedges; j++) {
                const V3 n1 = load3(vn + 3 * j), n2 = load3(vn + 3 * ((j + 1) % f.numedges));
                const V3 p1 = load3(L.vertexes3 + 3 * (size_t)face_vertex(L, f, j)), p2 = load3(L.vertexes3 + 3 * (size_t)face_vertex(L, f, j + 1));
                const V3 v1 = sub(p1, centre), v2 = sub(p2, centre), vspot = sub(spot, centre);
                const float aa = dot(v1, v1), bb = dot(v2, v2), ab = dot(v1, v2);
                const float a1 = (bb * dot(v1, vspot) - ab * dot(vspot, v2)) / (aa * bb - ab * ab);
                const float a2 = (dot(vspot, v2) - a1 * ab) / bb;
                if (a1 >= 0.0f && a2 >= 0.0f) {                              // inside the wedge centre-p1-p2
                    const float s = (1.0f - a1) - a2;                        // Go: the untyped 1.0 takes a1's type
                    V3 ph = scale(face_normal, s);
                    ph = add(ph, scale(n1, a1));
                    ph = add(ph, scale(n2, a2));
                    if (1.0e-20f > len(ph)) { vrad::set_error("Phong normal length out of bounds (face %d)", fn); return VRAD_E_INVALID; }
                    result = normalize(ph);
                    break;
                }
            }
        }
        for (int k = 0; k < 3; k++) normals3_out[3 * p + k] = result[k];
    }
    return VRAD_OK;
}

extern "C" int vrad_bsp_layout_lighting(const vrad_bsp_lumps* Lp, const int32_t* mins2, const int32_t* size2, vrad_dface* faces_out,
                                        int64_t* luxel_first, int64_t* lump_bytes) {
    if (!Lp || !mins2 || !size2 || !luxel_first || !lump_bytes) { vrad::set_error("vrad_bsp_layout_lighting: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    int64_t lux = 0, bytes = 0;
    for (int i = 0; i < L.n_faces; i++) {
        const vrad_dface& f = L.faces[i];
        luxel_first[i] = lux;
        vrad_dface o = f;
        for (int k = 0; k < 2; k++) { o.lm_mins[k] = mins2[2 * (size_t)i + k]; o.lm_size[k] = size2[2 * (size_t)i + k]; }
        if (!face_is_lit(L, f) || size2[2 * (size_t)i] < 0 || size2[2 * (size_t)i + 1] < 0) {
            o.lightofs = -1;
            for (int k = 0; k < 4; k++) o.styles[k] = 255;
        } else {
            // world.CalcFaceExtents (rad/world/face.go:66-87) logs such a face and carries on with the fatal error commented out; upstream
            // stops ("Bad surface extents - surface is too big to have a lightmap").  A lightmap larger than the format allows cannot be
            // laid out, so this is where the job stops here as well.
            if (size2[2 * (size_t)i] > 126 || size2[2 * (size_t)i + 1] > 126) {
                vrad::set_error("Bad surface extents - face %d is too big to have a lightmap (%d x %d luxels, limit 126)", i, size2[2 * (size_t)i], size2[2 * (size_t)i + 1]);
                return VRAD_E_INVALID;
            }
            const int64_t samples = (int64_t)(size2[2 * (size_t)i] + 1) * (size2[2 * (size_t)i + 1] + 1) * (face_is_bumped(L, f) ? 4 : 1);
            bytes += 4;                                                      // the style's average colour sits right before the samples
            if (bytes + samples * 4 > INT32_MAX) { vrad::set_error("vrad_bsp_layout_lighting: lighting lump would exceed 2 GiB at face %d", i); return VRAD_E_INVALID; }
            o.lightofs = (int32_t)bytes;
            o.styles[0] = 0; o.styles[1] = o.styles[2] = o.styles[3] = 255;
            bytes += samples * 4;
            lux += samples;
        }
        if (faces_out) faces_out[i] = o;
    }
    luxel_first[L.n_faces] = lux;
    *lump_bytes = bytes;
    return VRAD_OK;
}

extern "C" int vrad_bsp_face_luxels(const vrad_bsp_lumps* Lp, const int32_t* mins2, const int32_t* size2, const float* face_origins3,
                                    const int64_t* luxel_first, float* pos3, float* normal3, int32_t* luxel_face) {
    if (!Lp || !mins2 || !size2 || !luxel_first || !pos3 || !normal3) { vrad::set_error("vrad_bsp_face_luxels: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    for (int i = 0; i < L.n_faces; i++) {
        const int64_t count = luxel_first[i + 1] - luxel_first[i];
        if (count == 0) continue;
        const vrad_dface& f = L.faces[i];
        const vrad_texinfo& tx = L.texinfo[f.texinfo];
        const int w = size2[2 * (size_t)i] + 1, h = size2[2 * (size_t)i + 1] + 1;
        const int blocks = face_is_bumped(L, f) ? 4 : 1;
        if (count != (int64_t)w * h * blocks) { vrad::set_error("vrad_bsp_face_luxels: face %d has %lld luxels laid out, extents say %lld", i, (long long)count, (long long)w * h * blocks); return VRAD_E_INVALID; }
        const vrad_dplane& pl = L.planes[f.planenum];
        const V3 nrm = load3(pl.normal);
        const float dist = pl.dist;
        const V3 lv0 = load3(tx.lightmap_vecs[0]), lv1 = load3(tx.lightmap_vecs[1]);
        // a normal to the texture axes: points move along it without changing their (s,t); flipped towards the face normal
        V3 texnormal = normalize(cross(lv1, lv0));
        float distscale = dot(texnormal, nrm);
        if (distscale == 0.0f || distscale != distscale) { vrad::set_error("Texture axis perpendicular to face %d", i); return VRAD_E_INVALID; }
        if (distscale < 0.0f) { distscale = -distscale; texnormal = {{-texnormal[0], -texnormal[1], -texnormal[2]}}; }
        distscale = 1.0f / distscale;
        V3 l2w[2];
        l2w[0] = cross(lv1, nrm); l2w[0] = scale(l2w[0], 1.0f / dot(l2w[0], lv0));
        l2w[1] = cross(lv0, nrm); l2w[1] = scale(l2w[1], 1.0f / dot(l2w[1], lv1));
        V3 org;
        for (int k = 0; k < 3; k++) org[k] = (-(tx.lightmap_vecs[0][3] * l2w[0][k])) - (tx.lightmap_vecs[1][3] * l2w[1][k]);
        float d = dot(org, nrm) - dist;
        d *= distscale;
        for (int k = 0; k < 3; k++) org[k] = org[k] + ((-d) * texnormal[k]);
        if (face_origins3) org = add(org, load3(face_origins3 + 3 * (size_t)i));
        float bump9[9];
        if (blocks == 4) {
            const int rc = vrad_bump_normals(tx.texture_vecs[0], tx.texture_vecs[1], nrm.v, nrm.v, bump9);
            if (rc) return rc;
        }
        int64_t o = luxel_first[i];
        for (int b = 0; b < blocks; b++) {
            const V3 bn = b == 0 ? nrm : load3(bump9 + 3 * (b - 1));
            for (int t = 0; t < h; t++)
                for (int s = 0; s < w; s++, o++) {
                    const float us = (float)(mins2[2 * (size_t)i] + s), ut = (float)(mins2[2 * (size_t)i + 1] + t);
                    for (int k = 0; k < 3; k++) {
                        const float surf = (org[k] + (us * l2w[0][k])) + (ut * l2w[1][k]);
                        pos3[3 * o + k] = surf + nrm[k];                    // one unit off the surface
                        normal3[3 * o + k] = bn[k];
                    }
                    if (luxel_face) luxel_face[o] = i;
                }
        }
    }
    return VRAD_OK;
}

// ---- samples inside the face ----------------------------------------------------------------------------------------------------
namespace {

struct P2 { float s, t; };

// Sutherland-Hodgman against one axis-aligned half-plane: keep coord(axis) >= bound (keep_greater) or <= bound
int clip_axis(const P2* in, int n, int axis, float bound, bool keep_greater, P2* out) {
    int m = 0;
    for (int i = 0; i < n; i++) {
        const P2 a = in[i], b = in[(i + 1) % n];
        const float ca = axis ? a.t : a.s, cb = axis ? b.t : b.s;
        const bool ina = keep_greater ? ca >= bound : ca <= bound, inb = keep_greater ? cb >= bound : cb <= bound;
        if (ina) out[m++] = a;
        if (ina != inb) {
            const float u = (bound - ca) / (cb - ca);
            P2 x = {a.s + (u * (b.s - a.s)), a.t + (u * (b.t - a.t))};
            if (axis) x.t = bound; else x.s = bound;
            out[m++] = x;
        }
    }
    return m;
}

// area-weighted centroid of a polygon (fan from point 0); false when the polygon has no area
bool polygon_centroid(const P2* p, int n, P2& c) {
    float area2 = 0.0f, cs = 0.0f, ct = 0.0f;
    for (int i = 1; i + 1 < n; i++) {
        const float ax = p[i].s - p[0].s, ay = p[i].t - p[0].t, bx = p[i + 1].s - p[0].s, by = p[i + 1].t - p[0].t;
        const float cr = (ax * by) - (ay * bx);                              // twice the signed area of the fan triangle
        area2 += cr;
        cs += cr * ((p[0].s + p[i].s) + p[i + 1].s);
        ct += cr * ((p[0].t + p[i].t) + p[i + 1].t);
    }
    if (!(std::fabs(area2) > 1e-6f)) return false;
    c.s = cs / (3.0f * area2); c.t = ct / (3.0f * area2);
    return true;
}

P2 nearest_on_polygon(const P2* p, int n, P2 q) {
    P2 best = p[0];
    float best_d = 3.0e38f;
    for (int i = 0; i < n; i++) {
        const P2 a = p[i], b = p[(i + 1) % n];
        const float ex = b.s - a.s, ey = b.t - a.t;
        const float len2 = (ex * ex) + (ey * ey);
        float u = len2 > 0.0f ? (((q.s - a.s) * ex) + ((q.t - a.t) * ey)) / len2 : 0.0f;
        u = u < 0.0f ? 0.0f : (u > 1.0f ? 1.0f : u);
        const P2 x = {a.s + (u * ex), a.t + (u * ey)};
        const float d = ((x.s - q.s) * (x.s - q.s)) + ((x.t - q.t) * (x.t - q.t));
        if (d < best_d) { best_d = d; best = x; }
    }
    return best;
}

}  // namespace

extern "C" int vrad_bsp_place_samples(const vrad_bsp_lumps* Lp, const int32_t* mins2, const int32_t* size2, const float* face_origins3,
                                      const int64_t* luxel_first, float* pos3_inout, float* luxel_st2_out) {
    if (!Lp || !mins2 || !size2 || !luxel_first || !pos3_inout) { vrad::set_error("vrad_bsp_place_samples: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    for (int i = 0; i < L.n_faces; i++) {
        const int64_t count = luxel_first[i + 1] - luxel_first[i];
        if (count == 0) continue;
        const vrad_dface& f = L.faces[i];
        const vrad_texinfo& tx = L.texinfo[f.texinfo];
        const int w = size2[2 * (size_t)i] + 1, h = size2[2 * (size_t)i + 1] + 1;
        const int blocks = face_is_bumped(L, f) ? 4 : 1;
        if (count != (int64_t)w * h * blocks) { vrad::set_error("vrad_bsp_place_samples: face %d has %lld luxels laid out, extents say %lld", i, (long long)count, (long long)w * h * blocks); return VRAD_E_INVALID; }
        if (f.numedges < 3 || f.numedges > 64) { vrad::set_error("vrad_bsp_place_samples: face %d has %d edges", i, f.numedges); return VRAD_E_INVALID; }
        // the face winding in luxel space, relative to the lightmap mins
        P2 poly[64];
        for (int k = 0; k < f.numedges; k++) {
            const float* v = L.vertexes3 + 3 * (size_t)face_vertex(L, f, k);
            float st[2];
            for (int a = 0; a < 2; a++) {
                const float* lv = tx.lightmap_vecs[a];
                st[a] = ((((v[0] * lv[0]) + (v[1] * lv[1])) + (v[2] * lv[2])) + lv[3]) - (float)mins2[2 * (size_t)i + a];
            }
            poly[k] = {st[0], st[1]};
        }
        // luxel -> world, as vrad_bsp_face_luxels builds it
        const V3 nrm = load3(L.planes[f.planenum].normal);
        const V3 lv0 = load3(tx.lightmap_vecs[0]), lv1 = load3(tx.lightmap_vecs[1]);
        V3 l2w[2];
        l2w[0] = cross(lv1, nrm); l2w[0] = scale(l2w[0], 1.0f / dot(l2w[0], lv0));
        l2w[1] = cross(lv0, nrm); l2w[1] = scale(l2w[1], 1.0f / dot(l2w[1], lv1));
        for (int t = 0; t < h; t++)
            for (int s = 0; s < w; s++) {
                P2 a[72], b[72];
                // a cell that lies wholly inside the (convex) face keeps its grid point exactly
                bool whole = true;
                for (int cx = 0; cx < 2 && whole; cx++)
                    for (int cy = 0; cy < 2 && whole; cy++) {
                        const P2 q = {(float)s + (cx ? 0.5f : -0.5f), (float)t + (cy ? 0.5f : -0.5f)};
                        bool pos_side = false, neg_side = false;
                        for (int k = 0; k < f.numedges; k++) {
                            const P2 e0 = poly[k], e1 = poly[(k + 1) % f.numedges];
                            const float cr = ((e1.s - e0.s) * (q.t - e0.t)) - ((e1.t - e0.t) * (q.s - e0.s));
                            if (cr > 0.0f) pos_side = true; else if (cr < 0.0f) neg_side = true;
                        }
                        if (pos_side && neg_side) whole = false;
                    }
                if (whole) { if (luxel_st2_out) for (int blk = 0; blk < blocks; blk++) { const int64_t o = luxel_first[i] + (int64_t)blk * w * h + (int64_t)t * w + s; luxel_st2_out[2 * o] = (float)s; luxel_st2_out[2 * o + 1] = (float)t; } continue; }
                int n = clip_axis(poly, f.numedges, 0, (float)s - 0.5f, true, a);
                n = clip_axis(a, n, 0, (float)s + 0.5f, false, b);
                n = clip_axis(b, n, 1, (float)t - 0.5f, true, a);
                n = clip_axis(a, n, 1, (float)t + 0.5f, false, b);
                P2 c;
                if (n < 3 || !polygon_centroid(b, n, c)) c = nearest_on_polygon(poly, f.numedges, P2{(float)s, (float)t});
                const float ds = c.s - (float)s, dt = c.t - (float)t;      // how far the sample moved off the grid point, in luxels
                for (int blk = 0; blk < blocks; blk++) {
                    const int64_t o = luxel_first[i] + (int64_t)blk * w * h + (int64_t)t * w + s;
                    for (int k = 0; k < 3; k++) pos3_inout[3 * o + k] = (pos3_inout[3 * o + k] + (ds * l2w[0][k])) + (dt * l2w[1][k]);
                    if (luxel_st2_out) { luxel_st2_out[2 * o] = c.s; luxel_st2_out[2 * o + 1] = c.t; }
                }
            }
    }
    (void)face_origins3;
    return VRAD_OK;
}

extern "C" int vrad_color_to_rgbexp32(int64_t n, const float* rgb3, vrad_color_rgbexp32* out) {
    if (n < 0 || (n && (!rgb3 || !out))) { vrad::set_error("vrad_color_to_rgbexp32: bad arguments"); return VRAD_E_INVALID; }
    for (int64_t i = 0; i < n; i++) {
        const vrad::RgbExp c = vrad::pack_rgbexp32(rgb3[3 * i], rgb3[3 * i + 1], rgb3[3 * i + 2]);
        out[i].r = c.r; out[i].g = c.g; out[i].b = c.b; out[i].exponent = c.e;
    }
    return VRAD_OK;
}

extern "C" int vrad_color_from_rgbexp32(int64_t n, const vrad_color_rgbexp32* in, float* rgb3) {
    if (n < 0 || (n && (!in || !rgb3))) { vrad::set_error("vrad_color_from_rgbexp32: bad arguments"); return VRAD_E_INVALID; }
    for (int64_t i = 0; i < n; i++) {
        const vrad::RgbExp c = {in[i].r, in[i].g, in[i].b, in[i].exponent};
        vrad::unpack_rgbexp32(c, rgb3[3 * i], rgb3[3 * i + 1], rgb3[3 * i + 2]);
    }
    return VRAD_OK;
}

extern "C" int vrad_bsp_pack_lighting(const vrad_bsp_lumps* Lp, const int64_t* luxel_first, const vrad_color_rgbexp32* colors,
                                      uint8_t* lump_out, int64_t lump_bytes) {
    if (!Lp || !luxel_first || !colors || (lump_bytes && !lump_out) || lump_bytes < 0) { vrad::set_error("vrad_bsp_pack_lighting: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    std::memset(lump_out, 0, (size_t)lump_bytes);
    for (int i = 0; i < L.n_faces; i++) {
        const int64_t count = luxel_first[i + 1] - luxel_first[i];
        const vrad_dface& f = L.faces[i];
        if (count == 0 || f.lightofs < 0) continue;
        if (f.lightofs < 4 || (int64_t)f.lightofs + count * 4 > lump_bytes) { vrad::set_error("vrad_bsp_pack_lighting: face %d (%lld luxels at %d) does not fit the %lld-byte lump", i, (long long)count, f.lightofs, (long long)lump_bytes); return VRAD_E_INVALID; }
        std::memcpy(lump_out + f.lightofs, colors + luxel_first[i], (size_t)count * 4);
        // average colour of the flat block (the first (w+1)(h+1) samples), stored before the samples
        const int64_t flat = face_is_bumped(L, f) ? count / 4 : count;
        float sum[3] = {0, 0, 0};
        for (int64_t k = 0; k < flat; k++) {
            const vrad_color_rgbexp32& c = colors[luxel_first[i] + k];
            float r, g, b;
            vrad::unpack_rgbexp32({c.r, c.g, c.b, c.exponent}, r, g, b);
            sum[0] += r; sum[1] += g; sum[2] += b;
        }
        const float inv = 1.0f / (float)flat;
        const vrad::RgbExp avg = vrad::pack_rgbexp32(sum[0] * inv, sum[1] * inv, sum[2] * inv);
        uint8_t* a = lump_out + f.lightofs - 4;
        a[0] = avg.r; a[1] = avg.g; a[2] = avg.b; a[3] = (uint8_t)avg.e;
    }
    return VRAD_OK;
}

extern "C" int vrad_luxel_nearest_patch(int64_t n, const int32_t* luxel_face, const float* pos3, int n_patches, const int32_t* patch_face,
                                        const int32_t* child1, const float* origin3, int32_t* patch_out) {
    if (n < 0 || n_patches < 0 || (n && (!luxel_face || !pos3 || !patch_out)) || (n_patches && (!patch_face || !origin3))) {
        vrad::set_error("vrad_luxel_nearest_patch: bad arguments"); return VRAD_E_INVALID;
    }
    int32_t max_face = -1;
    for (int i = 0; i < n_patches; i++) if (patch_face[i] > max_face) max_face = patch_face[i];
    // leaf patches per face, CSR, ascending patch index
    std::vector<int32_t> first((size_t)max_face + 2, 0), list;
    for (int i = 0; i < n_patches; i++) if (patch_face[i] >= 0 && (!child1 || child1[i] == -1)) first[patch_face[i] + 1]++;
    for (int f = 0; f <= max_face; f++) first[f + 1] += first[f];
    list.resize((size_t)first[max_face + 1]);
    std::vector<int32_t> fill(first.begin(), first.end() - 1);
    for (int i = 0; i < n_patches; i++) if (patch_face[i] >= 0 && (!child1 || child1[i] == -1)) list[fill[patch_face[i]]++] = i;
    for (int64_t l = 0; l < n; l++) {
        const int f = luxel_face[l];
        int32_t best = -1;
        float best_d = 0.0f;
        if (f >= 0 && f <= max_face) {
            const V3 p = load3(pos3 + 3 * l);
            for (int32_t k = first[f]; k < first[f + 1]; k++) {
                const V3 d = sub(load3(origin3 + 3 * (size_t)list[k]), p);
                const float dd = dot(d, d);
                if (best < 0 || dd < best_d) { best = list[k]; best_d = dd; }      // the first of equally near patches wins
            }
        }
        patch_out[l] = best;
    }
    return VRAD_OK;
}
