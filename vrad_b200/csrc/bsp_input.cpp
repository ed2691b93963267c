// bsp_input.cpp -- from BSP lumps to the arrays the kernels take: the callers that sit
// directly in front of the hot path in the reference.  Pure host code; only the RADIAL-leaf branch of
// vrad_bsp_vis_for_light_environment reaches the device (through vrad_leafs_trace_to_sky).
//
// Reference map
//   cmd/tasks/loadbsp/main.go:186-340     ExtractBrushEntityShadowCasters, addBrushes, addBrushToRaytraceEnvironment,
//                                         addBrushesForRayTrace                       -> vrad_bsp_raytrace_triangles
//   cmd/tasks/loadbsp/brush/brush.go:7-36 GetBrushRecursive
//   vmath/polygon/winding.go:25-171       BaseWindingForPlane, ChopWindingInPlace
//   vmath/matrix/mat4.go:10-69            SetupMatrixOrgAngles, Mul4x3
//   rad/patches/build.go:21-65            MakePatches                                 -> vrad_bsp_face_patches
//   rad/world/face.go:92-114, point.go:12-46   WindingFromFace, RemoveColinearPoints
//   rad/patches/face.go:29-230            MakePatchForFace (texinfo-derived fields), IsSky, BaseLightForFace
//   rad/patches/subdivide.go:151-165      PreventSubdivision
//   rad/world/face.go:14-90               CalcFaceExtents                             -> vrad_bsp_face_extents
//   rad/start.go:21-66,100-111            luxel-density rescale, UpdateAllFaceLightmapExtents
//   rad/clustertable/nodes.go:21-36, build.go:9-31   MakeParents, BuildClusterTable
//   rad/lightmap/lightmap.go:284-422      BuildVisForLightEnvironment, MergeDLightVis, PVSCheck
//   rad/lightmap/vis.go:9-94              GetVisCache, DecompressVis (vrad_decompress_vis)
// Intent adopted where the literal text is defective: #15 the sky/displacement side test of
// main.go:255 is inverted; #16 winding.go:36 reads normal[1] for every axis, :78 overwrites Points[0] instead of
// writing Points[3], :149 loops forever, and Go's `case 0:` does not fall through to `case 1:` (:46-51) -- the C
// original is implemented; main.go:217 parses "*N" model names in base 8 and adds 1 -- the caller passes model
// indices; lightmap.go:337 skips the sky leafs a leaf CAN see (`0 != PVSCheck`) -- the test is upstream's `!PVSCheck`; rad/start.go:45 `tmp.Normalize().Len()`
// is always 1 -- the scale is the vector's length.
//
// Arithmetic: fp32 with Go's left-to-right evaluation and no FMA contraction (-ffp-contract=off); mgl32's Len and
// Normalize take the square root in double and round it back.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/vrad_bsp.h"

namespace vrad { void set_error(const char* fmt, ...); }

namespace {

constexpr int   kMaxPointsOnWinding = 64;             // common/constants/constants.go:32
constexpr float kMaxCoordInteger = 16384.0f;          // common/constants/constants.go:6
constexpr int   kMaxLightmapDim = 125;                // MAX_LIGHTMAP_DIM_WITHOUT_BORDER == MAX_DISP_..., constants.go:22,27

struct V3 { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
inline V3 sub(const V3& a, const V3& b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
inline V3 add(const V3& a, const V3& b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
inline V3 scale(const V3& a, float s) { return {{a[0] * s, a[1] * s, a[2] * s}}; }
inline float dot(const V3& a, const V3& b) { return ((a[0] * b[0]) + (a[1] * b[1])) + (a[2] * b[2]); }
inline V3 cross(const V3& a, const V3& b) { return {{(a[1] * b[2]) - (a[2] * b[1]), (a[2] * b[0]) - (a[0] * b[2]), (a[0] * b[1]) - (a[1] * b[0])}}; }
inline float len(const V3& a) { return (float)std::sqrt((double)(((a[0] * a[0]) + (a[1] * a[1])) + (a[2] * a[2]))); }
inline V3 normalize(const V3& a) { const float l = 1.0f / len(a); return {{a[0] * l, a[1] * l, a[2] * l}}; }
inline V3 load3(const float* p) { return {{p[0], p[1], p[2]}}; }

using Winding = std::vector<V3>;

// winding.go:25-83
Winding base_winding_for_plane(const V3& normal, float dist) {
    float best = -1.0f; int x = -1;
    for (int i = 0; i < 3; i++) { const float v = std::fabs(normal[i]); if (v > best) { x = i; best = v; } }
    V3 vup = {{0, 0, 0}};
    if (x == 2) vup[0] = 1.0f; else vup[2] = 1.0f;
    const float v = dot(vup, normal);
    for (int i = 0; i < 3; i++) vup[i] = vup[i] + ((-v) * normal[i]);          // vector.MA(&vup, -v, normal, &vup)
    vup = normalize(vup);
    const V3 org = scale(normal, dist);
    V3 vright = cross(vup, normal);
    vup = scale(vup, kMaxCoordInteger * 4);
    vright = scale(vright, kMaxCoordInteger * 4);
    Winding w(4);
    w[0] = add(sub(org, vright), vup);
    w[1] = add(add(org, vright), vup);
    w[2] = sub(add(org, vright), vup);
    w[3] = sub(sub(org, vright), vup);
    return w;
}

// winding.go:85-171: keep the part in front of (normal, dist).  Returns false when nothing is left.
bool chop_winding_in_place(Winding& w, const V3& normal, float dist, float epsilon) {
    const int n = (int)w.size();
    std::vector<float> dists(n + 1);
    std::vector<int> sides(n + 1);
    int counts[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) {
        float d = dot(w[i], normal);
        d -= dist;
        dists[i] = d;
        sides[i] = d > epsilon ? 0 : (d < -epsilon ? 1 : 2);            // SIDE_FRONT, SIDE_BACK, SIDE_ON
        counts[sides[i]]++;
    }
    sides[n] = sides[0]; dists[n] = dists[0];
    if (!counts[0]) { w.clear(); return false; }
    if (!counts[1]) return true;
    Winding f;
    f.reserve(n + 4);
    for (int i = 0; i < n; i++) {
        const V3& p1 = w[i];
        if (sides[i] == 2) { f.push_back(p1); continue; }
        if (sides[i] == 0) f.push_back(p1);
        if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
        const V3& p2 = w[(i + 1) % n];
        const float t = dists[i] / (dists[i] - dists[i + 1]);
        V3 mid;
        for (int j = 0; j < 3; j++) {
            if (normal[j] == 1.0f) mid[j] = dist;                          // avoid round off error when possible
            else if (normal[j] == -1.0f) mid[j] = -dist;
            else mid[j] = p1[j] + (t * (p2[j] - p1[j]));
        }
        f.push_back(mid);
    }
    w.swap(f);
    return true;
}

// brush.go:7-36 -- leafs in tree order, each brush once (first occurrence)
void brushes_under(const vrad_bsp_lumps& L, int node, std::vector<int>& list, std::vector<uint8_t>& seen) {
    std::vector<int> stack{node};
    while (!stack.empty()) {
        const int n = stack.back(); stack.pop_back();
        if (n < 0) {
            const vrad_dleaf& lf = L.leafs[-1 - n];
            for (int i = 0; i < lf.numleafbrushes; i++) {
                const int b = L.leafbrushes[lf.firstleafbrush + i];
                if (!seen[b]) { seen[b] = 1; list.push_back(b); }
            }
        } else {
            stack.push_back(L.nodes[n].children[1]);                       // children[0]'s subtree comes first
            stack.push_back(L.nodes[n].children[0]);
        }
    }
}

struct Xform { float m[12]; bool identity; };

// mat4.go:10-49 (rows 0..2 of the 4x4), trigonometry in double, entries rounded to fp32
Xform xform_from_origin_angles(const float origin[3], const float angles[3]) {
    const float kPi32 = (float)3.14159265358979323846;
    const double pitch = (double)(angles[0] * kPi32 / 180.0f), yaw = (double)(angles[1] * kPi32 / 180.0f), roll = (double)(angles[2] * kPi32 / 180.0f);
    const double sy = std::sin(yaw), cy = std::cos(yaw), sp = std::sin(pitch), cp = std::cos(pitch), sr = std::sin(roll), cr = std::cos(roll);
    Xform x;
    x.m[0] = (float)(cp * cy);  x.m[1] = (float)(sr * sp * cy + cr * -sy);  x.m[2] = (float)(cr * sp * cy + -sr * -sy);  x.m[3] = origin[0];
    x.m[4] = (float)(cp * sy);  x.m[5] = (float)(sr * sp * sy + cr * cy);   x.m[6] = (float)(cr * sp * sy + -sr * cy);   x.m[7] = origin[1];
    x.m[8] = (float)(-sp);      x.m[9] = (float)(sr * cp);                  x.m[10] = (float)(cr * cp);                  x.m[11] = origin[2];
    x.identity = false;
    return x;
}
inline V3 mul4x3(const Xform& x, const V3& p) {                            // mat4.go:63-69
    if (x.identity) return p;
    return {{((x.m[0] * p[0] + x.m[1] * p[1]) + x.m[2] * p[2]) + x.m[3],
             ((x.m[4] * p[0] + x.m[5] * p[1]) + x.m[6] * p[2]) + x.m[7],
             ((x.m[8] * p[0] + x.m[9] * p[1]) + x.m[10] * p[2]) + x.m[11]}};
}

struct TriSink {
    int max_tris; int32_t* ids; float* verts9; int n = 0;
    void put(int32_t id, const V3& a, const V3& b, const V3& c) {
        if (ids && verts9 && n < max_tris) {
            ids[n] = id;
            float* o = verts9 + 9 * (size_t)n;
            for (int k = 0; k < 3; k++) { o[k] = a[k]; o[3 + k] = b[k]; o[6 + k] = c[k]; }
        }
        n++;
    }
};

// main.go:237-277
void add_brush(const vrad_bsp_lumps& L, const vrad_dbrush& brush, const Xform& xf, TriSink& out) {
    if (!(brush.contents & VRAD_MASK_OPAQUE)) return;
    for (int i = 0; i < brush.numsides; i++) {
        const vrad_dbrushside& side = L.brushsides[brush.firstside + i];
        const int32_t tflags = side.texinfo >= 0 ? L.texinfo[side.texinfo].flags : 0;
        if ((tflags & VRAD_SURF_SKY) || side.dispinfo) continue;
        const vrad_dplane& pl = L.planes[side.planenum];
        Winding w = base_winding_for_plane(load3(pl.normal), pl.dist);
        bool alive = true;
        for (int j = 0; j < brush.numsides && alive; j++) {
            if (i == j) continue;
            const vrad_dbrushside& other = L.brushsides[brush.firstside + j];
            if (other.bevel) continue;
            const vrad_dplane& op = L.planes[other.planenum ^ 1];
            alive = chop_winding_in_place(w, load3(op.normal), op.dist, 0.0f);
        }
        if (!alive) continue;
        for (size_t j = 2; j < w.size(); j++)
            out.put(VRAD_TRACE_ID_OPAQUE, mul4x3(xf, w[0]), mul4x3(xf, w[j - 1]), mul4x3(xf, w[j]));
    }
}

inline int face_vertex(const vrad_bsp_lumps& L, const vrad_dface& f, int k) {    // lightmap.EdgeVertex, lightmap.go:267-282
    const int32_t se = L.surfedges[f.firstedge + k];
    return se < 0 ? L.edges[-(int64_t)se].v[1] : L.edges[se].v[0];
}

bool lumps_ok(const vrad_bsp_lumps* L, const char* who) {
    if (!L) { vrad::set_error("%s: bad arguments", who); return false; }
    return true;
}

int32_t vis_clusters(const vrad_bsp_lumps& L) {
    if (L.vis_len < 4 || !L.visdata) return 0;
    int32_t nc; std::memcpy(&nc, L.visdata, 4);
    return nc;
}

}  // namespace

extern "C" int vrad_bsp_raytrace_triangles(const vrad_bsp_lumps* Lp, int n_casters, const int32_t* caster_model, const float* caster_origin3,
                                           const float* caster_angles3, int max_tris, int32_t* ids, float* verts9, int* n_out) {
    if (!lumps_ok(Lp, "vrad_bsp_raytrace_triangles") || !n_out || n_casters < 0 || (n_casters && (!caster_model || !caster_origin3 || !caster_angles3))) {
        vrad::set_error("vrad_bsp_raytrace_triangles: bad arguments"); return VRAD_E_INVALID;
    }
    const vrad_bsp_lumps& L = *Lp;
    TriSink out{max_tris, ids, verts9};
    std::vector<int> list;
    std::vector<uint8_t> seen;
    // brush entities with vrad_brush_cast_shadows, in entity order (main.go:186-233)
    for (int c = 0; c < n_casters; c++) {
        const int m = caster_model[c];
        if (m <= 0 || m >= L.n_models) continue;                            // brushmodelForEntity returns nil (main.go:214-225)
        list.clear(); seen.assign((size_t)L.n_brushes, 0);
        brushes_under(L, L.models[m].headnode, list, seen);
        const Xform xf = xform_from_origin_angles(caster_origin3 + 3 * c, caster_angles3 + 3 * c);
        for (int b : list) add_brush(L, L.brushes[b], xf, out);
    }
    // the world: model 0's brushes, then its sky faces (main.go:279-340)
    if (L.n_models > 0) {
        Xform ident; ident.identity = true;
        list.clear(); seen.assign((size_t)L.n_brushes, 0);
        brushes_under(L, L.models[0].headnode, list, seen);
        for (int b : list) add_brush(L, L.brushes[b], ident, out);
        const vrad_dmodel& world = L.models[0];
        for (int i = 0; i < world.numfaces; i++) {
            const vrad_dface& f = L.faces[world.firstface + i];
            if (!(L.texinfo[f.texinfo].flags & VRAD_SURF_SKY)) continue;
            if (f.numedges > kMaxPointsOnWinding) { vrad::set_error("***** ERROR! MAX_POINTS_ON_WINDING reached! (face %d has %d edges)", world.firstface + i, f.numedges); return VRAD_E_INVALID; }
            V3 pts[kMaxPointsOnWinding];
            for (int j = 0; j < f.numedges; j++) pts[j] = load3(L.vertexes3 + 3 * (size_t)face_vertex(L, f, j));
            for (int j = 2; j < f.numedges; j++) out.put(VRAD_TRACE_ID_SKY, pts[0], pts[j - 1], pts[j]);
        }
    }
    *n_out = out.n;
    if (ids && verts9 && out.n > max_tris) { vrad::set_error("vrad_bsp_raytrace_triangles: %d triangles, room for %d", out.n, max_tris); return VRAD_E_NOMEM; }
    return VRAD_OK;
}

extern "C" int vrad_env_add_bsp(vrad_env* e, const vrad_bsp_lumps* L, int n_casters, const int32_t* caster_model, const float* caster_origin3,
                                const float* caster_angles3, int* n_added) {
    if (!e) { vrad::set_error("vrad_env_add_bsp: bad arguments"); return VRAD_E_INVALID; }
    int n = 0;
    int rc = vrad_bsp_raytrace_triangles(L, n_casters, caster_model, caster_origin3, caster_angles3, 0, nullptr, nullptr, &n);
    if (rc) return rc;
    std::vector<int32_t> ids((size_t)n);
    std::vector<float> verts(9 * (size_t)n);
    std::vector<uint8_t> flags((size_t)n, 0);
    if ((rc = vrad_bsp_raytrace_triangles(L, n_casters, caster_model, caster_origin3, caster_angles3, n, ids.data(), verts.data(), &n))) return rc;
    if (n && (rc = vrad_env_add_triangles(e, n, ids.data(), verts.data(), flags.data()))) return rc;
    if (n_added) *n_added = n;
    return VRAD_OK;
}

extern "C" int vrad_bsp_face_patches(const vrad_bsp_lumps* Lp, const float* model_origins3, float max_chop, int max_faces, int max_points,
                                     int* n_faces_out, int* n_points_out, vrad_face_patch* faces, float* points3, int32_t* face_number,
                                     float* reflectivity3, float* base_area, uint8_t* needs_bump, float* scale2) {
    if (!lumps_ok(Lp, "vrad_bsp_face_patches") || !n_faces_out || !n_points_out) { vrad::set_error("vrad_bsp_face_patches: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    int nf = 0, np = 0;
    const bool store = faces && points3;
    for (int m = 0; m < L.n_models; m++) {
        const vrad_dmodel& mod = L.models[m];
        const V3 origin = model_origins3 ? load3(model_origins3 + 3 * m) : V3{{0, 0, 0}};
        for (int j = 0; j < mod.numfaces; j++) {
            const int fn = mod.firstface + j;
            const vrad_dface& f = L.faces[fn];
            if (f.dispinfo != -1) continue;                                 // build.go:52: displacement faces make no face patch
            if (f.numedges > kMaxPointsOnWinding) { vrad::set_error("face %d has %d edges (MAX_POINTS_ON_WINDING = %d)", fn, f.numedges, kMaxPointsOnWinding); return VRAD_E_INVALID; }
            // world.WindingFromFace (face.go:92-114) + RemoveColinearPoints (point.go:12-46)
            V3 w[kMaxPointsOnWinding], kept[kMaxPointsOnWinding];
            const int n = f.numedges;
            for (int i = 0; i < n; i++) w[i] = add(load3(L.vertexes3 + 3 * (size_t)face_vertex(L, f, i)), origin);
            int nk = 0;
            for (int i = 0; i < n; i++) {
                const int nx = (i + 1) % n, pv = (i + n - 1) % n;
                const V3 v1 = normalize(sub(w[nx], w[i])), v2 = normalize(sub(w[i], w[pv]));
                if (dot(v1, v2) < 0.999f) kept[nk++] = w[i];
            }
            const vrad_texinfo& tx = L.texinfo[f.texinfo];
            const vrad_dtexdata& td = L.texdata[tx.texdata];
            if (store && nf < max_faces && np + nk <= max_points) {
                vrad_face_patch& o = faces[nf];
                std::memset(&o, 0, sizeof o);
                o.first_point = np; o.n_points = nk;
                for (int i = 0; i < nk; i++) for (int k = 0; k < 3; k++) points3[3 * (size_t)(np + i) + k] = kept[i][k];
                // patch.Plane (face.go:120-144): the face's plane (dface.planenum already names the facing plane), moved by the model origin
                const vrad_dplane& pl = L.planes[f.planenum];
                V3 nrm = load3(pl.normal);
                float dist = pl.dist;
                if (origin[0] != 0.0f || origin[1] != 0.0f || origin[2] != 0.0f) dist += dot(origin, nrm);
                for (int k = 0; k < 3; k++) o.normal[k] = nrm[k];
                o.plane_dist = dist;
                // Patch.Scale / chopScale (face.go:83-109): lengths of the texture / lightmap vectors
                float sc[2], chop_scale[2];
                for (int i = 0; i < 2; i++) {
                    float s = 0.0f, c = 0.0f;
                    for (int k = 0; k < 3; k++) { s += tx.texture_vecs[i][k] * tx.texture_vecs[i][k]; c += tx.lightmap_vecs[i][k] * tx.lightmap_vecs[i][k]; }
                    sc[i] = (float)std::sqrt((double)s); chop_scale[i] = (float)std::sqrt((double)c);
                }
                o.lux_scale = (chop_scale[0] + chop_scale[1]) / 2;
                o.chop = max_chop;
                o.sky = (tx.flags & VRAD_SURF_SKY) ? 1 : 0;
                o.no_subdivide = ((tx.flags & VRAD_SURF_NOCHOP) || ((tx.flags & VRAD_SURF_NOLIGHT) && !(tx.flags & VRAD_SURF_LIGHT))) ? 1 : 0;
                o.has_base_light = 0;                                       // texlights (lights.rad) stay with the Go driver
                if (face_number) face_number[nf] = fn;
                if (reflectivity3) for (int k = 0; k < 3; k++) { const float r = td.reflectivity[k] * 1.0f; reflectivity3[3 * (size_t)nf + k] = r > 0.99f ? 0.99f : r; }
                if (base_area) base_area[nf] = (float)(int32_t)((uint32_t)td.height * (uint32_t)td.width);       // Go's int32 product wraps (face.go:219)
                if (needs_bump) needs_bump[nf] = (tx.flags & VRAD_SURF_BUMPLIGHT) ? 1 : 0;
                if (scale2) { scale2[2 * (size_t)nf] = sc[0]; scale2[2 * (size_t)nf + 1] = sc[1]; }
            }
            nf++; np += nk;
        }
    }
    *n_faces_out = nf; *n_points_out = np;
    if (store && (nf > max_faces || np > max_points)) { vrad::set_error("vrad_bsp_face_patches: %d faces / %d points, room for %d / %d", nf, np, max_faces, max_points); return VRAD_E_NOMEM; }
    return VRAD_OK;
}

extern "C" int vrad_bsp_rescale_lightmap_vecs(int n_texinfo, vrad_texinfo* texinfo, float luxel_density) {
    if (n_texinfo < 0 || (n_texinfo && !texinfo)) { vrad::set_error("vrad_bsp_rescale_lightmap_vecs: bad arguments"); return VRAD_E_INVALID; }
    if (!(luxel_density < 1.0f)) return VRAD_OK;                            // start.go:22
    for (int i = 0; i < n_texinfo; i++)
        for (int j = 0; j < 2; j++) {
            V3 tmp = load3(texinfo[i].lightmap_vecs[j]);
            const float l = len(tmp);
            if (l == 0.0f) continue;
            float s = l;
            tmp = scale(tmp, 1.0f / l);                                     // scale = VectorNormalize(tmp)
            if (std::fabs(s) > luxel_density) {
                s = s < 0 ? -luxel_density : luxel_density;
                tmp = scale(tmp, s);
                for (int k = 0; k < 3; k++) texinfo[i].lightmap_vecs[j][k] = tmp[k];
            }
        }
    return VRAD_OK;
}

extern "C" int vrad_bsp_face_extents(const vrad_bsp_lumps* Lp, int32_t* mins2, int32_t* size2, int* n_oversize_out) {
    if (!lumps_ok(Lp, "vrad_bsp_face_extents") || !mins2 || !size2) { vrad::set_error("vrad_bsp_face_extents: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    int oversize = 0;
    for (int fn = 0; fn < L.n_faces; fn++) {
        const vrad_dface& f = L.faces[fn];
        const vrad_texinfo& tx = L.texinfo[f.texinfo];
        if (tx.flags & (VRAD_SURF_SKY | VRAD_SURF_NOLIGHT)) {               // start.go:104-106: non-lit texture keeps what the file says
            for (int i = 0; i < 2; i++) { mins2[2 * (size_t)fn + i] = f.lm_mins[i]; size2[2 * (size_t)fn + i] = f.lm_size[i]; }
            continue;
        }
        float mn[2] = {1e24f, 1e24f}, mx[2] = {-1e24f, -1e24f};
        for (int i = 0; i < f.numedges; i++) {
            const float* v = L.vertexes3 + 3 * (size_t)face_vertex(L, f, i);
            for (int j = 0; j < 2; j++) {
                const float val = ((v[0] * tx.lightmap_vecs[j][0] + v[1] * tx.lightmap_vecs[j][1]) + v[2] * tx.lightmap_vecs[j][2]) + tx.lightmap_vecs[j][3];
                if (val < mn[j]) mn[j] = val;
                if (val > mx[j]) mx[j] = val;
            }
        }
        const int max_dim = kMaxLightmapDim;                                // same constant for displacement and plain faces (constants.go:22,27)
        bool big = false;
        for (int i = 0; i < 2; i++) {
            mn[i] = (float)std::floor((double)mn[i]);
            mx[i] = (float)std::ceil((double)mx[i]);
            mins2[2 * (size_t)fn + i] = (int32_t)mn[i];
            size2[2 * (size_t)fn + i] = (int32_t)(mx[i] - mn[i]);
            if (size2[2 * (size_t)fn + i] > max_dim + 1) big = true;
        }
        if (big) oversize++;
    }
    if (n_oversize_out) *n_oversize_out = oversize;
    return VRAD_OK;
}

extern "C" int vrad_bsp_make_parents(const vrad_bsp_lumps* Lp, int32_t* node_parents, int32_t* leaf_parents) {
    if (!lumps_ok(Lp, "vrad_bsp_make_parents") || !node_parents || !leaf_parents) { vrad::set_error("vrad_bsp_make_parents: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    for (int i = 0; i < L.n_nodes; i++) node_parents[i] = -1;
    for (int i = 0; i < L.n_leafs; i++) leaf_parents[i] = -1;
    if (L.n_nodes == 0) return VRAD_OK;
    std::vector<int> stack{0};
    std::vector<uint8_t> seen((size_t)L.n_nodes, 0);
    seen[0] = 1;
    while (!stack.empty()) {                                               // nodes.go:21-36 from (0, -1)
        const int n = stack.back(); stack.pop_back();
        for (int k = 0; k < 2; k++) {
            const int c = L.nodes[n].children[k];
            if (c < 0) leaf_parents[-1 - c] = n;
            else {
                if (seen[c]) { vrad::set_error("vrad_bsp_make_parents: node %d is reached twice (not a tree)", c); return VRAD_E_INVALID; }
                seen[c] = 1; node_parents[c] = n; stack.push_back(c);
            }
        }
    }
    return VRAD_OK;
}

extern "C" int vrad_bsp_cluster_table(const vrad_bsp_lumps* Lp, int n_clusters, int32_t* first, int32_t* leafs) {
    if (!lumps_ok(Lp, "vrad_bsp_cluster_table") || n_clusters < 0 || !first || !leafs) { vrad::set_error("vrad_bsp_cluster_table: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    std::vector<int32_t> count((size_t)n_clusters + 1, 0);
    for (int j = 0; j < L.n_leafs; j++) { const int c = L.leafs[j].cluster; if (c >= 0 && c < n_clusters) count[c + 1]++; }
    first[0] = 0;
    for (int c = 0; c < n_clusters; c++) first[c + 1] = first[c] + count[c + 1];
    std::vector<int32_t> fill(first, first + n_clusters);
    for (int j = 0; j < L.n_leafs; j++) { const int c = L.leafs[j].cluster; if (c >= 0 && c < n_clusters) leafs[fill[c]++] = j; }
    return VRAD_OK;
}

extern "C" int vrad_bsp_vis_for_light_environment(vrad_env* env, const vrad_bsp_lumps* Lp, uint8_t* leaf_flags_out, uint8_t* sky_pvs_out, int* has_pvs) {
    if (!lumps_ok(Lp, "vrad_bsp_vis_for_light_environment") || !leaf_flags_out) { vrad::set_error("vrad_bsp_vis_for_light_environment: bad arguments"); return VRAD_E_INVALID; }
    const vrad_bsp_lumps& L = *Lp;
    const int32_t nc = vis_clusters(L);
    const size_t row = ((size_t)nc + 7) >> 3;
    const int32_t* byteofs = nc ? reinterpret_cast<const int32_t*>(L.visdata + 4) : nullptr;
    std::vector<uint8_t> pvs(row ? row : 1);
    // lightmap.GetVisCache(-1, cluster, &pvs) (vis.go:9-47): everything visible without vis data or for cluster < 0
    auto cluster_pvs = [&](int cluster, std::vector<uint8_t>& out) -> int {
        if (nc == 0 || cluster < 0) { std::fill(out.begin(), out.end(), (uint8_t)255); return VRAD_OK; }
        if (cluster >= nc) { vrad::set_error("leaf cluster %d outside the %d clusters of the visibility lump", cluster, nc); return VRAD_E_INVALID; }
        const int32_t ofs = byteofs[2 * cluster];
        if (ofs < 0 || ofs >= L.vis_len) { vrad::set_error("visofs == -1 (cluster %d)", cluster); return VRAD_E_INVALID; }
        const int64_t used = vrad_decompress_vis(L.visdata + ofs, L.vis_len - ofs, nc, out.data());
        return used < 0 ? (int)used : VRAD_OK;
    };
    auto pvs_check = [nc](const std::vector<uint8_t>& p, int cluster) -> bool {    // lightmap.go:413-422
        if (cluster < 0 || nc == 0) return true;
        return cluster < nc && (p[(size_t)cluster >> 3] & (1u << (cluster & 7))) != 0;
    };
    std::vector<uint8_t> flags((size_t)L.n_leafs);
    std::vector<uint8_t> merged(row, 0);
    bool any = false;
    int rc;
    // first pass (:286-310): leafs holding sky faces; their PVS rows merged into the sky lights' PVS
    for (int i = 0; i < L.n_leafs; i++) {
        const vrad_dleaf& lf = L.leafs[i];
        uint8_t fl = (uint8_t)(((uint16_t)lf.area_flags >> 9) & 0x7f);
        fl &= (uint8_t)~(VRAD_LEAF_FLAGS_SKY | VRAD_LEAF_FLAGS_SKY2D);
        for (int k = 0; k < lf.numleaffaces; k++) {
            const vrad_dface& f = L.faces[L.leaffaces[lf.firstleafface + k]];
            const int32_t tf = L.texinfo[f.texinfo].flags;
            if (tf & VRAD_SURF_SKY) {
                fl |= (tf & VRAD_SURF_SKY2D) ? VRAD_LEAF_FLAGS_SKY2D : VRAD_LEAF_FLAGS_SKY;
                if (row) {                                                  // MergeDLightVis / SetDLightVis (:399-411)
                    if ((rc = cluster_pvs(lf.cluster, pvs))) return rc;
                    for (size_t b = 0; b < row; b++) merged[b] |= pvs[b];
                }
                any = true;
                break;
            }
        }
        flags[i] = fl;
    }
    // second pass (:312-352): leafs that see a sky leaf
    std::vector<uint8_t> sees3d((size_t)L.n_leafs, 0), sees2d((size_t)L.n_leafs, 0);
    std::vector<int> sky_leafs;
    for (int i = 0; i < L.n_leafs; i++) if (flags[i] & (VRAD_LEAF_FLAGS_SKY | VRAD_LEAF_FLAGS_SKY2D)) sky_leafs.push_back(i);
    for (int i = 0; i < L.n_leafs; i++) {
        if (flags[i] & VRAD_LEAF_FLAGS_SKY) continue;
        if (L.leafs[i].contents & VRAD_CONTENTS_SOLID) continue;
        if ((rc = cluster_pvs(L.leafs[i].cluster, pvs))) return rc;
        for (int j : sky_leafs) {
            if (j == i) continue;
            if (!pvs_check(pvs, L.leafs[j].cluster)) continue;
            if (flags[j] & VRAD_LEAF_FLAGS_SKY2D) sees2d[i] = 1;
            if (flags[j] & VRAD_LEAF_FLAGS_SKY) { sees3d[i] = 1; break; }
        }
    }
    // third pass (:354-384); radial-vis leafs that saw nothing are traced
    std::vector<int> to_trace;
    for (int i = 0; i < L.n_leafs; i++) {
        if (flags[i] & VRAD_LEAF_FLAGS_SKY) continue;
        if (L.leafs[i].contents & VRAD_CONTENTS_SOLID) continue;
        if (sees2d[i]) flags[i] |= VRAD_LEAF_FLAGS_SKY2D;
        if (sees3d[i]) { flags[i] |= VRAD_LEAF_FLAGS_SKY; flags[i] &= (uint8_t)~VRAD_LEAF_FLAGS_SKY2D; }
        else if (flags[i] & VRAD_LEAF_FLAGS_RADIAL) to_trace.push_back(i);
    }
    if (!to_trace.empty()) {
        if (!env) { vrad::set_error("vrad_bsp_vis_for_light_environment: %zu LEAF_FLAGS_RADIAL leafs need CanLeafTraceToSky but no environment was given", to_trace.size()); return VRAD_E_STATE; }
        std::vector<int16_t> mn(3 * to_trace.size()), mx(3 * to_trace.size());
        std::vector<uint8_t> can(to_trace.size());
        for (size_t t = 0; t < to_trace.size(); t++)
            for (int k = 0; k < 3; k++) { mn[3 * t + k] = L.leafs[to_trace[t]].mins[k]; mx[3 * t + k] = L.leafs[to_trace[t]].maxs[k]; }
        if ((rc = vrad_leafs_trace_to_sky(env, (int)to_trace.size(), mn.data(), mx.data(), can.data()))) return rc;
        for (size_t t = 0; t < to_trace.size(); t++) if (can[t]) flags[to_trace[t]] |= VRAD_LEAF_FLAGS_SKY;
    }
    std::memcpy(leaf_flags_out, flags.data(), flags.size());
    if (sky_pvs_out && row) std::memcpy(sky_pvs_out, merged.data(), row);
    if (has_pvs) *has_pvs = (any && row) ? 1 : 0;
    return VRAD_OK;
}
