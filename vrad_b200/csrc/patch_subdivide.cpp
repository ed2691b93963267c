// patch_subdivide.cpp -- host side of the patch hierarchy: face patches and their
// recursive subdivision, the step that produces the Parent/Child1/Child2 trees the hierarchical transfer
// build (K2) and CollectLight (K4) work on.  Pure host code, no device needed.
//
// Reference map
//   rad/patches/face.go:29-197        MakePatchForFace: area (WindingArea), origin (WindingCenter), bounds, chop
//   rad/patches/subdivide.go:25-66    SubdividePatches: every face patch, in order
//   rad/patches/subdivide.go:167-248  SubdividePatch: widest axis in luxels, chop / minChop rule, "make more
//                                     square" rule, split at the bounds midpoint, depth-first recursion
//   rad/patches/subdivide.go:250-346  ClipWindingEpsilon (ON_EPSILON = 0.1, vmath/constants.go:8)
//   rad/patches/subdivide.go:352-406  CreateChildPatch: child copies the parent, new winding/area/origin/bounds,
//                                     edge-of-face chop rule
//   rad/patches/subdivide.go:409-437  WindingAreaAndBalancePoint
//   vmath/polygon/winding.go:217-260  WindingArea, WindingCenter, WindingBounds
// App. A intents applied: #16 (WindingCenter divides by a float point count and writes its result),
// #22 (WindingAreaAndBalancePoint writes the caller's centre).  GetPhongNormal (subdivide.go:385) is the
// plane normal here: the caller overwrites the children's normals with vrad_bsp_phong_normals (include/vrad_bsp.h) once PairEdges has run, as vrad_b200/bake.py does.
//
// Formulation: flat point/patch arrays and an explicit work stack instead of the reference's recursion over
// heap windings; child indices come out in the same order (child1's whole subtree before child2's).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/vrad_cuda.h"

namespace vrad { void set_error(const char* fmt, ...); }

namespace {

constexpr float kOnEpsilon = 0.1f;            // vmath/constants.go:8
constexpr int   kMaxPointsOnWinding = 64;     // common/constants/constants.go:32

struct V3 { float x, y, z; };
inline float comp(const V3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
inline void set_comp(V3& v, int k, float f) { (k == 0 ? v.x : (k == 1 ? v.y : v.z)) = f; }

struct Builder {
    std::vector<V3> pts;                                   // all windings, concatenated
    std::vector<int32_t> w_first, w_count;
    std::vector<V3> origin, normal, mins, maxs, face_mins, face_maxs;
    std::vector<float> plane_dist, area, chop, lux;
    std::vector<int32_t> parent, child1, child2, face;
    std::vector<uint8_t> sky, base_light;

    int add_patch() {
        const int i = (int)origin.size();
        origin.push_back({0, 0, 0}); normal.push_back({0, 0, 0}); mins.push_back({0, 0, 0}); maxs.push_back({0, 0, 0});
        face_mins.push_back({0, 0, 0}); face_maxs.push_back({0, 0, 0});
        plane_dist.push_back(0); area.push_back(0); chop.push_back(0); lux.push_back(0);
        parent.push_back(-1); child1.push_back(-1); child2.push_back(-1); face.push_back(-1);
        sky.push_back(0); base_light.push_back(0); w_first.push_back(0); w_count.push_back(0);
        return i;
    }
};

// mgl32.Vec3.Len: fp32 products and sums, the square root taken in double and rounded back
inline float vec_len(float x, float y, float z) { return (float)std::sqrt((double)(((x * x) + (y * y)) + (z * z))); }

// winding.go:241-260
void winding_bounds(const V3* p, int n, V3& mn, V3& mx) {
    mn = {99999.0f, 99999.0f, 99999.0f}; mx = {-99999.0f, -99999.0f, -99999.0f};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) {
            const float v = comp(p[i], j);
            if (v < comp(mn, j)) set_comp(mn, j, v);
            if (v > comp(mx, j)) set_comp(mx, j, v);
        }
}

// subdivide.go:409-437 (triangle fan from point 0; centre = area-weighted mean of the fan's centroids)
float area_and_balance_point(const V3* p, int n, V3& center) {
    center = {0, 0, 0};
    float total = 0.0f;
    for (int i = 2; i < n; i++) {
        const V3 d1 = {p[i - 1].x - p[0].x, p[i - 1].y - p[0].y, p[i - 1].z - p[0].z};
        const V3 d2 = {p[i].x - p[0].x, p[i].y - p[0].y, p[i].z - p[0].z};
        const float cx = (d1.y * d2.z) - (d1.z * d2.y), cy = (d1.z * d2.x) - (d1.x * d2.z), cz = (d1.x * d2.y) - (d1.y * d2.x);
        const float a = vec_len(cx, cy, cz);
        total += a;
        const float s = a / 3.0f;
        const V3* q[3] = {&p[i - 1], &p[i], &p[0]};
        for (int k = 0; k < 3; k++) {                       // vector.MA(center, area/3, point, center)
            center.x = center.x + (s * q[k]->x); center.y = center.y + (s * q[k]->y); center.z = center.z + (s * q[k]->z);
        }
    }
    if (total != 0.0f) { const float r = 1.0f / total; center.x = center.x * r; center.y = center.y * r; center.z = center.z * r; }
    return total * 0.5f;
}

// subdivide.go:250-346 for an axis-aligned split plane (normal = +axis): front = the side with larger coordinates
void clip_winding_axis(const V3* in, int n, int axis, float dist, std::vector<V3>& front, std::vector<V3>& back) {
    float dists[kMaxPointsOnWinding + 4];
    int sides[kMaxPointsOnWinding + 4];
    int counts[3] = {0, 0, 0};
    front.clear(); back.clear();
    if (n <= 0) return;                                    // an empty winding has no sides
    for (int i = 0; i < n; i++) {
        float dot = comp(in[i], axis);                     // Points[i].Dot(normal) with a unit axis normal
        // mgl32 Dot = p0*n0 + p1*n1 + p2*n2: the two zero products add +0 and leave the value unchanged
        dot -= dist;
        dists[i] = dot;
        sides[i] = dot > kOnEpsilon ? 0 : (dot < -kOnEpsilon ? 1 : 2);
        counts[sides[i]]++;
    }
    sides[n] = sides[0]; dists[n] = dists[0];
    if (!counts[0]) { back.assign(in, in + n); return; }
    if (!counts[1]) { front.assign(in, in + n); return; }
    for (int i = 0; i < n; i++) {
        const V3& p1 = in[i];
        if (sides[i] == 2) { front.push_back(p1); back.push_back(p1); continue; }
        if (sides[i] == 0) front.push_back(p1);
        if (sides[i] == 1) back.push_back(p1);
        if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
        const V3& p2 = in[(i + 1) % n];
        const float dot = dists[i] / (dists[i] - dists[i + 1]);
        V3 mid;
        for (int j = 0; j < 3; j++)                        // "avoid round off error when possible" (:317-325)
            set_comp(mid, j, j == axis ? dist : comp(p1, j) + (dot * (comp(p2, j) - comp(p1, j))));
        front.push_back(mid); back.push_back(mid);
    }
}

} // namespace

extern "C" int vrad_patches_subdivide(int n_faces, const vrad_face_patch* faces, const float* points3, float min_chop,
                                      int max_patches, int max_points, int* n_patches_out, int* n_points_out,
                                      float* origin3, float* normal3, float* plane_dist, float* area, float* mins3, float* maxs3,
                                      float* chop, int32_t* parent, int32_t* child1, int32_t* child2, int32_t* face,
                                      int32_t* wind_first, int32_t* wind_count, float* wind_points3) {
    if (n_faces < 0 || (n_faces > 0 && (!faces || !points3)) || !n_patches_out || !n_points_out) {
        vrad::set_error("vrad_patches_subdivide: bad arguments"); return VRAD_E_INVALID;
    }
    Builder B;
    // MakePatchForFace for every face, in order (face.go:29-197)
    for (int f = 0; f < n_faces; f++) {
        const vrad_face_patch& F = faces[f];
        if (F.n_points < 3 || F.n_points > kMaxPointsOnWinding || F.first_point < 0) {
            vrad::set_error("vrad_patches_subdivide: face %d has a winding of %d points", f, F.n_points); return VRAD_E_INVALID;
        }
        const V3* w = reinterpret_cast<const V3*>(points3) + F.first_point;
        float total = 0.0f;                                 // WindingArea, winding.go:217-229
        for (int i = 2; i < F.n_points; i++) {
            const V3 d1 = {w[i - 1].x - w[0].x, w[i - 1].y - w[0].y, w[i - 1].z - w[0].z};
            const V3 d2 = {w[i].x - w[0].x, w[i].y - w[0].y, w[i].z - w[0].z};
            total += vec_len((d1.y * d2.z) - (d1.z * d2.y), (d1.z * d2.x) - (d1.x * d2.z), (d1.x * d2.y) - (d1.y * d2.x));
        }
        const float a = total * 0.5f;
        if (a <= 0.0f) continue;                            // degenerate face: no patch (face.go:47-51)
        const int p = B.add_patch();
        B.w_first[p] = (int32_t)B.pts.size(); B.w_count[p] = F.n_points;
        B.pts.insert(B.pts.end(), w, w + F.n_points);
        V3 c = {0, 0, 0};                                   // WindingCenter, winding.go:231-239
        for (int i = 0; i < F.n_points; i++) { c.x = w[i].x + c.x; c.y = w[i].y + c.y; c.z = w[i].z + c.z; }
        const float sc = 1.0f / (float)F.n_points;
        B.origin[p] = {c.x * sc, c.y * sc, c.z * sc};
        B.normal[p] = {F.normal[0], F.normal[1], F.normal[2]};
        B.plane_dist[p] = F.plane_dist;
        B.area[p] = a; B.chop[p] = F.chop; B.lux[p] = F.lux_scale; B.face[p] = f;
        B.sky[p] = F.sky ? 1 : 0; B.base_light[p] = F.has_base_light ? 1 : 0;
        winding_bounds(w, F.n_points, B.face_mins[p], B.face_maxs[p]);
        B.mins[p] = B.face_mins[p]; B.maxs[p] = B.face_maxs[p];
    }
    // SubdividePatches (subdivide.go:50-66): each face patch in order; SubdividePatch depth first
    const int n_roots = (int)B.origin.size();
    std::vector<int> work;
    std::vector<V3> front, back, in;
    for (int r = 0; r < n_roots; r++) {
        if (faces[B.face[r]].no_subdivide) continue;        // PreventSubdivision / displacement face
        work.assign(1, r);
        while (!work.empty()) {
            const int p = work.back(); work.pop_back();
            if (B.sky[p]) continue;                         // never subdivide sky patches (:183-186)
            float total[3];
            for (int k = 0; k < 3; k++) total[k] = (comp(B.maxs[p], k) - comp(B.mins[p], k)) * B.lux[p];   // :191-192
            float widest = -1.0f; int axis = -1; bool split = false;
            for (int k = 0; k < 3; k++) {
                if (total[k] > widest) { axis = k; widest = total[k]; }
                if (total[k] >= B.chop[p] && total[k] >= min_chop) split = true;
            }
            if (!split && axis != -1) {                     // make more square (:204-212)
                if (total[axis] > total[(axis + 1) % 3] * 2.0f && total[axis] > total[(axis + 2) % 3] * 2.0f) {
                    if (B.chop[p] > min_chop) { split = true; B.chop[p] = std::fmax(min_chop, B.chop[p] / 2.0f); }
                }
            }
            if (!split) continue;
            const float dist = (comp(B.mins[p], axis) + comp(B.maxs[p], axis)) * 0.5f;                     // :221
            in.assign(B.pts.begin() + B.w_first[p], B.pts.begin() + B.w_first[p] + B.w_count[p]);
            clip_winding_axis(in.data(), (int)in.size(), axis, dist, front, back);
            if (front.size() > (size_t)kMaxPointsOnWinding || back.size() > (size_t)kMaxPointsOnWinding) {
                vrad::set_error("ClipWinding: MAX_POINTS_ON_WINDING"); return VRAD_E_INVALID;              // log.Fatal (:343-345)
            }
            V3 c1, c2;
            const float a1 = front.empty() ? 0.0f : area_and_balance_point(front.data(), (int)front.size(), c1);
            const float a2 = back.empty() ? 0.0f : area_and_balance_point(back.data(), (int)back.size(), c2);
            if (a1 == 0.0f || a2 == 0.0f) continue;         // "zero area child patch" (:229-232)
            int kids[2];
            for (int s = 0; s < 2; s++) {                   // CreateChildPatch (:352-406)
                const std::vector<V3>& w = s == 0 ? front : back;
                const int c = B.add_patch();
                kids[s] = c;
                B.normal[c] = B.normal[p]; B.plane_dist[c] = B.plane_dist[p]; B.chop[c] = B.chop[p]; B.lux[c] = B.lux[p];
                B.face[c] = B.face[p]; B.sky[c] = B.sky[p]; B.base_light[c] = B.base_light[p];
                B.face_mins[c] = B.face_mins[p]; B.face_maxs[c] = B.face_maxs[p];
                B.parent[c] = p;
                B.w_first[c] = (int32_t)B.pts.size(); B.w_count[c] = (int32_t)w.size();
                B.pts.insert(B.pts.end(), w.begin(), w.end());
                B.area[c] = s == 0 ? a1 : a2;
                B.origin[c] = s == 0 ? c1 : c2;
                winding_bounds(w.data(), (int)w.size(), B.mins[c], B.maxs[c]);
                if (B.base_light[c]) continue;              // don't check edges on surf lights (:389-392)
                float t[3];
                for (int k = 0; k < 3; k++) t[k] = (comp(B.maxs[c], k) - comp(B.mins[c], k)) * B.lux[c];
                if (B.chop[c] > min_chop && t[0] < B.chop[c] && t[1] < B.chop[c] && t[2] < B.chop[c]) {   // :395-403
                    for (int k = 0; k < 3; k++) {
                        if ((comp(B.face_maxs[c], k) == comp(B.maxs[c], k) || comp(B.face_mins[c], k) == comp(B.mins[c], k)) && t[k] > min_chop) {
                            B.chop[c] = std::fmax(min_chop, B.chop[c] / 2.0f);
                            break;
                        }
                    }
                }
            }
            B.child1[p] = kids[0]; B.child2[p] = kids[1];
            work.push_back(kids[1]); work.push_back(kids[0]);                                              // child1's subtree first (:246-247)
        }
    }
    const int np = (int)B.origin.size(), npt = (int)B.pts.size();
    *n_patches_out = np; *n_points_out = npt;
    if (np > max_patches || npt > max_points) {
        vrad::set_error("vrad_patches_subdivide: %d patches / %d winding points needed, capacity %d / %d", np, npt, max_patches, max_points);
        return VRAD_E_NOMEM;
    }
    auto put3 = [](float* dst, const std::vector<V3>& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(V3)); };
    auto put = [](auto* dst, const auto& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    put3(origin3, B.origin); put3(normal3, B.normal); put3(mins3, B.mins); put3(maxs3, B.maxs); put3(wind_points3, B.pts);
    put(plane_dist, B.plane_dist); put(area, B.area); put(chop, B.chop);
    put(parent, B.parent); put(child1, B.child1); put(child2, B.child2); put(face, B.face);
    put(wind_first, B.w_first); put(wind_count, B.w_count);
    return VRAD_OK;
}
