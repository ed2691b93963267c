// patches.cpp -- ORACLE (test infrastructure): face patches and their subdivision, restated literally
// (recursive, one heap winding per patch) from
//   rad/patches/face.go:29-197 (MakePatchForFace), rad/patches/subdivide.go:25-66 (SubdividePatches),
//   :167-248 (SubdividePatch), :250-346 (ClipWindingEpsilon), :352-406 (CreateChildPatch),
//   :409-437 (WindingAreaAndBalancePoint), vmath/polygon/winding.go:217-260 (WindingArea/Center/Bounds),
//   vmath/vector/vec3.go:7-17 (MA, Scale).
// App. A intents: #16 (WindingCenter: float division, result written), #22 (balance point written through).
// GetPhongNormal (subdivide.go:385) is taken as the plane normal (no face-neighbour smoothing here).
// mgl32: Sub/Add/Dot/Cross componentwise fp32; Len = float32(sqrt(float64(x*x+y*y+z*z))).
#include "oracle_impl.hpp"
#include <memory>

namespace orc {

struct Vec3 { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
static inline Vec3 sub(const Vec3& a, const Vec3& b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
static inline Vec3 cross(const Vec3& a, const Vec3& b) { return {{(a[1] * b[2]) - (a[2] * b[1]), (a[2] * b[0]) - (a[0] * b[2]), (a[0] * b[1]) - (a[1] * b[0])}}; }
static inline float dot(const Vec3& a, const Vec3& b) { return ((a[0] * b[0]) + (a[1] * b[1])) + (a[2] * b[2]); }
static inline float len(const Vec3& a) { return (float)sqrt((double)(((a[0] * a[0]) + (a[1] * a[1])) + (a[2] * a[2]))); }
static inline void MA(const Vec3& start, float scale, const Vec3& dir, Vec3& dest) {      // vec3.go:7-11
    dest[0] = start[0] + (scale * dir[0]); dest[1] = start[1] + (scale * dir[1]); dest[2] = start[2] + (scale * dir[2]);
}

typedef std::vector<Vec3> Winding;
typedef std::shared_ptr<Winding> WindingP;

struct HPatch {                       // the common/types/patch.go:9-64 fields this stage touches
    WindingP winding;
    Vec3 mins, maxs, face_mins, face_maxs, origin, normal;
    float plane_dist, area, chop, lux_scale;
    int parent, child1, child2, face;
    bool sky, base_light;
};

static const float ON_EPSILON = 0.1f;            // vmath/constants.go:8
static const int MAX_POINTS_ON_WINDING = 64;

static void WindingBounds(const Winding& w, Vec3& mins, Vec3& maxs) {                     // winding.go:241-260
    for (int j = 0; j < 3; j++) { mins[j] = 99999; maxs[j] = -99999; }
    for (size_t i = 0; i < w.size(); i++)
        for (int j = 0; j < 3; j++) {
            float v = w[i][j];
            if (v < mins[j]) mins[j] = v;
            if (v > maxs[j]) maxs[j] = v;
        }
}

static float WindingArea(const Winding& w) {                                               // winding.go:217-229
    float total = 0.0f;
    for (size_t i = 2; i < w.size(); i++) total += len(cross(sub(w[i - 1], w[0]), sub(w[i], w[0])));
    return total * 0.5f;
}

static float WindingAreaAndBalancePoint(const WindingP& w, Vec3& center) {                 // subdivide.go:409-437
    center = {{0, 0, 0}};
    if (!w) return 0.0f;
    float total = 0;
    for (size_t i = 2; i < w->size(); i++) {
        Vec3 d1 = sub((*w)[i - 1], (*w)[0]), d2 = sub((*w)[i], (*w)[0]);
        float area = len(cross(d1, d2));
        total += area;
        MA(center, area / 3.0f, (*w)[i - 1], center);
        MA(center, area / 3.0f, (*w)[i], center);
        MA(center, area / 3.0f, (*w)[0], center);
    }
    if (total != 0) { float s = 1.0f / total; center[0] = center[0] * s; center[1] = center[1] * s; center[2] = center[2] * s; }
    return total * 0.5f;
}

static int ClipWindingEpsilon(const Winding& in, const Vec3& normal, float dist, float epsilon, WindingP& front, WindingP& back) {   // :250-346
    float dists[MAX_POINTS_ON_WINDING + 4];
    int sides[MAX_POINTS_ON_WINDING + 4];
    int counts[3] = {0, 0, 0};
    int n = (int)in.size(), i;
    for (i = 0; i < n; i++) {
        float d = dot(in[i], normal);
        d -= dist;
        dists[i] = d;
        if (d > epsilon) sides[i] = 0; else if (d < -epsilon) sides[i] = 1; else sides[i] = 2;
        counts[sides[i]]++;
    }
    sides[i] = sides[0]; dists[i] = dists[0];
    back.reset(); front.reset();
    if (!counts[0]) { back = std::make_shared<Winding>(in); return 0; }
    if (!counts[1]) { front = std::make_shared<Winding>(in); return 0; }
    front = std::make_shared<Winding>(); back = std::make_shared<Winding>();
    for (i = 0; i < n; i++) {
        const Vec3& p1 = in[i];
        if (sides[i] == 2) { front->push_back(p1); back->push_back(p1); continue; }
        if (sides[i] == 0) front->push_back(p1);
        if (sides[i] == 1) back->push_back(p1);
        if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
        const Vec3& p2 = in[(i + 1) % n];
        float d = dists[i] / (dists[i] - dists[i + 1]);
        Vec3 mid;
        for (int j = 0; j < 3; j++) {
            if (normal[j] == 1) mid[j] = dist;
            else if (normal[j] == -1) mid[j] = -dist;
            else mid[j] = p1[j] + (d * (p2[j] - p1[j]));
        }
        front->push_back(mid); back->push_back(mid);
    }
    if ((int)front->size() > MAX_POINTS_ON_WINDING || (int)back->size() > MAX_POINTS_ON_WINDING) return -1;
    return 0;
}

struct Subdivider {
    std::vector<HPatch> patches;
    float minChop;
    bool failed = false;

    int CreateChildPatch(int nParentIndex, const WindingP& winding, float flArea, const Vec3& vecCenter) {   // :352-406
        patches.push_back(HPatch());
        int nChildIndex = (int)patches.size() - 1;
        HPatch& child = patches[nChildIndex];
        child = patches[nParentIndex];
        child.child1 = -1; child.child2 = -1; child.parent = nParentIndex;
        child.winding = winding; child.area = flArea; child.origin = vecCenter;
        WindingBounds(*child.winding, child.mins, child.maxs);
        if (child.base_light) return nChildIndex;
        Vec3 total = sub(child.maxs, child.mins);
        for (int k = 0; k < 3; k++) total[k] = total[k] * child.lux_scale;
        if (child.chop > minChop && total[0] < child.chop && total[1] < child.chop && total[2] < child.chop) {
            for (int i = 0; i < 3; i++) {
                if ((child.face_maxs[i] == child.maxs[i] || child.face_mins[i] == child.mins[i]) && total[i] > minChop) {
                    child.chop = fmaxf(minChop, child.chop / 2);
                    break;
                }
            }
        }
        return nChildIndex;
    }

    void SubdividePatch(int ndxPatch) {                                                     // :167-248
        if (failed) return;
        if (patches[ndxPatch].sky) return;
        float widest = -1; int widestAxis = -1; bool shouldSubDivide = false;
        Vec3 total = sub(patches[ndxPatch].maxs, patches[ndxPatch].mins);
        for (int k = 0; k < 3; k++) total[k] = total[k] * patches[ndxPatch].lux_scale;
        for (int i = 0; i < 3; i++) {
            if (total[i] > widest) { widestAxis = i; widest = total[i]; }
            if (total[i] >= patches[ndxPatch].chop && total[i] >= minChop) shouldSubDivide = true;
        }
        if (!shouldSubDivide && widestAxis != -1) {
            if (total[widestAxis] > total[(widestAxis + 1) % 3] * 2 && total[widestAxis] > total[(widestAxis + 2) % 3] * 2) {
                if (patches[ndxPatch].chop > minChop) {
                    shouldSubDivide = true;
                    patches[ndxPatch].chop = fmaxf(minChop, patches[ndxPatch].chop / 2);
                }
            }
        }
        if (!shouldSubDivide) return;
        Vec3 split = {{0, 0, 0}};
        split[widestAxis] = 1;
        float dist = (patches[ndxPatch].mins[widestAxis] + patches[ndxPatch].maxs[widestAxis]) * 0.5f;
        WindingP o1, o2;
        if (ClipWindingEpsilon(*patches[ndxPatch].winding, split, dist, ON_EPSILON, o1, o2)) { failed = true; return; }
        Vec3 center1, center2;
        float area1 = WindingAreaAndBalancePoint(o1, center1);
        float area2 = WindingAreaAndBalancePoint(o2, center2);
        if (area1 == 0 || area2 == 0) return;
        int c1 = CreateChildPatch(ndxPatch, o1, area1, center1);
        int c2 = CreateChildPatch(ndxPatch, o2, area2, center2);
        patches[ndxPatch].child1 = c1; patches[ndxPatch].child2 = c2;
        SubdividePatch(c1);
        SubdividePatch(c2);
    }
};

} // namespace orc

using namespace orc;

extern "C" {

// faces: the 36-byte vrad_face_patch record of include/vrad_cuda.h, passed as raw fields to keep the oracle
// independent of the product header.  Two-call pattern: outputs may be NULL to query the sizes.
int orc_patches_subdivide(int n_faces, const int32_t* first_point, const int32_t* n_points, const float* normal3,
                          const float* plane_dist_in, const float* lux_scale, const float* chop_in, const uint8_t* sky,
                          const uint8_t* no_subdivide, const uint8_t* has_base_light, const float* points3, float min_chop,
                          int* n_patches_out, int* n_points_out, float* origin3, float* normal3_out, float* plane_dist,
                          float* area, float* mins3, float* maxs3, float* chop, int32_t* parent, int32_t* child1,
                          int32_t* child2, int32_t* face, int32_t* wind_first, int32_t* wind_count, float* wind_points3) {
    Subdivider S;
    S.minChop = min_chop;
    for (int f = 0; f < n_faces; f++) {                                                     // MakePatchForFace
        WindingP w = std::make_shared<Winding>();
        for (int i = 0; i < n_points[f]; i++) {
            const float* p = points3 + 3 * (size_t)(first_point[f] + i);
            w->push_back({{p[0], p[1], p[2]}});
        }
        float a = WindingArea(*w);
        if (a <= 0) continue;                                                               // face.go:47-51
        HPatch P;
        P.parent = P.child1 = P.child2 = -1;
        P.area = a; P.sky = sky[f] != 0; P.lux_scale = lux_scale[f]; P.chop = chop_in[f];
        P.winding = w; P.plane_dist = plane_dist_in[f]; P.face = f; P.base_light = has_base_light[f] != 0;
        Vec3 c = {{0, 0, 0}};                                                               // WindingCenter
        for (size_t i = 0; i < w->size(); i++) { c[0] = (*w)[i][0] + c[0]; c[1] = (*w)[i][1] + c[1]; c[2] = (*w)[i][2] + c[2]; }
        float sc = 1.0f / (float)w->size();
        P.origin = {{c[0] * sc, c[1] * sc, c[2] * sc}};
        P.normal = {{normal3[3 * f], normal3[3 * f + 1], normal3[3 * f + 2]}};
        WindingBounds(*w, P.face_mins, P.face_maxs);
        P.mins = P.face_mins; P.maxs = P.face_maxs;
        S.patches.push_back(P);
    }
    int roots = (int)S.patches.size();
    for (int i = 0; i < roots; i++) {                                                       // SubdividePatches :50-66
        if (no_subdivide[S.patches[i].face]) continue;
        S.SubdividePatch(i);
    }
    if (S.failed) return -1;
    int np = (int)S.patches.size(), npt = 0;
    for (auto& p : S.patches) npt += (int)p.winding->size();
    *n_patches_out = np; *n_points_out = npt;
    if (!origin3) return 0;
    int wp = 0;
    for (int i = 0; i < np; i++) {
        const HPatch& p = S.patches[i];
        for (int k = 0; k < 3; k++) {
            origin3[3 * i + k] = p.origin[k]; normal3_out[3 * i + k] = p.normal[k]; mins3[3 * i + k] = p.mins[k]; maxs3[3 * i + k] = p.maxs[k];
        }
        plane_dist[i] = p.plane_dist; area[i] = p.area; chop[i] = p.chop;
        parent[i] = p.parent; child1[i] = p.child1; child2[i] = p.child2; face[i] = p.face;
        wind_first[i] = wp; wind_count[i] = (int)p.winding->size();
        for (auto& q : *p.winding) { wind_points3[3 * wp] = q[0]; wind_points3[3 * wp + 1] = q[1]; wind_points3[3 * wp + 2] = q[2]; wp++; }
    }
    return 0;
}

} // extern "C"
