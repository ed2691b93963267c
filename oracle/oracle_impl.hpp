// oracle_impl.hpp -- internal types of the CPU oracle (TEST INFRASTRUCTURE; see oracle.h).
#pragma once
#include "oracle.h"
#include <vector>
#include <cstdint>
#include <cmath>
#include <cstring>

namespace orc {

// maxps/minps-style selects: return b when either operand is NaN.  The CUDA path
// uses the same expressions so NaN-degenerate rays take the same branches.
static inline float min_sel(float a, float b) { return a < b ? a : b; }
static inline float max_sel(float a, float b) { return a > b ? a : b; }

struct KDNode {            // raytracer/cache/optimisedkdnode.go:15-54 (App. A #7: int32)
    int32_t children;      // (index<<2) | axis(0..2) or LEAF(3)
    float   split;         // split value; for leaves the triangle count as a float
};

struct TriGeom {           // raytracer/cache/triangle/trigeometrydata.go:3-12
    int32_t id;
    float   v[9];
    uint8_t flags;
    int8_t  tmp0, tmp1;
};

struct Patches {
    int n = 0;
    std::vector<float> origin, normal, plane_dist, area, refl;
    std::vector<int32_t> cluster;
    std::vector<uint8_t> flags;     // bit0 = sky
    // Patch.Parent / Child1 / Child2 / FaceNumber (common/types/patch.go:33,49-51); empty = flat (leaf patches only)
    std::vector<int32_t> parent, child1, child2, face;
    bool hier() const { return !child1.empty(); }
    // Patch.Winding (common/types/patch.go): first point / count per patch + points, clockwise seen from the front; empty = none
    std::vector<int32_t> wind_first, wind_count;
    std::vector<float> wind_pts;
    // Patch.NeedsBumpMap (common/types/patch.go:23) and the three bump normals of such a patch (normals[1..3]; normals[0] is
    // the flat patch normal); empty = no bump-mapped patches
    std::vector<uint8_t> needs_bump;
    std::vector<float> bump_normals;            // 9 per patch
    std::vector<float> total_bump;              // TotalLight.Light[1..3] (common/types/bumpLights.go:8-10), 9 per patch, from the last bounce call
};

struct Counters { int64_t nodes = 0, tris = 0, leaves = 0; };

// BSP point-location data read by trace.PointLeafnum (raytracer/trace/pointleaf.go:8-33) and
// clustertable.PointInLeaf (rad/clustertable/point.go:14-38): Nodes{PlaneNum, Children[2]},
// Planes{Normal, Distance, AxisType}, Leafs{Cluster, Area}, len(Areas).
struct Bsp {
    std::vector<int32_t> node_plane, node_children;      // children: 2 per node, negative = -1 - leaf
    std::vector<float>   plane_normal, plane_dist;
    std::vector<int32_t> plane_type;
    std::vector<int32_t> leaf_cluster, leaf_area;
    int n_areas = 0;
    bool set = false;
};

// cache/skycameras.go:8-40 + common/types/skycamera.go (Origin, WorldToSky, SkyToWorld, Area)
struct SkyCameras {
    std::vector<float>   origin, world_to_sky, sky_to_world;
    std::vector<int32_t> area;
    std::vector<int32_t> area_camera;                    // areaSkyCameras[], -1 = none
};

} // namespace orc

struct orc_env {
    std::vector<orc::TriGeom> geom;
    std::vector<orc_tri48>    tris;        // intersection format, same index as geom
    std::vector<orc::KDNode>  nodes;
    std::vector<int32_t>      tri_index;
    float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    int   max_depth = 0, n_leaves = 0;
    bool  built = false;
    double build_seconds = 0.0;
    orc::Counters counters;
    orc::Patches patches;
    std::vector<int64_t> rowptr;
    std::vector<int32_t> col;
    std::vector<float>   w;
    std::vector<float>   tri_color;   // Environment.TriangleColors (raytracer/environment.go:61-63), 3 per triangle
    orc::Bsp bsp;
    orc::SkyCameras cams;
    int light_trace_flags = 0;        // ORC_TL_CAN_RECURSE | ORC_TL_TEXTURE_SHADOWS for the light rays of orc_direct_light
};

namespace orc {

struct Hit { int32_t tri; float t; };

void tri_to_intersection_format(const TriGeom& g, orc_tri48& out);
void build_tree(orc_env* e);

// single-ray spec traversal.  any_hit_len >= 0 selects the visibility-only variant
// that stops at the first hit with t < any_hit_len.
Hit trace1(const orc_env* e, const float o[3], const float d[3], float tmin, float tmax,
           int32_t skip_id, Counters* ctr);
Hit trace_brute(const orc_env* e, const float o[3], const float d[3], float tmin, float tmax, int32_t skip_id);
void trace4(const orc_env* e, const float o[3][4], const float d[3][4], const float tmin[4], const float tmax[4],
            int32_t skip_id, int32_t hit_tri[4], float hit_t[4]);

// trace1 with the transparent-triangle rule (coverage != nullptr: CoverageCount semantics,
// raytracer/types/coverageCount.go:27-37); coverage == nullptr is exactly trace1.
Hit trace1_coverage(const orc_env* e, const float o[3], const float d[3], float tmin, float tmax,
                    int32_t skip_id, float* coverage);

// complete TestLineDoesHitSky on one segment (oracle/skytrace.cpp); leaf_point decides the recursion
float test_line_sky1(const orc_env* e, const float a[3], const float b[3], const float leaf_point[3], int flags, int32_t skip_id);
float test_line_fraction(const orc_env* e, const float a[3], const float b[3], int flags, int32_t skip_id);

// TestLine on one segment: returns 1 if visible.  mode: 0 spec, 2 brute.
int test_line1(const orc_env* e, const float a[3], const float b[3], int sky_mode, int mode);

} // namespace orc
