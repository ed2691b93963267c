// kd_builder.hpp -- host-side SAH kd-tree construction for libvradcuda (product code).
//
// Replaces Environment.SetupAccelerationStructure / RefineNode / CalculateCostsOfSplit
// (raytracer/environment.go:119-138, 238-387, 181-236) and
// OptimisedTriangle.ChangeIntoIntersectionFormat (raytracer/cache/optimisedtriangle.go:30-78).
// Output arrays use the reference's packed formats (raytracer/cache/optimisedkdnode.go:15-54,
// raytracer/cache/triangle/triintersectdata.go:3-22).
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/vrad_cuda.h"

namespace vrad {

struct KdTree {
    std::vector<int32_t> children;   // (index << 2) | axis, axis 3 = leaf
    std::vector<float>   split;      // split coordinate; for leaves the triangle count as a float
    std::vector<int32_t> tri_index;  // leaf triangle lists
    float bmin[3], bmax[3];
    int   max_depth = 0, n_leaves = 0;
};

// Builds the tree over n triangles given as 9 floats each.  Deterministic: the result does not
// depend on the number of host threads used.
void build_kd_tree(const float* verts9, int n, KdTree& out);

// Binned-SAH level-by-level builder (kd_fast.cu): same cost model, limits and output layout, candidates from 32 bins per axis.
// One source, two execution policies: the host's cores, or the device (stream = cudaStream_t as void*).  Returns a vrad_status;
// *why points at a static message on failure.
int build_kd_tree_binned_host(const float* verts9, int n, KdTree& out, const char** why);

// Triangle -> 48-byte intersection record (plane + two projected, normalised edge equations).
void make_intersection_records(const int32_t* ids, const float* verts9, const uint8_t* flags, int n, vrad_tri48* out);

// Structural validation of a tree in reference layout; returns the maximum leaf depth or -1
// (and fills err) if a child/triangle reference is out of range or the tree is not a tree.
int validate_kd_tree(int n_nodes, const int32_t* children, const float* split, int n_idx, const int32_t* tri_index,
                     int n_tris, int* n_leaves, const char** err);

} // namespace vrad
