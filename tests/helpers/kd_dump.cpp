// Test helper (CPU only): runs the PRODUCT host code (vrad_b200/csrc/kd_builder.cpp: SAH kd build, triangle
// precompute, tree validation) on a triangle file and dumps the results so that tests/test_kd_builder_cpu.py can
// compare them with the oracle without a GPU.
//   in : int64 n | n int32 ids | n*9 float32 vertices
//   out: int64 n_nodes, n_idx, max_depth, n_leaves, validate_depth | children | split | tri_index | 6 floats aabb | n*48 bytes
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../vrad_b200/csrc/kd_builder.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    long n = 0;
    if (fread(&n, 8, 1, f) != 1) return 2;
    std::vector<int32_t> ids(n);
    std::vector<float> v(9 * (size_t)n);
    if (fread(ids.data(), 4, n, f) != (size_t)n || fread(v.data(), 4, 9 * (size_t)n, f) != 9 * (size_t)n) return 2;
    fclose(f);
    vrad::KdTree t;
    vrad::build_kd_tree(v.data(), (int)n, t);
    std::vector<vrad_tri48> tris(n);
    vrad::make_intersection_records(ids.data(), v.data(), nullptr, (int)n, tris.data());
    int leaves = 0; const char* why = nullptr;
    long vdepth = vrad::validate_kd_tree((int)t.children.size(), t.children.data(), t.split.data(), (int)t.tri_index.size(),
                                         t.tri_index.data(), (int)n, &leaves, &why);
    // a corrupted copy must be rejected
    std::vector<int32_t> bad = t.children;
    if (bad.size() > 1) { bad[0] = (int32_t)((bad.size() + 5) << 2); }
    long vbad = bad.size() > 1 ? vrad::validate_kd_tree((int)bad.size(), bad.data(), t.split.data(), (int)t.tri_index.size(),
                                                        t.tri_index.data(), (int)n, nullptr, &why) : -1;
    FILE* o = fopen(argv[2], "wb");
    long hdr[6] = {(long)t.children.size(), (long)t.tri_index.size(), t.max_depth, t.n_leaves, vdepth, vbad};
    fwrite(hdr, 8, 6, o);
    fwrite(t.children.data(), 4, t.children.size(), o);
    fwrite(t.split.data(), 4, t.split.size(), o);
    fwrite(t.tri_index.data(), 4, t.tri_index.size(), o);
    float aabb[6] = {t.bmin[0], t.bmin[1], t.bmin[2], t.bmax[0], t.bmax[1], t.bmax[2]};
    fwrite(aabb, 4, 6, o);
    fwrite(tris.data(), 48, n, o);
    fclose(o);
    return leaves == t.n_leaves ? 0 : 3;
}
