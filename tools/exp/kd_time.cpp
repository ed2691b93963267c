// Development harness: times the product kd builder on a triangle file (n, then n*9 floats), prints stats.
#include <chrono>
#include <cstdio>
#include <vector>
#include "../../vrad_b200/csrc/kd_builder.hpp"
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); long n; fread(&n, 8, 1, f);
    std::vector<float> v(9 * n); fread(v.data(), 4, 9 * n, f); fclose(f);
    vrad::KdTree t;
    auto t0 = std::chrono::steady_clock::now();
    vrad::build_kd_tree(v.data(), (int)n, t);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("n=%ld nodes=%zu idx=%zu leaves=%d depth=%d  %.2f s\n", n, t.children.size(), t.tri_index.size(), t.n_leaves, t.max_depth, s);
    FILE* o = fopen(argv[2], "wb"); long nn = t.children.size(), ni = t.tri_index.size();
    fwrite(&nn, 8, 1, o); fwrite(&ni, 8, 1, o); fwrite(t.children.data(), 4, nn, o); fwrite(t.split.data(), 4, nn, o); fwrite(t.tri_index.data(), 4, ni, o); fclose(o);
    return 0;
}
