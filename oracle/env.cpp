// env.cpp -- ORACLE (test infrastructure): environment container + C entry points.
// Mirrors raytracer.Environment (raytracer/environment.go:28-69,119-138,422-432).
#include "oracle_impl.hpp"

extern "C" {

orc_env* orc_env_create(void) { return new orc_env(); }
void orc_env_destroy(orc_env* e) { delete e; }

int orc_env_add_triangles(orc_env* e, int n, const int32_t* ids, const float* verts9, const uint8_t* flags) {
    if (!e || n < 0 || e->built) return -1;
    for (int i = 0; i < n; i++) {
        orc::TriGeom g;
        g.id = ids[i];
        for (int k = 0; k < 9; k++) g.v[k] = verts9[9 * (size_t)i + k];
        g.flags = flags ? flags[i] : 0;
        g.tmp0 = g.tmp1 = 0;
        e->geom.push_back(g);
    }
    return 0;
}

int orc_env_build(orc_env* e) {
    if (!e || e->built) return -1;
    orc::build_tree(e);
    return 0;
}

int orc_env_sizes(orc_env* e, int* n_nodes, int* n_idx, int* n_tris, int* max_depth, int* n_leaves) {
    if (!e || !e->built) return -1;
    if (n_nodes) *n_nodes = (int)e->nodes.size();
    if (n_idx) *n_idx = (int)e->tri_index.size();
    if (n_tris) *n_tris = (int)e->tris.size();
    if (max_depth) *max_depth = e->max_depth;
    if (n_leaves) *n_leaves = e->n_leaves;
    return 0;
}

int orc_env_export(orc_env* e, int32_t* children, float* split, int32_t* tri_index, orc_tri48* tris, float aabb[6]) {
    if (!e || !e->built) return -1;
    for (size_t i = 0; i < e->nodes.size(); i++) {
        if (children) children[i] = e->nodes[i].children;
        if (split) split[i] = e->nodes[i].split;
    }
    if (tri_index) memcpy(tri_index, e->tri_index.data(), e->tri_index.size() * sizeof(int32_t));
    if (tris) memcpy(tris, e->tris.data(), e->tris.size() * sizeof(orc_tri48));
    if (aabb) for (int c = 0; c < 3; c++) { aabb[c] = e->bmin[c]; aabb[3 + c] = e->bmax[c]; }
    return 0;
}

// adopt a foreign tree over the same triangles (reference layout): lets the tests show that what a ray hits does not depend on which
// valid kd tree is walked.  The triangle records stay the oracle's own.
int orc_env_replace_tree(orc_env* e, int n_nodes, const int32_t* children, const float* split, int n_idx, const int32_t* tri_index, const float aabb[6]) {
    if (!e || !e->built || n_nodes <= 0 || !children || !split || (n_idx > 0 && !tri_index) || !aabb) return -1;
    e->nodes.resize((size_t)n_nodes);
    for (int i = 0; i < n_nodes; i++) { e->nodes[i].children = children[i]; e->nodes[i].split = split[i]; }
    e->tri_index.assign(tri_index, tri_index + n_idx);
    for (int c = 0; c < 3; c++) { e->bmin[c] = aabb[c]; e->bmax[c] = aabb[3 + c]; }
    return 0;
}

double orc_env_build_seconds(orc_env* e) { return e ? e->build_seconds : 0.0; }

} // extern "C"
