// kd_build.cpp -- ORACLE (test infrastructure): SAH kd-tree build and triangle
// precomputation, restating raytracer/environment.go:119-138 (SetupAccelerationStructure),
// :181-236 (CalculateCostsOfSplit), :238-387 (RefineNode),
// raytracer/cache/optimisedtriangle.go:30-78 (ChangeIntoIntersectionFormat), :83-103
// (ClassifyAgainstAxisSplit), vmath/polygon/edge.go:5-25 (GetEdgeEquation),
// vmath/polygon/surface.go:5-8 (BoxSurfaceArea) with SURVEY.md App. A fixes #1-#9.
// Deliberately the literal recursive O(n * candidates) formulation.
#include "oracle_impl.hpp"
#include <chrono>

namespace orc {

static const int   PLANECHECK_POSITIVE = 1, PLANECHECK_NEGATIVE = -1, PLANECHECK_STRADDLING = 0; // optimisedtriangle.go:10-12
static const float COST_OF_TRAVERSAL = 75.0f, COST_OF_INTERSECTION = 167.0f;                     // kdtree/constants.go:25-26
static const int   MAX_TREE_DEPTH = 21;                                                          // kdtree/constants.go:28

// optimisedtriangle.go:22-28 with App. A #1 (x,y,z of vertex i)
static inline float vertex(const TriGeom& g, int i, int axis) { return g.v[3 * i + axis]; }

// vmath/polygon/surface.go:5-8
static inline float box_surface_area(const float mn[3], const float mx[3]) {
    float d0 = mx[0] - mn[0], d1 = mx[1] - mn[1], d2 = mx[2] - mn[2];
    return 2.0f * (((d0 * d2) + (d0 * d1)) + (d1 * d2));
}

// optimisedtriangle.go:83-103
static inline int classify(const TriGeom& g, int axis, float split) {
    float mn = vertex(g, 0, axis), mx = mn;
    for (int v = 0; v < 3; v++) {
        float c = vertex(g, v, axis);
        mn = c < mn ? c : mn;
        mx = c > mx ? c : mx;
    }
    if (mn >= split) return PLANECHECK_POSITIVE;
    if (mx <= split) return PLANECHECK_NEGATIVE;
    if (mn == mx) return PLANECHECK_POSITIVE;
    return PLANECHECK_STRADDLING;
}

// vmath/polygon/edge.go:5-25
static void edge_equation(const float p1[3], const float p2[3], int c1, int c2, const float inside[3], float out[3]) {
    float nx = p1[c2] - p2[c2];
    float ny = p2[c1] - p1[c1];
    float d = -((nx * p1[c1]) + (ny * p1[c2]));
    float trial = ((inside[c1] * nx) + (inside[c2] * ny)) + d;
    if (trial < 0) { nx = -nx; ny = -ny; d = -d; trial = -trial; }
    out[0] = nx / trial; out[1] = ny / trial; out[2] = d / trial;
}

// optimisedtriangle.go:30-78.  mgl32 Cross/Normalize/Dot restated as fp32:
// Normalize = v * (1/sqrt(x*x+y*y+z*z)).
void tri_to_intersection_format(const TriGeom& g, orc_tri48& t) {
    const float* p1 = &g.v[0]; const float* p2 = &g.v[3]; const float* p3 = &g.v[6];
    float e1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    float e2[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
    float N[3] = {(e1[1] * e2[2]) - (e1[2] * e2[1]),
                  (e1[2] * e2[0]) - (e1[0] * e2[2]),
                  (e1[0] * e2[1]) - (e1[1] * e2[0])};
    float l = 1.0f / sqrtf(((N[0] * N[0]) + (N[1] * N[1])) + (N[2] * N[2]));
    N[0] = N[0] * l; N[1] = N[1] * l; N[2] = N[2] * l;
    int drop = 0;
    for (int c = 1; c < 3; c++) if (fabsf(N[c]) > fabsf(N[drop])) drop = c;
    t.d = ((N[0] * p1[0]) + (N[1] * p1[1])) + (N[2] * p1[2]);
    t.nx = N[0]; t.ny = N[1]; t.nz = N[2];
    t.id = g.id; t.flags = g.flags; t.unused = 0;
    t.sel0 = (uint8_t)((drop + 1) % 3);
    t.sel1 = (uint8_t)((drop + 2) % 3);
    edge_equation(p1, p2, t.sel0, t.sel1, p3, &t.e[0]);
    edge_equation(p2, p3, t.sel0, t.sel1, p1, &t.e[3]);
}

struct Builder {
    orc_env* e;

    // environment.go:181-236 with App. A #2 (out-params)
    float cost_of_split(int axis, const int* list, int n, const float mn[3], const float mx[3],
                        float& split, int& nl, int& nr, int& nb) {
        nl = nr = nb = 0;
        float min_c = 1.0e23f, max_c = -1.0e23f;
        for (int t = 0; t < n; t++) {
            TriGeom& g = e->geom[list[t]];
            for (int v = 0; v < 3; v++) {
                float c = vertex(g, v, axis);
                min_c = c < min_c ? c : min_c;
                max_c = c > max_c ? c : max_c;
            }
            int cls = classify(g, axis, split);
            if (cls == PLANECHECK_NEGATIVE) nl++;
            else if (cls == PLANECHECK_POSITIVE) nr++;
            else nb++;
            g.tmp0 = (int8_t)cls;
        }
        if (nl != 0 && nb == 0 && nr == 0) split = max_c;   // "grow" the empty side
        if (nr != 0 && nb == 0 && nl == 0) split = min_c;
        float lmx[3] = {mx[0], mx[1], mx[2]}, rmn[3] = {mn[0], mn[1], mn[2]};
        lmx[axis] = split; rmn[axis] = split;
        float sa_l = box_surface_area(mn, lmx);
        float sa_r = box_surface_area(rmn, mx);
        float isa = 1.0f / box_surface_area(mn, mx);
        return COST_OF_TRAVERSAL + COST_OF_INTERSECTION *
               (((float)nb + ((sa_l * isa) * (float)nl)) + ((sa_r * isa) * (float)nr));
    }

    void make_leaf(int node, const int* list, int n, int depth_level) {
        e->nodes[node].children = ORC_KDNODE_LEAF + ((int32_t)e->tri_index.size() << 2);
        e->nodes[node].split = (float)n;             // optimisedkdnode.go:51-54
        for (int t = 0; t < n; t++) e->tri_index.push_back(list[t]);
        e->n_leaves++;
        if (depth_level > e->max_depth) e->max_depth = depth_level;
    }

    // environment.go:238-387.  `level` is the true tree level (for stats); `depth` is the
    // reference's depth counter including the +100 trick (:378-380).
    void refine(int node, const int* list, int n, const float mn[3], const float mx[3], int depth, int level) {
        if (n < 3) { make_leaf(node, list, n, level); return; }
        float best_cost = 1.0e23f, best_split = 0.0f;
        int best_nl = 0, best_nr = 0, best_nb = 0, split_plane = 0;
        int tri_skip = 1 + (n / 10);
        for (int axis = 0; axis < 3; axis++) {
            for (int ts = -1; ts < n; ts += tri_skip) {
                for (int tv = 0; tv < 3; tv++) {
                    int tnl, tnr, tnb;
                    float tsplit;
                    if (ts == -1) {
                        tsplit = 0.5f * (mn[axis] + mx[axis]);            // App. A #3
                    } else {
                        tsplit = vertex(e->geom[list[ts]], tv, axis);
                        if (tsplit > mx[axis] || tsplit < mn[axis]) continue;
                    }
                    float cost = cost_of_split(axis, list, n, mn, mx, tsplit, tnl, tnr, tnb);
                    if (cost < best_cost) {
                        split_plane = axis; best_cost = cost;
                        best_nl = tnl; best_nr = tnr; best_nb = tnb; best_split = tsplit;
                        for (int t = 0; t < n; t++) { TriGeom& g = e->geom[list[t]]; g.tmp1 = g.tmp0; }
                    }
                    if (ts == -1) break;
                }
            }
        }
        float cost_no_split = (float)(167 * n);
        if (cost_no_split <= best_cost || depth > MAX_TREE_DEPTH) { make_leaf(node, list, n, level); return; }

        std::vector<int> nl(n);
        float lmx[3] = {mx[0], mx[1], mx[2]}, rmn[3] = {mn[0], mn[1], mn[2]};
        lmx[split_plane] = best_split; rmn[split_plane] = best_split;
        int n_left = 0, n_both = 0, n_right = 0;
        for (int t = 0; t < n; t++) {
            const TriGeom& g = e->geom[list[t]];
            if (g.tmp1 == PLANECHECK_NEGATIVE) nl[n_left++] = list[t];                   // App. A #4
            else if (g.tmp1 == PLANECHECK_POSITIVE) { n_right++; nl[n - n_right] = list[t]; }
            else { nl[best_nl + n_both] = list[t]; n_both++; }
        }
        int left = (int)e->nodes.size();
        e->nodes[node].children = split_plane + (left << 2);
        e->nodes[node].split = best_split;
        e->nodes.push_back(KDNode{0, 0.0f});
        e->nodes.push_back(KDNode{0, 0.0f});
        if (n < 20 && (best_nl == 0 || best_nr == 0)) depth += 100;
        refine(left, nl.data(), best_nl + best_nb, mn, lmx, depth + 1, level + 1);
        refine(left + 1, nl.data() + best_nl, best_nr + best_nb, rmn, mx, depth + 1, level + 1);
    }
};

void build_tree(orc_env* e) {
    auto t0 = std::chrono::steady_clock::now();
    e->nodes.clear(); e->tri_index.clear(); e->max_depth = 0; e->n_leaves = 0;
    e->nodes.push_back(KDNode{0, 0.0f});
    int n = (int)e->geom.size();
    std::vector<int> root(n);
    for (int i = 0; i < n; i++) root[i] = i;
    // raytracer/math/trianglelist.go:9-21 with App. A #6
    for (int c = 0; c < 3; c++) { e->bmin[c] = 1.0e23f; e->bmax[c] = -1.0e23f; }
    for (int i = 0; i < n; i++)
        for (int v = 0; v < 3; v++)
            for (int c = 0; c < 3; c++) {
                float x = vertex(e->geom[i], v, c);
                e->bmin[c] = x < e->bmin[c] ? x : e->bmin[c];
                e->bmax[c] = x > e->bmax[c] ? x : e->bmax[c];
            }
    Builder b{e};
    b.refine(0, root.data(), n, e->bmin, e->bmax, 0, 0);
    e->tris.resize(n);
    for (int i = 0; i < n; i++) tri_to_intersection_format(e->geom[i], e->tris[i]);
    e->built = true;
    e->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace orc
