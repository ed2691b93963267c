// kd_builder.cpp -- host SAH kd-tree build (product code; see kd_builder.hpp for the reference map).
//
// Same split candidates, cost arithmetic (fp32, one rounding per operation, evaluation order of
// raytracer/environment.go:229-233), tie-breaking (first strictly lower cost wins, :289) and
// node/leaf numbering (children appended adjacently at split time, left subtree first, :361-385)
// as the reference's recursive RefineNode with the SURVEY.md App. A corrections -- but organised
// for throughput: an explicit DFS work stack instead of recursion, per-triangle axis extents
// precomputed once in SoA form, per-node extents gathered into contiguous scratch so the
// classification counts are branch-free vectorisable loops, and all candidates of a large node
// costed in parallel (OpenMP) followed by a sequential first-minimum scan, which keeps the
// result independent of the thread count.
#include "kd_builder.hpp"
#include <cmath>
#include <cstring>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace vrad {

namespace {

constexpr float kCostTraversal = 75.0f;      // raytracer/kdtree/constants.go:25
constexpr float kCostIntersection = 167.0f;  // raytracer/kdtree/constants.go:26
constexpr int   kMaxTreeDepth = 21;          // raytracer/kdtree/constants.go:28
constexpr int   kLeaf = VRAD_KDNODE_LEAF;

inline float surface_area(const float mn[3], const float mx[3]) {       // vmath/polygon/surface.go:5-8
    float d0 = mx[0] - mn[0], d1 = mx[1] - mn[1], d2 = mx[2] - mn[2];
    return 2.0f * (((d0 * d2) + (d0 * d1)) + (d1 * d2));
}

struct Candidate {
    int   axis;
    float split0;       // value the triangles are classified against
    float split;        // after growing the empty side
    float cost;
    int   nl, nr, nb;
};

struct Task {
    int node;
    std::vector<int> list;
    float mn[3], mx[3];
    int depth;          // reference depth counter (with the +100 small-node trick)
    int level;          // true tree level
};

struct Extents {        // per-triangle axis extents, SoA
    std::vector<float> lo[3], hi[3];
};

} // namespace

void build_kd_tree(const float* verts9, int n, KdTree& out) {
    out.children.clear(); out.split.clear(); out.tri_index.clear();
    out.max_depth = 0; out.n_leaves = 0;
    out.children.reserve(4 * (size_t)n + 16); out.split.reserve(4 * (size_t)n + 16);
    out.tri_index.reserve(3 * (size_t)n + 16);

    Extents ext;
    for (int a = 0; a < 3; a++) { ext.lo[a].resize(n); ext.hi[a].resize(n); }
    for (int c = 0; c < 3; c++) { out.bmin[c] = 1.0e23f; out.bmax[c] = -1.0e23f; }
    for (int i = 0; i < n; i++) {
        const float* v = verts9 + 9 * (size_t)i;
        for (int a = 0; a < 3; a++) {
            float lo = v[a], hi = v[a];
            for (int k = 1; k < 3; k++) { float c = v[3 * k + a]; lo = c < lo ? c : lo; hi = c > hi ? c : hi; }
            ext.lo[a][i] = lo; ext.hi[a][i] = hi;
        }
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < 3; c++) {
                float x = v[3 * k + c];
                out.bmin[c] = x < out.bmin[c] ? x : out.bmin[c];
                out.bmax[c] = x > out.bmax[c] ? x : out.bmax[c];
            }
    }

    auto new_node = [&]() { out.children.push_back(0); out.split.push_back(0.0f); return (int)out.children.size() - 1; };
    auto make_leaf = [&](const Task& t) {
        out.children[t.node] = kLeaf + ((int32_t)out.tri_index.size() << 2);
        out.split[t.node] = (float)t.list.size();
        for (int tri : t.list) out.tri_index.push_back(tri);
        out.n_leaves++;
        out.max_depth = std::max(out.max_depth, t.level);
    };

    std::vector<Task> stack;
    {
        Task root;
        root.node = new_node();
        root.list.resize(n);
        for (int i = 0; i < n; i++) root.list[i] = i;
        for (int c = 0; c < 3; c++) { root.mn[c] = out.bmin[c]; root.mx[c] = out.bmax[c]; }
        root.depth = 0; root.level = 0;
        stack.push_back(std::move(root));
    }

    std::vector<float> lo[3], hi[3];
    std::vector<Candidate> cands;

    while (!stack.empty()) {
        Task t = std::move(stack.back());
        stack.pop_back();
        const int m = (int)t.list.size();
        if (m < 3) { make_leaf(t); continue; }

        // gather this node's extents into contiguous scratch; list-order vertex scan for the
        // grown-split extremes (environment.go:194-201)
        float min_c[3], max_c[3];
        for (int a = 0; a < 3; a++) {
            lo[a].resize(m); hi[a].resize(m);
            float mnc = 1.0e23f, mxc = -1.0e23f;
            for (int k = 0; k < m; k++) {
                int tri = t.list[k];
                lo[a][k] = ext.lo[a][tri]; hi[a][k] = ext.hi[a][tri];
                const float* v = verts9 + 9 * (size_t)tri;
                for (int j = 0; j < 3; j++) { float c = v[3 * j + a]; mnc = c < mnc ? c : mnc; mxc = c > mxc ? c : mxc; }
            }
            min_c[a] = mnc; max_c[a] = mxc;
        }

        // enumerate candidates in the reference's order (environment.go:267-307)
        cands.clear();
        const int tri_skip = 1 + (m / 10);
        for (int axis = 0; axis < 3; axis++) {
            for (int ts = -1; ts < m; ts += tri_skip) {
                if (ts == -1) {
                    cands.push_back({axis, 0.5f * (t.mn[axis] + t.mx[axis]), 0, 0, 0, 0, 0});
                    continue;
                }
                const float* v = verts9 + 9 * (size_t)t.list[ts];
                for (int tv = 0; tv < 3; tv++) {
                    float s = v[3 * tv + axis];
                    if (s > t.mx[axis] || s < t.mn[axis]) continue;
                    cands.push_back({axis, s, 0, 0, 0, 0, 0});
                }
            }
        }

        const float isa = 1.0f / surface_area(t.mn, t.mx);
        const int nc = (int)cands.size();
        const bool par = (long long)nc * m > 200000;
#pragma omp parallel for schedule(dynamic, 1) if (par)
        for (int ci = 0; ci < nc; ci++) {
            Candidate& c = cands[ci];
            const float s = c.split0;
            const float* l = lo[c.axis].data(); const float* h = hi[c.axis].data();
            int nr = 0, nl = 0;
            for (int k = 0; k < m; k++) {
                int r = l[k] >= s;                   // ClassifyAgainstAxisSplit: POSITIVE first
                nr += r;
                nl += (!r) & (h[k] <= s);
            }
            int nb = m - nr - nl;
            float sp = s;
            if (nl != 0 && nb == 0 && nr == 0) sp = max_c[c.axis];
            if (nr != 0 && nb == 0 && nl == 0) sp = min_c[c.axis];
            float lmx[3] = {t.mx[0], t.mx[1], t.mx[2]}, rmn[3] = {t.mn[0], t.mn[1], t.mn[2]};
            lmx[c.axis] = sp; rmn[c.axis] = sp;
            float sa_l = surface_area(t.mn, lmx), sa_r = surface_area(rmn, t.mx);
            c.cost = kCostTraversal + kCostIntersection *
                     (((float)nb + ((sa_l * isa) * (float)nl)) + ((sa_r * isa) * (float)nr));
            c.split = sp; c.nl = nl; c.nr = nr; c.nb = nb;
        }
        int best = -1;
        float best_cost = 1.0e23f;
        for (int ci = 0; ci < nc; ci++)
            if (cands[ci].cost < best_cost) { best_cost = cands[ci].cost; best = ci; }

        const float cost_no_split = (float)(167 * m);
        if (best < 0 || cost_no_split <= best_cost || t.depth > kMaxTreeDepth) { make_leaf(t); continue; }

        const Candidate bc = cands[best];
        // partition: left block in list order, straddlers next, right block filled from the end
        // (environment.go:343-358) -> right child sees [straddlers | right triangles reversed]
        std::vector<int> left_list, right_list;
        left_list.reserve(bc.nl + bc.nb); right_list.resize(bc.nb + bc.nr);
        std::vector<int> both; both.reserve(bc.nb);
        {
            const float* l = lo[bc.axis].data(); const float* h = hi[bc.axis].data();
            int n_right = 0;
            for (int k = 0; k < m; k++) {
                bool r = l[k] >= bc.split0;
                bool lft = !r && h[k] <= bc.split0;
                if (lft) left_list.push_back(t.list[k]);
                else if (r) { n_right++; right_list[bc.nb + bc.nr - n_right] = t.list[k]; }
                else both.push_back(t.list[k]);
            }
        }
        left_list.insert(left_list.end(), both.begin(), both.end());
        std::copy(both.begin(), both.end(), right_list.begin());

        int left = new_node();
        new_node();
        out.children[t.node] = bc.axis + (left << 2);
        out.split[t.node] = bc.split;
        int depth = t.depth;
        if (m < 20 && (bc.nl == 0 || bc.nr == 0)) depth += 100;       // environment.go:378-380

        Task lt, rt;
        lt.node = left; rt.node = left + 1;
        lt.list = std::move(left_list); rt.list = std::move(right_list);
        for (int c = 0; c < 3; c++) { lt.mn[c] = t.mn[c]; lt.mx[c] = t.mx[c]; rt.mn[c] = t.mn[c]; rt.mx[c] = t.mx[c]; }
        lt.mx[bc.axis] = bc.split; rt.mn[bc.axis] = bc.split;
        lt.depth = rt.depth = depth + 1;
        lt.level = rt.level = t.level + 1;
        stack.push_back(std::move(rt));       // right is processed after the whole left subtree
        stack.push_back(std::move(lt));
    }
}

namespace {
inline void edge_equation(const float p1[3], const float p2[3], int c1, int c2, const float inside[3], float out[3]) {
    // vmath/polygon/edge.go:5-25: line through p1,p2 in the (c1,c2) projection, positive towards
    // `inside`, scaled to 1 at `inside` (-> barycentric coordinate)
    float nx = p1[c2] - p2[c2];
    float ny = p2[c1] - p1[c1];
    float d = -((nx * p1[c1]) + (ny * p1[c2]));
    float trial = ((inside[c1] * nx) + (inside[c2] * ny)) + d;
    if (trial < 0) { nx = -nx; ny = -ny; d = -d; trial = -trial; }
    out[0] = nx / trial; out[1] = ny / trial; out[2] = d / trial;
}
} // namespace

void make_intersection_records(const int32_t* ids, const float* verts9, const uint8_t* flags, int n, vrad_tri48* out) {
#pragma omp parallel for schedule(static) if (n > 50000)
    for (int i = 0; i < n; i++) {
        const float* p1 = verts9 + 9 * (size_t)i; const float* p2 = p1 + 3; const float* p3 = p1 + 6;
        float e1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        float e2[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
        float N[3] = {(e1[1] * e2[2]) - (e1[2] * e2[1]), (e1[2] * e2[0]) - (e1[0] * e2[2]), (e1[0] * e2[1]) - (e1[1] * e2[0])};
        float l = 1.0f / sqrtf(((N[0] * N[0]) + (N[1] * N[1])) + (N[2] * N[2]));
        N[0] = N[0] * l; N[1] = N[1] * l; N[2] = N[2] * l;
        int drop = 0;
        for (int c = 1; c < 3; c++) if (fabsf(N[c]) > fabsf(N[drop])) drop = c;
        vrad_tri48& t = out[i];
        t.nx = N[0]; t.ny = N[1]; t.nz = N[2];
        t.d = ((N[0] * p1[0]) + (N[1] * p1[1])) + (N[2] * p1[2]);
        t.id = ids[i];
        t.sel0 = (uint8_t)((drop + 1) % 3); t.sel1 = (uint8_t)((drop + 2) % 3);
        t.flags = flags ? flags[i] : 0; t.unused = 0;
        edge_equation(p1, p2, t.sel0, t.sel1, p3, &t.e[0]);
        edge_equation(p2, p3, t.sel0, t.sel1, p1, &t.e[3]);
    }
}

int validate_kd_tree(int n_nodes, const int32_t* children, const float* split, int n_idx, const int32_t* tri_index,
                     int n_tris, int* n_leaves, const char** err) {
    static const char* e_range = "kd tree: child index out of range";
    static const char* e_leaf = "kd tree: leaf triangle range out of bounds";
    static const char* e_tri = "kd tree: triangle index out of range";
    static const char* e_cycle = "kd tree: node reachable twice or cycle";
    if (n_nodes < 1) { if (err) *err = e_range; return -1; }
    for (int i = 0; i < n_idx; i++)
        if (tri_index[i] < 0 || tri_index[i] >= n_tris) { if (err) *err = e_tri; return -1; }
    std::vector<uint8_t> seen(n_nodes, 0);
    std::vector<std::pair<int, int>> st;
    st.push_back({0, 0});
    int max_depth = 0, leaves = 0;
    while (!st.empty()) {
        auto [node, depth] = st.back(); st.pop_back();
        if (seen[node]) { if (err) *err = e_cycle; return -1; }
        seen[node] = 1;
        int32_t c = children[node];
        if ((c & 3) == kLeaf) {
            int start = c >> 2;
            float cntf = split[node];
            if (!(cntf >= 0.0f) || cntf > (float)n_idx) { if (err) *err = e_leaf; return -1; }
            int cnt = (int)cntf;
            if (start < 0 || (long long)start + cnt > n_idx) { if (err) *err = e_leaf; return -1; }
            leaves++; max_depth = std::max(max_depth, depth);
        } else {
            int left = c >> 2;
            if (left <= 0 || left + 1 >= n_nodes) { if (err) *err = e_range; return -1; }
            st.push_back({left + 1, depth + 1}); st.push_back({left, depth + 1});
        }
    }
    if (n_leaves) *n_leaves = leaves;
    return max_depth;
}

} // namespace vrad
