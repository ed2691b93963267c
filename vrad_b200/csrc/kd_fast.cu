// kd_fast.cu -- binned-SAH kd-tree construction, level by level, for the device ("GPU kd build (binned SAH)
// for the 1M-tri case") and -- from the same source -- for the host's cores.
//
// What it stands for: the reference's RTE_FLAGS_FAST_TREE_GENERATION (raytracer/constants.go:5), a flag the Go code declares and
// never reads; the exact builder (RefineNode / CalculateCostsOfSplit, raytracer/environment.go:181-387; here kd_builder.cpp)
// tries every triSkip-th vertex as a split candidate and is O(n * candidates) per node.  This builder keeps the reference's cost
// model and limits -- cost = COST_OF_TRAVERSAL + COST_OF_INTERSECTION * (nBoth + SA_L/SA * nLeft + SA_R/SA * nRight)
// (environment.go:229-233, raytracer/kdtree/constants.go:25-27), leaf when COST_OF_INTERSECTION * n <= best (:310), n < 3 (:241),
// depth > MAX_TREE_DEPTH = 21 (constants.go:28), left child gets left + both, right child gets both + right (:381-385), on-plane
// triangles go right (optimisedtriangle.go:99-101) -- but takes its candidates from 32 equal bins per axis of the node's box and
// counts triangles by their (clipped) bounding boxes.  Output is the reference's packed layout (OptimisedKDNode, TriangleIndexList),
// so the traversal kernels run on it unchanged.  Away from knife edges the closest hit of a ray does not depend on which valid
// tree is walked (ties resolve by triangle index); a ray grazing a triangle edge that lies in a split plane touches a leaf in one
// point and may or may not visit it.  Parity therefore: K1 on this tree is bit-exact against the oracle's tracer walking THIS tree
// (OracleEnv.replace_tree), and equal to the exact tree's results except on such rays (tests/test_kd_fast_cpu.py).
//
// Formulation: breadth first.  A level holds its active nodes (box, output node index, range of triangle references) and one
// flat array of references grouped by node.  Per level: (1) thread per reference adds its clipped box to the 3 x 32 start / end
// bins of its node (integer atomics); (2) thread per node sweeps the 93 candidate planes; (3) thread per reference classifies
// itself against its node's plane; (4) two exclusive scans over the reference flags give every reference its slot in the left /
// right child -- a stable partition, so the tree does not depend on thread scheduling; (5) thread per node emits the packed node
// and its two children, thread per reference moves itself (or, in a leaf, writes its triangle into TriangleIndexList).
// Steps are functors run through an execution policy: a generic CUDA kernel + cub::DeviceScan on the device, OpenMP loops on the
// host.  The host policy is what the CPU tests exercise (the tree it makes is identical to the device's by construction: integer
// atomics and stable scans are order-independent, and device code is compiled with -fmad=false like the host's -ffp-contract=off).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cub/cub.cuh>
#include "env_internal.cuh"
#include "kd_builder.hpp"

namespace vrad {
namespace kdfast {

constexpr int   kBins = 32;
constexpr int   kMaxDepth = 21;            // raytracer/kdtree/constants.go:28
constexpr float kCostTraversal = 75.0f;    // :25
constexpr float kCostIntersection = 167.0f;  // :27

#if defined(__CUDACC__)
#define KD_HD __host__ __device__ __forceinline__
#else
#define KD_HD inline
#endif

KD_HD void atomic_inc(int32_t* p) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, 1);
#else
    __atomic_fetch_add(p, 1, __ATOMIC_RELAXED);
#endif
}

// order-preserving map float -> uint32, so that the largest coordinate of a set is an integer atomicMax (order independent)
KD_HD uint32_t float_key(float f) {
#if defined(__CUDA_ARCH__)
    const uint32_t u = __float_as_uint(f);
#else
    uint32_t u; std::memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
KD_HD float key_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}
KD_HD void atomic_max_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    uint32_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
#endif
}

struct LevelNode {
    float lo[3], hi[3];
    int32_t node;                 // index in the output node arrays
    int32_t ref_begin, ref_end;   // this node's slice of the level's reference array
    int32_t depth;
};

struct Split { int32_t axis; float pos; };      // axis -1 = leaf

// ---- the steps ---------------------------------------------------------------------------------------------------------------

struct TriBounds {                 // per triangle, once
    const float* verts9; float* tmin; float* tmax;
    KD_HD void operator()(int64_t t) const {
        const float* v = verts9 + 9 * t;
        for (int a = 0; a < 3; a++) {
            float mn = v[a], mx = v[a];
            if (v[3 + a] < mn) mn = v[3 + a]; if (v[3 + a] > mx) mx = v[3 + a];
            if (v[6 + a] < mn) mn = v[6 + a]; if (v[6 + a] > mx) mx = v[6 + a];
            tmin[3 * t + a] = mn; tmax[3 * t + a] = mx;
        }
    }
};

struct InitRefs {
    int32_t* ref_tri; int32_t* ref_slot;
    KD_HD void operator()(int64_t r) const { ref_tri[r] = (int32_t)r; ref_slot[r] = 0; }
};

// Clamped in float BEFORE the conversion: for a denormal box width inv is +inf, the product +inf or NaN, and converting that
// is undefined on the host (cvttss2si gives INT_MIN -> bin 0) but saturates on the device (bin 31) -- host and device policies
// would build different trees from the same triangles.  NaN compares false both times and lands in bin 0 everywhere.
KD_HD int bin_of(float x, float lo, float inv) {
    const float f = (x - lo) * inv;
    if (!(f > 0.0f)) return 0;
    if (f >= (float)(kBins - 1)) return kBins - 1;
    return (int)f;
}

struct BinRefs {
    const LevelNode* nodes; const int32_t* ref_tri; const int32_t* ref_slot; const float* tmin; const float* tmax;
    int32_t* bin_lo; int32_t* bin_hi;          // [slot][axis][bin]
    uint32_t* bin_end_max;                     // float_key of the largest (clipped) end coordinate among the references ending in the bin; 0 = none
    KD_HD void operator()(int64_t r) const {
        const int slot = ref_slot[r], t = ref_tri[r];
        const LevelNode& nd = nodes[slot];
        for (int a = 0; a < 3; a++) {
            const float w = nd.hi[a] - nd.lo[a];
            int b0 = 0, b1 = 0;
            float mx = nd.hi[a];
            if (w > 0.0f) {
                const float inv = (float)kBins / w;
                float mn = tmin[3 * (int64_t)t + a];
                mx = tmax[3 * (int64_t)t + a];
                if (mn < nd.lo[a]) mn = nd.lo[a];
                if (mx > nd.hi[a]) mx = nd.hi[a];
                b0 = bin_of(mn, nd.lo[a], inv); b1 = bin_of(mx, nd.lo[a], inv);
                if (b1 < b0) b1 = b0;
            }
            atomic_inc(&bin_lo[((int64_t)slot * 3 + a) * kBins + b0]);
            atomic_inc(&bin_hi[((int64_t)slot * 3 + a) * kBins + b1]);
            atomic_max_u32(&bin_end_max[((int64_t)slot * 3 + a) * kBins + b1], float_key(mx));
        }
    }
};

KD_HD float box_area(const float lo[3], const float hi[3]) {        // polygon.BoxSurfaceArea, vmath/polygon/surface.go:5-8
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * ((dx * dy + dy * dz) + dz * dx);
}

struct ChooseSplit {
    const LevelNode* nodes; const int32_t* bin_lo; const int32_t* bin_hi; const uint32_t* bin_end_max; Split* split; int force_leaf;
    KD_HD void operator()(int64_t slot) const {
        const LevelNode& nd = nodes[slot];
        const int n = nd.ref_end - nd.ref_begin;
        Split s = {-1, 0.0f};
        if (!force_leaf && n >= 3 && nd.depth <= kMaxDepth) {
            float best = kCostIntersection * (float)n;               // the cost of staying a leaf (environment.go:310)
            const float sa = box_area(nd.lo, nd.hi);
            if (sa > 0.0f) {
                const float inv_sa = 1.0f / sa;
                for (int a = 0; a < 3; a++) {
                    const float w = nd.hi[a] - nd.lo[a];
                    if (!(w > 0.0f)) continue;
                    const int32_t* bl = bin_lo + ((int64_t)slot * 3 + a) * kBins;
                    const int32_t* bh = bin_hi + ((int64_t)slot * 3 + a) * kBins;
                    const uint32_t* be = bin_end_max + ((int64_t)slot * 3 + a) * kBins;
                    int n_start = 0, n_end = 0;                      // references starting / ending in bins 0..p
                    uint32_t end_key = 0;                            // largest end coordinate in bins 0..p
                    for (int p = 0; p < kBins - 1; p++) {
                        n_start += bl[p]; n_end += bh[p];
                        if (be[p] > end_key) end_key = be[p];
                        // The plane after bin p.  Pulled back from the bin boundary to the largest coordinate at which a reference
                        // of bins 0..p ends, when that lies in bin p: everything counted as "left only" still is (end <= plane), and
                        // references that start between there and the boundary stop straddling -- on grid-aligned geometry the
                        // plane lands on the shared vertex coordinates, as the exact builder's vertex candidates do
                        // (environment.go:265-307).
                        float pos = nd.lo[a] + w * ((float)(p + 1) / (float)kBins);
                        if (bh[p] > 0) {
                            const float snapped = key_float(end_key);
                            const float bin_floor = nd.lo[a] + w * ((float)p / (float)kBins);
                            if (snapped > bin_floor && snapped < pos) pos = snapped;
                        }
                        if (!(pos > nd.lo[a] && pos < nd.hi[a])) continue;
                        const int n_left_only = n_end;               // end before the plane
                        const int n_right_only = n - n_start;        // start after the plane
                        const int n_both = n - n_left_only - n_right_only;
                        float lhi[3] = {nd.hi[0], nd.hi[1], nd.hi[2]}, rlo[3] = {nd.lo[0], nd.lo[1], nd.lo[2]};
                        lhi[a] = pos; rlo[a] = pos;
                        const float sa_l = box_area(nd.lo, lhi), sa_r = box_area(rlo, nd.hi);
                        const float cost = kCostTraversal + kCostIntersection * (((float)n_both + (sa_l * inv_sa) * (float)n_left_only) + (sa_r * inv_sa) * (float)n_right_only);
                        if (cost < best) { best = cost; s.axis = a; s.pos = pos; }
                    }
                }
            }
        }
        split[slot] = s;
    }
};

// which children a reference goes to: bit 0 = left, bit 1 = right.  Uses the clipped box, so a reference always lands somewhere.
KD_HD int classify(const LevelNode& nd, const Split& s, const float* tmin, const float* tmax, int t) {
    float mn = tmin[3 * (int64_t)t + s.axis], mx = tmax[3 * (int64_t)t + s.axis];
    if (mn < nd.lo[s.axis]) mn = nd.lo[s.axis];
    if (mx > nd.hi[s.axis]) mx = nd.hi[s.axis];
    const int left = mn < s.pos ? 1 : 0;
    const int right = (mx > s.pos || !left) ? 2 : 0;                 // on the plane: right (optimisedtriangle.go:99-101)
    return left | right;
}

struct Classify {
    const LevelNode* nodes; const Split* split; const int32_t* ref_tri; const int32_t* ref_slot; const float* tmin; const float* tmax;
    int32_t* flag_l; int32_t* flag_r;
    KD_HD void operator()(int64_t r) const {
        const int slot = ref_slot[r];
        const Split s = split[slot];
        int c = 0;
        if (s.axis >= 0) c = classify(nodes[slot], s, tmin, tmax, ref_tri[r]);
        flag_l[r] = c & 1; flag_r[r] = (c >> 1) & 1;
    }
};

struct NodeCounts {                 // per node: is it split, how many leaf references, how many references the children take
    const LevelNode* nodes; const Split* split; const int32_t* scan_l; const int32_t* scan_r;
    int32_t* is_split; int32_t* leaf_refs; int32_t* child_refs;
    KD_HD void operator()(int64_t slot) const {
        const LevelNode& nd = nodes[slot];
        const bool sp = split[slot].axis >= 0;
        is_split[slot] = sp ? 1 : 0;
        leaf_refs[slot] = sp ? 0 : nd.ref_end - nd.ref_begin;
        child_refs[slot] = sp ? (scan_l[nd.ref_end] - scan_l[nd.ref_begin]) + (scan_r[nd.ref_end] - scan_r[nd.ref_begin]) : 0;
    }
};

struct EmitNodes {
    const LevelNode* nodes; const Split* split; const int32_t* scan_l; const int32_t* scan_r;
    const int32_t* pair_index; const int32_t* leaf_ofs; const int32_t* child_ofs;       // exclusive scans over the node flags above
    int32_t node_base, idx_base;                                                       // output nodes / TriangleIndexList entries so far
    int32_t* out_children; float* out_split; LevelNode* next;
    KD_HD void operator()(int64_t slot) const {
        const LevelNode& nd = nodes[slot];
        const Split s = split[slot];
        if (s.axis < 0) {            // leaf: (start << 2) | LEAF, count as a float (optimisedkdnode.go:29-32,44-54)
            out_children[nd.node] = (int32_t)(((uint32_t)(idx_base + leaf_ofs[slot]) << 2) | 3u);
            out_split[nd.node] = (float)(nd.ref_end - nd.ref_begin);
            return;
        }
        const int pair = pair_index[slot];
        const int left_node = node_base + 2 * pair;                   // children adjacent, right = left + 1 (optimisedkdnode.go:39-41)
        out_children[nd.node] = (int32_t)(((uint32_t)left_node << 2) | (uint32_t)s.axis);
        out_split[nd.node] = s.pos;
        const int nl = scan_l[nd.ref_end] - scan_l[nd.ref_begin], nr = scan_r[nd.ref_end] - scan_r[nd.ref_begin];
        LevelNode l = nd, r = nd;
        l.hi[s.axis] = s.pos; r.lo[s.axis] = s.pos;
        l.node = left_node; r.node = left_node + 1;
        l.depth = r.depth = nd.depth + 1;
        l.ref_begin = child_ofs[slot]; l.ref_end = l.ref_begin + nl;
        r.ref_begin = l.ref_end; r.ref_end = r.ref_begin + nr;
        next[2 * pair] = l; next[2 * pair + 1] = r;
    }
};

struct MoveRefs {
    const LevelNode* nodes; const Split* split; const int32_t* ref_tri; const int32_t* ref_slot;
    const int32_t* flag_l; const int32_t* flag_r; const int32_t* scan_l; const int32_t* scan_r;
    const int32_t* pair_index; const int32_t* leaf_ofs; const int32_t* child_ofs; int32_t idx_base;
    int32_t* out_tri_index; int32_t* next_tri; int32_t* next_slot;
    KD_HD void operator()(int64_t r) const {
        const int slot = ref_slot[r], t = ref_tri[r];
        const LevelNode& nd = nodes[slot];
        if (split[slot].axis < 0) { out_tri_index[idx_base + leaf_ofs[slot] + ((int32_t)r - nd.ref_begin)] = t; return; }
        const int pair = pair_index[slot];
        const int nl = scan_l[nd.ref_end] - scan_l[nd.ref_begin];
        if (flag_l[r]) { const int d = child_ofs[slot] + (scan_l[r] - scan_l[nd.ref_begin]); next_tri[d] = t; next_slot[d] = 2 * pair; }
        if (flag_r[r]) { const int d = child_ofs[slot] + nl + (scan_r[r] - scan_r[nd.ref_begin]); next_tri[d] = t; next_slot[d] = 2 * pair + 1; }
    }
};

// ---- execution policies --------------------------------------------------------------------------------------------------------

template <class F> __global__ void __launch_bounds__(256) k_for_each(int64_t n, F f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}

struct DeviceExec {
    cudaStream_t stream;
    int launches = 0;
    cudaError_t err = cudaSuccess;
    void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
    ~DeviceExec() { if (scan_tmp) cudaFree(scan_tmp); }
    void note(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }
    void* alloc(size_t bytes) { void* p = nullptr; note(cudaMalloc(&p, bytes ? bytes : 1)); return p; }
    void free(void* p) { if (p) cudaFree(p); }
    void zero(void* p, size_t bytes) { if (bytes) note(cudaMemsetAsync(p, 0, bytes, stream)); }
    void upload(void* d, const void* h, size_t bytes) { if (bytes) note(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream)); }
    void download(void* h, const void* d, size_t bytes) { if (bytes) note(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream)); note(cudaStreamSynchronize(stream)); }
    template <class F> void for_each(int64_t n, F f) {
        if (n <= 0) return;
        k_for_each<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, f);
        note(cudaGetLastError()); launches++;
    }
    // out[0..n] = exclusive prefix sums of in[0..n-1] (out has n + 1 entries: the last is the total)
    void scan(const int32_t* in, int32_t* out, int64_t n) {
        zero(out, 4);
        if (n <= 0) return;
        size_t need = 0;
        cub::DeviceScan::InclusiveSum(nullptr, need, in, out + 1, (int)n, stream);
        if (need > scan_tmp_bytes) { if (scan_tmp) cudaFree(scan_tmp); note(cudaMalloc(&scan_tmp, need)); scan_tmp_bytes = need; }
        note(cub::DeviceScan::InclusiveSum(scan_tmp, need, in, out + 1, (int)n, stream)); launches++;
    }
};

struct HostExec {
    int launches = 0;
    cudaError_t err = cudaSuccess;
    void* alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
    void free(void* p) { std::free(p); }
    void zero(void* p, size_t bytes) { if (bytes) std::memset(p, 0, bytes); }
    void upload(void* d, const void* h, size_t bytes) { if (bytes) std::memcpy(d, h, bytes); }
    void download(void* h, const void* d, size_t bytes) { if (bytes) std::memcpy(h, d, bytes); }
    template <class F> void for_each(int64_t n, F f) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) f(i);
    }
    void scan(const int32_t* in, int32_t* out, int64_t n) {
        int32_t acc = 0;
        out[0] = 0;
        for (int64_t i = 0; i < n; i++) { acc += in[i]; out[i + 1] = acc; }
    }
};

// ---- the driver loop -------------------------------------------------------------------------------------------------------------

template <class Exec>
int build(Exec& ex, const float* verts9_exec /* in the policy's memory */, int n, const float scene_lo[3], const float scene_hi[3], KdTree& out, const char** why) {
    const int64_t cap_refs = std::max<int64_t>((int64_t)n * 8, 1 << 16);        // references alive on one level
    const int64_t cap_nodes = std::max<int64_t>((int64_t)n * 8, 1 << 12);       // output nodes
    const int64_t cap_idx = std::max<int64_t>((int64_t)n * 16, 1 << 16);        // TriangleIndexList entries
    const int64_t cap_level = cap_refs;                                          // active nodes on one level (each holds >= 0 references; bounded below)
    auto A = [&](size_t bytes) { return ex.alloc(bytes); };
    float* tmin = (float*)A(12 * (size_t)std::max(n, 1)); float* tmax = (float*)A(12 * (size_t)std::max(n, 1));
    int32_t* ref_tri[2] = {(int32_t*)A(4 * cap_refs), (int32_t*)A(4 * cap_refs)};
    int32_t* ref_slot[2] = {(int32_t*)A(4 * cap_refs), (int32_t*)A(4 * cap_refs)};
    int32_t* flag_l = (int32_t*)A(4 * cap_refs); int32_t* flag_r = (int32_t*)A(4 * cap_refs);
    int32_t* scan_l = (int32_t*)A(4 * (cap_refs + 1)); int32_t* scan_r = (int32_t*)A(4 * (cap_refs + 1));
    LevelNode* level[2] = {(LevelNode*)A(sizeof(LevelNode) * cap_level), (LevelNode*)A(sizeof(LevelNode) * cap_level)};
    Split* split = (Split*)A(sizeof(Split) * cap_level);
    int32_t* bin_lo = nullptr; int32_t* bin_hi = nullptr; uint32_t* bin_end_max = nullptr; int64_t bin_cap = 0;    // grown per level: 3 * 32 counters per active node
    int32_t* is_split = (int32_t*)A(4 * cap_level); int32_t* leaf_refs = (int32_t*)A(4 * cap_level); int32_t* child_refs = (int32_t*)A(4 * cap_level);
    int32_t* pair_index = (int32_t*)A(4 * (cap_level + 1)); int32_t* leaf_ofs = (int32_t*)A(4 * (cap_level + 1)); int32_t* child_ofs = (int32_t*)A(4 * (cap_level + 1));
    int32_t* out_children = (int32_t*)A(4 * cap_nodes); float* out_split = (float*)A(4 * cap_nodes);
    int32_t* out_idx = (int32_t*)A(4 * cap_idx);
    int rc = VRAD_OK;
    auto fail = [&](int code, const char* msg) { rc = code; *why = msg; };

    bool have_all = true;
    for (const void* p : {(const void*)tmin, (const void*)tmax, (const void*)ref_tri[0], (const void*)ref_tri[1], (const void*)ref_slot[0], (const void*)ref_slot[1],
                          (const void*)flag_l, (const void*)flag_r, (const void*)scan_l, (const void*)scan_r, (const void*)level[0], (const void*)level[1],
                          (const void*)split, (const void*)is_split, (const void*)leaf_refs, (const void*)child_refs, (const void*)pair_index,
                          (const void*)leaf_ofs, (const void*)child_ofs, (const void*)out_children, (const void*)out_split, (const void*)out_idx})
        have_all = have_all && p != nullptr;
    if (!have_all || ex.err != cudaSuccess) { fail(VRAD_E_NOMEM, "binned kd build: out of memory for the level buffers"); }   // no kernel runs on a null buffer

    if (rc == VRAD_OK) {
    ex.for_each(n, TriBounds{verts9_exec, tmin, tmax});
    ex.for_each(n, InitRefs{ref_tri[0], ref_slot[0]});
    LevelNode root;
    for (int a = 0; a < 3; a++) { root.lo[a] = scene_lo[a]; root.hi[a] = scene_hi[a]; }
    root.node = 0; root.ref_begin = 0; root.ref_end = n; root.depth = 0;
    ex.upload(level[0], &root, sizeof root);
    }
    int64_t n_active = rc == VRAD_OK ? 1 : 0, n_refs = n, n_nodes = 1, n_idx = 0;
    int cur = 0, max_depth = 0, n_leaves = 0;
    for (int depth = 0; n_active > 0 && rc == VRAD_OK; depth++) {
        if (depth > kMaxDepth + 2) { fail(VRAD_E_UNSUPPORTED, "binned kd build: level loop did not terminate"); break; }
        if (n_active * 3 * kBins > bin_cap) {
            ex.free(bin_lo); ex.free(bin_hi); ex.free(bin_end_max);
            bin_cap = n_active * 3 * kBins;
            bin_lo = (int32_t*)A(4 * (size_t)bin_cap); bin_hi = (int32_t*)A(4 * (size_t)bin_cap); bin_end_max = (uint32_t*)A(4 * (size_t)bin_cap);
            if (!bin_lo || !bin_hi || !bin_end_max || ex.err != cudaSuccess) { fail(VRAD_E_NOMEM, "binned kd build: out of memory for the bins"); break; }
        }
        int32_t totals[3] = {0, 0, 0};                                            // split nodes, leaf references, child references
        for (int force_leaf = 0; force_leaf < 2; force_leaf++) {
            if (!force_leaf) {
                ex.zero(bin_lo, 4 * (size_t)n_active * 3 * kBins); ex.zero(bin_hi, 4 * (size_t)n_active * 3 * kBins);
                ex.zero(bin_end_max, 4 * (size_t)n_active * 3 * kBins);
                ex.for_each(n_refs, BinRefs{level[cur], ref_tri[cur], ref_slot[cur], tmin, tmax, bin_lo, bin_hi, bin_end_max});
            }
            ex.for_each(n_active, ChooseSplit{level[cur], bin_lo, bin_hi, bin_end_max, split, force_leaf});
            ex.for_each(n_refs, Classify{level[cur], split, ref_tri[cur], ref_slot[cur], tmin, tmax, flag_l, flag_r});
            ex.scan(flag_l, scan_l, n_refs); ex.scan(flag_r, scan_r, n_refs);
            ex.for_each(n_active, NodeCounts{level[cur], split, scan_l, scan_r, is_split, leaf_refs, child_refs});
            ex.scan(is_split, pair_index, n_active); ex.scan(leaf_refs, leaf_ofs, n_active); ex.scan(child_refs, child_ofs, n_active);
            ex.download(&totals[0], pair_index + n_active, 4); ex.download(&totals[1], leaf_ofs + n_active, 4); ex.download(&totals[2], child_ofs + n_active, 4);
            if (ex.err != cudaSuccess) break;
            // out of room for the next level: finish this level as leaves instead (the tree stays valid, only shallower)
            if (totals[2] <= cap_refs && 2 * (int64_t)totals[0] <= cap_level && n_nodes + 2 * (int64_t)totals[0] <= cap_nodes) break;
        }
        if (ex.err != cudaSuccess) { fail(VRAD_E_CUDA, cudaGetErrorString(ex.err)); break; }
        if (n_idx + totals[1] > cap_idx) { fail(VRAD_E_NOMEM, "binned kd build: triangle index list exceeds 16 entries per triangle"); break; }
        ex.for_each(n_active, EmitNodes{level[cur], split, scan_l, scan_r, pair_index, leaf_ofs, child_ofs, (int32_t)n_nodes, (int32_t)n_idx,
                                        out_children, out_split, level[cur ^ 1]});
        ex.for_each(n_refs, MoveRefs{level[cur], split, ref_tri[cur], ref_slot[cur], flag_l, flag_r, scan_l, scan_r, pair_index, leaf_ofs, child_ofs,
                                     (int32_t)n_idx, out_idx, ref_tri[cur ^ 1], ref_slot[cur ^ 1]});
        const int64_t leaves_here = n_active - totals[0];
        if (leaves_here > 0) { n_leaves += (int)leaves_here; max_depth = depth; }
        n_nodes += 2 * (int64_t)totals[0]; n_idx += totals[1];
        n_active = 2 * (int64_t)totals[0]; n_refs = totals[2];
        cur ^= 1;
    }
    if (rc == VRAD_OK) {
        out.children.resize((size_t)n_nodes); out.split.resize((size_t)n_nodes); out.tri_index.resize((size_t)n_idx);
        ex.download(out.children.data(), out_children, 4 * (size_t)n_nodes);
        ex.download(out.split.data(), out_split, 4 * (size_t)n_nodes);
        ex.download(out.tri_index.data(), out_idx, 4 * (size_t)n_idx);
        if (ex.err != cudaSuccess) fail(VRAD_E_CUDA, cudaGetErrorString(ex.err));
        for (int a = 0; a < 3; a++) { out.bmin[a] = scene_lo[a]; out.bmax[a] = scene_hi[a]; }
        out.max_depth = max_depth; out.n_leaves = n_leaves;
    }
    for (void* p : {(void*)tmin, (void*)tmax, (void*)ref_tri[0], (void*)ref_tri[1], (void*)ref_slot[0], (void*)ref_slot[1], (void*)flag_l, (void*)flag_r,
                    (void*)scan_l, (void*)scan_r, (void*)level[0], (void*)level[1], (void*)split, (void*)bin_lo, (void*)bin_hi, (void*)bin_end_max, (void*)is_split,
                    (void*)leaf_refs, (void*)child_refs, (void*)pair_index, (void*)leaf_ofs, (void*)child_ofs, (void*)out_children, (void*)out_split, (void*)out_idx})
        ex.free(p);
    return rc;
}

// scene box exactly as the exact builder computes it (kd_builder.cpp: min / max over all vertices)
void scene_bounds(const float* verts9, int n, float lo[3], float hi[3]) {
    for (int a = 0; a < 3; a++) { lo[a] = 1.0e23f; hi[a] = -1.0e23f; }           // environment.go:391-392
    for (int64_t i = 0; i < 3 * (int64_t)n; i++)
        for (int a = 0; a < 3; a++) { const float v = verts9[3 * i + a]; if (v < lo[a]) lo[a] = v; if (v > hi[a]) hi[a] = v; }
    if (n == 0) for (int a = 0; a < 3; a++) lo[a] = hi[a] = 0.0f;
}

}  // namespace kdfast

// host cores
int build_kd_tree_binned_host(const float* verts9, int n, KdTree& out, const char** why) {
    float lo[3], hi[3];
    kdfast::scene_bounds(verts9, n, lo, hi);
    kdfast::HostExec ex;
    return kdfast::build(ex, verts9, n, lo, hi, out, why);
}

// the device; verts9 is host memory (the environment's triangle list), uploaded once
int build_kd_tree_binned_device(cudaStream_t stream, const float* verts9, int n, KdTree& out, int* launches, const char** why) {
    float lo[3], hi[3];
    kdfast::scene_bounds(verts9, n, lo, hi);
    kdfast::DeviceExec ex{stream};
    float* d_verts = (float*)ex.alloc(36 * (size_t)std::max(n, 1));
    ex.upload(d_verts, verts9, 36 * (size_t)n);
    int rc = ex.err == cudaSuccess ? kdfast::build(ex, d_verts, n, lo, hi, out, why) : VRAD_E_CUDA;
    if (ex.err != cudaSuccess && rc == VRAD_OK) { rc = VRAD_E_CUDA; *why = cudaGetErrorString(ex.err); }
    ex.free(d_verts);
    if (launches) *launches = ex.launches;
    return rc;
}

}  // namespace vrad
