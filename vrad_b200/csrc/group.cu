// group.cu -- the in-process multi-GPU handle (vrad_env_create_multi).
//
// The reference is one goroutine behind a package-level singleton (raytracer/environment.go:17-25;
// common/constants/constants.go:43): its driver cannot start one process per GPU.  SURVEY 8(b) therefore puts the
// devices INSIDE the handle: the caller holds one vrad_env*, the handle owns one child environment per device
// (rank r of world n) and runs every call on one worker thread per child.  Rays, luxels and index pairs are split
// into contiguous ranges (no collective); patches and the scene are replicated; transfer rows are sharded and the
// per-bounce radiance exchange is the fused peer-store kernel of k4_bounce.cu, with the children's collectives going
// through LocalGroup (comm.cu) instead of NCCL and peer buffers mapped by cudaDeviceEnablePeerAccess instead of IPC.
// Entry points that are not listed here return VRAD_E_UNSUPPORTED on such a handle.
#include "env_internal.cuh"
#include <functional>
#include <thread>

namespace vrad {


// fn(child, rank) on one thread per child; the first failure is reported with its rank
static int group_run(vrad_env* g, const std::function<int(vrad_env*, int)>& fn) {
    LocalGroup& G = *g->multi;
    const int world = G.world;
    std::vector<int> rc(world, 0);
    std::vector<std::string> msg(world);
    std::vector<std::thread> th;
    th.reserve(world);
    for (int r = 0; r < world; r++)
        th.emplace_back([&, r]() {
            cudaSetDevice(G.ranks[r]->cfg.device);
            rc[r] = fn(G.ranks[r], r);
            if (rc[r]) { msg[r] = vrad_last_error(); G.fail(); }
        });
    for (auto& t : th) t.join();
    G.reset();
    for (int r = 0; r < world; r++)
        if (rc[r]) {
            // a rank that only saw a peer fail says so; prefer the message of the rank that failed first-hand
            int best = r;
            for (int q = 0; q < world; q++) if (rc[q] && msg[q].find("in-process group failed") == std::string::npos) { best = q; break; }
            set_error("device %d (rank %d of %d): %s", G.ranks[best]->cfg.device, best, world, msg[best].c_str());
            return rc[best];
        }
    return VRAD_OK;
}

static int group_all(vrad_env* g, const std::function<int(vrad_env*)>& fn) {
    return group_run(g, [&](vrad_env* c, int) -> int { return fn(c); });
}

// contiguous range of rank r of n items, boundaries on multiples of `align`
static void range_of(int64_t n, int world, int r, int64_t align, int64_t& a, int64_t& b) {
    auto cut = [&](int q) { const int64_t x = n / world * q + std::min<int64_t>(q, n % world); return q == world ? n : x / align * align; };
    a = cut(r); b = cut(r + 1);
}

static int reject_device_ptrs(const char* what, std::initializer_list<const void*> ptrs) {
    for (const void* p : ptrs)
        if (p && is_device_ptr(p)) { set_error("%s: a multi-GPU handle takes host buffers (a device buffer belongs to one device)", what); return VRAD_E_INVALID; }
    return 0;
}

int group_destroy(vrad_env* g) {
    LocalGroup* G = g->multi;
    for (vrad_env* c : G->ranks) if (c) { c->group = nullptr; vrad_env_destroy(c); }
    delete G;
    g->multi = nullptr;
    return 0;
}

int group_set_option(vrad_env* g, const char* name, int value) { return group_all(g, [&](vrad_env* c) -> int { return vrad_env_set_option(c, name, value); }); }

int group_add_triangles(vrad_env* g, int n, const int32_t* ids, const float* verts9, const uint8_t* flags) {
    for (vrad_env* c : g->multi->ranks) { int rc = vrad_env_add_triangles(c, n, ids, verts9, flags); if (rc) return rc; }
    return VRAD_OK;
}

// the tree is built once (rank 0, on the host) and adopted by the other ranks in the reference's own layouts
static int group_share_tree(vrad_env* g) {
    vrad_env* c0 = g->multi->ranks[0];
    const KdTree& T = c0->tree;
    float aabb[6];
    for (int c = 0; c < 3; c++) { aabb[c] = T.bmin[c]; aabb[3 + c] = T.bmax[c]; }
    return group_run(g, [&](vrad_env* c, int r) -> int {
        if (r == 0) return (int)VRAD_OK;
        int rc = vrad_env_upload_tree(c, (int)T.children.size(), T.children.data(), T.split.data(), (int)T.tri_index.size(), T.tri_index.data(),
                                      (int)c0->h_tris.size(), c0->h_tris.data(), aabb);
        if (rc == VRAD_OK && !c0->h_colors.empty()) rc = vrad_env_set_triangle_colors(c, (int)(c0->h_colors.size() / 3), c0->h_colors.data());
        return rc;
    });
}

int group_build(vrad_env* g, int fast, int where) {
    vrad_env* c0 = g->multi->ranks[0];
    cudaSetDevice(c0->cfg.device);
    int rc = fast ? vrad_env_build_fast(c0, where) : vrad_env_build(c0);
    if (rc) return rc;
    return group_share_tree(g);
}

int group_upload_tree(vrad_env* g, int n_nodes, const int32_t* children, const float* split, int n_idx, const int32_t* tri_index, int n_tris,
                      const vrad_tri48* tris, const float aabb[6]) {
    return group_all(g, [&](vrad_env* c) -> int { return vrad_env_upload_tree(c, n_nodes, children, split, n_idx, tri_index, n_tris, tris, aabb); });
}

int group_set_triangle_colors(vrad_env* g, int n, const float* rgb3) { return group_all(g, [&](vrad_env* c) -> int { return vrad_env_set_triangle_colors(c, n, rgb3); }); }
int group_points_upload(vrad_env* g, int64_t n, const float* xyz3) {
    int rc = reject_device_ptrs("vrad_points_upload", {xyz3});
    return rc ? rc : group_all(g, [&](vrad_env* c) -> int { return vrad_points_upload(c, n, xyz3); });
}
int group_set_sky_dirs(vrad_env* g, int n, const float* dirs3) { return group_all(g, [&](vrad_env* c) -> int { return vrad_set_sky_dirs(c, n, dirs3); }); }
int group_set_light_trace_flags(vrad_env* g, int flags) { return group_all(g, [&](vrad_env* c) -> int { return vrad_set_light_trace_flags(c, flags); }); }

int group_last_timing(vrad_env* g, float* ms, int* launches) {
    float best = 0.f; int nl = 0;
    for (vrad_env* c : g->multi->ranks) {
        float m = 0.f; int l = 0;
        cudaSetDevice(c->cfg.device);
        int rc = vrad_env_last_timing(c, &m, &l);
        if (rc) return rc;
        best = std::max(best, m); nl += l;
    }
    if (ms) *ms = best;
    if (launches) *launches = nl;
    return VRAD_OK;
}

// ---- K1: each rank takes a contiguous range of the batch (boundaries on 32 segments: whole words of the bit vector) ----
int group_test_lines(vrad_env* g, int64_t n, const float* a, const float* b, int sky_mode, uint32_t* bits) {
    int rc = reject_device_ptrs("vrad_test_lines", {a, b, bits});
    if (rc) return rc;
    const int world = g->multi->world;
    return group_run(g, [&](vrad_env* c, int r) -> int {
        int64_t s0, s1;
        range_of(n, world, r, 32, s0, s1);
        const int64_t m = s1 - s0;
        if (m <= 0) return (int)VRAD_OK;
        if (!c->built) { set_error("vrad_test_lines: acceleration structure not built"); return (int)VRAD_E_STATE; }
        void* d_o;
        const size_t wb = (size_t)((m + 31) / 32) * 4;
        int rcc = scratch_get(c, 2, wb, &d_o);
        if (rcc) return rcc;
        // x[n] y[n] z[n] blocks of the whole batch: this rank reads [s0, s1) of each (host stride n)
        if ((rcc = launch_test_lines_pipelined(c, m, a + s0, b + s0, n, nullptr, sky_mode, (uint32_t*)d_o))) return rcc;
        VRAD_CUDA_CHECK(cudaMemcpyAsync(bits + (s0 >> 5), d_o, wb, cudaMemcpyDeviceToHost, c->stream));
        int bad = 0;
        if ((rcc = read_bad_index_count(c, &bad))) return rcc;        // synchronises
        if (bad) { set_error("vrad_test_lines: the segment copy did not arrive on the device"); return (int)VRAD_E_CUDA; }
        return (int)VRAD_OK;
    });
}

int group_test_lines_indexed(vrad_env* g, int64_t n, const int32_t* pairs2, int sky_mode, uint32_t* bits) {
    int rc = reject_device_ptrs("vrad_test_lines_indexed", {pairs2, bits});
    if (rc) return rc;
    const int world = g->multi->world;
    return group_run(g, [&](vrad_env* c, int r) -> int {
        int64_t s0, s1;
        range_of(n, world, r, 32, s0, s1);
        return s1 > s0 ? vrad_test_lines_indexed(c, s1 - s0, pairs2 + 2 * s0, sky_mode, bits + (s0 >> 5)) : (int)VRAD_OK;
    });
}

int group_trace_rays(vrad_env* g, int64_t n, const float* ox, const float* oy, const float* oz, const float* dx, const float* dy, const float* dz,
                     const float* tmin, const float* tmax, int32_t skip_id, int32_t* hit_tri, int32_t* hit_sid, float* hit_t) {
    int rc = reject_device_ptrs("vrad_trace_rays", {ox, oy, oz, dx, dy, dz, tmin, tmax, hit_tri, hit_sid, hit_t});
    if (rc) return rc;
    const int world = g->multi->world;
    return group_run(g, [&](vrad_env* c, int r) -> int {
        int64_t s0, s1;
        range_of(n, world, r, 1, s0, s1);
        if (s1 <= s0) return (int)VRAD_OK;
        return vrad_trace_rays(c, s1 - s0, ox + s0, oy + s0, oz + s0, dx + s0, dy + s0, dz + s0, tmin ? tmin + s0 : nullptr, tmax + s0, skip_id,
                               hit_tri ? hit_tri + s0 : nullptr, hit_sid ? hit_sid + s0 : nullptr, hit_t ? hit_t + s0 : nullptr);
    });
}

// ---- patches, K2, K3, K4 ----
int group_patches_upload(vrad_env* g, int n, const float* origin3, const float* normal3, const float* plane_dist, const float* area,
                         const float* reflectivity3, const int32_t* cluster, const uint8_t* flags) {
    return group_all(g, [&](vrad_env* c) -> int { return vrad_patches_upload(c, n, origin3, normal3, plane_dist, area, reflectivity3, cluster, flags); });
}
int group_set_hierarchy(vrad_env* g, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face) {
    return group_all(g, [&](vrad_env* c) -> int { return vrad_patches_set_hierarchy(c, n, parent, child1, child2, face); });
}

int group_set_bump(vrad_env* g, int n, const uint8_t* needs_bump, const float* bump_normals9) {
    return group_all(g, [&](vrad_env* c) -> int { return vrad_patches_set_bump(c, n, needs_bump, bump_normals9); });
}

int group_set_windings(vrad_env* g, int n, const int32_t* first, const int32_t* count, int n_points, const float* points3) {
    return group_all(g, [&](vrad_env* c) -> int { return vrad_patches_set_windings(c, n, first, count, n_points, points3); });
}

int group_build_transfers(vrad_env* g, int n_clusters, const uint8_t* pvs, int64_t* nnz_out) {
    std::vector<int64_t> nnz(g->multi->world, 0);
    int rc = group_run(g, [&](vrad_env* c, int r) -> int { return vrad_build_transfers(c, n_clusters, pvs, &nnz[r]); });
    if (rc) return rc;
    int64_t sum = 0;
    for (int64_t v : nnz) sum += v;
    if (nnz_out) *nnz_out = sum;
    return VRAD_OK;
}

int group_transfers_info(vrad_env* g, int64_t* row0, int64_t* row1, int64_t* nnz) {
    int64_t sum = 0, last = 0;
    for (vrad_env* c : g->multi->ranks) {
        int64_t a, b, z;
        int rc = vrad_transfers_info(c, &a, &b, &z);
        if (rc) return rc;
        sum += z; last = b;
    }
    if (row0) *row0 = 0;
    if (row1) *row1 = last;
    if (nnz) *nnz = sum;
    return VRAD_OK;
}

// the rows of all ranks, in rank (= row) order
int group_transfers_download(vrad_env* g, int64_t* rowptr, int32_t* col, float* w) {
    int64_t pos = 0, row = 0;
    for (vrad_env* c : g->multi->ranks) {
        int64_t a, b, z;
        int rc = vrad_transfers_info(c, &a, &b, &z);
        if (rc) return rc;
        cudaSetDevice(c->cfg.device);
        std::vector<int64_t> rp(b - a + 1);
        if ((rc = vrad_transfers_download(c, rp.data(), col ? col + pos : nullptr, w ? w + pos : nullptr))) return rc;
        if (rowptr) for (int64_t i = 0; i < b - a; i++) rowptr[row + i] = pos + rp[i];
        row += b - a; pos += z;
    }
    if (rowptr) rowptr[row] = pos;
    return VRAD_OK;
}

int group_direct_light(vrad_env* g, int64_t n, const float* pos3, const float* normal3, int n_lights, const vrad_light* lights, float* rgb_out) {
    int rc = reject_device_ptrs("vrad_direct_light", {pos3, normal3, rgb_out});
    if (rc) return rc;
    const int world = g->multi->world;
    return group_run(g, [&](vrad_env* c, int r) -> int {
        int64_t s0, s1;
        range_of(n, world, r, 1, s0, s1);
        return s1 > s0 ? vrad_direct_light(c, s1 - s0, pos3 + 3 * s0, normal3 + 3 * s0, n_lights, lights, rgb_out + 3 * s0) : (int)VRAD_OK;
    });
}

int group_bounce(vrad_env* g, const float* emit0_rgb, int n_bounces, int early_out, float* total_rgb_out, float added_last[3], int* bounces_done) {
    int rc = reject_device_ptrs("vrad_bounce", {emit0_rgb, total_rgb_out});
    if (rc) return rc;
    // every rank ends with the complete result; rank 0's copy goes to the caller
    return group_run(g, [&](vrad_env* c, int r) -> int {
        return vrad_bounce(c, emit0_rgb, n_bounces, early_out, r == 0 ? total_rgb_out : nullptr, r == 0 ? added_last : nullptr, r == 0 ? bounces_done : nullptr);
    });
}

} // namespace vrad
using namespace vrad;

extern "C" {

int vrad_env_create_multi(const vrad_multi_config* cfg, vrad_env** out) {
    if (!out) { set_error("vrad_env_create_multi: out is NULL"); return VRAD_E_INVALID; }
    *out = nullptr;
    if (!cfg || cfg->n_devices < 1 || cfg->n_devices > kMaxWorld) { set_error("vrad_env_create_multi: n_devices must be 1..%d", kMaxWorld); return VRAD_E_INVALID; }
    vrad_env* g = new (std::nothrow) vrad_env();
    LocalGroup* G = new (std::nothrow) LocalGroup();
    if (!g || !G) { delete g; delete G; return VRAD_E_NOMEM; }
    G->world = cfg->n_devices;
    g->multi = G;
    g->cfg = vrad_config{cfg->devices[0], 0, cfg->n_devices, cfg->flags};
    for (int r = 0; r < cfg->n_devices; r++) {
        for (int q = 0; q < r; q++) if (cfg->devices[q] == cfg->devices[r]) G->shares_device = true;
        vrad_config c{cfg->devices[r], r, cfg->n_devices, cfg->flags};
        vrad_env* child = nullptr;
        int rc = vrad_env_create(&c, &child);
        if (rc) { const std::string why = vrad_last_error(); group_destroy(g); delete g; set_error("vrad_env_create_multi: device %d: %s", cfg->devices[r], why.c_str()); return rc; }
        child->group = G;
        G->ranks.push_back(child);
    }
    *out = g;
    return VRAD_OK;
}

} // extern "C"
