// vrad_env.cu -- handle lifetime, geometry/tree management and the K1 entry points of the C-ABI.
// Reference map per function: see include/vrad_cuda.h.
#include "env_internal.cuh"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdarg>
#include <new>

namespace vrad {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int scratch_get(vrad_env* e, int slot, size_t bytes, void** out) {
    if ((size_t)slot >= e->scratch.size()) e->scratch.resize(slot + 1);
    if (e->scratch[slot].alloc(bytes ? bytes : 1) != 0) { set_error("out of device memory (%zu bytes)", bytes); return VRAD_E_NOMEM; }
    *out = e->scratch[slot].p;
    return 0;
}

int stage_in(vrad_env* e, int slot, const void* p, size_t bytes, const void** dev_out, bool* was_host) {
    if (!p) { *dev_out = nullptr; if (was_host) *was_host = false; return 0; }
    if (is_device_ptr(p)) { *dev_out = p; if (was_host) *was_host = false; return 0; }
    void* d = nullptr;
    int rc = scratch_get(e, slot, bytes, &d);
    if (rc) return rc;
    VRAD_CUDA_CHECK(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, e->stream));
    *dev_out = d;
    if (was_host) *was_host = true;
    return 0;
}

int stage_out(vrad_env* e, int slot, void* p, size_t bytes, void** dev_out, bool* was_host) {
    if (!p) { *dev_out = nullptr; *was_host = false; return 0; }
    if (is_device_ptr(p)) { *dev_out = p; *was_host = false; return 0; }
    int rc = scratch_get(e, slot, bytes, dev_out);
    if (rc) return rc;
    *was_host = true;
    return 0;
}

int finish_out(vrad_env* e, void* host_p, const void* dev_p, size_t bytes, bool was_host) {
    if (!was_host || !host_p) return 0;
    VRAD_CUDA_CHECK(cudaMemcpyAsync(host_p, dev_p, bytes, cudaMemcpyDeviceToHost, e->stream));
    return 0;
}

void timing_begin(vrad_env* e) { cudaEventRecord(e->ev0, e->stream); }
void timing_end(vrad_env* e, int launches) { cudaEventRecord(e->ev1, e->stream); e->last_launches = launches; e->last_ms = -1.0f; }

int sync_if_needed(vrad_env* e, bool any_host) {
    if (any_host || !e->async) {
        VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
        VRAD_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

int upload_top_levels(vrad_env* e);

static int upload_scene(vrad_env* e) {
    const KdTree& T = e->tree;
    const size_t nn = T.children.size(), ni = T.tri_index.size(), nt = e->h_tris.size();
    if (e->d_nodes.alloc(nn) || e->d_tri_index.alloc(ni ? ni : 1) || e->d_q0.alloc(nt ? nt : 1) ||
        e->d_q1.alloc(nt ? nt : 1) || e->d_q2.alloc(nt ? nt : 1)) {
        set_error("out of device memory for scene"); return VRAD_E_NOMEM;
    }
    std::vector<int2> nodes(nn);
    for (size_t i = 0; i < nn; i++) { nodes[i].x = T.children[i]; memcpy(&nodes[i].y, &T.split[i], 4); }
    std::vector<float4> q0(nt), q1(nt), q2(nt);
    for (size_t i = 0; i < nt; i++) {
        const vrad_tri48& t = e->h_tris[i];
        q0[i] = make_float4(t.nx, t.ny, t.nz, t.d);
        q1[i] = make_float4(t.e[0], t.e[1], t.e[2], t.e[3]);
        int32_t sel = (int32_t)t.sel0 | ((int32_t)t.sel1 << 8) | ((int32_t)t.flags << 16);
        float idf, self;
        memcpy(&idf, &t.id, 4); memcpy(&self, &sel, 4);
        q2[i] = make_float4(t.e[4], t.e[5], idf, self);
    }
    VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_nodes.p, nodes.data(), nn * sizeof(int2), cudaMemcpyHostToDevice, e->stream));
    if (ni) VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_tri_index.p, T.tri_index.data(), ni * 4, cudaMemcpyHostToDevice, e->stream));
    if (nt) {
        VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_q0.p, q0.data(), nt * 16, cudaMemcpyHostToDevice, e->stream));
        VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_q1.p, q1.data(), nt * 16, cudaMemcpyHostToDevice, e->stream));
        VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_q2.p, q2.data(), nt * 16, cudaMemcpyHostToDevice, e->stream));
    }
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    DevScene& S = e->scene;
    S.nodes = e->d_nodes.p; S.tri_index = e->d_tri_index.p; S.q0 = e->d_q0.p; S.q1 = e->d_q1.p; S.q2 = e->d_q2.p;
    for (int c = 0; c < 3; c++) { S.bmin[c] = T.bmin[c]; S.bmax[c] = T.bmax[c]; }
    S.n_nodes = (int)nn; S.n_idx = (int)ni; S.n_tris = (int)nt;
    S.tri_cov = nullptr;
    e->built = true;
    int rct = upload_top_levels(e);
    if (rct) return rct;
    return upload_triangle_coverage(e);
}

// colour.X per triangle -> device (coverage of transparent triangles, coverageCount.go:29)
int upload_triangle_coverage(vrad_env* e) {
    if (!e->built || e->h_colors.empty()) return 0;
    const size_t nt = e->h_tris.size();
    if (e->h_colors.size() != 3 * nt) { set_error("triangle colours: %zu given for %zu triangles", e->h_colors.size() / 3, nt); return VRAD_E_INVALID; }
    if (e->d_tri_cov.alloc(nt ? nt : 1)) { set_error("out of device memory for triangle colours"); return VRAD_E_NOMEM; }
    std::vector<float> cov(nt);
    for (size_t i = 0; i < nt; i++) cov[i] = e->h_colors[3 * i];
    VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_tri_cov.p, cov.data(), nt * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    e->scene.tri_cov = e->d_tri_cov.p;
    return 0;
}

// The top of the tree in breadth-first order for shared-memory staging (DevScene::top): nodes are taken level by level
// while both children of a node still fit the budget (EnvOptions::k1_top nodes, at most 2047 = 16 KB).  A staged node whose
// children are staged too points at their slots, flagged with kTopRef; every other word is the node's own.
int upload_top_levels(vrad_env* e) {
    DevScene& S = e->scene;
    S.top = nullptr; S.n_top = 0;
    const int budget = std::min(e->opt.k1_top, 2047);
    const KdTree& T = e->tree;
    if (budget < 3 || T.children.empty()) return 0;
    std::vector<int> order(1, 0), slot_of;            // order[slot] = tree node
    std::vector<int2> top;
    std::vector<int> first_child_slot(1, -1);
    for (size_t head = 0; head < order.size(); head++) {
        const int n = order[head], word = T.children[n];
        if ((word & 3) == 3) continue;                                    // leaf: nothing to expand
        if ((int)order.size() + 2 > budget) continue;                     // children stay in global memory
        first_child_slot[head] = (int)order.size();
        order.push_back(word >> 2); order.push_back((word >> 2) + 1);
        first_child_slot.push_back(-1); first_child_slot.push_back(-1);
    }
    top.resize(order.size());
    for (size_t k = 0; k < order.size(); k++) {
        const int n = order[k];
        int word = T.children[n];
        if (first_child_slot[k] >= 0) word = ((first_child_slot[k] | kTopRef) << 2) | (word & 3);
        top[k].x = word; memcpy(&top[k].y, &T.split[n], 4);
    }
    if (e->d_top.alloc(top.size())) { set_error("out of device memory for the staged tree levels"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_top.p, top.data(), top.size() * sizeof(int2), cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    S.top = e->d_top.p; S.n_top = (int)top.size();
    return 0;
}

} // namespace vrad

using namespace vrad;

extern "C" void vrad_comm_destroy_internal(vrad_env*);   // comm.cu
namespace vrad { int build_gather_plan(vrad_env* e, const int32_t* rowlen, int64_t nloc); int upload_top_levels(vrad_env* e); }

extern "C" {

const char* vrad_last_error(void) { return g_err; }
const char* vrad_version(void) { return "vrad-b200 0.1 (sm_100a)"; }

int vrad_env_create(const vrad_config* cfg, vrad_env** out) {
    if (!out) { set_error("vrad_env_create: out is NULL"); return VRAD_E_INVALID; }
    *out = nullptr;
    vrad_config c{0, 0, 1, 0};
    if (cfg) c = *cfg;
    if (c.world < 1 || c.rank < 0 || c.rank >= c.world) { set_error("vrad_env_create: bad rank/world %d/%d", c.rank, c.world); return VRAD_E_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("vrad_env_create: no CUDA device (this library has no CPU fallback)");
        return VRAD_E_CUDA;
    }
    if (c.device < 0 || c.device >= ndev) { set_error("vrad_env_create: device %d out of range (%d devices)", c.device, ndev); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(c.device));
    vrad_env* e = new (std::nothrow) vrad_env();
    if (!e) return VRAD_E_NOMEM;
    // any failure below destroys what was created so far (streams, events, the handle) before the status goes back
    struct Guard { vrad_env* e; ~Guard() { if (e) vrad_env_destroy(e); } } guard{e};
    e->cfg = c;
    auto env_int = [](const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; };
    e->opt.k1_sort = env_int("VRAD_K1_SORT", e->opt.k1_sort);
    e->opt.k1_top = env_int("VRAD_K1_TOP", e->opt.k1_top);
    e->opt.k1_key = env_int("VRAD_K1_KEY", e->opt.k1_key);
    e->opt.k1_stream = env_int("VRAD_K1_STREAM", e->opt.k1_stream);
    e->opt.k1_bpsm = env_int("VRAD_K1_BPSM", e->opt.k1_bpsm);
    e->opt.k1_sort_bits = env_int("VRAD_K1_SORT_BITS", e->opt.k1_sort_bits);
    e->opt.k4_seg = env_int("VRAD_K4_SEG", e->opt.k4_seg);
    { const char* v = getenv("VRAD_K4_ORDER"); e->opt.k4_long_first = v && std::string(v) == "long"; }
    e->opt.k4_persist = env_int("VRAD_K4_PERSIST", e->opt.k4_persist);
    e->opt.k4_block = env_int("VRAD_K4_BLOCK", e->opt.k4_block);
    e->opt.k4_pool = env_int("VRAD_K4_POOL", e->opt.k4_pool);
    e->opt.k4_hier_p2p = env_int("VRAD_K4_HIER_P2P", e->opt.k4_hier_p2p);
    e->opt.k4_l2_mb = env_int("VRAD_K4_L2_MB", e->opt.k4_l2_mb);
    e->opt.k2_stream = env_int("VRAD_K2_STREAM", e->opt.k2_stream);
    e->opt.k4_pack = env_int("VRAD_K4_PACK", e->opt.k4_pack);
    e->opt.k4_short = env_int("VRAD_K4_SHORT", e->opt.k4_short);
    e->opt.k4_bk_rows = env_int("VRAD_K4_BK_ROWS", e->opt.k4_bk_rows);
    e->opt.k4_bk_min_items = env_int("VRAD_K4_BK_MIN_ITEMS", e->opt.k4_bk_min_items);
    e->opt.k4_items = env_int("VRAD_K4_ITEMS", e->opt.k4_items);
    e->opt.k4_pdl = env_int("VRAD_K4_PDL", e->opt.k4_pdl);
    e->opt.k4_graph = env_int("VRAD_K4_GRAPH", e->opt.k4_graph);
    e->opt.k4_sim_peers = env_int("VRAD_K4_SIM_PEERS", e->opt.k4_sim_peers);
    cudaDeviceProp prop;
    VRAD_CUDA_CHECK(cudaGetDeviceProperties(&prop, c.device));
    e->sm_count = prop.multiProcessorCount;
    VRAD_CUDA_CHECK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    VRAD_CUDA_CHECK(cudaEventCreate(&e->ev0));
    VRAD_CUDA_CHECK(cudaEventCreate(&e->ev1));
    VRAD_CUDA_CHECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int s = 0; s < 2; s++) {
        VRAD_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_copied[s], cudaEventDisableTiming));
        VRAD_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_done[s], cudaEventDisableTiming));
    }
    VRAD_CUDA_CHECK(cudaMallocHost((void**)&e->h_one, 4));
    *e->h_one = 1u;
    guard.e = nullptr;
    *out = e;
    return VRAD_OK;
}

void vrad_env_destroy(vrad_env* e) {
    if (e && e->multi) { vrad::group_destroy(e); delete e; return; }
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    vrad_comm_destroy_internal(e);
    e->d_nodes.release(); e->d_tri_index.release(); e->d_q0.release(); e->d_q1.release(); e->d_q2.release();
    e->d_tri_cov.release(); e->d_top.release(); e->d_points.release(); e->d_bsp_nodes.release(); e->d_bsp_planes.release(); e->d_cams.release();
    e->d_leaf_cluster.release(); e->d_leaf_area.release(); e->d_area_camera.release();
    for (auto& s : e->scratch) s.release();
    e->patches.origin_area.release(); e->patches.normal_dist.release(); e->patches.refl.release(); e->patches.cluster.release();
    e->patches.tree.release(); e->patches.collect_ids.release(); e->patches.collect_ptr.release(); e->patches.collect_ent.release(); e->patches.leaf_rows.release(); e->patches.child2.release();
    e->patches.bump_normals.release(); e->patches.bump_rows.release(); for (int b = 0; b < 3; b++) e->patches.total_bump[b].release();
    e->transfers.rowptr.release(); e->transfers.rowlen.release(); e->transfers.tr.release();
    e->d_sky_dirs.release(); e->d_er[0].release(); e->d_er[1].release(); e->d_total.release(); e->d_partials.release(); e->d_add.release();
    if (e->bounce_graph.exec) cudaGraphExecDestroy(e->bounce_graph.exec);
    e->peers.d_sink.release();
    e->transfers.items.release(); e->transfers.item_slot.release(); e->transfers.block_ptr.release(); e->transfers.part_sum.release(); e->transfers.row_ctr.release();
    for (int s = 0; s < 2; s++) {
        e->d_stage[s].release();
        if (e->ev_copied[s]) cudaEventDestroy(e->ev_copied[s]);
        if (e->ev_done[s]) cudaEventDestroy(e->ev_done[s]);
    }
    if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
    if (e->h_one) cudaFreeHost(e->h_one);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

int vrad_env_set_stream(vrad_env* e, void* cuda_stream) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_env_set_stream");
    if (!e) return VRAD_E_INVALID;
    cudaStreamSynchronize(e->stream);
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return VRAD_OK;
}

int vrad_env_set_option(vrad_env* e, const char* name, int value) {
    VRAD_MULTI(e, group_set_option(e, name, value));
    if (!e || !name) { set_error("vrad_env_set_option: bad arguments"); return VRAD_E_INVALID; }
    const std::string n(name);
    EnvOptions& o = e->opt;
    if (n == "k1_sort") o.k1_sort = value;
    else if (n == "k1_key") o.k1_key = value;
    else if (n == "k1_stream") o.k1_stream = value;
    else if (n == "k1_bpsm") o.k1_bpsm = value;
    else if (n == "k1_sort_bits") o.k1_sort_bits = value;
    else if (n == "k1_top") { o.k1_top = value; if (e->built) { VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device)); VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream)); return upload_top_levels(e); } }
    else if (n == "k4_items") o.k4_items = value;
    else if (n == "k4_hier_p2p") o.k4_hier_p2p = value;
    else if (n == "k4_l2_mb") o.k4_l2_mb = value;
    else if (n == "k2_stream") o.k2_stream = value;
    else if (n == "k4_bk_rows" || n == "k4_pack" || n == "k4_seg" || n == "k4_long_first" || n == "k4_persist" || n == "k4_block" || n == "k4_pool") {
        (n == "k4_bk_rows" ? o.k4_bk_rows : n == "k4_pack" ? o.k4_pack : n == "k4_seg" ? o.k4_seg : (n == "k4_persist" ? o.k4_persist : (n == "k4_block" ? o.k4_block : (n == "k4_pool" ? o.k4_pool : o.k4_long_first)))) = value;
        if (e->transfers.ready) {           // re-plan the resident rows
            VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
            const int64_t nloc = e->transfers.row1 - e->transfers.row0;
            std::vector<int32_t> rl(nloc ? nloc : 1);
            if (nloc) VRAD_CUDA_CHECK(cudaMemcpy(rl.data(), e->transfers.rowlen.p, (size_t)nloc * 4, cudaMemcpyDeviceToHost));
            return build_gather_plan(e, rl.data(), nloc);
        }
    }
    else if (n == "k4_short") o.k4_short = value;
    else if (n == "k4_pdl") o.k4_pdl = value;
    else if (n == "k4_graph") o.k4_graph = value;
    else if (n == "k4_sim_peers") o.k4_sim_peers = value;
    else { set_error("vrad_env_set_option: unknown option '%s'", name); return VRAD_E_INVALID; }
    return VRAD_OK;
}

int vrad_env_set_async(vrad_env* e, int async) {
    if (e && e->multi) return VRAD_OK;      // calls on a multi-GPU handle take host buffers and are synchronous
    if (!e) return VRAD_E_INVALID;
    e->async = async != 0;
    return VRAD_OK;
}

int vrad_env_last_timing(vrad_env* e, float* kernel_ms, int* n_launches) {
    VRAD_MULTI(e, group_last_timing(e, kernel_ms, n_launches));
    if (!e) return VRAD_E_INVALID;
    if (e->last_ms < 0.0f) {
        VRAD_CUDA_CHECK(cudaEventSynchronize(e->ev1));
        VRAD_CUDA_CHECK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    }
    if (kernel_ms) *kernel_ms = e->last_ms;
    if (n_launches) *n_launches = e->last_launches;
    return VRAD_OK;
}

void* vrad_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); set_error("vrad_host_alloc(%zu) failed", bytes); return nullptr; }
    return p;
}
void vrad_host_free(void* p) { if (p) cudaFreeHost(p); }

int vrad_env_add_triangles(vrad_env* e, int n, const int32_t* ids, const float* verts9, const uint8_t* flags) {
    VRAD_MULTI(e, group_add_triangles(e, n, ids, verts9, flags));
    if (!e || n < 0 || (n > 0 && (!ids || !verts9))) { set_error("vrad_env_add_triangles: bad arguments"); return VRAD_E_INVALID; }
    if (e->built) { set_error("vrad_env_add_triangles: acceleration structure already built"); return VRAD_E_STATE; }
    e->h_ids.insert(e->h_ids.end(), ids, ids + n);
    e->h_verts.insert(e->h_verts.end(), verts9, verts9 + 9 * (size_t)n);
    if (flags) e->h_flags.insert(e->h_flags.end(), flags, flags + n); else e->h_flags.insert(e->h_flags.end(), n, 0);
    return VRAD_OK;
}

int vrad_env_set_triangle_colors(vrad_env* e, int n, const float* rgb3) {
    VRAD_MULTI(e, group_set_triangle_colors(e, n, rgb3));
    if (!e || n < 0 || (n > 0 && !rgb3)) { set_error("vrad_env_set_triangle_colors: bad arguments"); return VRAD_E_INVALID; }
    const size_t have = e->built ? e->h_tris.size() : e->h_ids.size();
    if ((size_t)n != have) { set_error("vrad_env_set_triangle_colors: %d colours for %zu triangles", n, have); return VRAD_E_INVALID; }
    e->h_colors.assign(rgb3, rgb3 + 3 * (size_t)n);
    if (!e->built) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    return upload_triangle_coverage(e);
}

int vrad_env_build(vrad_env* e) {
    VRAD_MULTI(e, group_build(e, 0, 0));
    if (!e) return VRAD_E_INVALID;
    if (e->built) { set_error("vrad_env_build: already built"); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    auto t0 = std::chrono::steady_clock::now();
    const int n = (int)e->h_ids.size();
    build_kd_tree(e->h_verts.data(), n, e->tree);
    e->h_tris.resize(n);
    make_intersection_records(e->h_ids.data(), e->h_verts.data(), e->h_flags.data(), n, e->h_tris.data());
    e->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (e->tree.max_depth >= kStackMax) { set_error("vrad_env_build: tree depth %d exceeds %d", e->tree.max_depth, kStackMax - 1); return VRAD_E_UNSUPPORTED; }
    return upload_scene(e);
}

int vrad_env_build_fast(vrad_env* e, int where) {
    VRAD_MULTI(e, group_build(e, 1, where));
    if (!e || (where != VRAD_BUILD_ON_DEVICE && where != VRAD_BUILD_ON_HOST && where != VRAD_BUILD_AUTO)) { set_error("vrad_env_build_fast: bad arguments"); return VRAD_E_INVALID; }
    if (e->built) { set_error("vrad_env_build_fast: already built"); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    auto t0 = std::chrono::steady_clock::now();
    const int n = (int)e->h_ids.size();
    const char* why = "";
    int launches = 0;
    // below ~10k triangles the level-by-level device build is all launch latency (r02: 996 triangles, 8.5 ms on the device against
    // 2.2 ms for the exact builder on the host): the same functors run on the host's cores instead -- the same tree by construction
    constexpr int kDeviceBuildMin = 10000;
    if (where == VRAD_BUILD_AUTO) where = n < kDeviceBuildMin ? VRAD_BUILD_ON_HOST : VRAD_BUILD_ON_DEVICE;
    int rc = where == VRAD_BUILD_ON_HOST ? build_kd_tree_binned_host(e->h_verts.data(), n, e->tree, &why)
                                         : build_kd_tree_binned_device(e->stream, e->h_verts.data(), n, e->tree, &launches, &why);
    if (rc) { set_error("vrad_env_build_fast: %s", why); return rc; }
    // the tree came from kernels: check it like a foreign tree before any traversal kernel walks it
    int leaves = 0;
    const int depth = validate_kd_tree((int)e->tree.children.size(), e->tree.children.data(), e->tree.split.data(), (int)e->tree.tri_index.size(),
                                       e->tree.tri_index.data(), n, &leaves, &why);
    if (depth < 0) { set_error("vrad_env_build_fast: built tree is not valid: %s", why); return VRAD_E_UNSUPPORTED; }
    e->tree.max_depth = depth; e->tree.n_leaves = leaves;
    e->h_tris.resize(n);
    make_intersection_records(e->h_ids.data(), e->h_verts.data(), e->h_flags.data(), n, e->h_tris.data());
    e->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    e->last_launches = launches;
    if (e->tree.max_depth >= kStackMax) { set_error("vrad_env_build_fast: tree depth %d exceeds %d", e->tree.max_depth, kStackMax - 1); return VRAD_E_UNSUPPORTED; }
    return upload_scene(e);
}

int vrad_kd_build_binned_host(int n, const float* verts9, int max_nodes, int max_idx, int32_t* children, float* split, int32_t* tri_index,
                              int* n_nodes, int* n_idx, float aabb[6], int* max_depth) {
    if (n < 0 || (n > 0 && !verts9) || !n_nodes || !n_idx) { set_error("vrad_kd_build_binned_host: bad arguments"); return VRAD_E_INVALID; }
    KdTree t;
    const char* why = "";
    int rc = build_kd_tree_binned_host(verts9, n, t, &why);
    if (rc) { set_error("vrad_kd_build_binned_host: %s", why); return rc; }
    int leaves = 0;
    const int depth = validate_kd_tree((int)t.children.size(), t.children.data(), t.split.data(), (int)t.tri_index.size(), t.tri_index.data(), n, &leaves, &why);
    if (depth < 0) { set_error("vrad_kd_build_binned_host: built tree is not valid: %s", why); return VRAD_E_UNSUPPORTED; }
    *n_nodes = (int)t.children.size(); *n_idx = (int)t.tri_index.size();
    if (max_depth) *max_depth = depth;
    if (aabb) for (int c = 0; c < 3; c++) { aabb[c] = t.bmin[c]; aabb[3 + c] = t.bmax[c]; }
    if (children && split && tri_index) {
        if (*n_nodes > max_nodes || *n_idx > max_idx) { set_error("vrad_kd_build_binned_host: %d nodes / %d indices, room for %d / %d", *n_nodes, *n_idx, max_nodes, max_idx); return VRAD_E_NOMEM; }
        memcpy(children, t.children.data(), 4 * t.children.size()); memcpy(split, t.split.data(), 4 * t.split.size());
        if (!t.tri_index.empty()) memcpy(tri_index, t.tri_index.data(), 4 * t.tri_index.size());
    }
    return VRAD_OK;
}

int vrad_env_upload_tree(vrad_env* e, int n_nodes, const int32_t* children, const float* split, int n_idx,
                         const int32_t* tri_index, int n_tris, const vrad_tri48* tris, const float aabb[6]) {
    VRAD_MULTI(e, group_upload_tree(e, n_nodes, children, split, n_idx, tri_index, n_tris, tris, aabb));
    if (!e || !children || !split || (n_idx > 0 && !tri_index) || (n_tris > 0 && !tris) || !aabb) { set_error("vrad_env_upload_tree: bad arguments"); return VRAD_E_INVALID; }
    if (e->built) { set_error("vrad_env_upload_tree: already built"); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const char* why = nullptr;
    int leaves = 0;
    int depth = validate_kd_tree(n_nodes, children, split, n_idx, tri_index, n_tris, &leaves, &why);
    if (depth < 0) { set_error("vrad_env_upload_tree: %s", why); return VRAD_E_INVALID; }
    if (depth >= kStackMax) { set_error("vrad_env_upload_tree: tree depth %d exceeds %d", depth, kStackMax - 1); return VRAD_E_UNSUPPORTED; }
    for (int i = 0; i < n_tris; i++)
        if (tris[i].sel0 > 2 || tris[i].sel1 > 2) { set_error("vrad_env_upload_tree: triangle %d has bad coordinate selectors", i); return VRAD_E_INVALID; }
    e->tree.children.assign(children, children + n_nodes);
    e->tree.split.assign(split, split + n_nodes);
    e->tree.tri_index.assign(tri_index, tri_index + n_idx);
    for (int c = 0; c < 3; c++) { e->tree.bmin[c] = aabb[c]; e->tree.bmax[c] = aabb[3 + c]; }
    e->tree.max_depth = depth; e->tree.n_leaves = leaves;
    e->h_tris.assign(tris, tris + n_tris);
    return upload_scene(e);
}

int vrad_env_stats(vrad_env* e, int* n_nodes, int* n_idx, int* n_tris, int* max_depth, int* n_leaves, float aabb[6], double* build_seconds) {
    VRAD_MULTI_RANK0(e);
    if (!e) return VRAD_E_INVALID;
    if (!e->built) { set_error("vrad_env_stats: not built"); return VRAD_E_STATE; }
    if (n_nodes) *n_nodes = (int)e->tree.children.size();
    if (n_idx) *n_idx = (int)e->tree.tri_index.size();
    if (n_tris) *n_tris = (int)e->h_tris.size();
    if (max_depth) *max_depth = e->tree.max_depth;
    if (n_leaves) *n_leaves = e->tree.n_leaves;
    if (aabb) for (int c = 0; c < 3; c++) { aabb[c] = e->tree.bmin[c]; aabb[3 + c] = e->tree.bmax[c]; }
    if (build_seconds) *build_seconds = e->build_seconds;
    return VRAD_OK;
}

int vrad_env_download_tree(vrad_env* e, int32_t* children, float* split, int32_t* tri_index, vrad_tri48* tris) {
    VRAD_MULTI_RANK0(e);
    if (!e) return VRAD_E_INVALID;
    if (!e->built) { set_error("vrad_env_download_tree: not built"); return VRAD_E_STATE; }
    // read back from the device copy so the test sees what the kernels see
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const size_t nn = e->scene.n_nodes, ni = e->scene.n_idx, nt = e->scene.n_tris;
    std::vector<int2> nodes(nn);
    VRAD_CUDA_CHECK(cudaMemcpy(nodes.data(), e->d_nodes.p, nn * sizeof(int2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nn; i++) {
        if (children) children[i] = nodes[i].x;
        if (split) memcpy(&split[i], &nodes[i].y, 4);
    }
    if (tri_index && ni) VRAD_CUDA_CHECK(cudaMemcpy(tri_index, e->d_tri_index.p, ni * 4, cudaMemcpyDeviceToHost));
    if (tris && nt) {
        std::vector<float4> q0(nt), q1(nt), q2(nt);
        VRAD_CUDA_CHECK(cudaMemcpy(q0.data(), e->d_q0.p, nt * 16, cudaMemcpyDeviceToHost));
        VRAD_CUDA_CHECK(cudaMemcpy(q1.data(), e->d_q1.p, nt * 16, cudaMemcpyDeviceToHost));
        VRAD_CUDA_CHECK(cudaMemcpy(q2.data(), e->d_q2.p, nt * 16, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nt; i++) {
            vrad_tri48& t = tris[i];
            t.nx = q0[i].x; t.ny = q0[i].y; t.nz = q0[i].z; t.d = q0[i].w;
            t.e[0] = q1[i].x; t.e[1] = q1[i].y; t.e[2] = q1[i].z; t.e[3] = q1[i].w;
            t.e[4] = q2[i].x; t.e[5] = q2[i].y;
            int32_t sel;
            memcpy(&t.id, &q2[i].z, 4); memcpy(&sel, &q2[i].w, 4);
            t.sel0 = sel & 0xff; t.sel1 = (sel >> 8) & 0xff; t.flags = (sel >> 16) & 0xff; t.unused = 0;
        }
    }
    return VRAD_OK;
}

int vrad_trace_rays(vrad_env* e, int64_t n, const float* ox, const float* oy, const float* oz, const float* dx,
                    const float* dy, const float* dz, const float* tmin, const float* tmax, int32_t skip_id,
                    int32_t* hit_tri, int32_t* hit_sid, float* hit_t) {
    VRAD_MULTI(e, group_trace_rays(e, n, ox, oy, oz, dx, dy, dz, tmin, tmax, skip_id, hit_tri, hit_sid, hit_t));
    if (!e || n < 0 || (n > 0 && (!ox || !oy || !oz || !dx || !dy || !dz || !tmax))) { set_error("vrad_trace_rays: bad arguments"); return VRAD_E_INVALID; }
    if (!e->built) { set_error("vrad_trace_rays: acceleration structure not built"); return VRAD_E_STATE; }
    if (n == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const size_t b = (size_t)n * 4;
    const void* in[8]; bool hin[8]; bool any_host = false;
    const float* src[8] = {ox, oy, oz, dx, dy, dz, tmin, tmax};
    for (int k = 0; k < 8; k++) { int rc = stage_in(e, k, src[k], b, &in[k], &hin[k]); if (rc) return rc; any_host |= hin[k]; }
    void* o_tri; void* o_sid; void* o_t; bool h0, h1, h2;
    int rc;
    if ((rc = stage_out(e, 8, hit_tri, b, &o_tri, &h0))) return rc;
    if ((rc = stage_out(e, 9, hit_sid, b, &o_sid, &h1))) return rc;
    if ((rc = stage_out(e, 10, hit_t, b, &o_t, &h2))) return rc;
    any_host |= h0 | h1 | h2;
    rc = launch_trace_rays(e, n, (const float*)in[0], (const float*)in[1], (const float*)in[2], (const float*)in[3],
                           (const float*)in[4], (const float*)in[5], (const float*)in[6], (const float*)in[7], skip_id,
                           (int32_t*)o_tri, (int32_t*)o_sid, (float*)o_t, nullptr);
    if (rc) return rc;
    if ((rc = finish_out(e, hit_tri, o_tri, b, h0))) return rc;
    if ((rc = finish_out(e, hit_sid, o_sid, b, h1))) return rc;
    if ((rc = finish_out(e, hit_t, o_t, b, h2))) return rc;
    return sync_if_needed(e, any_host);
}

int vrad_trace4(vrad_env* e, const float origin_xyz4[12], const float dir_xyz4[12], const float tmin[4], const float tmax[4],
                int32_t skip_id, int32_t hit_ids[4], float hit_dist[4], float normal_xyz4[12]) {
    VRAD_MULTI_RANK0(e);            // one packet: a latency call, any rank answers it
    if (!e || !origin_xyz4 || !dir_xyz4 || !tmin || !tmax || !hit_ids || !hit_dist) { set_error("vrad_trace4: bad arguments"); return VRAD_E_INVALID; }
    if (!e->built) { set_error("vrad_trace4: acceleration structure not built"); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    // one FourRays packet == a 4-ray batch: lanes are independent rays (FourVectors is x[4] y[4] z[4])
    float h_in[32];
    memcpy(h_in, origin_xyz4, 48); memcpy(h_in + 12, dir_xyz4, 48); memcpy(h_in + 24, tmin, 16); memcpy(h_in + 28, tmax, 16);
    const void* d_in; bool hh;
    int rc = stage_in(e, 0, h_in, sizeof(h_in), &d_in, &hh);
    if (rc) return rc;
    // the 80-byte result block (ids, distances, normals) always lands in the handle's scratch: hit_ids is a 16-byte buffer of the
    // caller's, host or device, and only its own 4 values are written through it
    void* d_out;
    if ((rc = scratch_get(e, 1, 4 * (4 + 4 + 12), &d_out))) return rc;
    const float* f = (const float*)d_in;
    int32_t* o_tri = (int32_t*)d_out; float* o_t = (float*)d_out + 4; float* o_n = (float*)d_out + 8;
    rc = launch_trace_rays(e, 4, f, f + 4, f + 8, f + 12, f + 16, f + 20, f + 24, f + 28, skip_id, o_tri, nullptr, o_t, o_n);
    if (rc) return rc;
    float h_out[20];
    VRAD_CUDA_CHECK(cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    auto put = [&](void* dst, const float* src, size_t bytes) -> cudaError_t {
        if (is_device_ptr(dst)) return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
        memcpy(dst, src, bytes); return cudaSuccess;
    };
    VRAD_CUDA_CHECK(put(hit_ids, h_out, 16));
    VRAD_CUDA_CHECK(put(hit_dist, h_out + 4, 16));
    if (normal_xyz4) VRAD_CUDA_CHECK(put(normal_xyz4, h_out + 8, 48));
    return VRAD_OK;
}

int vrad_test_lines(vrad_env* e, int64_t n, const float* start_xyz_soa, const float* stop_xyz_soa, int sky_mode, uint32_t* vis_bits) {
    VRAD_MULTI(e, group_test_lines(e, n, start_xyz_soa, stop_xyz_soa, sky_mode, vis_bits));
    if (!e || n < 0 || (n > 0 && (!start_xyz_soa || !stop_xyz_soa || !vis_bits))) { set_error("vrad_test_lines: bad arguments"); return VRAD_E_INVALID; }
    if (!e->built) { set_error("vrad_test_lines: acceleration structure not built"); return VRAD_E_STATE; }
    if (n == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const size_t b = (size_t)n * 12, wb = (size_t)((n + 31) / 32) * 4;
    const void *d_a, *d_b; bool ha, hb, ho;
    int rc;
    if (n >= ((int64_t)1 << 22) && !is_device_ptr(start_xyz_soa) && !is_device_ptr(stop_xyz_soa)) {
        // large host batch: overlap the H2D copies with the traversal
        void* d_o;
        if ((rc = stage_out(e, 2, vis_bits, wb, &d_o, &ho))) return rc;
        if ((rc = launch_test_lines_pipelined(e, n, start_xyz_soa, stop_xyz_soa, n, nullptr, sky_mode, (uint32_t*)d_o))) return rc;
        if ((rc = finish_out(e, vis_bits, d_o, wb, ho))) return rc;
        int bad = 0;
        if ((rc = read_bad_index_count(e, &bad))) return rc;         // synchronises; non-zero: the kernel gave up waiting for its input
        if (bad) { set_error("vrad_test_lines: the segment copy did not arrive on the device"); return VRAD_E_CUDA; }
        return VRAD_OK;
    }
    if ((rc = stage_in(e, 0, start_xyz_soa, b, &d_a, &ha))) return rc;
    if ((rc = stage_in(e, 1, stop_xyz_soa, b, &d_b, &hb))) return rc;
    void* d_o;
    if ((rc = stage_out(e, 2, vis_bits, wb, &d_o, &ho))) return rc;
    if ((rc = launch_test_lines(e, n, (const float*)d_a, (const float*)d_b, sky_mode, (uint32_t*)d_o))) return rc;
    if ((rc = finish_out(e, vis_bits, d_o, wb, ho))) return rc;
    return sync_if_needed(e, ha | hb | ho);
}

/* endpoint table for vrad_test_lines_indexed: patch origins, light origins, luxel samples ... (xyz interleaved) */
int vrad_points_upload(vrad_env* e, int64_t n_points, const float* xyz3) {
    VRAD_MULTI(e, group_points_upload(e, n_points, xyz3));
    if (!e || n_points <= 0 || !xyz3) { set_error("vrad_points_upload: bad arguments"); return VRAD_E_INVALID; }
    if (n_points > 0x7fffffff) { set_error("vrad_points_upload: %lld points exceed the int32 index range", (long long)n_points); return VRAD_E_INVALID; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    if (e->d_points.alloc((size_t)n_points)) { set_error("out of device memory for %lld points", (long long)n_points); return VRAD_E_NOMEM; }
    std::vector<float> host;
    const float* src = xyz3;
    if (is_device_ptr(xyz3)) {
        host.resize(3 * (size_t)n_points);
        VRAD_CUDA_CHECK(cudaMemcpy(host.data(), xyz3, host.size() * 4, cudaMemcpyDeviceToHost));
        src = host.data();
    }
    std::vector<float4> p4((size_t)n_points);
    for (int64_t i = 0; i < n_points; i++) p4[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
    VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_points.p, p4.data(), p4.size() * sizeof(float4), cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    e->n_points = n_points;
    return VRAD_OK;
}

int vrad_test_lines_indexed(vrad_env* e, int64_t n, const int32_t* pairs2, int sky_mode, uint32_t* vis_bits) {
    VRAD_MULTI(e, group_test_lines_indexed(e, n, pairs2, sky_mode, vis_bits));
    if (!e || n < 0 || (n > 0 && (!pairs2 || !vis_bits))) { set_error("vrad_test_lines_indexed: bad arguments"); return VRAD_E_INVALID; }
    if (!e->built) { set_error("vrad_test_lines_indexed: acceleration structure not built"); return VRAD_E_STATE; }
    if (e->n_points == 0) { set_error("vrad_test_lines_indexed: no point table (vrad_points_upload first)"); return VRAD_E_STATE; }
    if (n == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const size_t b = (size_t)n * 8, wb = (size_t)((n + 31) / 32) * 4;
    int rc; bool hp, ho; void* d_o;
    const bool host_pairs = !is_device_ptr(pairs2);
    // Indices are checked on the device, next to the traversal (which clamps them, so a bad index is never an
    // out-of-bounds read): a host pass over the pairs would cost more than the PCIe copy it precedes.
    if ((rc = stage_out(e, 2, vis_bits, wb, &d_o, &ho))) return rc;
    if (host_pairs && n >= ((int64_t)1 << 22)) {
        if ((rc = launch_test_lines_pipelined(e, n, nullptr, nullptr, 0, pairs2, sky_mode, (uint32_t*)d_o))) return rc;
        if ((rc = finish_out(e, vis_bits, d_o, wb, ho))) return rc;
        int bad = 0;
        if ((rc = read_bad_index_count(e, &bad))) return rc;         // synchronises
        if (bad) { set_error("vrad_test_lines_indexed: %d indices outside the %lld-point table", bad, (long long)e->n_points); return VRAD_E_INVALID; }
        return VRAD_OK;
    }
    const void* d_p;
    if ((rc = stage_in(e, 0, pairs2, b, &d_p, &hp))) return rc;
    if (host_pairs || !e->async) {           // async callers with device buffers: the kernels clamp the indices (memory-safe), nothing is read back
        int bad = 0;
        if ((rc = check_pairs_on_device(e, n, (const int32_t*)d_p, &bad))) return rc;
        if (bad) { set_error("vrad_test_lines_indexed: %d indices outside the %lld-point table", bad, (long long)e->n_points); return VRAD_E_INVALID; }
    }
    if ((rc = launch_test_lines_indexed(e, n, (const int32_t*)d_p, sky_mode, (uint32_t*)d_o))) return rc;
    if ((rc = finish_out(e, vis_bits, d_o, wb, ho))) return rc;
    return sync_if_needed(e, hp | ho);
}

} // extern "C"
