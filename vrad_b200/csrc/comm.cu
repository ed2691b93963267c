// comm.cu -- multi-GPU plumbing: one handle per process/GPU; the only data-path collective is the
// per-bounce radiance all-gather of K4 (plus a 3-float all-reduce of `added`).  NCCL is resolved at
// run time with dlopen("libnccl.so.2") so that a process that already loaded NCCL (e.g. through
// torch.distributed) shares that copy and a single-GPU user needs no NCCL at all.
// The reference has no communication layer of any kind.
#include "env_internal.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <string>
#include <vector>

namespace vrad {

void comm_close_peers(vrad_env* e);

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm_t*, int, nccl_uid, int);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_allgather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_bcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);

static struct {
    void* h = nullptr;
    fn_get_uid get_uid; fn_init_rank init_rank; fn_destroy destroy; fn_errstr errstr;
    fn_allgather allgather; fn_allreduce allreduce; fn_bcast bcast;
} g_nccl;

constexpr int kNcclFloat = 7;   // ncclFloat32
constexpr int kNcclSum = 0;     // ncclSum
constexpr int kNcclInt32 = 2;   // ncclInt32
constexpr int kNcclInt64 = 4;   // ncclInt64

static int load_nccl() {
    if (g_nccl.h) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return VRAD_E_COMM; }
    g_nccl.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    g_nccl.allgather = (fn_allgather)dlsym(h, "ncclAllGather");
    g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_nccl.bcast = (fn_bcast)dlsym(h, "ncclBroadcast");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.errstr || !g_nccl.allgather || !g_nccl.allreduce || !g_nccl.bcast) {
        set_error("libnccl.so.2 lacks a required symbol"); dlclose(h); return VRAD_E_COMM;
    }
    g_nccl.h = h;
    return 0;
}

#define VRAD_NCCL_CHECK(expr)                                                                  \
    do {                                                                                       \
        int _r = (expr);                                                                       \
        if (_r != 0) { set_error("%s failed: %s", #expr, g_nccl.errstr(_r)); return VRAD_E_COMM; } \
    } while (0)

// ---- collectives of an in-process group (LocalGroup): host memory + a barrier instead of NCCL ------------------------------
#define VRAD_GROUP_BARRIER(G) do { if (!(G).barrier()) { set_error("a rank of the in-process group failed"); return VRAD_E_COMM; } } while (0)

static int group_allgather_rows(vrad_env* e, float4* buf, const int64_t* bounds) {
    LocalGroup& G = *e->group;
    const int rank = e->cfg.rank;
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));        // my rows are in place before anybody reads them
    G.ptr[rank][0] = buf;
    VRAD_GROUP_BARRIER(G);
    for (int r = 0; r < G.world; r++) {
        const int64_t cnt = bounds[r + 1] - bounds[r];
        if (r == rank || cnt <= 0) continue;
        VRAD_CUDA_CHECK(cudaMemcpyPeerAsync(buf + bounds[r], e->cfg.device, (const float4*)G.ptr[r][0] + bounds[r], G.ranks[r]->cfg.device,
                                            (size_t)cnt * sizeof(float4), e->stream));
    }
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    VRAD_GROUP_BARRIER(G);                                    // nobody overwrites its rows while a peer still copies them
    return 0;
}

static int group_exchange_rows(vrad_env* e, int64_t row0, int64_t row1, int64_t* all2) {
    LocalGroup& G = *e->group;
    G.rows[e->cfg.rank][0] = row0; G.rows[e->cfg.rank][1] = row1;
    VRAD_GROUP_BARRIER(G);
    for (int r = 0; r < G.world; r++) { all2[2 * r] = G.rows[r][0]; all2[2 * r + 1] = G.rows[r][1]; }
    VRAD_GROUP_BARRIER(G);
    return 0;
}

static int group_allreduce_i32(vrad_env* e, int32_t* d, size_t n) {
    LocalGroup& G = *e->group;
    const int rank = e->cfg.rank;
    G.i32[rank].resize(n);
    VRAD_CUDA_CHECK(cudaMemcpyAsync(G.i32[rank].data(), d, n * 4, cudaMemcpyDeviceToHost, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    VRAD_GROUP_BARRIER(G);
    std::vector<int32_t> sum(n, 0);
    for (int r = 0; r < G.world; r++) for (size_t i = 0; i < n; i++) sum[i] += G.i32[r][i];
    VRAD_CUDA_CHECK(cudaMemcpyAsync(d, sum.data(), n * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    VRAD_GROUP_BARRIER(G);
    return 0;
}

static int group_allreduce3(vrad_env* e, float* d3) {
    LocalGroup& G = *e->group;
    const int rank = e->cfg.rank;
    VRAD_CUDA_CHECK(cudaMemcpyAsync(G.f3[rank], d3, 12, cudaMemcpyDeviceToHost, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    VRAD_GROUP_BARRIER(G);
    float s[3] = {0.f, 0.f, 0.f};
    for (int r = 0; r < G.world; r++) for (int c = 0; c < 3; c++) s[c] += G.f3[r][c];        // rank order: the same sum on every rank
    VRAD_CUDA_CHECK(cudaMemcpyAsync(d3, s, 12, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    VRAD_GROUP_BARRIER(G);
    return 0;
}

// peer buffers of an in-process group: plain pointers once peer access is on.  Ranks that share a device (single-GPU tests) do
// not use the fused exchange: the in-kernel barrier needs the ranks' kernels to run side by side, which one device does not promise.
static int group_setup_peers(vrad_env* e, size_t n_pad) {
    LocalGroup& G = *e->group;
    PeerLinks& P = e->peers;
    const int world = G.world, rank = e->cfg.rank;
    if (P.ready && P.n_pad == n_pad) return 0;
    if (G.shares_device || world > kMaxWorld) return 0;
    if (P.d_flags.alloc(kFlagWords)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemsetAsync(P.d_flags.p, 0, kFlagWords * sizeof(uint32_t), e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    int ok = 1;
    for (int r = 0; r < world && ok; r++) {
        if (r == rank) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, e->cfg.device, G.ranks[r]->cfg.device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
        const cudaError_t ce = cudaDeviceEnablePeerAccess(G.ranks[r]->cfg.device, 0);
        if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
        cudaGetLastError();
    }
    G.ptr[rank][0] = e->d_er[0].p; G.ptr[rank][1] = e->d_er[1].p; G.ptr[rank][2] = P.d_flags.p; G.ok[rank] = ok;
    VRAD_GROUP_BARRIER(G);
    bool all_ok = true;
    for (int r = 0; r < world; r++) all_ok = all_ok && G.ok[r];
    PeerTable tbl{};
    for (int r = 0; r < world; r++) {
        P.er[0][r] = tbl.er[0][r] = (float4*)G.ptr[r][0]; P.er[1][r] = tbl.er[1][r] = (float4*)G.ptr[r][1];
        P.flags[r] = tbl.flags[r] = (uint32_t*)G.ptr[r][2];
    }
    tbl.world = world; tbl.rank = rank;
    VRAD_GROUP_BARRIER(G);
    if (!all_ok) { P.ready = false; return 0; }
    if (P.d_table.alloc(1)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpy(P.d_table.p, &tbl, sizeof(tbl), cudaMemcpyHostToDevice));
    P.ready = true; P.simulated = false; P.n_pad = n_pad; P.table_world = world;
    return 0;
}

// Row-indexed all-gather for arbitrary contiguous row blocks: rank r owns rows [bounds[r], bounds[r+1]) of the
// float4 array `buf` (indexed by global row) and broadcasts them in place.  `world` small broadcasts; used for
// the final `total` gather and as the exchange when peer mapping is unavailable.
int comm_allgather_rows(vrad_env* e, float4* buf, const int64_t* bounds) {
    if (e->group) return group_allgather_rows(e, buf, bounds);
    if (!e->nccl_comm) { set_error("communicator not initialised"); return VRAD_E_STATE; }
    for (int r = 0; r < e->cfg.world; r++) {
        const int64_t cnt = bounds[r + 1] - bounds[r];
        if (cnt <= 0) continue;
        VRAD_NCCL_CHECK(g_nccl.bcast(buf + bounds[r], buf + bounds[r], (size_t)cnt * 4, kNcclFloat, r, (nccl_comm_t)e->nccl_comm, e->stream));
    }
    return 0;
}

// every rank contributes its [row0,row1); returns the world+1 boundaries, checked to tile [0,n) in rank order
int comm_exchange_bounds(vrad_env* e, int64_t row0, int64_t row1, int64_t n, int64_t* bounds_out) {
    const int world = e->cfg.world;
    std::vector<int64_t> all(2 * (size_t)world);
    if (e->group) { int rcg = group_exchange_rows(e, row0, row1, all.data()); if (rcg) return rcg; }
    else {
        if (!e->nccl_comm) { set_error("communicator not initialised"); return VRAD_E_STATE; }
        DevBuf<int64_t> d;
        if (d.alloc(2 * (size_t)world)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
        const int64_t mine[2] = {row0, row1};
        VRAD_CUDA_CHECK(cudaMemcpyAsync(d.p + 2 * e->cfg.rank, mine, 16, cudaMemcpyHostToDevice, e->stream));
        VRAD_NCCL_CHECK(g_nccl.allgather(d.p + 2 * e->cfg.rank, d.p, 2, kNcclInt64, (nccl_comm_t)e->nccl_comm, e->stream));
        VRAD_CUDA_CHECK(cudaMemcpyAsync(all.data(), d.p, 16 * (size_t)world, cudaMemcpyDeviceToHost, e->stream));
        VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
        d.release();
    }
    int64_t expect = 0;
    for (int r = 0; r < world; r++) {
        if (all[2 * r] != expect || all[2 * r + 1] < all[2 * r]) {
            set_error("transfer rows of the ranks do not tile [0,%lld): rank %d holds [%lld,%lld), expected start %lld",
                      (long long)n, r, (long long)all[2 * r], (long long)all[2 * r + 1], (long long)expect);
            return VRAD_E_STATE;
        }
        bounds_out[r] = all[2 * r];
        expect = all[2 * r + 1];
    }
    if (expect != n) { set_error("transfer rows of the ranks end at %lld, not %lld", (long long)expect, (long long)n); return VRAD_E_STATE; }
    bounds_out[world] = n;
    return 0;
}

// in-place sum of an int32 device array over all ranks
int comm_allreduce_i32(vrad_env* e, int32_t* d, size_t n) {
    if (e->group) return group_allreduce_i32(e, d, n);
    if (!e->nccl_comm) { set_error("communicator not initialised"); return VRAD_E_STATE; }
    VRAD_NCCL_CHECK(g_nccl.allreduce(d, d, n, kNcclInt32, kNcclSum, (nccl_comm_t)e->nccl_comm, e->stream));
    return 0;
}

int comm_allreduce3(vrad_env* e, float* d3) {
    if (e->group) return group_allreduce3(e, d3);
    if (!e->nccl_comm) { set_error("communicator not initialised"); return VRAD_E_STATE; }
    VRAD_NCCL_CHECK(g_nccl.allreduce(d3, d3, 3, kNcclFloat, kNcclSum, (nccl_comm_t)e->nccl_comm, e->stream));
    return 0;
}

// Exchange CUDA IPC handles of er[0], er[1] and the flag words through the NCCL communicator and map
// every peer's buffers.  Collective: all ranks call it with the same n_pad.  Returns VRAD_OK with
// peers.ready == false when peer mapping is unavailable (the caller then keeps the NCCL all-gather).
int comm_setup_peers(vrad_env* e, size_t n_pad) {
    if (e->group) return group_setup_peers(e, n_pad);
    PeerLinks& P = e->peers;
    const int world = e->cfg.world, rank = e->cfg.rank;
    if (P.ready && P.n_pad == n_pad) return 0;
    if (world > kMaxWorld) return 0;
    static const bool force_nccl = [] { const char* v = getenv("VRAD_K4_EXCHANGE"); return v && std::string(v) == "nccl"; }();
    if (force_nccl) return 0;
    comm_close_peers(e);
    if (P.d_flags.alloc(kFlagWords)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemsetAsync(P.d_flags.p, 0, kFlagWords * sizeof(uint32_t), e->stream));
    struct Handles { cudaIpcMemHandle_t h[3]; int ok; int pad[15]; };
    static_assert(sizeof(Handles) % 4 == 0, "handle record must be a whole number of floats");
    std::vector<Handles> all(world);
    Handles mine{};
    mine.ok = cudaIpcGetMemHandle(&mine.h[0], e->d_er[0].p) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.h[1], e->d_er[1].p) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.h[2], P.d_flags.p) == cudaSuccess;
    cudaGetLastError();
    DevBuf<unsigned char> d_all;
    if (d_all.alloc(sizeof(Handles) * world)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpyAsync(d_all.p + sizeof(Handles) * rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice, e->stream));
    VRAD_NCCL_CHECK(g_nccl.allgather(d_all.p + sizeof(Handles) * rank, d_all.p, sizeof(Handles) / 4, kNcclFloat, (nccl_comm_t)e->nccl_comm, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all.p, sizeof(Handles) * world, cudaMemcpyDeviceToHost, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    d_all.release();
    bool ok = true;
    for (int r = 0; r < world; r++) ok = ok && all[r].ok;
    for (int r = 0; r < world && ok; r++) {
        if (r == rank) { P.er[0][r] = e->d_er[0].p; P.er[1][r] = e->d_er[1].p; P.flags[r] = P.d_flags.p; continue; }
        for (int k = 0; k < 3; k++) {
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            P.opened[k][r] = ptr;
        }
        P.er[0][r] = (float4*)P.opened[0][r]; P.er[1][r] = (float4*)P.opened[1][r]; P.flags[r] = (uint32_t*)P.opened[2][r];
    }
    // every rank must take the same path: agree on success with a 3-float all-reduce (min via negated sum)
    DevBuf<float> d_ok;
    if (d_ok.alloc(3)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    float h_ok[3] = {ok ? 0.0f : 1.0f, 0.0f, 0.0f};
    VRAD_CUDA_CHECK(cudaMemcpyAsync(d_ok.p, h_ok, 12, cudaMemcpyHostToDevice, e->stream));
    VRAD_NCCL_CHECK(g_nccl.allreduce(d_ok.p, d_ok.p, 3, kNcclFloat, kNcclSum, (nccl_comm_t)e->nccl_comm, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(h_ok, d_ok.p, 12, cudaMemcpyDeviceToHost, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    d_ok.release();
    if (h_ok[0] != 0.0f) { comm_close_peers(e); return 0; }
    PeerTable tbl{};
    for (int r = 0; r < world; r++) { tbl.er[0][r] = P.er[0][r]; tbl.er[1][r] = P.er[1][r]; tbl.flags[r] = P.flags[r]; }
    tbl.world = world; tbl.rank = rank;
    if (P.d_table.alloc(1)) { set_error("out of device memory"); return VRAD_E_NOMEM; }
    VRAD_CUDA_CHECK(cudaMemcpy(P.d_table.p, &tbl, sizeof(tbl), cudaMemcpyHostToDevice));
    P.ready = true; P.simulated = false; P.n_pad = n_pad; P.table_world = world;
    return 0;
}

void comm_close_peers(vrad_env* e) {
    PeerLinks& P = e->peers;
    for (int k = 0; k < 3; k++)
        for (int r = 0; r < kMaxWorld; r++)
            if (P.opened[k][r]) { cudaIpcCloseMemHandle(P.opened[k][r]); P.opened[k][r] = nullptr; }
    P.ready = false; P.n_pad = 0;
}

} // namespace vrad
using namespace vrad;

extern "C" {

int vrad_comm_unique_id(void* out128) {
    if (!out128) return VRAD_E_INVALID;
    int rc = load_nccl();
    if (rc) return rc;
    nccl_uid id;
    VRAD_NCCL_CHECK(g_nccl.get_uid(&id));
    memcpy(out128, id.internal, 128);
    return VRAD_OK;
}

int vrad_comm_init(vrad_env* e, const void* unique_id128) {
    if (e && e->multi) return VRAD_OK;      // an in-process handle needs no communicator id
    if (!e || !unique_id128) return VRAD_E_INVALID;
    if (e->cfg.world == 1) return VRAD_OK;
    if (e->nccl_comm) { set_error("vrad_comm_init: already initialised"); return VRAD_E_STATE; }
    int rc = load_nccl();
    if (rc) return rc;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    nccl_uid id;
    memcpy(id.internal, unique_id128, 128);
    nccl_comm_t c = nullptr;
    VRAD_NCCL_CHECK(g_nccl.init_rank(&c, e->cfg.world, id, e->cfg.rank));
    e->nccl_comm = c;
    return VRAD_OK;
}

void vrad_comm_destroy_internal(vrad_env* e) {
    if (e) { comm_close_peers(e); e->peers.d_flags.release(); e->peers.d_table.release(); }
    if (e && e->nccl_comm && g_nccl.h) { g_nccl.destroy((nccl_comm_t)e->nccl_comm); e->nccl_comm = nullptr; }
}

} // extern "C"
