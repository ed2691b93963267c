// env_internal.cuh -- the vrad_env handle and small host helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/vrad_cuda.h"
#include "common.cuh"
#include "kd_builder.hpp"

namespace vrad {

void set_error(const char* fmt, ...);

#define VRAD_CUDA_CHECK(expr)                                                                       \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            vrad::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return VRAD_E_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

// Owning device buffer (raw cudaMalloc; freed with the handle).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        release();
        if (count == 0) return 0;
        if (cudaMalloc((void**)&p, count * sizeof(T)) != cudaSuccess) { p = nullptr; n = 0; cudaGetLastError(); return -1; }
        n = count;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct PatchesDev {
    int n = 0;
    DevBuf<float4> origin_area;     // origin.xyz, area
    DevBuf<float4> normal_dist;     // normal.xyz, plane_dist
    DevBuf<float4> refl;            // reflectivity.rgb, sky flag (1.0 = sky)
    DevBuf<int32_t> cluster;
    // Patch.Winding (common/types/patch.go): {first point, count} per patch + the points, clockwise seen from the front; empty = MakeTransfer
    // uses the differential form factor everywhere (vrad_patches_set_windings)
    DevBuf<int2> wind; DevBuf<float4> wind_pts; bool has_windings = false;
    std::vector<int32_t> h_cluster;
    std::vector<uint8_t> h_flags;
    std::vector<float> h_normal;                // for the winding orientation check
    std::vector<float> h_area, h_refl;          // host copies for the hierarchy checks / collect weights
    // patch hierarchy (Patch.Parent / Child1 / FaceNumber, common/types/patch.go:33,49-51); empty = flat
    bool hier = false;
    DevBuf<int4> tree;                          // {parent, child1, face, cluster of the face root} per patch
    std::vector<int32_t> h_root_cluster;
    // CollectLight for interior patches, flattened: interior patch p = sum over the leaves of its subtree of
    // w(p, leaf) * value(leaf), w = product of the area fractions along the path (vrad.cpp CollectLight, App. B.4)
    int n_interior = 0, n_collect_long = 0;     // collect rows are ordered long (>= 128 leaves) first
    DevBuf<int32_t> collect_ids;                // interior patch numbers
    DevBuf<int64_t> collect_ptr;                // n_interior + 1 offsets into collect_ent
    DevBuf<int2>    collect_ent;                // {leaf patch, weight bits}
    std::vector<int32_t> h_child1, h_parent;
    // bump-mapped patches (Patch.NeedsBumpMap, common/types/patch.go:23): flags, normals[1..3] as three float4 per patch, the
    // local rows that gather bump sums, and TotalLight.Light[1..3] (common/types/bumpLights.go:8-10) indexed by global patch
    bool bump = false;
    std::vector<uint8_t> h_needs_bump;
    DevBuf<float4> bump_normals;                // [N][3]
    DevBuf<float4> total_bump[3];
    DevBuf<int32_t> bump_rows;
    int n_bump_rows = 0;
    int64_t bump_rows_row0 = -1, bump_rows_row1 = -1;
    DevBuf<int32_t> child2;
    DevBuf<int32_t> leaf_rows;                  // local row numbers of the leaf patches in [leaf_rows_row0, leaf_rows_row1)
    int n_leaf_rows = 0;
    int64_t leaf_rows_row0 = -1, leaf_rows_row1 = -1;
};

struct TransfersDev {
    int64_t row0 = 0, row1 = 0;     // rows owned by this rank
    int64_t nnz = 0;                // logical (unpadded) entries
    int64_t nnz_padded = 0;         // device entries (rows padded to multiples of 4)
    DevBuf<int64_t> rowptr;         // padded offsets, row1-row0+1 entries, in units of entries
    DevBuf<int32_t> rowlen;         // logical row lengths
    DevBuf<int2>    tr;             // {col, w bits} pairs == the reference's Transfer struct (transfer.go:3-6)
    bool ready = false;
    bool rows_ascending = true;     // columns strictly ascending inside every row (always so from vrad_build_transfers; checked by vrad_transfers_upload)
    // Gather plan (K4): one warp per work item.  A row longer than `seg` entries is cut into parts of `seg` entries
    // (one item each) so that no warp carries a 30 us row into the tail of a 40 us kernel; the parts' sums meet in
    // part_sum[] and the part that arrives last (row_ctr) adds them in part order and runs the row's epilogue.
    //   item int4 = {local row, entries | n_parts << 16 | part index << 24, first entry (position in tr) lo, hi}
    DevBuf<int4>    items;
    DevBuf<int32_t> item_slot;      // per item: first part slot of its row in part_sum (-1: the row is one item)
    DevBuf<int32_t> block_ptr;      // n_blocks + 1 offsets into items: the items of each block of the gather grid
    DevBuf<float4>  part_sum;       // one slot per part of a split row
    DevBuf<int32_t> row_ctr;        // arrivals per local row (split rows only; the finisher resets it)
    int n_items = 0, n_slots = 0, n_blocks = 0, seg_shift = 11, plan_warps = 8, pool_begin = 0;
    // Packed transfer streams (K4, k4_pack): the same entries as tr[], split into a weight stream (f32) and a column stream (u16,
    // relative to its segment's base) -- 6 bytes per transfer instead of 8.  A segment is a run of a row's entries whose columns
    // lie within one 65,536-column window counted from the row's first column (and at most pk_max_seg entries); it starts on a
    // 64-entry boundary of both streams (padding: weight 0, column offset 0), and inside each group of 64 the entries are interleaved
    // (positions 2L, 2L+1 = entries L, 32+L) so that one vector load hands a lane two entries 32 apart.
    //   segment int4 = {first entry (position in the packed streams) lo, hi, padded entries, column base}
    DevBuf<float>    pk_w;
    DevBuf<uint16_t> pk_c;
    DevBuf<int4>     pk_segs;
    DevBuf<int32_t>  pk_seg_ptr;    // row1-row0+1 offsets into pk_segs
    int64_t pk_entries = 0; int pk_n_segs = 0; bool packed = false;
    bool plan_packed = false;       // the items of the plan below are segments of the packed streams
    bool plan_blocked = false;      // ... or blocks of rows of the block-row streams
    // Block-row streams (K4, k4_pack = 2): kBlockRows consecutive rows (global row numbers, aligned) share ONE column list -- the union of
    // their columns -- with kBlockRows weights per entry (0 where a row does not have the column): rows of neighbouring patches see almost
    // the same emitters (C4 map: union of 4 rows = 1.14 x one row), so 4 rows cost 1.14 x (2 + 16) bytes instead of 4 x 6, and -- what
    // matters more -- ONE 16-byte er[] gather serves 4 rows.  Segments as for the packed streams (column windows of 65,536 from the
    // block's first column, at most bk_max_seg entries, padded to 32 entries).
    //   segment int4 = {first entry (position in bk_c / bk_w) lo, hi, padded entries, column base}
    DevBuf<float>    bk_w;          // bk_rows weights per entry: those of the block's rows for this column
    DevBuf<uint16_t> bk_c;
    DevBuf<int4>     bk_segs;
    DevBuf<int32_t>  bk_seg_ptr;    // per block of rows: offsets into bk_segs (n_row_blocks + 1)
    int64_t bk_entries = 0; int bk_n_segs = 0, bk_n_blocks = 0, bk_rows = 4; bool blocked = false;
    int64_t plan_serial = 0;        // bumped by every re-plan (invalidates the captured bounce graph)
    int64_t rows_serial = 0;        // bumped when the resident rows change (vrad_build_transfers / vrad_transfers_upload: collective calls)
};

// BSP point-location data on the device (trace.PointLeafnum, clustertable.PointInLeaf) and the sky cameras.
//   nodes   int4   [n_nodes]  {plane, child0, child1, plane axis type}  -- one 128-bit load per step
//   planes  float4 [n_planes] {normal.xyz, dist}
struct DevBsp {
    const int4*    nodes;
    const float4*  planes;
    const int32_t* leaf_cluster;
    const int32_t* leaf_area;
    const int32_t* area_camera;    // areaSkyCameras[n_areas], -1 = none (cache/skycameras.go:10)
    const float4*  cams;           // {origin.xyz, WorldToSky} per sky camera
    int n_nodes, n_leafs, n_areas, n_cams;
};

constexpr int kMaxWorld = 8;

// the bounce loop of a call, captured as a CUDA graph (k4_bounce.cu); valid while every captured argument is unchanged
struct GraphCache {
    cudaGraphExec_t exec = nullptr;
    int n_bounces = 0, n_items = 0;
    const void *items = nullptr, *er0 = nullptr, *total = nullptr, *add = nullptr, *tr = nullptr;
    int64_t row0 = 0, tag = -1;
    bool p2p = false;
};

// Peer-memory view of the radiance buffers (multi-GPU K4): pointers into every rank's er[0]/er[1]
// and flag words, obtained through CUDA IPC (one process per GPU).  Slot `rank` is the local buffer.
struct PeerTable;
// flag words of one rank (device memory, mapped by every peer):
//   [0, kMaxWorld)      arrival words: word p = the last bounce epoch rank p has finished storing into this rank's buffers
//   [kFlagBase]         epoch of the bounce before the first one of the current vrad_bounce call (written by the owner only)
//   [kFlagTicket]       blocks of the running gather that have finished (the last one signals and resets it)
//   [kFlagPool]         items of the common pool handed out in the running gather (reset by its last block)
//   [kFlagError]        set when a wait gave up (a peer never signalled): the call returns VRAD_E_COMM instead of hanging
constexpr int kFlagBase = kMaxWorld, kFlagTicket = kMaxWorld + 1, kFlagError = kMaxWorld + 2, kFlagPool = kMaxWorld + 3, kFlagWords = 2 * kMaxWorld;
struct PeerLinks {
    bool      ready = false;
    bool      simulated = false;             // VRAD_K4_SIM_PEERS: every "peer" is this device (single-GPU tuning of the N-rank slice)
    size_t    n_pad = 0;
    int       table_world = 0;               // ranks in the peer table (== cfg.world, or 1 for a single-GPU handle run with k4_items)
    float4*   er[2][kMaxWorld] = {};
    uint32_t* flags[kMaxWorld] = {};
    void*     opened[3][kMaxWorld] = {};     // mappings to close
    DevBuf<uint32_t> d_flags;                // kFlagWords words, layout above
    DevBuf<float4>   d_sink;                 // simulated peers: where the rows "sent" to the other ranks land
    DevBuf<PeerTable> d_table;
};

// the same pointers as one record in device memory, read by k4_gather_items<true> and k4_peer_wait
struct PeerTable {
    float4*   er[2][kMaxWorld];
    uint32_t* flags[kMaxWorld];              // flags[p][rank] = this rank's arrival word in rank p's flag block
    int world, rank;
};

} // namespace vrad

#include <condition_variable>
#include <mutex>
namespace vrad {
// Ranks that live in ONE process (vrad_env_create_multi): the handle the caller holds owns one child environment per device and
// runs every collective call on one thread per child.  The children's collectives (row balance, block bounds, the fallback
// radiance exchange, `added`, peer-buffer discovery) go through this object -- shared host memory, a generation barrier,
// cudaMemcpyPeerAsync -- instead of NCCL, and peer buffers are plain pointers after cudaDeviceEnablePeerAccess (no IPC).
struct LocalGroup {
    int world = 0;
    std::vector<vrad_env*> ranks;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t gen = 0;
    bool failed = false;                 // a rank left a collective call with an error: the others must not wait for it
    // exchange slots, written by their rank before a barrier and read by the others after it
    void* ptr[8][4] = {};
    std::vector<int32_t> i32[8];
    float f3[8][3] = {};
    int64_t rows[8][2] = {};
    int ok[8] = {};
    bool shares_device = false;          // two ranks on one device (single-GPU tests): no in-kernel barrier between them
    // returns false when a rank failed
    bool barrier() {
        std::unique_lock<std::mutex> lk(m);
        if (failed) return false;
        const uint64_t g = gen;
        if (++arrived == world) { arrived = 0; gen++; cv.notify_all(); return true; }
        cv.wait(lk, [&] { return gen != g || failed; });
        return !failed;
    }
    void fail() { std::lock_guard<std::mutex> lk(m); failed = true; cv.notify_all(); }
    void reset() { std::lock_guard<std::mutex> lk(m); failed = false; arrived = 0; }
};
// tuning switches of a handle (vrad_env_set_option); the environment variables of the same meaning give the defaults
struct EnvOptions {
    int k1_sort = -1;      // VRAD_K1_SORT: order segment batches before tracing; -1 = batches of >= 65536 segments, 0 never, 1 always
    int k1_key = -1;       // VRAD_K1_KEY: layout of the sort key (k1_trace.cu: -1 by the scene box, 0 start-major Morton, 1 6-D Morton, 2 cubic start cells)
    int k1_sort_bits = 30; // VRAD_K1_SORT_BITS: leading key bits that take part in the radix sort (8..30; one pass per 8 bits)
    int k1_bpsm = 12;      // VRAD_K1_BPSM: resident blocks per SM of the streaming traversal kernel: 12 (40 registers), 10 (51), 8 (64)
    int k1_stream = 1;     // VRAD_K1_STREAM: unordered batches also go through the persistent streaming kernel (0 = the per-chunk kernels)
    int k1_top = 0;        // VRAD_K1_TOP: stage the top levels of the kd tree in shared memory (0 = off, else node budget)
    int k4_seg = 16384;    // VRAD_K4_SEG: entries per gather work item (rows longer than this are split)
    int k4_long_first = 0; // VRAD_K4_ORDER=long: work items longest first
    int k4_block = 0;      // VRAD_K4_BLOCK: threads per block of the multi-GPU gather: 192 (6 warps, 6 blocks/SM, 56 registers), 256 (8 warps, 5 blocks/SM, 48 registers), 0 = 256 with the packed streams, 192 with the pairs
    int k2_stream = 0;     // VRAD_K2_STREAM: pass A of the transfer build as compacted ray queues with lane refill (experiment, slower: DESIGN section 7; 0 = one ray slot per (row, candidate) thread)
    int k4_bk_min_items = 80; // VRAD_K4_BK_MIN_ITEMS: tenths of a block-row item per (SM x 8 warps) below which several ranks stay on the packed streams
    int k4_bk_rows = 4;    // VRAD_K4_BK_ROWS: rows per block of the block-row streams (2 or 4)
    int k4_short = -1;     // VRAD_K4_SHORT: the short-row gather (8 lanes per row) on one GPU: -1 = where rows average < 400 transfers, 0 = never, 1 = always
    int k4_pack = 2;       // VRAD_K4_PACK: 2 = gather from the block-row streams where neighbouring rows share their columns and the work items stay fine enough (else as 1), 3 / 4 = block rows with a warp / a thread block per work item whatever the item count (2 chooses), 1 = gather from the packed 6-byte streams where rows are (nearly) one segment each, 0 = from the {col,w} pairs; 9 = pack whatever the segment count
    int k4_l2_mb = 0;      // VRAD_K4_L2_MB: MB of L2 set aside for the head of the transfer stream of the multi-GPU gather (0 = off)
    int k4_hier_p2p = 1;   // VRAD_K4_HIER_P2P: patch hierarchy on several GPUs: leaf rows by peer stores (0 = all-gather pass per bounce)
    int k4_pool = 25;      // VRAD_K4_POOL: percent of the work left out of the persistent blocks' ranges for whoever finishes early
    int k4_items = 0;      // VRAD_K4_ITEMS: run the multi-GPU (work-item) gather on a single-GPU handle too (a one-rank peer table)
    int k4_persist = 1;    // VRAD_K4_PERSIST: one gather block per resident slot over equal-work item ranges (0 = 8 items per block)
    int k4_pdl = 1;        // VRAD_K4_PDL: chain the bounces of the multi-GPU gather with programmatic dependent launch
    int k4_graph = 1;      // VRAD_K4_GRAPH: replay the bounce loop as a CUDA graph
    int k4_sim_peers = 0;  // VRAD_K4_SIM_PEERS: single-GPU stand-in for the peers of a multi-GPU handle (timing only)
};
} // namespace vrad

struct vrad_env {
    vrad_config cfg{};
    vrad::EnvOptions opt{};
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    bool async = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.0f;
    int last_launches = 0;
    int sm_count = 148;

    // host geometry (Environment.OptimizedTriangleList before SetupAccelerationStructure)
    std::vector<int32_t> h_ids;
    std::vector<float>   h_verts;
    std::vector<uint8_t> h_flags;
    // host copies of the built tree in reference layout
    vrad::KdTree tree;
    std::vector<vrad_tri48> h_tris;
    double build_seconds = 0.0;
    bool built = false;

    // device scene
    vrad::DevBuf<int2> d_nodes;
    vrad::DevBuf<int32_t> d_tri_index;
    vrad::DevBuf<float4> d_q0, d_q1, d_q2;
    vrad::DevBuf<float> d_tri_cov;     // colour.X per triangle (coverage of transparent triangles)
    vrad::DevBuf<int2> d_top;          // top tree levels in breadth-first order (DevScene::top)
    vrad::DevBuf<float4> d_points;     // resident endpoint table of vrad_test_lines_indexed (xyz, pad)
    int64_t n_points = 0;
    std::vector<float> h_colors;       // Environment.TriangleColors, 3 per triangle
    vrad::DevScene scene{};

    // BSP lumps + sky cameras (f2)
    vrad::DevBuf<int4> d_bsp_nodes;
    vrad::DevBuf<float4> d_bsp_planes, d_cams;
    vrad::DevBuf<int32_t> d_leaf_cluster, d_leaf_area, d_area_camera;
    vrad::DevBsp bsp{};
    bool bsp_ready = false;
    std::vector<int32_t> h_cam_area, h_area_camera;
    std::vector<float> h_cam_w2s;

    // second stream + double-buffered staging for the pipelined host-buffer paths
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    vrad::DevBuf<float> d_stage[2];
    uint32_t* h_one = nullptr;         // pinned word holding 1: the source of the chunk-arrival flag copies

    // scratch for staging host pointers
    std::vector<vrad::DevBuf<unsigned char>> scratch;

    vrad::PatchesDev patches;
    vrad::TransfersDev transfers;
    vrad::DevBuf<float> d_sky_dirs; int n_sky_dirs = 0;
    int light_trace_flags = 0;         // VRAD_TL_CAN_RECURSE | VRAD_TL_TEXTURE_SHADOWS for the light rays of K3

    // bounce state
    vrad::DevBuf<float4> d_er[2];      // emit*refl (rgb, pad), full N (padded to world*rows_per_rank)
    vrad::DevBuf<float4> d_total;      // accumulated bounced light for local rows
    vrad::DevBuf<float>  d_partials;   // per-block partial sums of `added`
    vrad::DevBuf<float4> d_add;        // light added per local row by the last gather (deterministic `added` reduction)
    void* nccl_comm = nullptr;         // ncclComm_t
    vrad::LocalGroup* group = nullptr; // child of an in-process multi-GPU handle: its collectives go through the group
    vrad::LocalGroup* multi = nullptr; // the in-process multi-GPU handle itself (owns the group and its children; no device state of its own)
    vrad::PeerLinks peers;             // K4 fused exchange over NVLink peer memory
    vrad::GraphCache bounce_graph;
    size_t l2_set_aside_req = 0, l2_set_aside = 0, l2_window_max = 0, l2_window_bytes = 0;   // L2 residency of the transfer stream (k4_l2_mb)
    float l2_hit_ratio = 0.0f;
    int64_t bounds_serial = -1;        // rows_serial the cached block boundaries belong to
    int64_t bounds[vrad::kMaxWorld + 1] = {};
};

namespace vrad {

bool is_device_ptr(const void* p);
// returns a device pointer for `p` (n bytes); host memory is copied H2D on e->stream into scratch slot `slot`
int stage_in(vrad_env* e, int slot, const void* p, size_t bytes, const void** dev_out, bool* was_host);
// returns a device pointer to write results into; if `p` is host memory a scratch buffer is used
int stage_out(vrad_env* e, int slot, void* p, size_t bytes, void** dev_out, bool* was_host);
int finish_out(vrad_env* e, void* host_p, const void* dev_p, size_t bytes, bool was_host);
// plain device scratch of `bytes` in slot `slot` (grown on demand, owned by the handle)
int scratch_get(vrad_env* e, int slot, size_t bytes, void** out);
void timing_begin(vrad_env* e);
void timing_end(vrad_env* e, int launches);
int sync_if_needed(vrad_env* e, bool any_host);
inline bool has_comm(const vrad_env* e) { return e->nccl_comm != nullptr || e->group != nullptr; }

// kernel launchers implemented in the k*.cu files
int launch_trace_rays(vrad_env* e, int64_t n, const float* ox, const float* oy, const float* oz, const float* dx,
                      const float* dy, const float* dz, const float* tmin, const float* tmax, int32_t skip_id,
                      int32_t* hit_tri, int32_t* hit_sid, float* hit_t, float* normal_soa);
int launch_test_lines(vrad_env* e, int64_t n, const float* start_soa, const float* stop_soa, int sky_mode, uint32_t* bits);
int launch_test_lines_pipelined(vrad_env* e, int64_t n, const float* h_a, const float* h_b, int64_t host_stride, const int32_t* h_pairs, int sky_mode, uint32_t* d_bits);
int launch_test_lines_indexed(vrad_env* e, int64_t n, const int32_t* pairs, int sky_mode, uint32_t* bits);
int check_pairs_on_device(vrad_env* e, int64_t n, const int32_t* d_pairs, int* bad_out);
int read_bad_index_count(vrad_env* e, int* bad_out);
int upload_triangle_coverage(vrad_env* e);
struct KdTree;
int build_kd_tree_binned_device(cudaStream_t stream, const float* verts9, int n, KdTree& out, int* launches, const char** why);

} // namespace vrad

// in-process multi-GPU handle (group.cu): a handle with e->multi set owns one child per device
namespace vrad {
int group_destroy(vrad_env* g);
int group_set_option(vrad_env* g, const char* name, int value);
int group_add_triangles(vrad_env* g, int n, const int32_t* ids, const float* verts9, const uint8_t* flags);
int group_build(vrad_env* g, int fast, int where);
int group_upload_tree(vrad_env* g, int n_nodes, const int32_t* children, const float* split, int n_idx, const int32_t* tri_index, int n_tris, const vrad_tri48* tris, const float aabb[6]);
int group_set_triangle_colors(vrad_env* g, int n, const float* rgb3);
int group_points_upload(vrad_env* g, int64_t n, const float* xyz3);
int group_set_sky_dirs(vrad_env* g, int n, const float* dirs3);
int group_set_light_trace_flags(vrad_env* g, int flags);
int group_last_timing(vrad_env* g, float* ms, int* launches);
int group_test_lines(vrad_env* g, int64_t n, const float* a, const float* b, int sky_mode, uint32_t* bits);
int group_test_lines_indexed(vrad_env* g, int64_t n, const int32_t* pairs2, int sky_mode, uint32_t* bits);
int group_trace_rays(vrad_env* g, int64_t n, const float* ox, const float* oy, const float* oz, const float* dx, const float* dy, const float* dz,
                     const float* tmin, const float* tmax, int32_t skip_id, int32_t* hit_tri, int32_t* hit_sid, float* hit_t);
int group_patches_upload(vrad_env* g, int n, const float* origin3, const float* normal3, const float* plane_dist, const float* area, const float* reflectivity3, const int32_t* cluster, const uint8_t* flags);
int group_set_hierarchy(vrad_env* g, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face);
int group_build_transfers(vrad_env* g, int n_clusters, const uint8_t* pvs, int64_t* nnz_out);
int group_transfers_info(vrad_env* g, int64_t* row0, int64_t* row1, int64_t* nnz);
int group_transfers_download(vrad_env* g, int64_t* rowptr, int32_t* col, float* w);
int group_direct_light(vrad_env* g, int64_t n, const float* pos3, const float* normal3, int n_lights, const vrad_light* lights, float* rgb_out);
int group_set_bump(vrad_env* g, int n, const uint8_t* needs_bump, const float* bump_normals9);
int group_set_windings(vrad_env* g, int n, const int32_t* first, const int32_t* count, int n_points, const float* points3);
int group_bounce(vrad_env* g, const float* emit0_rgb, int n_bounces, int early_out, float* total_rgb_out, float added_last[3], int* bounces_done);
inline int group_unsupported(const char* what) { set_error("%s is not available on a multi-GPU handle (vrad_env_create_multi)", what); return VRAD_E_UNSUPPORTED; }
}
// a multi-GPU handle: forward to the group; queries that any rank can answer go to rank 0
#define VRAD_MULTI(e, call) do { if ((e) && (e)->multi) return vrad::call; } while (0)
#define VRAD_MULTI_RANK0(e) do { if ((e) && (e)->multi) { (e) = (e)->multi->ranks[0]; cudaSetDevice((e)->cfg.device); } } while (0)
#define VRAD_MULTI_UNSUPPORTED(e, name) do { if ((e) && (e)->multi) return vrad::group_unsupported(name); } while (0)
