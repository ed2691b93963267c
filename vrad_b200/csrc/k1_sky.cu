// k1_sky.cu -- the complete trace.TestLineDoesHitSky on the device, plus the BSP
// point-location kernels and sky-camera bookkeeping it needs.
//
// Reference map
//   raytracer/trace/testline.go:18-94     TestLineDoesHitSky: skip id (:36), sky-id occlusion rule (:42-51),
//                                         coverage (:52-55), 3D-skybox recursion (:57-89), clamp (:91-93)
//   raytracer/trace/pointleaf.go:8-33     PointLeafnum
//   rad/clustertable/point.go:10-38       ClusterFromPoint / PointInLeaf
//   rad/cameras/skycamera.go:10-49        ProcessSkyCameras; cache/skycameras.go:8-40 camera tables
//   raytracer/types/coverageCount.go:16-48  CoverageCount (device form: common.cuh `Coverage`)
//   rad/lightmap/lightmap.go:425-451      CanLeafTraceToSky
// Arithmetic order follows oracle/skytrace.cpp statement for statement (results are bit-identical).
//
// Mapping: one segment per lane traced to completion (the recursion pass re-traces only the segments that
// are still not fully occluded and whose leaf's area has no sky camera -- a small, spatially coherent
// subset).  HBM traffic: 24 B in + 4 B out per segment for the primary pass; the recursion pass re-reads
// 24 B + 4 B and rewrites 4 B for the segments of warps that recurse.
#include "env_internal.cuh"
#include "sky.cuh"

namespace vrad {

constexpr int kSkyBlock = 128;

__global__ void __launch_bounds__(256)
k_point_leafnum(DevBsp B, int64_t n, const float* __restrict__ pts, int32_t* __restrict__ out, int want_cluster) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (want_cluster) out[i] = __ldg(&B.leaf_cluster[point_in_leaf(B, x, y, z)]);
    else out[i] = point_leafnum(B, x, y, z);
}

// pass 1: occlusion of every segment (final fraction when no recursion pass follows)
template <bool COVER>
__global__ void __launch_bounds__(kSkyBlock)
k1_sky_primary(DevScene S, int64_t n, const float* __restrict__ a, const float* __restrict__ b, int skip_id,
               int finish, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int64_t j = valid ? i : 0;
    bool degenerate;
    const float occ = primary_occlusion<COVER>(S, valid, __ldcs(&a[j]), __ldcs(&a[n + j]), __ldcs(&a[2 * n + j]),
                                               __ldcs(&b[j]), __ldcs(&b[n + j]), __ldcs(&b[2 * n + j]), skip_id, degenerate);
    if (valid) out[i] = finish ? finish_fraction(occ) : occ;
}

// pass 2: 3D-skybox recursion (testline.go:57-89) + the final clamp.  io holds the occlusion from pass 1.
template <bool COVER>
__global__ void __launch_bounds__(kSkyBlock)
k1_sky_recurse(DevScene S, DevBsp B, int64_t n, const float* __restrict__ a, const float* __restrict__ b,
               int skip_id, int packet_leaf, float* __restrict__ io) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int64_t j = valid ? i : 0;
    float occ = valid ? io[j] : 1.0f;
    const float ax = a[j], ay = a[n + j], az = a[2 * n + j];
    float dx = b[j] - ax, dy = b[n + j] - ay, dz = b[2 * n + j] - az;
    const bool degenerate = (((dx * dx) + (dy * dy)) + (dz * dz)) == 0.0f;       // zero-length: visible, no recursion
    bool recurse = valid && !degenerate && occ < 1.0f;
    if (recurse) {
        const int64_t li = packet_leaf ? (j & ~(int64_t)3) : j;                  // start.Vec(0), testline.go:63
        const int leaf = point_leafnum(B, a[li], a[n + li], a[2 * n + li]);
        recurse = false;
        if (leaf >= 0 && leaf < B.n_leafs) {
            const int area = __ldg(&B.leaf_area[leaf]);                          // :65-66
            if (area >= 0 && area < B.n_areas) recurse = __ldg(&B.area_camera[area]) < 0;   // :67
        }
    }
    if (__any_sync(0xffffffffu, recurse)) {
        float magsq = dx * dx;                                                   // fourvectors.go:71-78
        magsq = (dy * dy) + magsq;
        magsq = (dz * dz) + magsq;
        const float rs = (float)(1.0 / sqrt((double)magsq));                     // simd.go:162-169
        dx = dx * rs; dy = dy * rs; dz = dz * rs;
        for (int c = 0; c < B.n_cams; c++) {                                     // :69
            const float4 cam = __ldg(&B.cams[c]);
            const float sx = cam.x + (ax * cam.w), sy = cam.y + (ay * cam.w), sz = cam.z + (az * cam.w);   // :73-76
            const float ex = (dx * kMaxTraceLength) + sx, ey = (dy * kMaxTraceLength) + sy, ez = (dz * kMaxTraceLength) + sz;   // :78-80
            bool deg2;
            const float occ2 = primary_occlusion<COVER>(S, recurse, sx, sy, sz, ex, ey, ez, skip_id, deg2);  // :81, canRecurse=false
            if (recurse) {
                const float fv2 = deg2 ? 1.0f : finish_fraction(occ2);
                occ = occ + 1.0f;                                                // :82
                occ = occ - fv2;                                                 // :83
            }
        }
    }
    if (valid) io[i] = finish_fraction(occ);
}

// CanLeafTraceToSky front end (lightmap.go:428-441): segments leaf centre -> centre - dir*MAX_TRACE_LENGTH
__global__ void __launch_bounds__(256)
k_leaf_sky_segments(int n_leafs, int n_dirs, const int16_t* __restrict__ mins, const int16_t* __restrict__ maxs,
                    const float* __restrict__ dirs, float* __restrict__ a, float* __restrict__ b) {
    const int64_t n = (int64_t)n_leafs * n_dirs;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int leaf = (int)(i / n_dirs), d = (int)(i % n_dirs);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float c = (float)((int)mins[3 * leaf + k] + (int)maxs[3 * leaf + k]) * 0.5f;
        a[k * n + i] = c;
        b[k * n + i] = (dirs[3 * d + k] * (-kMaxTraceLength)) + c;
    }
}

// lightmap.go:445-447: any direction with fractionVisible > 0; one warp per leaf
__global__ void __launch_bounds__(256)
k_leaf_sky_reduce(int n_leafs, int n_dirs, const float* __restrict__ fv, uint8_t* __restrict__ out) {
    const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= n_leafs) return;
    bool any = false;
    for (int d = lane; d < n_dirs; d += 32) any |= fv[(int64_t)warp * n_dirs + d] > 0.0f;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) out[warp] = any ? 1 : 0;
}

static int launch_sky(vrad_env* e, int64_t n, const float* d_a, const float* d_b, int flags, int32_t skip_id, float* d_out, int* launches) {
    const bool cover = (flags & VRAD_TL_TEXTURE_SHADOWS) != 0;
    const bool recurse = (flags & VRAD_TL_CAN_RECURSE) && e->bsp_ready && e->bsp.n_cams > 0;
    const int grid = (int)((n + kSkyBlock - 1) / kSkyBlock);
    if (cover) k1_sky_primary<true><<<grid, kSkyBlock, 0, e->stream>>>(e->scene, n, d_a, d_b, skip_id, recurse ? 0 : 1, d_out);
    else k1_sky_primary<false><<<grid, kSkyBlock, 0, e->stream>>>(e->scene, n, d_a, d_b, skip_id, recurse ? 0 : 1, d_out);
    *launches += 1;
    if (recurse) {
        const int pl = (flags & VRAD_TL_PACKET_LEAF) ? 1 : 0;
        if (cover) k1_sky_recurse<true><<<grid, kSkyBlock, 0, e->stream>>>(e->scene, e->bsp, n, d_a, d_b, skip_id, pl, d_out);
        else k1_sky_recurse<false><<<grid, kSkyBlock, 0, e->stream>>>(e->scene, e->bsp, n, d_a, d_b, skip_id, pl, d_out);
        *launches += 1;
    }
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

static int launch_points(vrad_env* e, int64_t n, const float* pts3, int32_t* out, int want_cluster, const char* who) {
    if (!e || n < 0 || (n > 0 && (!pts3 || !out))) { set_error("%s: bad arguments", who); return VRAD_E_INVALID; }
    if (!e->bsp_ready) { set_error("%s: no BSP uploaded (vrad_bsp_upload)", who); return VRAD_E_STATE; }
    if (n == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const void* d_p; void* d_o; bool hp, ho;
    int rc;
    if ((rc = stage_in(e, 0, pts3, (size_t)n * 12, &d_p, &hp))) return rc;
    if ((rc = stage_out(e, 1, out, (size_t)n * 4, &d_o, &ho))) return rc;
    timing_begin(e);
    k_point_leafnum<<<(int)((n + 255) / 256), 256, 0, e->stream>>>(e->bsp, n, (const float*)d_p, (int32_t*)d_o, want_cluster);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    if ((rc = finish_out(e, out, d_o, (size_t)n * 4, ho))) return rc;
    return sync_if_needed(e, hp | ho);
}

} // namespace vrad

using namespace vrad;

extern "C" {

int vrad_bsp_upload(vrad_env* e, int n_nodes, const int32_t* node_plane, const int32_t* node_children2, int n_planes,
                    const float* plane_normal3, const float* plane_dist, const int32_t* plane_type, int n_leafs,
                    const int32_t* leaf_cluster, const int32_t* leaf_area, int n_areas) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_bsp_upload");
    if (!e || n_nodes < 0 || n_planes < 0 || n_leafs < 1 || n_areas < 0 || !leaf_cluster || !leaf_area ||
        (n_nodes > 0 && (!node_plane || !node_children2)) || (n_planes > 0 && (!plane_normal3 || !plane_dist || !plane_type))) {
        set_error("vrad_bsp_upload: bad arguments"); return VRAD_E_INVALID;
    }
    for (int i = 0; i < n_nodes; i++) {
        if (node_plane[i] < 0 || node_plane[i] >= n_planes) { set_error("vrad_bsp_upload: node %d has plane %d outside [0,%d)", i, node_plane[i], n_planes); return VRAD_E_INVALID; }
        for (int c = 0; c < 2; c++) {
            const int ch = node_children2[2 * i + c];
            // children of a BSP lump are written after their parent; requiring it rules out cycles in PointLeafnum's loop
            if (ch >= n_nodes || -1 - ch >= n_leafs || (ch >= 0 && ch <= i)) { set_error("vrad_bsp_upload: node %d has bad child %d", i, ch); return VRAD_E_INVALID; }
        }
    }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    std::vector<int4> nodes(n_nodes);
    for (int i = 0; i < n_nodes; i++) nodes[i] = make_int4(node_plane[i], node_children2[2 * i], node_children2[2 * i + 1], plane_type[node_plane[i]]);
    std::vector<float4> planes(n_planes);
    for (int i = 0; i < n_planes; i++) planes[i] = make_float4(plane_normal3[3 * i], plane_normal3[3 * i + 1], plane_normal3[3 * i + 2], plane_dist[i]);
    std::vector<int32_t> area_camera(n_areas > 0 ? n_areas : 1, -1);
    if (e->d_bsp_nodes.alloc(n_nodes ? n_nodes : 1) || e->d_bsp_planes.alloc(n_planes ? n_planes : 1) || e->d_leaf_cluster.alloc(n_leafs) ||
        e->d_leaf_area.alloc(n_leafs) || e->d_area_camera.alloc(area_camera.size()) || e->d_cams.alloc(1)) {
        set_error("out of device memory for the BSP lumps"); return VRAD_E_NOMEM;
    }
    if (n_nodes) VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_bsp_nodes.p, nodes.data(), (size_t)n_nodes * sizeof(int4), cudaMemcpyHostToDevice, e->stream));
    if (n_planes) VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_bsp_planes.p, planes.data(), (size_t)n_planes * sizeof(float4), cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_leaf_cluster.p, leaf_cluster, (size_t)n_leafs * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_leaf_area.p, leaf_area, (size_t)n_leafs * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaMemcpyAsync(e->d_area_camera.p, area_camera.data(), area_camera.size() * 4, cudaMemcpyHostToDevice, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    DevBsp& B = e->bsp;
    B.nodes = e->d_bsp_nodes.p; B.planes = e->d_bsp_planes.p; B.leaf_cluster = e->d_leaf_cluster.p; B.leaf_area = e->d_leaf_area.p;
    B.area_camera = e->d_area_camera.p; B.cams = e->d_cams.p;
    B.n_nodes = n_nodes; B.n_leafs = n_leafs; B.n_areas = n_areas; B.n_cams = 0;
    e->h_cam_area.clear(); e->h_cam_w2s.clear();
    e->h_area_camera.assign(n_areas, -1);
    e->bsp_ready = true;
    return VRAD_OK;
}

int vrad_point_leafnum(vrad_env* e, int64_t n, const float* pts3, int32_t* leaf_out) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_point_leafnum");
    return launch_points(e, n, pts3, leaf_out, 0, "vrad_point_leafnum");
}

int vrad_cluster_from_point(vrad_env* e, int64_t n, const float* pts3, int32_t* cluster_out) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_cluster_from_point");
    return launch_points(e, n, pts3, cluster_out, 1, "vrad_cluster_from_point");
}

int vrad_sky_cameras_set(vrad_env* e, int n, const float* origin3, const float* scale, int* n_kept_out) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_sky_cameras_set");
    if (!e || n < 0 || (n > 0 && (!origin3 || !scale))) { set_error("vrad_sky_cameras_set: bad arguments"); return VRAD_E_INVALID; }
    if (!e->bsp_ready) { set_error("vrad_sky_cameras_set: no BSP uploaded (vrad_bsp_upload)"); return VRAD_E_STATE; }
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    std::vector<int32_t> leaf(n > 0 ? n : 1, -1);
    if (n > 0) {   // PointLeafnum(&origin), skycamera.go:27 -- on the device like every other point query
        std::vector<float> pts(origin3, origin3 + 3 * (size_t)n);
        int rc = launch_points(e, n, pts.data(), leaf.data(), 0, "vrad_sky_cameras_set");
        if (rc) return rc;
        VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    std::vector<int32_t> leaf_area(e->bsp.n_leafs);
    VRAD_CUDA_CHECK(cudaMemcpy(leaf_area.data(), e->d_leaf_area.p, (size_t)e->bsp.n_leafs * 4, cudaMemcpyDeviceToHost));
    std::vector<float4> cams;
    e->h_cam_area.clear(); e->h_cam_w2s.clear();
    e->h_area_camera.assign(e->bsp.n_areas, -1);                                   // :12-14
    for (int i = 0; i < n; i++) {
        int area = -1;
        if (leaf[i] >= 0 && leaf[i] < e->bsp.n_leafs) area = leaf_area[leaf[i]];   // :30-32
        if (scale[i] > 0.0f) {                                                     // :35
            const float w2s = 1.0f / scale[i];                                     // :38
            cams.push_back(make_float4(origin3[3 * i], origin3[3 * i + 1], origin3[3 * i + 2], w2s));
            e->h_cam_area.push_back(area); e->h_cam_w2s.push_back(w2s);
            if (area >= 0 && area < e->bsp.n_areas) e->h_area_camera[area] = (int)cams.size() - 1;   // :41-43
        }
    }
    if (e->d_cams.alloc(cams.size() ? cams.size() : 1)) { set_error("out of device memory for sky cameras"); return VRAD_E_NOMEM; }
    if (!cams.empty()) VRAD_CUDA_CHECK(cudaMemcpy(e->d_cams.p, cams.data(), cams.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if (e->bsp.n_areas > 0) VRAD_CUDA_CHECK(cudaMemcpy(e->d_area_camera.p, e->h_area_camera.data(), (size_t)e->bsp.n_areas * 4, cudaMemcpyHostToDevice));
    e->bsp.cams = e->d_cams.p;
    e->bsp.n_cams = (int)cams.size();
    if (n_kept_out) *n_kept_out = (int)cams.size();
    return VRAD_OK;
}

int vrad_sky_cameras_get(vrad_env* e, int* n_cameras, int32_t* cam_area, float* world_to_sky, int32_t* area_camera) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_sky_cameras_get");
    if (!e) return VRAD_E_INVALID;
    if (!e->bsp_ready) { set_error("vrad_sky_cameras_get: no BSP uploaded"); return VRAD_E_STATE; }
    if (n_cameras) *n_cameras = (int)e->h_cam_area.size();
    for (size_t i = 0; i < e->h_cam_area.size(); i++) {
        if (cam_area) cam_area[i] = e->h_cam_area[i];
        if (world_to_sky) world_to_sky[i] = e->h_cam_w2s[i];
    }
    if (area_camera) for (size_t i = 0; i < e->h_area_camera.size(); i++) area_camera[i] = e->h_area_camera[i];
    return VRAD_OK;
}

int vrad_test_lines_sky(vrad_env* e, int64_t n, const float* start_xyz_soa, const float* stop_xyz_soa, int flags,
                        int32_t static_prop_to_skip, float* fraction_visible) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_test_lines_sky");
    if (!e || n < 0 || (n > 0 && (!start_xyz_soa || !stop_xyz_soa || !fraction_visible))) { set_error("vrad_test_lines_sky: bad arguments"); return VRAD_E_INVALID; }
    if (!e->built) { set_error("vrad_test_lines_sky: acceleration structure not built"); return VRAD_E_STATE; }
    if (n == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const size_t b = (size_t)n * 12;
    const void *d_a, *d_b; void* d_o; bool ha, hb, ho;
    int rc;
    if ((rc = stage_in(e, 0, start_xyz_soa, b, &d_a, &ha))) return rc;
    if ((rc = stage_in(e, 1, stop_xyz_soa, b, &d_b, &hb))) return rc;
    if ((rc = stage_out(e, 2, fraction_visible, (size_t)n * 4, &d_o, &ho))) return rc;
    const int32_t skip_id = VRAD_TRACE_ID_STATICPROP | static_prop_to_skip;      // testline.go:36
    int launches = 0;
    timing_begin(e);
    rc = launch_sky(e, n, (const float*)d_a, (const float*)d_b, flags, skip_id, (float*)d_o, &launches);
    timing_end(e, launches);
    if (rc) return rc;
    if ((rc = finish_out(e, fraction_visible, d_o, (size_t)n * 4, ho))) return rc;
    return sync_if_needed(e, ha | hb | ho);
}

int vrad_leafs_trace_to_sky(vrad_env* e, int n_leafs, const int16_t* mins3, const int16_t* maxs3, uint8_t* can_out) {
    VRAD_MULTI_UNSUPPORTED(e, "vrad_leafs_trace_to_sky");
    if (!e || n_leafs < 0 || (n_leafs > 0 && (!mins3 || !maxs3 || !can_out))) { set_error("vrad_leafs_trace_to_sky: bad arguments"); return VRAD_E_INVALID; }
    if (!e->built) { set_error("vrad_leafs_trace_to_sky: acceleration structure not built"); return VRAD_E_STATE; }
    if (e->n_sky_dirs <= 0) { set_error("vrad_leafs_trace_to_sky: no sky directions set (vrad_set_sky_dirs)"); return VRAD_E_STATE; }
    if (n_leafs == 0) return VRAD_OK;
    VRAD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const int nd = e->n_sky_dirs;
    const int64_t n = (int64_t)n_leafs * nd;
    const void *d_mins, *d_maxs; void *d_seg, *d_fv, *d_o; bool h0, h1, ho;
    int rc;
    if ((rc = stage_in(e, 0, mins3, (size_t)n_leafs * 6, &d_mins, &h0))) return rc;
    if ((rc = stage_in(e, 1, maxs3, (size_t)n_leafs * 6, &d_maxs, &h1))) return rc;
    if ((rc = stage_out(e, 2, can_out, (size_t)n_leafs, &d_o, &ho))) return rc;
    if ((rc = scratch_get(e, 3, (size_t)n * 24, &d_seg))) return rc;
    if ((rc = scratch_get(e, 4, (size_t)n * 4, &d_fv))) return rc;
    float* a = (float*)d_seg; float* b = a + 3 * n;
    int launches = 0;
    timing_begin(e);
    k_leaf_sky_segments<<<(int)((n + 255) / 256), 256, 0, e->stream>>>(n_leafs, nd, (const int16_t*)d_mins, (const int16_t*)d_maxs, e->d_sky_dirs.p, a, b);
    launches++;
    rc = launch_sky(e, n, a, b, VRAD_TL_CAN_RECURSE, VRAD_TRACE_ID_STATICPROP | -1, (float*)d_fv, &launches);   // lightmap.go:444
    if (!rc) {
        k_leaf_sky_reduce<<<(int)(((int64_t)n_leafs * 32 + 255) / 256), 256, 0, e->stream>>>(n_leafs, nd, (const float*)d_fv, (uint8_t*)d_o);
        launches++;
    }
    timing_end(e, launches);
    if (rc) return rc;
    VRAD_CUDA_CHECK(cudaGetLastError());
    if ((rc = finish_out(e, can_out, d_o, (size_t)n_leafs, ho))) return rc;
    return sync_if_needed(e, h0 | h1 | ho);
}

// ---- host-only PVS helpers (rad/lightmap/vis.go) -------------------------------------------

// DecompressVis (vis.go:54-94; SURVEY App. A #23: the literal loop cannot terminate correctly, the
// intent is the standard run-length code: a non-zero byte is copied, a zero byte is followed by the
// number of zero bytes it stands for)
int64_t vrad_decompress_vis(const uint8_t* in, int64_t in_len, int n_clusters, uint8_t* out_row) {
    if (!in || !out_row || n_clusters < 0 || in_len < 0) { set_error("vrad_decompress_vis: bad arguments"); return VRAD_E_INVALID; }
    const int row = (n_clusters + 7) >> 3;                                         // :61
    int64_t ip = 0;
    int op = 0;
    while (op < row) {
        if (ip >= in_len) { set_error("vrad_decompress_vis: input ends after %d of %d bytes", op, row); return VRAD_E_INVALID; }
        if (in[ip]) { out_row[op++] = in[ip++]; continue; }                        // :70-75
        if (ip + 1 >= in_len) { set_error("vrad_decompress_vis: truncated run"); return VRAD_E_INVALID; }
        int c = in[ip + 1];                                                        // :77
        if (c == 0) { set_error("DecompressVis: 0 repeat"); return VRAD_E_INVALID; }   // :78-80 (log.Fatalf)
        ip += 2;
        if (op + c > row) c = row - op;                                            // :83-86 overrun is clamped
        while (c-- > 0) out_row[op++] = 0;                                         // :88-92
    }
    return ip;
}

int vrad_pvs_from_vis_lump(int n_clusters, const int32_t* byteofs2, const uint8_t* visdata, int64_t vis_len, uint8_t* pvs_out) {
    if (n_clusters < 0 || (n_clusters > 0 && (!byteofs2 || !visdata || !pvs_out))) { set_error("vrad_pvs_from_vis_lump: bad arguments"); return VRAD_E_INVALID; }
    const int row = (n_clusters + 7) >> 3;
    std::vector<uint8_t> bits(row ? row : 1);
    for (int c = 0; c < n_clusters; c++) {
        uint8_t* out = pvs_out + (size_t)c * n_clusters;
        const int32_t ofs = byteofs2[2 * c];                                       // DVIS_PVS, vis.go:33
        if (ofs == -1) { memset(out, 0, n_clusters); continue; }                   // reference: log.Fatalf("visofs == -1") (:35-37)
        if (ofs < 0 || ofs >= vis_len) { set_error("vrad_pvs_from_vis_lump: cluster %d has offset %d outside the lump (%lld bytes)", c, ofs, (long long)vis_len); return VRAD_E_INVALID; }
        const int64_t used = vrad_decompress_vis(visdata + ofs, vis_len - ofs, n_clusters, bits.data());
        if (used < 0) return (int)used;
        for (int k = 0; k < n_clusters; k++) out[k] = (bits[k >> 3] >> (k & 7)) & 1;   // PVSCheck, lightmap.go:413-421
    }
    return VRAD_OK;
}

} // extern "C"
