/*
 * vrad_cuda.h -- C-ABI of libvradcuda.so: the B200 (sm_100a) drop-in for the data-parallel hot
 * path of VRADiant (Galaco/VRAD).  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * repository root).  The Go driver keeps cache/ BSP loading, rad/ entry points and the
 * raytracer/trace call surface and binds these symbols through cgo.
 *
 * Conventions
 *   - every function returns 0 on success, a negative vrad_status on failure;
 *     vrad_last_error() returns a thread-local message.  The reference has no error returns
 *     (failures are log.Fatal or log.Panicln, e.g. raytracer/trace/testline.go:21); the Go shim
 *     turns a non-zero status into log.Fatalf.
 *   - a handle is NOT thread-safe (the reference is single-goroutine: common/constants/constants.go:43).
 *   - one handle drives ONE GPU.  Multi-GPU = one handle per process/GPU with (rank, world) set in
 *     vrad_config; rays/luxels/patch rows are sharded by rank and the only data-path collective is
 *     the per-bounce radiance all-gather (vrad_comm_init).
 *   - data pointers may be pageable host, pinned host (vrad_host_alloc) or device memory; the
 *     library inspects them with cudaPointerGetAttributes.  Host pointers are copied in/out
 *     before the call returns (cgo rule: C must not retain Go pointers).
 *   - there is no CPU fallback: without a CUDA device vrad_env_create fails with VRAD_E_CUDA.
 */
#ifndef VRAD_CUDA_H
#define VRAD_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    VRAD_OK = 0,
    VRAD_E_INVALID = -1,     /* bad argument / call order */
    VRAD_E_CUDA = -2,        /* CUDA runtime error (message in vrad_last_error) */
    VRAD_E_NOMEM = -3,
    VRAD_E_STATE = -4,       /* e.g. trace before build */
    VRAD_E_COMM = -5,        /* NCCL error */
    VRAD_E_UNSUPPORTED = -6  /* e.g. a kd tree deeper than the device traversal stack */
} vrad_status;

/* raytracer/constants.go:5-18 */
#define VRAD_TRACE_ID_SKY        0x01000000
#define VRAD_TRACE_ID_OPAQUE     0x02000000
#define VRAD_TRACE_ID_STATICPROP 0x04000000
#define VRAD_KDNODE_LEAF 3
/* triangle flag (TriIntersectData.NFlags, raytracer/cache/triangle/triintersectdata.go:20): upstream
 * FCACHETRI_TRANSPARENT.  Such a triangle is an ordinary blocker unless the trace runs with texture
 * shadows (VRAD_TL_TEXTURE_SHADOWS), where it goes through the coverage rule of
 * raytracer/types/coverageCount.go:16-48 evaluated on the device. */
#define VRAD_TRI_TRANSPARENT 0x01

typedef struct vrad_env vrad_env;   /* opaque; stands in for raytracer.Environment (raytracer/environment.go:28-39) */

typedef struct {
    int device;      /* CUDA device ordinal driven by this handle */
    int rank;        /* shard index of this handle, 0..world-1 */
    int world;       /* number of cooperating handles (1 = single GPU) */
    int flags;       /* VRAD_CFG_* */
} vrad_config;
#define VRAD_CFG_DEFAULT 0

/* raytracer/cache/triangle/triintersectdata.go:3-22 -- 48-byte intersection record */
typedef struct {
    float nx, ny, nz, d;
    int32_t id;
    float e[6];
    uint8_t sel0, sel1, flags, unused;
} vrad_tri48;

/* common/types/light.go:10-44 + worldlight fields set in rad/lightmap/lights.go:71-81,216-256,259-341 */
typedef struct {
    int32_t type;            /* emittype: 0 surface, 1 point, 2 spotlight, 3 skylight, 5 skyambient */
    float origin[3];
    float intensity[3];
    float normal[3];
    float stopdot, stopdot2, exponent, radius;
    float constant_attn, linear_attn, quadratic_attn;
    float start_fade, end_fade, cap_dist;
    int32_t flags;
    float pad[3];
} vrad_light;               /* 96 bytes */

/* ---- lifetime --------------------------------------------------------------------------- */
/* raytracer.GetEnvironment / NewEnvironment (raytracer/environment.go:17-25,434-442) */
int  vrad_env_create(const vrad_config* cfg, vrad_env** out);
void vrad_env_destroy(vrad_env*);
const char* vrad_last_error(void);
/* run the library's work on the caller's CUDA stream (cudaStream_t as void*); NULL = own stream */
int  vrad_env_set_stream(vrad_env*, void* cuda_stream);
/* async != 0: calls whose data pointers are all device memory return without synchronising */
int  vrad_env_set_async(vrad_env*, int async);
/* tuning switches, by name (defaults come from the environment variables in parentheses):
 *   "k1_sort"       (VRAD_K1_SORT)      order segment batches by start cell / direction octant / end cell before tracing:
 *                                       -1 = batches of >= 65536 segments on scenes of at most 4 MB (default), 0 = never, 1 = always
 *   "k1_key"        (VRAD_K1_KEY)       layout of the 30-bit sort key: -1 by the shape of the scene box (default), 0 start-major Morton,
 *                                       1 six-dimensional Morton, 2 cubic start cells
 *   "k1_stream"     (VRAD_K1_STREAM)    unordered batches through the persistent streaming kernel (default 1; 0 = per-chunk kernels)
 *   "k1_top"        (VRAD_K1_TOP)       stage the top of the kd tree in shared memory: 0 = off (default), n = node budget
 *   "k4_seg"        (VRAD_K4_SEG)       entries per gather work item; longer transfer rows are split (default 16384)
 *   "k4_long_first" (VRAD_K4_ORDER=long) gather work items longest first
 *   "k4_persist"    (VRAD_K4_PERSIST)   gather grid = one block per resident slot over equal-work item ranges (default 1);
 *                                       0 = 8 items per block, as many blocks as that takes
 *   "k4_pool"       (VRAD_K4_POOL)      percent of the gather work kept out of the persistent blocks' ranges, for whoever
 *                                       finishes its range early (default 25)
 *   "k4_items"      (VRAD_K4_ITEMS)     run the multi-GPU (work-item) gather kernel on a single-GPU handle too (default 0)
 *   "k4_block"      (VRAD_K4_BLOCK)     threads per gather block: 256 (5 blocks per SM) or 192 (6 per SM, more registers);
 *                                       0 (default) = 256 with the packed streams, 192 with the pairs
 *   "k4_pack"       (VRAD_K4_PACK)      form the gather streams the transfers in: 2 (default) = block rows where neighbouring rows share
 *                                       their columns (4 rows per column list), else packed 6-byte entries, else the {col,w} pairs;
 *                                       1 = packed or pairs; 0 = pairs; 3 = block rows regardless of the work-item count; 9 = packed
 *                                       regardless of the segment count (vrad_transfers_layout reports what is in use)
 *   "k4_bk_rows"    (VRAD_K4_BK_ROWS)   rows per block of the block-row streams: 4 (default) or 2
 *   "k4_short"      (VRAD_K4_SHORT)     the short-row gather (8 lanes per row) on one GPU: -1 (default) = where rows average under
 *                                       400 transfers, 0 = never, 1 = always
 *   "k1_bpsm"       (VRAD_K1_BPSM)      resident blocks per SM of the streaming traversal kernel: 12 (default), 10, 8
 *   "k1_sort_bits"  (VRAD_K1_SORT_BITS) leading bits of the 30-bit order key that take part in the radix sort (default 30)
 *   "k2_stream"     (VRAD_K2_STREAM)    K2 pass A as per-warp ray queues with lane refill (default 0: measured slower)
 *   "k4_pdl"        (VRAD_K4_PDL)       multi-GPU gather: chain the bounces by programmatic dependent launch (default 1)
 *   "k4_graph"      (VRAD_K4_GRAPH)     replay the bounce loop as a CUDA graph (default 1)
 *   "k4_sim_peers"  (VRAD_K4_SIM_PEERS) world > 1 without a communicator: this device stands in for every peer -- timing
 *                                       of one rank's slice on one GPU; the light is NOT valid (default 0)
 * Results do not depend on any of them (K4: within its 1e-4 tolerance). */
int  vrad_env_set_option(vrad_env*, const char* name, int value);
/* device time (ms, CUDA events on the launching stream) and kernel-launch count of the last call */
int  vrad_env_last_timing(vrad_env*, float* kernel_ms, int* n_launches);
/* pinned staging buffers for the batched calls */
void* vrad_host_alloc(size_t bytes);
void  vrad_host_free(void* p);

/* ---- geometry + acceleration structure -------------------------------------------------- */
/* Environment.AddTriangleWithMaterial (raytracer/environment.go:45-69): ids int32, 9 floats per triangle, flags */
int  vrad_env_add_triangles(vrad_env*, int n, const int32_t* ids, const float* verts9, const uint8_t* flags);
/* Environment.SetupAccelerationStructure (raytracer/environment.go:119-138): host SAH build
 * (RefineNode :238-387, CalculateCostsOfSplit :181-236), ChangeIntoIntersectionFormat
 * (raytracer/cache/optimisedtriangle.go:30-78), upload */
int  vrad_env_build(vrad_env*);
/* RTE_FLAGS_FAST_TREE_GENERATION (raytracer/constants.go:5 -- declared by the reference, read nowhere): the binned-SAH builder.
 * Same cost model, leaf rules, depth limit and packed output as vrad_env_build, split candidates from 32 bins per axis of the
 * node's box instead of every triSkip-th vertex; built level by level on the device (where = VRAD_BUILD_ON_DEVICE: one thread per
 * triangle reference / per node, integer atomics and stable scans, so the tree does not depend on scheduling) or with the same
 * code on the host's cores (VRAD_BUILD_ON_HOST).  The tree differs from vrad_env_build's; closest hits and visibility do not
 * (ties resolve by triangle index), except for rays that graze a triangle edge lying in a split plane, where any kd tracer's answer
 * depends on the planes.  The finished tree is validated like an uploaded one before any kernel walks it. */
#define VRAD_BUILD_ON_DEVICE 0
#define VRAD_BUILD_ON_HOST   1
#define VRAD_BUILD_AUTO      2   /* the device from 10,000 triangles up, the host's cores below (where the device build is all launch latency) */
int  vrad_env_build_fast(vrad_env*, int where);
/* the same builder without an environment (host-only helper; no device needed): arrays in reference layout.  children / split /
 * tri_index may be NULL to size the buffers (*n_nodes, *n_idx); aabb and max_depth may be NULL. */
int  vrad_kd_build_binned_host(int n, const float* verts9, int max_nodes, int max_idx, int32_t* children, float* split, int32_t* tri_index,
                               int* n_nodes, int* n_idx, float aabb[6], int* max_depth);
/* adopt a tree built elsewhere, in the reference's own layouts (OptimisedKDNode, TriangleIndexList, TriIntersectData) */
int  vrad_env_upload_tree(vrad_env*, int n_nodes, const int32_t* children, const float* split,
                          int n_idx, const int32_t* tri_index, int n_tris, const vrad_tri48* tris,
                          const float aabb[6]);
int  vrad_env_stats(vrad_env*, int* n_nodes, int* n_idx, int* n_tris, int* max_depth, int* n_leaves,
                    float aabb[6], double* build_seconds);
/* read the tree back in reference layout (Environment.OptimizedKDTree / TriangleIndexList /
 * OptimizedTriangleList, raytracer/environment.go:33-35; GetTriangle :422-424) */
int  vrad_env_download_tree(vrad_env*, int32_t* children, float* split, int32_t* tri_index, vrad_tri48* tris);

/* ---- K1: ray casting --------------------------------------------------------------------- */
/* Environment.Trace4Rays (raytracer/environment.go:140-145): one FourRays packet
 * (raytracer/types/fourrays.go:8-11) -> RayTracingResult (raytracer/types/result.go:8-12).
 * origin/dir/normal are x[4] y[4] z[4].  Latency path; kept for call-surface fidelity.
 * (callback = nil: transparent triangles block like any other, as upstream does without a callback.) */
int  vrad_trace4(vrad_env*, const float origin_xyz4[12], const float dir_xyz4[12], const float tmin[4],
                 const float tmax[4], int32_t skip_id, int32_t hit_ids[4], float hit_dist[4], float normal_xyz4[12]);
/* batched Trace4Rays: n rays SoA.  tmin may be NULL (0).  hit_tri = index into the triangle list or -1,
 * hit_sid = NTriangleID of that triangle (or -1), hit_t = distance (1e23 on a miss).  Outputs may be NULL.
 * A handle traces exactly the n rays it is given; with several GPUs the caller gives each rank its own range. */
int  vrad_trace_rays(vrad_env*, int64_t n, const float* ox, const float* oy, const float* oz,
                     const float* dx, const float* dy, const float* dz, const float* tmin, const float* tmax,
                     int32_t skip_id, int32_t* hit_tri, int32_t* hit_sid, float* hit_t);
/* trace.TestLine / trace.TestLineDoesHitSky (raytracer/trace/testline.go:18-94, without the 3D-skybox
 * recursion of :57-89 -- vrad_test_lines_sky is the complete form): n segments, SoA blocks x[n] y[n] z[n]; vis_bits bit (i%32) of word i/32, 1 = visible.
 * sky_mode 0: any hit occludes; 1: a nearest hit on a TRACE_ID_SKY triangle does not occlude (:46-48). */
int  vrad_test_lines(vrad_env*, int64_t n, const float* start_xyz_soa, const float* stop_xyz_soa,
                     int sky_mode, uint32_t* vis_bits);

/* The same test for segments given as pairs of indices into a point table that is resident on the device -- how the
 * lighting stages name their shadow segments: patch origin -> light origin, patch -> patch, leaf centre -> sky sample
 * (rad/lightmap/lightmap.go:435-447 builds its FourVectors from such a fixed start and a table of directions; rad/patches and
 * rad/lightmap hold the patch origins and light origins the segments run between).  vrad_points_upload copies n_points
 * xyz triples once (patch origins, light origins, luxel samples ...); vrad_test_lines_indexed then takes 8 bytes per segment
 * ({int32 start, int32 stop}, interleaved) instead of 24: a host batch is bound by PCIe at 24 B/segment, not by the kernel.
 * Results are identical to vrad_test_lines on the same coordinates.  An index outside the table is VRAD_E_INVALID (checked
 * on the host for host buffers; device buffers are checked by a device pass unless the handle is async, where the kernels
 * clamp the indices instead). */
int  vrad_points_upload(vrad_env*, int64_t n_points, const float* xyz3);
int  vrad_test_lines_indexed(vrad_env*, int64_t n, const int32_t* pairs2, int sky_mode, uint32_t* vis_bits);

/* Environment.TriangleColors (raytracer/environment.go:61-63, GetTriangleColor :430-432): 3 floats per
 * triangle, in the order the triangles were added.  colour.X is the coverage CoverageCount adds for a
 * transparent triangle (coverageCount.go:29).  May be called before or after the build. */
int  vrad_env_set_triangle_colors(vrad_env*, int n, const float* rgb3);

/* ---- BSP point location, sky cameras, full TestLineDoesHitSky ------ */
/* The lumps trace.PointLeafnum (raytracer/trace/pointleaf.go:8-33) and clustertable.PointInLeaf
 * (rad/clustertable/point.go:14-38) walk: Nodes{PlaneNum, Children[2]} (negative child = -1-leaf),
 * Planes{Normal, Distance, AxisType}, Leafs{Cluster, Area}; n_areas = len(Areas). */
int  vrad_bsp_upload(vrad_env*, int n_nodes, const int32_t* node_plane, const int32_t* node_children2,
                     int n_planes, const float* plane_normal3, const float* plane_dist, const int32_t* plane_type,
                     int n_leafs, const int32_t* leaf_cluster, const int32_t* leaf_area, int n_areas);
/* trace.PointLeafnum for n points (xyz interleaved) */
int  vrad_point_leafnum(vrad_env*, int64_t n, const float* pts3, int32_t* leaf_out);
/* clustertable.ClusterFromPoint (rad/clustertable/point.go:10-12) for n points: TEST_EPSILON-tolerant
 * descent that prefers the front child unless it ends in a cluster -1 leaf */
int  vrad_cluster_from_point(vrad_env*, int64_t n, const float* pts3, int32_t* cluster_out);
/* cameras.ProcessSkyCameras (rad/cameras/skycamera.go:10-49): one entry per sky_camera entity (origin,
 * scale); cameras with scale <= 0 are dropped; fills cache.skyCameras / areaSkyCameras
 * (cache/skycameras.go:8-40).  n_kept_out may be NULL.  Needs vrad_bsp_upload first. */
int  vrad_sky_cameras_set(vrad_env*, int n, const float* origin3, const float* scale, int* n_kept_out);
/* read back: per kept camera its area and WorldToSky, and areaSkyCameras[n_areas]; pointers may be NULL */
int  vrad_sky_cameras_get(vrad_env*, int* n_cameras, int32_t* cam_area, float* world_to_sky, int32_t* area_camera);

#define VRAD_TL_CAN_RECURSE      1   /* canRecurse (testline.go:18,58): clip into the 3D sky boxes */
#define VRAD_TL_TEXTURE_SHADOWS  2   /* textureShadows (testline.go:14,32-34,52-55): transparent-triangle coverage */
#define VRAD_TL_PACKET_LEAF      4   /* the leaf/area of every group of 4 segments comes from its first
                                        segment, as the FourVectors form does (start.Vec(0), testline.go:63) */
/* trace.TestLineDoesHitSky (raytracer/trace/testline.go:18-94), complete: skip id
 * TRACE_ID_STATICPROP|static_prop_to_skip (:36), sky-id test (:42-51), coverage (:52-55), 3D-skybox
 * recursion through every sky camera (:57-89), fractionVisible = 1 - clamp(occlusion) (:91-93).
 * start/stop: SoA blocks x[n] y[n] z[n]; fraction_visible: n floats. */
int  vrad_test_lines_sky(vrad_env*, int64_t n, const float* start_xyz_soa, const float* stop_xyz_soa, int flags,
                         int32_t static_prop_to_skip, float* fraction_visible);
/* lightmap.CanLeafTraceToSky (rad/lightmap/lightmap.go:425-451) for n_leafs BSP leafs: from the centre of
 * Mins/Maxs (int16 triples) along every sky direction set with vrad_set_sky_dirs (vmath.Anorms), with
 * recursion; can_out[i] = 1 when any direction has fractionVisible > 0. */
int  vrad_leafs_trace_to_sky(vrad_env*, int n_leafs, const int16_t* mins3, const int16_t* maxs3, uint8_t* can_out);

/* ---- PVS ------------------------------------------------------------ */
/* lightmap.DecompressVis (rad/lightmap/vis.go:54-94): one run-length coded PVS row of the visibility
 * lump -> (n_clusters+7)/8 bytes.  Host-only helper (no device needed).  Returns the number of input
 * bytes consumed, or a negative vrad_status. */
int64_t vrad_decompress_vis(const uint8_t* in, int64_t in_len, int n_clusters, uint8_t* out_row);
/* The whole visibility lump -> the n_clusters x n_clusters byte matrix vrad_build_transfers takes:
 * byteofs2 = dvis ByteOffset[cluster][DVIS_PVS, DVIS_PAS] (lightmap.GetVisCache, vis.go:9-47); a cluster
 * with offset -1 sees nothing.  Host-only helper. */
int  vrad_pvs_from_vis_lump(int n_clusters, const int32_t* byteofs2, const uint8_t* visdata, int64_t vis_len, uint8_t* pvs_out);

/* ---- patch hierarchy --------------------------------------------- */
/* One face as MakePatchForFace sees it (rad/patches/face.go:29-197). */
typedef struct {
    int32_t first_point, n_points;   /* the face winding: points3[3*first_point ...], n_points <= 64 */
    float   normal[3], plane_dist;   /* patch.Plane (face.go:119-142) */
    float   lux_scale;               /* patch.LuxScale = mean length of the two lightmap vectors (face.go:83-109); 1/16 by default */
    float   chop;                    /* patch.Chop = maxChop (face.go:24,111) */
    uint8_t sky;                     /* patch.Sky = IsSky(f) (face.go:106) */
    uint8_t no_subdivide;            /* PreventSubdivision (subdivide.go:151-165) or a displacement face (:60) */
    uint8_t has_base_light;          /* BaseLight != 0: no edge-of-face chop rule (subdivide.go:389-392) */
    uint8_t pad;
} vrad_face_patch;                   /* 36 bytes */
/* patches.MakePatchForFace + SubdividePatches / SubdividePatch / ClipWindingEpsilon / CreateChildPatch /
 * WindingAreaAndBalancePoint (rad/patches/face.go:29-197, subdivide.go:25-66,167-437).  Host-only helper (no
 * device needed).  Patches come out in the reference's order: one root per non-degenerate face, then the children
 * depth first (child1's subtree before child2's).  min_chop = minChop (face.go:25).  Output arrays may be NULL;
 * *n_patches_out / *n_points_out always receive the sizes; VRAD_E_NOMEM when max_patches / max_points are too
 * small (call once with 0 capacities to size the buffers). */
int  vrad_patches_subdivide(int n_faces, const vrad_face_patch* faces, const float* points3, float min_chop,
                            int max_patches, int max_points, int* n_patches_out, int* n_points_out,
                            float* origin3, float* normal3, float* plane_dist, float* area, float* mins3, float* maxs3,
                            float* chop, int32_t* parent, int32_t* child1, int32_t* child2, int32_t* face,
                            int32_t* wind_first, int32_t* wind_count, float* wind_points3);

/* ---- patches, K2 transfers, K3 direct light, K4 bounce ---------------------------------- */
/* fields of common/types/patch.go:9-64 the kernels read (leaf patches only) */
int  vrad_patches_upload(vrad_env*, int n, const float* origin3, const float* normal3, const float* plane_dist,
                         const float* area, const float* reflectivity3, const int32_t* cluster, const uint8_t* flags);
/* Patch.Parent / Child1 / Child2 / FaceNumber (common/types/patch.go:33,49-51) of the patches uploaded before, as
 * patches.SubdividePatches leaves them (vrad_patches_subdivide): children after their parent, in pairs, carrying the
 * parent's reflectivity.  face may be NULL.  Switches the stages to the hierarchical form: vrad_build_transfers builds
 * rows for leaf patches only and picks emitters by the top-down walk of upstream's TestPatchToPatch (descend while
 * |origin_i - origin_j|^2 / 16 < area_j; faces other than the receiver's own), and vrad_bounce runs CollectLight for
 * interior patches (area-weighted average of the two children).  Clusters given to vrad_patches_upload apply to every
 * patch; the PVS test uses the cluster of an emitter's face root. */
int  vrad_patches_set_hierarchy(vrad_env*, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face);
/* Patch.Winding (common/types/patch.go; the polygons vrad_patches_subdivide returns as wind_first / wind_count / wind_points3) of the
 * patches uploaded before.  With windings set, MakeTransfer (vrad_build_transfers) switches from the differential-to-differential
 * form factor to the polygon-to-differential contour integral for emitters that are large for their distance
 * (pi * 0.04 * |delta|^2 < area_j; SURVEY App. B.3 "optional" -- upstream vismat.cpp, absent from the reference: parity unpinned,
 * pinned by the closed form for a rectangle instead).  Patches with count < 3 keep the differential form.  Points run clockwise
 * seen from the patch's front (BSP face convention); a winding the other way round is reversed on upload.  n = 0 removes the
 * windings again.  vrad_patches_upload clears them. */
int  vrad_patches_set_windings(vrad_env*, int n, const int32_t* first, const int32_t* count, int n_points, const float* points3);
/* Bump-mapped patches (Patch.NeedsBumpMap, common/types/patch.go:23; BumpLights = NUM_BUMP_VECTS + 1 light values per patch,
 * common/types/bumpLights.go:8-10, common/constants/constants.go:33).  vrad_bump_normals is upstream's GetBumpNormals for one
 * face (texture S/T vectors, flat and phong normal -> the three bump-basis normals; host-only).  After vrad_patches_set_bump
 * (flag + 9 floats per patch; normals[0] is the patch normal) vrad_bounce also accumulates TotalLight.Light[1..3] for the
 * bump-mapped leaf patches -- each transfer projected on the bump normals as upstream's GatherLight does -- read back with
 * vrad_bounce_bump_totals (9 floats per patch: Light[1], Light[2], Light[3]; zero for other patches).  Light[0] (the flat
 * value, which is what a patch re-emits) is vrad_bounce's total_rgb_out.  Single GPU only for now. */
int  vrad_bump_normals(const float s_vect[3], const float t_vect[3], const float flat_normal[3], const float phong_normal[3], float out9[9]);
int  vrad_patches_set_bump(vrad_env*, int n, const uint8_t* needs_bump, const float* bump_normals9);
int  vrad_bounce_bump_totals(vrad_env*, float* out9);
/* patch-to-patch visibility + form factor -> transfer lists (common/types/transfer.go:3-6;
 * Patch.NumTransfers/Transfers patch.go:60-61).  pvs: n_clusters x n_clusters bytes (non-zero = visible)
 * or NULL (host memory).  Builds and keeps resident the CSR rows owned by this rank.  nnz_out = local nnz.
 * world > 1: ranks own contiguous row blocks that tile [0,N) in rank order.  If vrad_comm_init was called the
 * call is COLLECTIVE and the blocks are balanced by estimated transfers (every 16th candidate pair tested, summed
 * over ranks); otherwise equal blocks of ceil(N/world) rows.  vrad_transfers_info reports the block. */
int  vrad_build_transfers(vrad_env*, int n_clusters, const uint8_t* pvs, int64_t* nnz_out);
/* adopt prebuilt transfer rows [row0,row1): rowptr has row1-row0+1 entries starting at 0.  world > 1: any
 * contiguous blocks that tile [0,N) in rank order (vrad_bounce verifies this collectively). */
int  vrad_transfers_upload(vrad_env*, int64_t row0, int64_t row1, const int64_t* rowptr, const int32_t* col, const float* w);
int  vrad_transfers_info(vrad_env*, int64_t* row0, int64_t* row1, int64_t* nnz);
/* how the resident rows are stored for the gather (any pointer may be NULL): entries of the {col,w} pair array (rows padded to 4);
 * entries and segments of the packed streams (f32 weight + u16 column offset from a per-segment base, 6 bytes per entry, segments
 * padded to 64; 0 / 0 when unused -- hierarchy, rows over many column windows, k4_pack = 0); entries of the block-row streams
 * (block_rows = 2 or 4 consecutive rows share one column list: u16 column offset + block_rows f32 weights per entry of the union; 0 when unused).
 * The gather reads the block-row streams if present, else the packed streams, else the pairs.
 * In-process multi-GPU handle: sums over the ranks. */
int  vrad_transfers_layout(vrad_env*, int64_t* pair_entries, int64_t* packed_entries, int64_t* packed_segments, int64_t* block_entries, int64_t* block_rows);
int  vrad_transfers_download(vrad_env*, int64_t* rowptr, int32_t* col, float* w);
/* rows [row_begin,row_end) of the resident lists (global row numbers, must be owned by this rank): rowptr gets
 * row_end-row_begin+1 offsets starting at 0; col/w must hold `capacity` entries (VRAD_E_INVALID if too small).
 * For matrices too large to download whole (C5: 5.7e9 transfers). */
int  vrad_transfers_download_rows(vrad_env*, int64_t row_begin, int64_t row_end, int64_t* rowptr, int32_t* col, float* w, int64_t capacity);
/* sky-ambient sample directions (vmath.Anorms, vmath/constants.go:15,21-184) */
int  vrad_set_sky_dirs(vrad_env*, int n, const float* dirs3);
/* ---- light records from entities and patches (host-only; the producers of vrad_direct_light's inputs) -------- */
/* Entity.LightForString (common/types/entity.go:103-158): "r g b [scale]" (or one number, or two 4-tuples LDR HDR) ->
 * linear RGB intensity.  VRAD_E_INVALID for negative components or an unknown form (rgb_out = 0). */
int  vrad_light_for_string(const char* value, float rgb_out[3]);
/* The key-values the light parsers read, already converted the way Entity.FloatForKey / VectorForKey / LightForKey do. */
typedef struct {
    int32_t classname;                 /* 0 "light", 1 "light_spot", 2 "light_environment" (lights.go:98-112) */
    float   origin[3];
    int32_t light_ok; float light[3];  /* LightForKey("_light") */
    int32_t has_target; float target_origin[3];     /* "target" -> FindTargetEntity(...).origin (lights.go:184-194) */
    float   angles[3], pitch, angle;   /* "angles" (pitch yaw roll), "pitch", "angle" (-1 up, -2 down) */
    float   inner_cone, cone, exponent;             /* "_inner_cone", "_cone", "_exponent" (degrees) */
    float   fifty_percent_distance, zero_percent_distance; int32_t hardfalloff;
    float   constant_attn, linear_attn, quadratic_attn, distance;
    int32_t ambient_ok; float ambient[3];           /* light_environment: LightForKey("_ambient") */
} vrad_light_entity;                   /* 124 bytes */
/* CreateDirectLights, entity part (rad/lightmap/lights.go:90-113): ParseLightPoint / ParseLightSpot / ParseLightEnvironment
 * + ParseLightGeneric, SetLightFalloffParams, SetupLightNormalFromProps (:173-426) and the quadratic fit of
 * vmath/quadratic/solver.go.  One record per light / light_spot, two (sky + sky ambient) for the first light_environment,
 * in entity order.  *n_out = number of lights; VRAD_E_NOMEM when max_out is too small. */
int  vrad_lights_from_entities(int n, const vrad_light_entity* ents, int max_out, vrad_light* out, int* n_out);
/* CreateDirectLights, surface part (lights.go:49-82): one EMIT_SURFACE light per leaf patch whose average BaseLight
 * reaches light_threshold (lightThreshold, :25); scale2 = Patch.Scale[2]; child1 may be NULL (all leaves). */
int  vrad_lights_from_patches(int n, const float* origin3, const float* normal3, const float* base_light3, const float* area,
                              const float* scale2, const float* base_area, const int32_t* child1, float light_threshold,
                              int max_out, vrad_light* out, int* n_out);

/* How K3's light rays are tested: 0 (default) = binary TestLine, sky lights with the sky-id rule only;
 * VRAD_TL_CAN_RECURSE = sky lights go through the 3D-skybox recursion (canRecurse = true, as
 * lightmap.CanLeafTraceToSky calls it, rad/lightmap/lightmap.go:444; needs vrad_bsp_upload + vrad_sky_cameras_set);
 * VRAD_TL_TEXTURE_SHADOWS = every light ray accumulates transparent-triangle coverage (testline.go:14,32-34,52-55)
 * and contributes dot * fractionVisible. */
int  vrad_set_light_trace_flags(vrad_env*, int flags);
/* per-luxel direct lighting with shadow rays (north_star K3; light parameters per vrad_light) */
int  vrad_direct_light(vrad_env*, int64_t n_luxels, const float* pos3, const float* normal3,
                       int n_lights, const vrad_light* lights, float* rgb_out);
/* iterative bounce gather over the resident transfers (north_star K4).  emit0_rgb: N*3 initial patch
 * radiance; total_rgb_out: N*3 accumulated bounced light (all rows, gathered across ranks);
 * added_last: RGB sum emitted by the last bounce; early_out: stop when all of added < 1. */
int  vrad_bounce(vrad_env*, const float* emit0_rgb, int n_bounces, int early_out, float* total_rgb_out,
                 float added_last[3], int* bounces_done);

/* ---- multi-GPU plumbing ------------------------------------------------------------------ */
/* ONE handle that drives several GPUs of this process ("multi-GPU is internal to the handle, invisible to Go"):
 * the reference is a single goroutine behind a package-level singleton (raytracer/environment.go:17-25;
 * common/constants/constants.go:43), so its driver cannot start one process per GPU.  The handle owns one environment per listed
 * device and runs every call on one worker thread per device: geometry, patches and the point table are replicated (the kd tree
 * is built once and adopted by the other devices); vrad_test_lines / _indexed / vrad_trace_rays / vrad_direct_light split their
 * batch into one contiguous range per device; vrad_build_transfers shards the patch rows (balanced by estimated transfers) and
 * vrad_bounce exchanges the radiance rows every bounce with the fused peer-store kernel (cudaDeviceEnablePeerAccess; no NCCL, no
 * IPC), returning the complete result.  Data pointers must be HOST memory on such a handle; calls are synchronous.  Available:
 * vrad_env_destroy / set_option / last_timing / add_triangles / set_triangle_colors / build / build_fast / upload_tree / stats /
 * download_tree / vrad_trace4 (device 0), vrad_trace_rays, vrad_test_lines, vrad_points_upload, vrad_test_lines_indexed,
 * vrad_patches_upload, vrad_patches_set_hierarchy, vrad_build_transfers, vrad_transfers_info / _download, vrad_set_sky_dirs,
 * vrad_set_light_trace_flags, vrad_direct_light, vrad_bounce; the others return VRAD_E_UNSUPPORTED.  The same device may be
 * listed twice (tests on a single-GPU box): such ranks exchange radiance by device copies instead of the in-kernel barrier. */
typedef struct {
    int n_devices;       /* 1..8 */
    int devices[8];      /* CUDA device ordinals, in rank order */
    int flags;           /* VRAD_CFG_* */
} vrad_multi_config;
int  vrad_env_create_multi(const vrad_multi_config* cfg, vrad_env** out);
/* One process per GPU instead (e.g. under torchrun): one handle per process with (rank, world) in vrad_config and a shared id -- */
/* 128-byte NCCL unique id: rank 0 calls vrad_comm_unique_id, the driver distributes it, every rank
 * calls vrad_comm_init before vrad_bounce. */
int  vrad_comm_unique_id(void* out128);
int  vrad_comm_init(vrad_env*, const void* unique_id128);

/* library / build identification */
const char* vrad_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VRAD_CUDA_H */
