/*
 * oracle.h -- CPU ORACLE for the vrad-b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the executable specification the CUDA path is checked against.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under vrad_b200/ links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (Galaco/VRAD, /root/reference) ships no golden
 * vectors or known-answer tests for this path, its Trace4Rays is a stub
 * (raytracer/environment.go:140-145), its radiosity stages are absent, and no Go
 * toolchain exists in the build image, so the reference cannot be executed.  The
 * oracle restates
 *   - the reference's data layouts      (raytracer/cache/optimisedkdnode.go:15-54,
 *                                        raytracer/cache/triangle/triintersectdata.go:3-22,
 *                                        raytracer/types/fourrays.go:8-11, result.go:8-12,
 *                                        common/types/patch.go:9-64, transfer.go:3-6, light.go:10-44)
 *   - the reference's SAH kd build      (raytracer/environment.go:119-138,181-236,238-387)
 *   - its triangle precomputation       (raytracer/cache/optimisedtriangle.go:30-103,
 *                                        vmath/polygon/edge.go:5-25, surface.go:5-8)
 *   - its trace call surface            (raytracer/trace/testline.go:18-94)
 * with the SURVEY.md App. A defect corrections, and encodes Source-SDK-2013
 * semantics (uncited) for the stubbed/absent stages.
 * Independent arbiters: the brute-force tracer and the analytic known-answer
 * tests in tests/.
 *
 * Arithmetic contract: IEEE-754 binary32, round-to-nearest-even per operation,
 * no FMA contraction (compile with -ffp-contract=off), exact division and sqrt
 * (vmath/ssemath/simd/simd.go:110-117,162-178 use exact 1/x and math.Sqrt).
 */
#ifndef VRAD_ORACLE_H
#define VRAD_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* raytracer/constants.go:9-16 */
#define ORC_TRACE_ID_SKY        0x01000000
#define ORC_TRACE_ID_OPAQUE     0x02000000
#define ORC_TRACE_ID_STATICPROP 0x04000000
#define ORC_KDNODE_LEAF 3

/* raytracer/cache/triangle/triintersectdata.go:3-22 -- 48 bytes */
typedef struct {
    float nx, ny, nz, d;
    int32_t id;
    float e[6];
    uint8_t sel0, sel1, flags, unused;
} orc_tri48;

/* light record shared with include/vrad_cuda.h (vrad_light): 96 bytes.
 * Fields follow common/types/light.go:10-44 + worldlight fields set in
 * rad/lightmap/lights.go:71-81,216-256,259-341. */
typedef struct {
    int32_t type;            /* 0 surface, 1 point, 2 spotlight, 3 skylight, 5 skyambient (bsp emittype_t order) */
    float origin[3];
    float intensity[3];
    float normal[3];
    float stopdot, stopdot2, exponent, radius;
    float constant_attn, linear_attn, quadratic_attn;
    float start_fade, end_fade, cap_dist;
    int32_t flags;
    float pad[3];
} orc_light;

typedef struct orc_env orc_env;

orc_env* orc_env_create(void);
void     orc_env_destroy(orc_env*);
/* Environment.AddTriangleWithMaterial, raytracer/environment.go:45-69 */
int  orc_env_add_triangles(orc_env*, int n, const int32_t* ids, const float* verts9, const uint8_t* flags);
/* Environment.SetupAccelerationStructure, raytracer/environment.go:119-138 */
int  orc_env_build(orc_env*);
int  orc_env_sizes(orc_env*, int* n_nodes, int* n_idx, int* n_tris, int* max_depth, int* n_leaves);
int  orc_env_export(orc_env*, int32_t* children, float* split, int32_t* tri_index, orc_tri48* tris, float aabb[6]);
int  orc_env_replace_tree(orc_env*, int n_nodes, const int32_t* children, const float* split, int n_idx, const int32_t* tri_index, const float aabb[6]);
double orc_env_build_seconds(orc_env*);

/* nearest hit over ALL triangles (ground truth; O(n_tris) per ray) */
int  orc_trace_brute(orc_env*, int64_t n, const float* ox, const float* oy, const float* oz,
                     const float* dx, const float* dy, const float* dz,
                     const float* tmin, const float* tmax, int32_t skip_id,
                     int32_t* hit_tri, int32_t* hit_sid, float* hit_t, int threads);
/* THE SPEC: single-ray kd traversal (scalar form) */
int  orc_trace1(orc_env*, int64_t n, const float* ox, const float* oy, const float* oz,
                const float* dx, const float* dy, const float* dz,
                const float* tmin, const float* tmax, int32_t skip_id,
                int32_t* hit_tri, int32_t* hit_sid, float* hit_t, int threads);
/* 4-wide packet traversal with FourRays semantics (the timed "reference CPU path").
 * Rays are taken 4 at a time in input order; n need not be a multiple of 4. */
int  orc_trace4(orc_env*, int64_t n, const float* ox, const float* oy, const float* oz,
                const float* dx, const float* dy, const float* dz,
                const float* tmin, const float* tmax, int32_t skip_id,
                int32_t* hit_tri, int32_t* hit_sid, float* hit_t, int threads);
/* one FourRays packet, exact RayTracingResult layout (raytracer/types/result.go:8-12) */
int  orc_trace4_packet(orc_env*, const float origin_xyz4[12], const float dir_xyz4[12],
                       const float tmin[4], const float tmax[4], int32_t skip_id,
                       int32_t hit_ids[4], float hit_dist[4], float normal_xyz4[12]);
/* TestLine / TestLineDoesHitSky, raytracer/trace/testline.go:22-51,91-93 (no skybox recursion).
 * start/stop: SoA blocks x[n] y[n] z[n].  vis_bits: bit i of word i/32; 1 = visible.
 * mode 0 = single-ray spec, 1 = packet tracer, 2 = brute force. */
int  orc_test_lines(orc_env*, int64_t n, const float* start_soa, const float* stop_soa,
                    int sky_mode, uint32_t* vis_bits, int mode, int threads);
/* counters for the last orc_trace1 call with threads==1 (nodes visited, triangles tested) */
int  orc_trace_counters(orc_env*, int64_t* nodes_visited, int64_t* tris_tested, int64_t* leaves_visited);

/* ---- radiosity stages ---- */
int  orc_patches_set(orc_env*, int n, const float* origin3, const float* normal3, const float* plane_dist,
                     const float* area, const float* reflectivity3, const int32_t* cluster, const uint8_t* flags);
/* Patch.Parent / Child1 / Child2 / FaceNumber (common/types/patch.go:33,49-51) for the patches set before: switches
 * K2 to the hierarchical candidate walk (vismat.cpp TestPatchToPatch) and K4's CollectLight to
 * the parent/child form (App. B.4).  face may be NULL. */
int  orc_patches_set_hierarchy(orc_env*, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face);
/* Patch.Winding of the patches set before (n = 0 removes them): MakeTransfer then uses the polygon-to-differential form factor for
 * emitters that are large for their distance (pi * 0.04 * |delta|^2 < area_j; SURVEY App. B.3 "optional").  Clockwise seen from the
 * patch's front; a winding the other way round is reversed. */
int  orc_patches_set_windings(orc_env*, int n, const int32_t* first, const int32_t* count, int n_points, const float* points3);
/* K2: builds CSR transfers.  pvs: n_clusters x n_clusters bytes (nonzero = visible) or NULL. */
int  orc_build_transfers(orc_env*, int n_clusters, const uint8_t* pvs, int64_t* nnz_out, int threads);
int  orc_transfers_get(orc_env*, int64_t* rowptr, int32_t* col, float* w);
/* one row (incl. MakeScales) for spot checks on maps whose full matrix is too large for the CPU; returns the entry count */
int64_t orc_transfer_row(orc_env*, int row, int n_clusters, const uint8_t* pvs, int32_t* col_out, float* w_out, int64_t cap);
/* sky-ambient sample directions (the 162 `Anorms`, vmath/constants.go:15,21-184), copied */
int  orc_set_sky_dirs(int n, const float* dirs3);
/* light rays of orc_direct_light: 0 = binary TestLine; ORC_TL_CAN_RECURSE / ORC_TL_TEXTURE_SHADOWS = complete form */
int  orc_env_set_light_trace_flags(orc_env*, int flags);
/* K3: direct light per luxel; rgb_out 3 floats per luxel */
int  orc_direct_light(orc_env*, int64_t n_luxels, const float* pos3, const float* normal3,
                      int n_lights, const orc_light* lights, float* rgb_out, int threads);
/* upstream GetBumpNormals: the three bump-basis normals of a face (texture S/T vectors, flat and phong normal) */
int  orc_bump_normals(const float s_vect[3], const float t_vect[3], const float flat_normal[3], const float phong_normal[3], float out9[9]);
/* Patch.NeedsBumpMap (common/types/patch.go:23) + the bump normals; orc_bounce then also accumulates TotalLight.Light[1..3]
 * (common/types/bumpLights.go:8-10) for the bump-mapped leaf patches, read back with orc_bounce_bump_totals (9 floats per patch) */
int  orc_patches_set_bump(orc_env*, int n, const uint8_t* needs_bump, const float* bump_normals9);
int  orc_bounce_bump_totals(orc_env*, float* out9);
/* K4: bounce.  emit0_rgb: N*3.  total_rgb_out: N*3 (accumulated bounced light, excludes emit0). */
int  orc_bounce(orc_env*, const float* emit0_rgb, int n_bounces, int early_out,
                float* total_rgb_out, float added_last[3], int* bounces_done, int threads);
/* one gather iteration on caller-supplied CSR (for sharded tests/timing): out[r-row0] for rows [row0,row1) */
int  orc_gather_rows(int64_t row0, int64_t row1, const int64_t* rowptr, const int32_t* col, const float* w,
                     const float* emit_rgb, const float* refl_rgb, float* out_rgb, int threads);

/* ---- full TestLineDoesHitSky call surface (oracle/skytrace.cpp) ---- */
#define ORC_TL_CAN_RECURSE     1   /* canRecurse, raytracer/trace/testline.go:18,58 */
#define ORC_TL_TEXTURE_SHADOWS 2   /* textureShadows, testline.go:14,32-34,52-55 */
#define ORC_TL_PACKET_LEAF     4   /* leaf of every group of 4 segments from its first segment (start.Vec(0), testline.go:63) */
/* Environment.TriangleColors (raytracer/environment.go:61-63; GetTriangleColor :430-432): 3 floats per triangle */
int  orc_env_set_triangle_colors(orc_env*, int n, const float* rgb3);
/* BSP point-location data (bsp Nodes / Planes / Leafs lumps as the reference's cache holds them) */
int  orc_bsp_set(orc_env*, int n_nodes, const int32_t* node_plane, const int32_t* node_children2, int n_planes,
                 const float* plane_normal3, const float* plane_dist, const int32_t* plane_type, int n_leafs,
                 const int32_t* leaf_cluster, const int32_t* leaf_area, int n_areas);
/* trace.PointLeafnum, raytracer/trace/pointleaf.go:8-33 */
int  orc_point_leafnum(orc_env*, int64_t n, const float* pts3, int32_t* leaf_out);
/* clustertable.ClusterFromPoint / PointInLeaf, rad/clustertable/point.go:10-38 */
int  orc_cluster_from_point(orc_env*, int64_t n, const float* pts3, int32_t* cluster_out);
/* cameras.ProcessSkyCameras, rad/cameras/skycamera.go:10-49; returns the number of cameras kept */
int  orc_sky_cameras_set(orc_env*, int n, const float* origin3, const float* scale);
int  orc_sky_cameras_get(orc_env*, int32_t* cam_area, float* world_to_sky, int32_t* area_camera);
/* trace.TestLineDoesHitSky per segment, raytracer/trace/testline.go:18-94 */
int  orc_test_lines_sky(orc_env*, int64_t n, const float* start_soa, const float* stop_soa, int flags,
                        int32_t static_prop_to_skip, float* fraction_visible, int threads);
/* lightmap.CanLeafTraceToSky, rad/lightmap/lightmap.go:425-451 (dirs = vmath.Anorms) */
int  orc_leafs_trace_to_sky(orc_env*, int n_leafs, const int16_t* mins3, const int16_t* maxs3, int n_dirs,
                            const float* dirs3, uint8_t* can_out, int threads);
/* lightmap.DecompressVis, rad/lightmap/vis.go:54-94: one PVS row; returns input bytes consumed or <0 */
int64_t orc_decompress_vis(const uint8_t* in, int64_t in_len, int n_clusters, uint8_t* out_row);

int  orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
