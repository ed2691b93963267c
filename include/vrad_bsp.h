/*
 * vrad_bsp.h -- the BSP side of the hot path: the .bsp container, the lump
 * records the path consumes, and the host code that turns lumps into the arrays the kernels take
 * (triangles for K1, face patches for K2/K4, luxels for K3) and the baked light back into the
 * lighting lump.  Part of libvradcuda.so; everything here except vrad_lightmap_finalize and
 * vrad_bsp_vis_for_light_environment (with LEAF_FLAGS_RADIAL leafs) is host-only and runs without a GPU.
 *
 * The Go driver keeps its own loader (cache.BuildLumpCache, cache/bsp.go:51-91, on github.com/galaco/bsp);
 * it fills a vrad_bsp_lumps from cache.LumpCache (the cgo source is in integration/go/).  Programs that have
 * no loader (the C++ driver, the tests, the benchmark scenes) use the vrad_bspfile_* container below.
 *
 * Record layouts are the on-disk ones of Source BSP version 20 (the format github.com/galaco/bsp reads,
 * Gopkg.lock:24-28); they are little-endian and naturally aligned, so a lump is an array of these structs.
 */
#ifndef VRAD_BSP_H
#define VRAD_BSP_H

#include <stddef.h>
#include <stdint.h>
#include "vrad_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lump indices used by the path (bsp.LUMP_* in cache/bsp.go:59-81) */
enum {
    VRAD_LUMP_ENTITIES = 0, VRAD_LUMP_PLANES = 1, VRAD_LUMP_TEXDATA = 2, VRAD_LUMP_VERTEXES = 3, VRAD_LUMP_VISIBILITY = 4,
    VRAD_LUMP_NODES = 5, VRAD_LUMP_TEXINFO = 6, VRAD_LUMP_FACES = 7, VRAD_LUMP_LIGHTING = 8, VRAD_LUMP_LEAFS = 10,
    VRAD_LUMP_EDGES = 12, VRAD_LUMP_SURFEDGES = 13, VRAD_LUMP_MODELS = 14, VRAD_LUMP_LEAFFACES = 16, VRAD_LUMP_LEAFBRUSHES = 17,
    VRAD_LUMP_BRUSHES = 18, VRAD_LUMP_BRUSHSIDES = 19, VRAD_LUMP_AREAS = 20, VRAD_LUMP_AREAPORTALS = 21,
    VRAD_LUMP_VERTNORMALS = 30, VRAD_LUMP_VERTNORMALINDICES = 31, VRAD_LUMP_TEXDATA_STRING_DATA = 43,
    VRAD_LUMP_TEXDATA_STRING_TABLE = 44, VRAD_LUMP_LIGHTING_HDR = 53, VRAD_LUMP_FACES_HDR = 58, VRAD_LUMP_MAP_FLAGS = 59,
    VRAD_HEADER_LUMPS = 64
};
#define VRAD_BSP_IDENT   0x50534256   /* "VBSP" */
#define VRAD_BSP_VERSION 20

/* surface / contents / leaf flags (github.com/galaco/bsp/flags, .../primitives/leaf) */
#define VRAD_SURF_LIGHT     0x0001
#define VRAD_SURF_SKY2D     0x0002
#define VRAD_SURF_SKY       0x0004
#define VRAD_SURF_NOLIGHT   0x0400
#define VRAD_SURF_BUMPLIGHT 0x0800
#define VRAD_SURF_NOCHOP    0x4000
#define VRAD_CONTENTS_SOLID    0x1
#define VRAD_CONTENTS_OPAQUE   0x80
#define VRAD_CONTENTS_MOVEABLE 0x4000
#define VRAD_MASK_OPAQUE (VRAD_CONTENTS_SOLID | VRAD_CONTENTS_MOVEABLE | VRAD_CONTENTS_OPAQUE)
#define VRAD_LEAF_FLAGS_SKY    0x01
#define VRAD_LEAF_FLAGS_RADIAL 0x02
#define VRAD_LEAF_FLAGS_SKY2D  0x04

typedef struct { float normal[3]; float dist; int32_t type; } vrad_dplane;                       /* 20 bytes */
typedef struct { uint16_t v[2]; } vrad_dedge;                                                    /*  4 */
typedef struct {
    uint16_t planenum; uint8_t side, on_node; int32_t firstedge; int16_t numedges, texinfo, dispinfo, fog_volume;
    uint8_t styles[4]; int32_t lightofs; float area; int32_t lm_mins[2], lm_size[2]; int32_t orig_face;
    uint16_t num_prims, first_prim; uint32_t smoothing_groups;
} vrad_dface;                                                                                    /* 56 */
typedef struct { float texture_vecs[2][4]; float lightmap_vecs[2][4]; int32_t flags, texdata; } vrad_texinfo;   /* 72 */
typedef struct { float reflectivity[3]; int32_t name_id, width, height, view_width, view_height; } vrad_dtexdata; /* 32 */
typedef struct { float mins[3], maxs[3], origin[3]; int32_t headnode, firstface, numfaces; } vrad_dmodel;       /* 48 */
typedef struct { int32_t planenum; int32_t children[2]; int16_t mins[3], maxs[3]; uint16_t firstface, numfaces; int16_t area, pad; } vrad_dnode; /* 32 */
typedef struct {
    int32_t contents; int16_t cluster; int16_t area_flags;        /* area = low 9 bits, flags = high 7 bits */
    int16_t mins[3], maxs[3]; uint16_t firstleafface, numleaffaces, firstleafbrush, numleafbrushes; int16_t leaf_water_data, pad;
} vrad_dleaf;                                                                                    /* 32 (lump version 1) */
typedef struct { int32_t firstside, numsides, contents; } vrad_dbrush;                           /* 12 */
typedef struct { uint16_t planenum; int16_t texinfo, dispinfo, bevel; } vrad_dbrushside;         /*  8 */
typedef struct { uint8_t r, g, b; int8_t exponent; } vrad_color_rgbexp32;                        /*  4 */

/* The lumps cache.LumpCache holds (cache/bsp.go:21-47), as typed views.  visdata = the raw visibility lump
 * (int32 numclusters, int32 byteofs[numclusters][2], run-length rows; cache.LumpCache.VisDataRaw). */
typedef struct {
    int32_t n_planes;      const vrad_dplane* planes;
    int32_t n_vertexes;    const float* vertexes3;
    int32_t n_edges;       const vrad_dedge* edges;
    int32_t n_surfedges;   const int32_t* surfedges;
    int32_t n_faces;       const vrad_dface* faces;
    int32_t n_texinfo;     const vrad_texinfo* texinfo;
    int32_t n_texdata;     const vrad_dtexdata* texdata;
    int32_t n_models;      const vrad_dmodel* models;
    int32_t n_nodes;       const vrad_dnode* nodes;
    int32_t n_leafs;       const vrad_dleaf* leafs;
    int32_t n_leaffaces;   const uint16_t* leaffaces;
    int32_t n_leafbrushes; const uint16_t* leafbrushes;
    int32_t n_brushes;     const vrad_dbrush* brushes;
    int32_t n_brushsides;  const vrad_dbrushside* brushsides;
    int32_t n_areas;       int32_t pad0;
    int64_t vis_len;       const uint8_t* visdata;
} vrad_bsp_lumps;

/* ---- .bsp container ---------------------------------------------------------------------- */
/* loadBSP (cmd/tasks/loadbsp/main.go:163-170: bsp.NewReader(file).Read()) and the writer the finish task
 * leaves commented out (cmd/tasks/finish/main.go:15-18).  Header = ident, version, 64 x {fileofs, filelen,
 * version, fourCC}, mapRevision; lumps are kept as opaque byte strings and written back 4-byte aligned in
 * index order.  Pointers returned by vrad_bspfile_get_lump / _lumps stay valid until that lump is set again
 * or the file is closed. */
typedef struct vrad_bspfile vrad_bspfile;
int  vrad_bspfile_create(int map_revision, vrad_bspfile** out);
int  vrad_bspfile_open(const char* path, vrad_bspfile** out);
int  vrad_bspfile_get_lump(vrad_bspfile*, int lump, const void** data, int64_t* len, int* lump_version);
int  vrad_bspfile_set_lump(vrad_bspfile*, int lump, const void* data, int64_t len, int lump_version);
int  vrad_bspfile_save(vrad_bspfile*, const char* path);
void vrad_bspfile_close(vrad_bspfile*);
/* cache.SetTargetFaces (cmd/tasks/loadbsp/main.go:79-89): which face lump the job lights.  hdr = 0: LUMP_FACES / LUMP_LIGHTING;
 * hdr != 0: LUMP_FACES_HDR / LUMP_LIGHTING_HDR (an empty HDR face lump is seeded from LUMP_FACES).  vrad_bspfile_lumps views the
 * chosen lump from then on; the two lump numbers to write the results to come back through the out pointers (may be NULL). */
int  vrad_bspfile_set_target_faces(vrad_bspfile*, int hdr, int* face_lump_out, int* lighting_lump_out);
/* typed views of the lumps above; VRAD_E_INVALID when a lump's length is not a multiple of its record size,
 * the leaf lump is not version 1, or an index stored in one lump points outside another. */
int  vrad_bspfile_lumps(vrad_bspfile*, vrad_bsp_lumps* out);

/* The checks vrad_bspfile_lumps runs, for lumps that come from elsewhere (the Go loader's cache.LumpCache): every index one lump
 * stores into another is in range, counts have data, the node lump is a forest under the models' head nodes.  The functions below
 * assume lumps that passed. */
int  vrad_bsp_validate(const vrad_bsp_lumps*);

/* ---- lumps -> K1 triangles ----------------------------------------------------------------- */
/* addBrushesForRayTrace + addBrushToRaytraceEnvironment + brush.GetBrushRecursive
 * (cmd/tasks/loadbsp/main.go:233-340, cmd/tasks/loadbsp/brush/brush.go:7-36) with polygon.BaseWindingForPlane /
 * ChopWindingInPlace (vmath/polygon/winding.go:25-171): every opaque brush of model 0 -> its non-sky,
 * non-displacement sides as triangle fans with id TRACE_ID_OPAQUE, then every SURF_SKY face of model 0 as a fan
 * with id TRACE_ID_SKY -- in that order, so triangle indices equal the reference's.  (the
 * literal text keeps only the sky sides and its winding code cannot run; upstream's rule is implemented.)
 * Before the world come the brush entities with "vrad_brush_cast_shadows" (ExtractBrushEntityShadowCasters, main.go:186-211),
 * in entity order: caster_model = index of the entity's brush model ("*N"), caster_origin3 / caster_angles3 = its "origin" /
 * "angles" keys (pitch yaw roll, degrees), applied with matrix.SetupMatrixOrgAngles + Mul4x3 (vmath/matrix/mat4.go:10-69).
 * ids / verts9 may be NULL to size the buffers; VRAD_E_NOMEM when max_tris is too small. */
int  vrad_bsp_raytrace_triangles(const vrad_bsp_lumps*, int n_casters, const int32_t* caster_model, const float* caster_origin3,
                                 const float* caster_angles3, int max_tris, int32_t* ids, float* verts9, int* n_out);
/* the same, straight into an environment (Environment.AddTriangle for each, raytracer/environment.go:41-43) */
int  vrad_env_add_bsp(vrad_env*, const vrad_bsp_lumps*, int n_casters, const int32_t* caster_model, const float* caster_origin3,
                      const float* caster_angles3, int* n_added);

/* ---- lumps -> face patches ------------------------------------------------------------------ */
/* patches.MakePatches + world.WindingFromFace + world.RemoveColinearPoints + the texinfo-derived fields of
 * patches.MakePatchForFace / BaseLightForFace / PreventSubdivision (rad/patches/build.go:21-65,
 * rad/world/face.go:92-114, rad/world/point.go:12-46, rad/patches/face.go:29-230, subdivide.go:151-165): one record
 * per non-displacement face of every model, in face order, ready for vrad_patches_subdivide.  model_origins3 =
 * the "origin" key of each model's entity (n_models x 3; NULL = all zero).  Per output face: its face number, the
 * reflectivity (texdata reflectivity clamped to 0.99), BaseArea (texdata width x height), NeedsBumpMap
 * (SURF_BUMPLIGHT) and the lightmap/texture scales Patch.Scale.  Degenerate faces (area <= 0) still get a record
 * (vrad_patches_subdivide drops them as MakePatchForFace does).  The plane is Planes[f.Planenum] as in the reference
 * (face.go:120): the face lump names the facing plane of a plane pair, `side` is not consulted.
 * Output arrays may be NULL to size the buffers. */
int  vrad_bsp_face_patches(const vrad_bsp_lumps*, const float* model_origins3, float max_chop,
                           int max_faces, int max_points, int* n_faces_out, int* n_points_out,
                           vrad_face_patch* faces, float* points3, int32_t* face_number, float* reflectivity3,
                           float* base_area, uint8_t* needs_bump, float* scale2);
/* ---- texture lights (lights.rad) ------------------------------------------------------------ */
typedef struct { char name[116]; float value[3]; } vrad_texlight;     /* types.TexLight (common/types/texlight.go:5-9); 128 bytes */
/* lights_rad.Reader.Read + lightForString + forceTextureShadowsOnModel (common/parser/lights-rad/reader.go:19-192): the text of a
 * lights.rad file -> the texlight table (a later definition of a name overrides an earlier one; at most MAX_TEXLIGHTS = 128) and the
 * "noshadow" materials / "forcetextureshadow" models, returned as NUL-terminated names back to back in names_out (first the
 * n_noshadow materials, then the n_forced models; *names_len = bytes used; names_out may be NULL to size).  hdr != 0 keeps "hdr:"
 * lines and drops "ldr:" lines.  out may be NULL to count. */
int  vrad_texlights_parse(const char* text, int64_t len, int hdr, int max_out, vrad_texlight* out, int* n_out,
                          char* names_out, int64_t names_cap, int* n_noshadow, int* n_forced, int64_t* names_len);
/* patches.BaseLightForFace / LightForTexture (rad/patches/face.go:208-280) for the faces vrad_bsp_face_patches returned (face_number):
 * the material name of each face (LUMP_TEXDATA_STRING_TABLE / _DATA; cubemap-patched names "maps/<map_name>/<original>_%d_%d_%d" are
 * reduced to <original>) looked up in the texlight table -> Patch.BaseLight (base_light3_out, zero when the material emits nothing).
 * faces_inout (may be NULL): emitting faces get has_base_light = 1 and, unless SURF_NOCHOP, may be subdivided (face.go:158-163 sets
 * SURF_LIGHT on their texinfo).  The result feeds vrad_lights_from_patches. */
int  vrad_bsp_apply_texlights(const vrad_bsp_lumps*, const int32_t* string_table, int n_strings, const char* string_data, int64_t string_len,
                              const char* map_name, int n_texlights, const vrad_texlight* texlights,
                              int n_faces, const int32_t* face_number, vrad_face_patch* faces_inout, float* base_light3_out);

/* rad.Start's luxel-density rescale (rad/start.go:21-64): lightmap vectors longer than luxel_density luxels per unit are
 * shortened to it (in place; no-op for luxel_density >= 1).  Call before vrad_bsp_face_extents, as
 * UpdateAllFaceLightmapExtents (:100-111) does. */
int  vrad_bsp_rescale_lightmap_vecs(int n_texinfo, vrad_texinfo* texinfo, float luxel_density);
/* world.CalcFaceExtents (rad/world/face.go:14-90) for every face: lightmap mins and size in luxels;
 * *n_oversize_out (may be NULL) counts faces larger than MAX_LIGHTMAP_DIM_WITHOUT_BORDER + 1 = 126 luxels
 * -- common/constants/constants.go:22,27 give displacement and plain faces the same limit, 125 + 1), which the reference logs.
 * SURF_SKY / SURF_NOLIGHT faces keep the extents stored in the face (UpdateAllFaceLightmapExtents, rad/start.go:104-106). */
int  vrad_bsp_face_extents(const vrad_bsp_lumps*, int32_t* mins2, int32_t* size2, int* n_oversize_out);

/* ---- tree / cluster tables ------------------------------------------------------------------- */
/* clustertable.MakeParents(0, -1) (rad/clustertable/nodes.go:21-36): parent node of every node and leaf */
int  vrad_bsp_make_parents(const vrad_bsp_lumps*, int32_t* node_parents, int32_t* leaf_parents);
/* clustertable.BuildClusterTable (rad/clustertable/build.go:9-31): the leafs of every cluster, ascending, as
 * CSR (first[n_clusters + 1], leafs[n_leafs]) */
int  vrad_bsp_cluster_table(const vrad_bsp_lumps*, int n_clusters, int32_t* first, int32_t* leafs);
/* lightmap.BuildVisForLightEnvironment + MergeDLightVis + PVSCheck (rad/lightmap/lightmap.go:284-422): the leaf
 * flags (LEAF_FLAGS_SKY / SKY2D on leafs that hold or see sky faces) and the merged PVS of the sky light and
 * the sky ambient light ((n_clusters + 7) / 8 bytes, *has_pvs = 0 when no leaf holds a sky face).  Leafs with
 * LEAF_FLAGS_RADIAL that see no sky leaf are traced with lightmap.CanLeafTraceToSky through `env`
 * (vrad_leafs_trace_to_sky); env may be NULL when the map has no such leafs (VRAD_E_STATE otherwise). */
int  vrad_bsp_vis_for_light_environment(vrad_env* env, const vrad_bsp_lumps*, uint8_t* leaf_flags_out,
                                        uint8_t* sky_pvs_out, int* has_pvs);

/* ---- smoothing normals ------------------------------------------------------------------------ */
/* lightmap.PairEdges (rad/lightmap/lightmap.go:37-216): per face-vertex smoothed normals (sum of numedges
 * entries, face order) and the neighbour lists (CSR: first[n_faces + 1], at most 64 per face).
 * smoothing_threshold = cos of the crease angle (0.7071067 by default, :29).  neighbours may be NULL. */
int  vrad_bsp_pair_edges(const vrad_bsp_lumps*, float smoothing_threshold, float* vertex_normals3,
                         int32_t* neighbour_first, int32_t* neighbours, int max_neighbours);
/* lightmap.SaveVertexNormals + NormalList.FindOrAddNormal (lightmap.go:218-265, normallist.go:12-49): the
 * LUMP_VERTNORMALS / LUMP_VERTNORMALINDICES contents: unique normals (1e-5 squared distance, 8x8x8 grid) and
 * one index per face-vertex.  normals3 may be NULL to size; returns VRAD_E_NOMEM when max_normals is too small. */
int  vrad_bsp_save_vertex_normals(int n_face_vertices, const float* vertex_normals3, int max_normals,
                                  float* normals3, uint16_t* indices, int* n_normals_out);
/* lightmap.GetPhongNormal (rad/lightmap/normallist.go:52-143) for n points: face[i] is the face a point lies
 * on, centroids3 = cache.faceCentroids (rad/patches/face.go:151: the root patch's origin minus the face offset).
 * Points outside every centre-edge wedge keep the face normal. */
int  vrad_bsp_phong_normals(const vrad_bsp_lumps*, float smoothing_threshold, const float* vertex_normals3,
                            const float* centroids3, int64_t n, const int32_t* face, const float* points3, float* normals3_out);

/* ---- luxels and the lighting lump -------------------------------------------------------------- */
/* Lay the lighting lump out and point the faces at it (what upstream's PrecompLightmapOffsets does, UNCITED -- absent
 * from the reference).  Per lit face, in face order: 4 bytes for the style's average colour, then
 * (size[0]+1) x (size[1]+1) samples of 4 bytes -- four such blocks (flat + NUM_BUMP_VECTS, common/constants/constants.go:33)
 * for SURF_BUMPLIGHT faces.  faces_out (n_faces records, may be NULL) = the face lump with lightofs pointing at the first
 * sample, styles = {0,255,255,255}, lm_mins / lm_size from vrad_bsp_face_extents; SURF_SKY / SURF_NOLIGHT faces get
 * lightofs -1.  luxel_first (n_faces + 1) = offset of each face's samples in the luxel arrays K3 / K5 work on (unlit faces:
 * empty range); *lump_bytes = size of the lump.  A lit face larger than 126 luxels on a side (vrad_bsp_face_extents counts them) is an
 * error here: "Bad surface extents", as upstream has it (the reference logs and goes on, rad/world/face.go:66-87). */
int  vrad_bsp_layout_lighting(const vrad_bsp_lumps*, const int32_t* mins2, const int32_t* size2, vrad_dface* faces_out,
                              int64_t* luxel_first, int64_t* lump_bytes);
/* Sample positions for K3 (upstream InitLightinfo / CalcPoints, UNCITED; without upstream's nudging of samples that fall off
 * the face): for each lit face the luxel corners mins + (s,t), t major, mapped to the face plane through the inverse of the
 * lightmap vectors and lifted one unit along the face normal.  normal3 = the face normal (Planes[f.Planenum]); the three
 * extra blocks of a bump-mapped face repeat the positions with the bump-basis normals (vrad_bump_normals), so that K3 run on
 * the block gives the light on that basis vector.  face_origins3 = cache.faceOffsets (the model origin per face) or NULL. */
int  vrad_bsp_face_luxels(const vrad_bsp_lumps*, const int32_t* mins2, const int32_t* size2, const float* face_origins3,
                          const int64_t* luxel_first, float* pos3, float* normal3, int32_t* luxel_face);
/* Move every sample of vrad_bsp_face_luxels onto its face (own rule; upstream's BuildFacesamples chops the face winding into
 * luxel-sized pieces and lights their centres, UNCITED): the sample of luxel (s,t) becomes the centroid of the part of the face that
 * lies in the luxel's cell [s-1/2, s+1/2] x [t-1/2, t+1/2]; a luxel whose cell misses the face takes the nearest point of the face
 * outline.  Grid points that hang over the face's edge would otherwise sit inside the neighbouring brush and come out black.
 * pos3_inout = the positions of vrad_bsp_face_luxels (moved in place, still one unit off the surface; the blocks of a bump-mapped
 * face move alike); luxel_st2_out (may be NULL) = the sample's lightmap coordinates relative to the mins. */
int  vrad_bsp_place_samples(const vrad_bsp_lumps*, const int32_t* mins2, const int32_t* size2, const float* face_origins3,
                            const int64_t* luxel_first, float* pos3_inout, float* luxel_st2_out);
/* upstream VectorToColorRGBExp32 (UNCITED): linear RGB -> 8-bit mantissas with a shared power-of-two exponent (largest
 * component brought into [128,255]).  Host-only; the device kernel behind vrad_lightmap_finalize runs the same inline
 * function (vrad_b200/csrc/rgbexp.cuh).  Negative / NaN components count as 0; below 2^-120 everything encodes as 0. */
int  vrad_color_to_rgbexp32(int64_t n, const float* rgb3, vrad_color_rgbexp32* out);
/* inverse (upstream ColorRGBExp32ToVector): component * 2^exponent */
int  vrad_color_from_rgbexp32(int64_t n, const vrad_color_rgbexp32* in, float* rgb3);
/* K5 -- final light per luxel on the device (upstream FinalLightFace, UNCITED): out = RGBExp32(direct + indirect), negative
 * components clamped to zero.  indirect3 may be NULL.  Pointers may be host or device memory. */
int  vrad_lightmap_finalize(vrad_env*, int64_t n, const float* direct3, const float* indirect3, vrad_color_rgbexp32* out);
/* Which patch lights which luxel: the leaf patch (child1[i] == -1; child1 NULL = all patches are leaves) of the luxel's own face
 * whose origin is nearest to the sample point; -1 when the face has no patch.  Own rule -- upstream's FinalLightFace filters the
 * patch lights of a face and its neighbours through a radial kernel (BuildPatchRadial / SampleRadial, UNCITED and not restated);
 * the nearest-patch lookup is the piecewise-constant form of it.  Host-only. */
int  vrad_luxel_nearest_patch(int64_t n, const int32_t* luxel_face, const float* pos3, int n_patches, const int32_t* patch_face,
                              const int32_t* child1, const float* origin3, int32_t* patch_out);
/* Bounced light per luxel by upstream's radial filter (radial.cpp: BuildPatchRadial / AddBouncedToRadial / SampleRadial; UNCITED,
 * absent from the reference).  vrad_bsp_radial_entries (host) lists, per lit face, the leaf patches of the face and of its
 * smoothing neighbours (vrad_bsp_pair_edges; NULL = own patches only) with their position and extent in that face's luxel space
 * (relative to the lightmap mins; extents at least one luxel, stored as reciprocals).  vrad_luxel_radial_light (device) then gives
 * every luxel sum(r * TotalLight(patch)) / sum(r) with r = 2 - ((cs-s)/ds)^2 - ((ct-t)/dt)^2 over the entries with r > 0 (zero when no
 * patch reaches it); the extra blocks of a bump-mapped face take patch_bump9 (vrad_bounce_bump_totals; NULL = the flat totals).
 * luxel_first / size2 / entry_first are host arrays; the rest may be host or device memory.  _host runs the same code on the
 * host's cores (no device needed). */
typedef struct { int32_t patch; float s, t, inv_ds, inv_dt; } vrad_radial_entry;      /* 20 bytes */
int  vrad_bsp_radial_entries(const vrad_bsp_lumps*, const int32_t* mins2, const float* face_origins3,
                             int n_patches, const int32_t* patch_face, const int32_t* child1, const float* origin3,
                             const int32_t* wind_first, const int32_t* wind_count, const float* wind_points3,
                             const int32_t* neighbour_first, const int32_t* neighbours,
                             int64_t max_entries, int64_t* entry_first, vrad_radial_entry* entries, int64_t* n_entries_out);
int  vrad_luxel_radial_light(vrad_env*, int64_t n, const int32_t* luxel_face, int n_faces, const int64_t* luxel_first, const int32_t* size2,
                             const int64_t* entry_first, const vrad_radial_entry* entries, int n_patches, const float* patch_total3,
                             const float* patch_bump9, float* indirect3_out);
int  vrad_luxel_radial_light_host(int64_t n, const int32_t* luxel_face, int n_faces, const int64_t* luxel_first, const int32_t* size2,
                                  const int64_t* entry_first, const vrad_radial_entry* entries, int n_patches, const float* patch_total3,
                                  const float* patch_bump9, float* indirect3_out);
/* K5 with the bounced light looked up per luxel: out = RGBExp32(direct + patch_total[luxel_patch]) (luxel_patch -1: direct only).
 * patch_total3 = the N x 3 totals vrad_bounce returns.  Device pointers must be 16-byte aligned. */
int  vrad_lightmap_finalize_patches(vrad_env*, int64_t n, const float* direct3, const int32_t* luxel_patch, int n_patches,
                                    const float* patch_total3, vrad_color_rgbexp32* out);
/* Scatter the packed luxels into the lighting lump laid out by vrad_bsp_layout_lighting (lumps->faces must be the faces_out
 * of that call) and fill each face's average colour from its flat block. */
int  vrad_bsp_pack_lighting(const vrad_bsp_lumps*, const int64_t* luxel_first, const vrad_color_rgbexp32* colors,
                            uint8_t* lump_out, int64_t lump_bytes);

#ifdef __cplusplus
}
#endif
#endif /* VRAD_BSP_H */
