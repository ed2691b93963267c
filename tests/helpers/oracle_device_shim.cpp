// Test helper (CPU only): the DEVICE entry points of libvradcuda.so that integration/cpp/vrad_bake.hpp calls, re-implemented on the CPU
// oracle (oracle/_build/liboracle.so).  Linked INTO the test binary (drive_main.cpp + this file), it takes precedence over the library's
// own definitions, so the C++ bake -- bake::Light, bake::Finish, bake::BakeFile -- runs end to end without a GPU and
// tests/test_bsp_cpu.py can compare its lit .bsp with what vrad_b200/bake.py produces on the oracle's environment.
// TEST INFRASTRUCTURE: this is the only place outside tests/*.py where product-side host code is linked against the oracle; nothing
// under vrad_b200/ or integration/ refers to it.  The host-only entry points (vrad_bsp_*, vrad_lights_*, vrad_color_*, ...) still come
// from libvradcuda.so.
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../../include/vrad_bsp.h"
#include "../../oracle/oracle.h"

struct vrad_env { orc_env* o; };
static const int kThreads = 4;

extern "C" {

int vrad_env_create(const vrad_config*, vrad_env** out) { *out = new vrad_env{orc_env_create()}; return 0; }
void vrad_env_destroy(vrad_env* e) { if (e) { orc_env_destroy(e->o); delete e; } }
int vrad_env_add_triangles(vrad_env* e, int n, const int32_t* ids, const float* verts9, const uint8_t* flags) { return orc_env_add_triangles(e->o, n, ids, verts9, flags); }
int vrad_env_build(vrad_env* e) { return orc_env_build(e->o); }
int vrad_patches_upload(vrad_env* e, int n, const float* origin3, const float* normal3, const float* plane_dist, const float* area, const float* refl3,
                        const int32_t* cluster, const uint8_t* flags) { return orc_patches_set(e->o, n, origin3, normal3, plane_dist, area, refl3, cluster, flags); }
int vrad_patches_set_hierarchy(vrad_env* e, int n, const int32_t* parent, const int32_t* child1, const int32_t* child2, const int32_t* face) {
    return orc_patches_set_hierarchy(e->o, n, parent, child1, child2, face);
}
int vrad_build_transfers(vrad_env* e, int n_clusters, const uint8_t* pvs, int64_t* nnz_out) { return orc_build_transfers(e->o, n_clusters, pvs, nnz_out, kThreads); }
static std::vector<float> g_sky_dirs;
int vrad_set_sky_dirs(vrad_env*, int n, const float* dirs3) { g_sky_dirs.assign(dirs3, dirs3 + 3 * static_cast<size_t>(n)); return orc_set_sky_dirs(n, dirs3); }
// lightmap.CanLeafTraceToSky for a batch of leafs: what vrad_bsp_vis_for_light_environment (host code in libvradcuda.so) calls for
// LEAF_FLAGS_RADIAL leafs -- the library's call lands here because the executable's definition comes first in symbol lookup
int vrad_leafs_trace_to_sky(vrad_env* e, int n_leafs, const int16_t* mins3, const int16_t* maxs3, uint8_t* can_out) {
    return orc_leafs_trace_to_sky(e->o, n_leafs, mins3, maxs3, static_cast<int>(g_sky_dirs.size() / 3), g_sky_dirs.data(), can_out, kThreads);
}
int vrad_bsp_upload(vrad_env* e, int n_nodes, const int32_t* node_plane, const int32_t* node_children2, int n_planes, const float* plane_normal3,
                    const float* plane_dist, const int32_t* plane_type, int n_leafs, const int32_t* leaf_cluster, const int32_t* leaf_area, int n_areas) {
    return orc_bsp_set(e->o, n_nodes, node_plane, node_children2, n_planes, plane_normal3, plane_dist, plane_type, n_leafs, leaf_cluster, leaf_area, n_areas);
}
int vrad_cluster_from_point(vrad_env* e, int64_t n, const float* pts3, int32_t* cluster_out) { return orc_cluster_from_point(e->o, n, pts3, cluster_out); }
int vrad_direct_light(vrad_env* e, int64_t n, const float* pos3, const float* normal3, int n_lights, const vrad_light* lights, float* rgb_out) {
    static_assert(sizeof(vrad_light) == sizeof(orc_light), "light records must match");
    return orc_direct_light(e->o, n, pos3, normal3, n_lights, reinterpret_cast<const orc_light*>(lights), rgb_out, kThreads);
}
int vrad_bounce(vrad_env* e, const float* emit0, int n_bounces, int early_out, float* total, float added[3], int* done) {
    return orc_bounce(e->o, emit0, n_bounces, early_out, total, added, done, kThreads);
}
int vrad_patches_set_bump(vrad_env* e, int n, const uint8_t* needs_bump, const float* bump_normals9) { return orc_patches_set_bump(e->o, n, needs_bump, bump_normals9); }
int vrad_bounce_bump_totals(vrad_env* e, float* out9) { return orc_bounce_bump_totals(e->o, out9); }
int vrad_luxel_radial_light(vrad_env*, int64_t n, const int32_t* luxel_face, int n_faces, const int64_t* luxel_first, const int32_t* size2, const int64_t* entry_first,
                            const vrad_radial_entry* entries, int n_patches, const float* patch_total3, const float* patch_bump9, float* out) {
    return vrad_luxel_radial_light_host(n, luxel_face, n_faces, luxel_first, size2, entry_first, entries, n_patches, patch_total3, patch_bump9, out);
}
int vrad_lightmap_finalize(vrad_env*, int64_t n, const float* direct3, const float* indirect3, vrad_color_rgbexp32* out) {
    std::vector<float> sum(3 * static_cast<size_t>(n));
    for (size_t i = 0; i < sum.size(); i++) sum[i] = direct3[i] + (indirect3 ? indirect3[i] : 0.0f);
    return vrad_color_to_rgbexp32(n, sum.data(), out);
}

}  // extern "C"
