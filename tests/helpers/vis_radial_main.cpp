// Test helper (CPU only): lightmap.BuildVisForLightEnvironment with LEAF_FLAGS_RADIAL leafs, whose CanLeafTraceToSky branch needs the
// tracer.  Linked with oracle_device_shim.cpp, so the library's host code calls the oracle where it would call the device.
//   vis_radial map.bsp anorms.txt  ->  one line: the leaf flags
#include <cstdio>
#include "../../integration/cpp/vrad_bake.hpp"

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    loadbsp::Bsp bsp(argv[1]);
    const loadbsp::RayTraceTriangles t = loadbsp::BrushesForRayTrace(bsp.lumps);
    raytracer::Environment env(0);
    std::vector<uint8_t> triFlags(t.ids.size(), 0);
    raytracer::fatal_on(vrad_env_add_triangles(env.handle(), (int)t.ids.size(), t.ids.data(), t.verts9.data(), triFlags.data()), "vrad_env_add_triangles");
    raytracer::fatal_on(vrad_env_build(env.handle()), "vrad_env_build");
    const std::vector<float> dirs = bake::ReadSkyDirs(argv[2]);
    raytracer::fatal_on(vrad_set_sky_dirs(env.handle(), (int)(dirs.size() / 3), dirs.data()), "vrad_set_sky_dirs");
    std::vector<uint8_t> flags((size_t)bsp.lumps.n_leafs + 1), pvs(4096);
    int has = 0;
    raytracer::fatal_on(vrad_bsp_vis_for_light_environment(env.handle(), &bsp.lumps, flags.data(), pvs.data(), &has), "vrad_bsp_vis_for_light_environment");
    for (int i = 0; i < bsp.lumps.n_leafs; i++) std::printf("%d ", flags[i]);
    std::printf("\n");
    return 0;
}
