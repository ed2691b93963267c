// drive_main.cpp -- tiny C++ driver that uses the host mirror the way the reference's loadbsp step uses
// raytracer.Environment (cmd/tasks/loadbsp/main.go:276,148): add geometry, build, trace packets.
// Prints one line per packet lane; tests/test_gpu_cpp_driver.py compares the output with the oracle.
#include <cstdio>
#include <cstring>
#include <vector>
#include "vrad_environment.hpp"
#include "vrad_bsp.h"
#include "vrad_bake.hpp"

// `drive --bsp map.bsp`: the host-only half of loadbsp.Main + rad.Start on a real file (cmd/tasks/loadbsp/main.go:56-150,
// rad/start.go:66-98): read the lumps, make the ray-trace triangles, the face patches and their subdivision, the face extents
// and the lighting-lump layout, and print the counts.  Needs no GPU; tests/test_bsp_cpu.py checks the numbers.
static int bsp_summary(const char* path) {
    vrad_bspfile* probe = nullptr;
    if (vrad_bspfile_open(path, &probe)) { std::fprintf(stderr, "%s\n", vrad_last_error()); return 1; }     // a missing file is an error message, not an abort
    vrad_bspfile_close(probe);
    loadbsp::Bsp bsp(path);
    const vrad_bsp_lumps& L = bsp.lumps;
    const loadbsp::RayTraceTriangles tris = loadbsp::BrushesForRayTrace(L);
    const patches::FacePatches fp = patches::MakePatches(L);
    patches::PatchTree t = patches::SubdividePatches(fp.faces, fp.points3);
    int leaves = 0;
    for (int i = 0; i < t.size(); i++) leaves += t.child1[i] == -1;
    const lightmap::FaceNeighbours fn = lightmap::PairEdges(L);
    std::vector<int32_t> mins(2 * (size_t)L.n_faces), size(2 * (size_t)L.n_faces);
    std::vector<int64_t> first((size_t)L.n_faces + 1);
    int64_t lump_bytes = 0;
    int oversize = 0;
    if (vrad_bsp_face_extents(&L, mins.data(), size.data(), &oversize) ||
        vrad_bsp_layout_lighting(&L, mins.data(), size.data(), nullptr, first.data(), &lump_bytes)) {
        std::fprintf(stderr, "%s\n", vrad_last_error()); return 1;
    }
    std::printf("bsp faces %d brushes %d triangles %d patches %d leaves %d luxels %lld lighting_bytes %lld oversize %d neighbours %d\n",
                L.n_faces, L.n_brushes, (int)tris.ids.size(), t.size(), leaves, (long long)first[L.n_faces], (long long)lump_bytes, oversize, (int)fn.neighbours.size());
    return 0;
}

// `drive --prepare map.bsp [switches]`: the host half of the bake (bake::Prepare) with a checksum per array; tests/test_bsp_cpu.py compares them with
// what vrad_b200/bake.py prepares from the same file.  No GPU.
// trailing switches shared by --prepare and --bake: [-lights f] [-bounce n] [-luxeldensity x] [-smooth deg] [-chop c] [-maxchop c] [-hdr] [-fast] [-textureshadows]
static bake::Options parse_options(int argc, char** argv, int first) {
    bake::Options o;
    for (int i = first; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : "0"; };
        if (a == "-lights") o.lightsRad = next();
        else if (a == "-bounce") o.bounce = std::atoi(next());
        else if (a == "-luxeldensity") o.luxelDensity = static_cast<float>(std::atof(next()));
        else if (a == "-smooth") o.smoothDegrees = static_cast<float>(std::atof(next()));
        else if (a == "-chop") o.chop = static_cast<float>(std::atof(next()));
        else if (a == "-maxchop") o.maxChop = static_cast<float>(std::atof(next()));
        else if (a == "-hdr") o.hdr = true;
        else if (a == "-fast") o.fast = true;
        else if (a == "-textureshadows") o.textureShadows = true;
        else { std::fprintf(stderr, "unknown switch %s\n", a.c_str()); std::exit(2); }
    }
    return o;
}

static int bake_prepare(const char* path, const bake::Options& opt) {
    const char* lights_rad = opt.lightsRad.empty() ? nullptr : opt.lightsRad.c_str();
    vrad_bspfile* probe = nullptr;
    if (vrad_bspfile_open(path, &probe)) { std::fprintf(stderr, "%s\n", vrad_last_error()); return 1; }
    vrad_bspfile_close(probe);
    loadbsp::Bsp bsp(path, opt.hdr);
    const void* ent = nullptr; int64_t len = 0;
    vrad_bspfile_get_lump(bsp.file, VRAD_LUMP_ENTITIES, &ent, &len, nullptr);
    std::string text(static_cast<const char*>(ent), static_cast<size_t>(len));
    while (!text.empty() && text.back() == '\0') text.pop_back();
    bake::Prepared P;
    bake::TexLights tex;
    if (lights_rad) { bake::LoadTexLights(bsp, path, lights_rad, tex); tex.hdr = opt.hdr; }
    bake::Prepare(bsp.lumps, text, P, lights_rad ? &tex : nullptr, opt.chop, opt.maxChop, opt.SmoothingThreshold(), opt.luxelDensity);
    auto line = [](const char* name, uint64_t sum, size_t count) { std::printf("%s %llu %zu\n", name, (unsigned long long)sum, count); };
    line("tri_ids", bake::Checksum(P.tris.ids), P.tris.ids.size()); line("tri_verts", bake::Checksum(P.tris.verts9), P.tris.verts9.size());
    line("origin", bake::Checksum(P.tree.origin), P.tree.origin.size()); line("normal", bake::Checksum(P.tree.normal), P.tree.normal.size());
    line("plane_dist", bake::Checksum(P.tree.plane_dist), P.tree.plane_dist.size()); line("area", bake::Checksum(P.tree.area), P.tree.area.size());
    line("parent", bake::Checksum(P.tree.parent), P.tree.parent.size()); line("child1", bake::Checksum(P.tree.child1), P.tree.child1.size());
    line("face_of_patch", bake::Checksum(P.face_of_patch), P.face_of_patch.size()); line("cluster", bake::Checksum(P.cluster), P.cluster.size());
    line("flags", bake::Checksum(P.flags), P.flags.size()); line("refl", bake::Checksum(P.refl3), P.refl3.size());
    line("needs_bump", bake::Checksum(P.needs_bump), P.needs_bump.size()); line("bump_basis", bake::Checksum(P.bump_basis9), P.bump_basis9.size());
    line("pvs", bake::Checksum(P.pvs), P.pvs.size()); line("sky_pvs", bake::Checksum(P.sky_pvs), P.sky_pvs.size());
    line("lights", bake::Checksum(P.lights), P.lights.size());
    line("lm_mins", bake::Checksum(P.mins2), P.mins2.size()); line("lm_size", bake::Checksum(P.size2), P.size2.size());
    line("lit_faces", bake::Checksum(P.lit_faces), P.lit_faces.size()); line("luxel_first", bake::Checksum(P.luxel_first), P.luxel_first.size());
    line("lux_pos", bake::Checksum(P.lux_pos3), P.lux_pos3.size()); line("lux_normal", bake::Checksum(P.lux_normal3), P.lux_normal3.size());
    line("lux_face", bake::Checksum(P.lux_face), P.lux_face.size());
    line("radial_first", bake::Checksum(P.radial_first), P.radial_first.size()); line("radial_entries", bake::Checksum(P.radial), P.radial.size());
    std::printf("lump_bytes %lld\n", (long long)P.lump_bytes);
    return 0;
}

// `drive --bake in.bsp out.bsp anorms.txt [switches]`: the whole job on the GPU (bake::BakeFile); prints what tests/test_gpu_zz_bsp_bake.py checks
static int bake_file(const char* in, const char* out, const char* anorms, const bake::Options& opt) {
    const bake::Lit lit = bake::BakeFile(in, out, anorms, 0, opt);
    std::printf("baked transfers %lld bounces %d direct %llu emit %llu total %llu bump %llu\n", (long long)lit.nnz, lit.bounces,
                (unsigned long long)bake::Checksum(lit.direct3), (unsigned long long)bake::Checksum(lit.emit3), (unsigned long long)bake::Checksum(lit.total3),
                (unsigned long long)bake::Checksum(lit.bump9));
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 3 && !std::strcmp(argv[1], "--bsp")) return bsp_summary(argv[2]);
    if (argc >= 3 && !std::strcmp(argv[1], "--prepare")) return bake_prepare(argv[2], parse_options(argc, argv, 3));
    if (argc >= 5 && !std::strcmp(argv[1], "--bake")) return bake_file(argv[2], argv[3], argv[4], parse_options(argc, argv, 5));
    raytracer::Environment env;
    // a 256^3 room (6 inward quads) with two box occluders, ids as loadbsp assigns them
    env.AddAxisAlignedRectangularSolid(raytracer::TRACE_ID_OPAQUE, {0, 0, 0}, {256, 256, 256});
    env.AddAxisAlignedRectangularSolid(raytracer::TRACE_ID_OPAQUE + 16, {40, 40, 0}, {100, 90, 64});
    env.AddAxisAlignedRectangularSolid(raytracer::TRACE_ID_OPAQUE + 32, {150, 120, 0}, {220, 200, 128});
    env.SetupAccelerationStructure();
    const float dirs[4][3] = {{0.6f, 0.0f, -0.8f}, {-0.6f, 0.48f, -0.64f}, {0.0f, 1.0f, 0.0f}, {0.36f, 0.48f, 0.8f}};
    for (int p = 0; p < 4; p++) {
        raytracer::FourRays rays;
        raytracer::Flt4x tmin{0, 0, 0, 0}, tmax{1000, 1000, 1000, 1000};
        for (int l = 0; l < 4; l++) {
            rays.Origin.X[l] = 20.0f + 50.0f * l + 3.0f * p; rays.Origin.Y[l] = 30.0f + 40.0f * p; rays.Origin.Z[l] = 200.0f;
            rays.Direction.X[l] = dirs[(l + p) % 4][0]; rays.Direction.Y[l] = dirs[(l + p) % 4][1]; rays.Direction.Z[l] = dirs[(l + p) % 4][2];
        }
        raytracer::RayTracingResult res;
        env.Trace4Rays(rays, tmin, tmax, &res);
        for (int l = 0; l < 4; l++)
            std::printf("%d %d %d %.9g %.9g %.9g %.9g\n", p, l, res.HitIds[l], res.HitDistance[l],
                        res.SurfaceNormal.X[l], res.SurfaceNormal.Y[l], res.SurfaceNormal.Z[l]);
    }
    raytracer::FourVectors a, b;
    for (int l = 0; l < 4; l++) { a.X[l] = 10; a.Y[l] = 10 + 60.0f * l; a.Z[l] = 10; b.X[l] = 250; b.Y[l] = 240 - 50.0f * l; b.Z[l] = 20 + 60.0f * l; }
    raytracer::Flt4x vis;
    trace::TestLineDoesHitSky(env, a, b, &vis);
    std::printf("vis %g %g %g %g\n", vis[0], vis[1], vis[2], vis[3]);

    // second scenario: the hand-made sky scene of vrad_b200/scenes.py (mini_sky_scene): floor, TRACE_ID_SKY ceiling, a static prop
    // (id 7) and three transparent panes with coverage colours, traced with textureShadows on through the FourVectors mirror
    {
        raytracer::Environment sky;
        auto quad = [&](int32_t id, float x0, float x1, float y0, float y1, float z, uint16_t flags, float cov) {
            sky.AddTriangleWithMaterial(id, {x0, y0, z}, {x1, y0, z}, {x1, y1, z}, {cov, 0, 0}, flags, 0);
            sky.AddTriangleWithMaterial(id, {x0, y0, z}, {x1, y1, z}, {x0, y1, z}, {cov, 0, 0}, flags, 0);
        };
        quad(raytracer::TRACE_ID_OPAQUE, -512, 512, -512, 512, 0, 0, 1);
        quad(raytracer::TRACE_ID_SKY, -512, 512, -512, 512, 512, 0, 1);
        quad(raytracer::TRACE_ID_STATICPROP | 7, -60, 60, -60, 60, 300, 0, 1);
        quad(raytracer::TRACE_ID_OPAQUE, 100, 200, -50, 50, 400, VRAD_TRI_TRANSPARENT, 0.25f);
        quad(raytracer::TRACE_ID_OPAQUE, 100, 200, -50, 50, 420, VRAD_TRI_TRANSPARENT, 0.5f);
        quad(raytracer::TRACE_ID_OPAQUE, 150, 200, -50, 50, 440, VRAD_TRI_TRANSPARENT, 0.5f);
        sky.SetupAccelerationStructure();
        struct Case { float a[3], b[3]; bool shadows; int prop; };
        const Case cases[] = {{{0, 0, 100}, {0, 0, 5000}, false, -1}, {{0, 0, 100}, {0, 0, 5000}, false, 7}, {{120, 0, 100}, {120, 0, 5000}, false, -1},
                              {{120, 0, 100}, {120, 0, 5000}, true, -1}, {{120, 0, 100}, {120, 0, 410}, true, -1}, {{170, 0, 100}, {170, 0, 5000}, true, -1},
                              {{170, 0, 100}, {170, 0, 430}, true, -1}};
        for (const Case& c : cases) {
            raytracer::FourVectors s4, e4;
            for (int l = 0; l < 4; l++) { s4.X[l] = c.a[0]; s4.Y[l] = c.a[1]; s4.Z[l] = c.a[2]; e4.X[l] = c.b[0]; e4.Y[l] = c.b[1]; e4.Z[l] = c.b[2]; }
            trace::textureShadows = c.shadows;
            raytracer::Flt4x fv;
            trace::TestLineDoesHitSky(sky, s4, e4, &fv, true, c.prop);
            std::printf("sky %.9g %.9g %.9g %.9g\n", fv[0], fv[1], fv[2], fv[3]);
        }
        raytracer::Vec3 col = sky.GetTriangleColor(6);
        std::printf("colour %g\n", col[0]);
    }
    // third scenario (host only): MakePatchForFace + SubdividePatches on one 256 x 128 face, 16 units per luxel, chop 4
    {
        std::vector<float> pts = {0, 0, 0, 256, 0, 0, 256, 128, 0, 0, 128, 0};
        vrad_face_patch f{};
        f.first_point = 0; f.n_points = 4; f.normal[2] = 1.0f; f.plane_dist = 0.0f; f.lux_scale = 1.0f / 16.0f; f.chop = 4.0f;
        patches::PatchTree t = patches::SubdividePatches({f}, pts);
        int leaves = 0; float leaf_area = 0;
        for (int i = 0; i < t.size(); i++) if (t.child1[i] == -1) { leaves++; leaf_area += t.area[i]; }
        std::printf("patches %d leaves %d leaf_area %g child1 %d child2 %d\n", t.size(), leaves, leaf_area, t.child1[0], t.child2[0]);
    }
    return 0;
}
