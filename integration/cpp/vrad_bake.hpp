// vrad_bake.hpp -- the whole job in C++ over the C-ABI: .bsp in -> lit .bsp out.  The host side of the reference is compiled code (Go, no
// toolchain in the image), so this is its C++ stand-in; vrad_b200/bake.py is the same sequence in Python and the tests hold the two
// against each other (tests/test_bsp_cpu.py: `drive --prepare` digests == bake.prepare's).
//
// Call order = the reference's tasks: loadbsp.Main (cmd/tasks/loadbsp/main.go:38-160: lumps, entities, shadow casters, ray-trace
// environment) -> rad.Start (rad/start.go:21-98: MakeParents / cluster table, MakePatches, PairEdges, SubdividePatches,
// CreateDirectLights) -> the RadWorld step the reference comments out (cmd/tasks/computerad/main.go:7: transfers, direct light,
// bounces) -> finish (cmd/tasks/finish/main.go:8-40: the lighting lump and the file).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "vrad_environment.hpp"

namespace bake {

using raytracer::fatal_on;
using Entity = std::vector<std::pair<std::string, std::string>>;            // key/value pairs in file order

// the entity lump: `{ "key" "value" ... }` blocks (what vmf.NewReader(...).Read() yields, loadbsp/main.go:176-181)
inline std::vector<Entity> ParseEntities(const std::string& text) {
    std::vector<Entity> ents;
    Entity cur;
    bool open = false;
    size_t i = 0;
    while (i < text.size()) {
        const char c = text[i];
        if (c == '{') { cur.clear(); open = true; i++; }
        else if (c == '}') { if (open) ents.push_back(cur); open = false; i++; }
        else if (c == '"') {
            const size_t j = text.find('"', i + 1);
            if (j == std::string::npos) break;
            const size_t k0 = text.find('"', j + 1);
            if (k0 == std::string::npos) break;
            const size_t k1 = text.find('"', k0 + 1);
            if (k1 == std::string::npos) break;
            if (open) cur.emplace_back(text.substr(i + 1, j - i - 1), text.substr(k0 + 1, k1 - k0 - 1));
            i = k1 + 1;
        } else i++;
    }
    return ents;
}
// Entity.ValueForKey (common/types/entity.go): a later duplicate of a key wins
inline const std::string* Value(const Entity& e, const char* key) {
    const std::string* v = nullptr;
    for (const auto& kv : e) if (kv.first == key) v = &kv.second;
    return v;
}
inline void Floats(const std::string* s, int n, float* out) {               // FloatForKey / VectorForKey: missing numbers are 0
    for (int k = 0; k < n; k++) out[k] = 0.0f;
    if (!s) return;
    std::istringstream in(*s);
    for (int k = 0; k < n; k++) { double v; if (!(in >> v)) break; out[k] = static_cast<float>(v); }
}

// light / light_spot / light_environment -> the records vrad_lights_from_entities takes
inline std::vector<vrad_light_entity> LightEntities(const std::vector<Entity>& ents) {
    std::vector<vrad_light_entity> out;
    for (const Entity& e : ents) {
        const std::string* cls = Value(e, "classname");
        if (!cls) continue;
        int code = *cls == "light" ? 0 : (*cls == "light_spot" ? 1 : (*cls == "light_environment" ? 2 : -1));
        if (code < 0) continue;
        vrad_light_entity r;
        std::memset(&r, 0, sizeof r);
        r.classname = code;
        Floats(Value(e, "origin"), 3, r.origin);
        if (const std::string* v = Value(e, "_light")) r.light_ok = vrad_light_for_string(v->c_str(), r.light) == 0;
        if (const std::string* v = Value(e, "_ambient")) r.ambient_ok = vrad_light_for_string(v->c_str(), r.ambient) == 0;
        if (const std::string* target = Value(e, "target"))
            for (const Entity& o : ents) { const std::string* name = Value(o, "targetname"); if (name && *name == *target) { r.has_target = 1; Floats(Value(o, "origin"), 3, r.target_origin); } }
        Floats(Value(e, "angles"), 3, r.angles);
        Floats(Value(e, "pitch"), 1, &r.pitch); Floats(Value(e, "angle"), 1, &r.angle);
        Floats(Value(e, "_inner_cone"), 1, &r.inner_cone); Floats(Value(e, "_cone"), 1, &r.cone); Floats(Value(e, "_exponent"), 1, &r.exponent);
        Floats(Value(e, "_fifty_percent_distance"), 1, &r.fifty_percent_distance); Floats(Value(e, "_zero_percent_distance"), 1, &r.zero_percent_distance);
        Floats(Value(e, "_constant_attn"), 1, &r.constant_attn); Floats(Value(e, "_linear_attn"), 1, &r.linear_attn); Floats(Value(e, "_quadratic_attn"), 1, &r.quadratic_attn);
        Floats(Value(e, "_distance"), 1, &r.distance);
        float hf; Floats(Value(e, "_hardfalloff"), 1, &hf); r.hardfalloff = static_cast<int32_t>(hf);
        out.push_back(r);
    }
    return out;
}

struct Prepared {
    std::vector<Entity> ents;
    loadbsp::RayTraceTriangles tris;
    patches::PatchTree tree;
    std::vector<int32_t> face_of_patch, cluster;
    std::vector<uint8_t> flags;                      // bit 0 = sky
    std::vector<uint8_t> needs_bump;                 // Patch.NeedsBumpMap (SURF_BUMPLIGHT)
    std::vector<float> bump_basis9;                  // upstream GetBumpNormals per bump-mapped patch (zero elsewhere)
    std::vector<float> refl3, face_origin3, centroids3;
    std::vector<uint8_t> pvs;                        // n_clusters x n_clusters (empty: no vis data)
    int n_clusters = 0;
    std::vector<uint8_t> sky_pvs;                    // merged PVS of the sky leafs (empty: none)
    std::vector<vrad_light> lights;
    std::vector<int32_t> mins2, size2;
    std::vector<vrad_dface> lit_faces;
    std::vector<vrad_texinfo> texinfo;               // the rescaled copy when -luxeldensity < 1 (empty otherwise)
    vrad_bsp_lumps lumps{};                          // the file's lumps with `faces` pointing at lit_faces
    std::vector<int64_t> luxel_first, radial_first;
    int64_t lump_bytes = 0;
    std::vector<float> lux_pos3, lux_normal3;
    std::vector<int32_t> lux_face;
    std::vector<vrad_radial_entry> radial;
    lightmap::FaceNeighbours neighbours;
};

// trace.PointLeafnum (raytracer/trace/pointleaf.go:8-33) on the host, for the per-model lookups
inline int PointCluster(const vrad_bsp_lumps& L, const float p[3]) {
    int node = L.models[0].headnode;
    while (node >= 0) {
        const vrad_dnode& nd = L.nodes[node];
        const vrad_dplane& pl = L.planes[nd.planenum];
        const float d = pl.type < 3 ? p[pl.type] - pl.dist : ((pl.normal[0] * p[0] + pl.normal[1] * p[1]) + pl.normal[2] * p[2]) - pl.dist;
        node = nd.children[d < 0 ? 1 : 0];
    }
    return L.leafs[-1 - node].cluster;
}

// everything the device stages take, from the lumps (host code of the library; no GPU needed)
// the command-line switches of the reference that reach this path (cmd/args.go:53-93)
struct Options {
    int   bounce = 8;                 // -bounce
    float luxelDensity = 1.0f;        // -luxeldensity
    float smoothDegrees = 45.0f;      // -smooth
    float chop = 4.0f, maxChop = 4.0f;   // -chop, -maxchop
    bool  hdr = false;                // -hdr
    bool  fast = false;               // -fast: here the binned kd build (RTE_FLAGS_FAST_TREE_GENERATION)
    bool  textureShadows = false;     // -textureshadows
    bool  polygonFormFactor = false;  // MakeTransfer: polygon-to-differential form factor for near pairs (vrad_patches_set_windings; upstream's rule, not in the reference)
    std::string lightsRad;            // -lights (path of a lights.rad; empty = none)
    float SmoothingThreshold() const { return smoothDegrees == 45.0f ? 0.7071067f : static_cast<float>(std::cos(static_cast<double>(smoothDegrees) * 3.14159265358979323846 / 180.0)); }
};

// texlights: the text of a lights.rad file plus the texdata string lumps (LUMP_TEXDATA_STRING_TABLE / _DATA) and the map's name
struct TexLights { std::string radText, mapName; const int32_t* stringTable = nullptr; int nStrings = 0; const char* stringData = nullptr; int64_t stringLen = 0; bool hdr = false; };

inline void Prepare(const vrad_bsp_lumps& Lfile, const std::string& entityText, Prepared& P, const TexLights* tex = nullptr,
                    float minChop = 4.0f, float maxChop = 4.0f, float smoothing = 0.7071067f, float luxelDensity = 1.0f) {
    vrad_bsp_lumps L = Lfile;
    if (luxelDensity < 1.0f) {                                            // rad.Start (rad/start.go:21-66): luxels no denser than -luxeldensity
        P.texinfo.assign(Lfile.texinfo, Lfile.texinfo + Lfile.n_texinfo);
        fatal_on(vrad_bsp_rescale_lightmap_vecs(Lfile.n_texinfo, P.texinfo.data(), luxelDensity), "vrad_bsp_rescale_lightmap_vecs");
        L.texinfo = P.texinfo.data();
    }
    P.ents = ParseEntities(entityText);
    // ExtractBrushEntityShadowCasters (main.go:186-211) + the "origin" of every brush model's entity (MakePatches, build.go:38-45)
    std::vector<int32_t> casterModel; std::vector<float> casterOrigin, casterAngles, modelOrigins(3 * static_cast<size_t>(L.n_models), 0.0f);
    for (const Entity& e : P.ents) {
        const std::string* model = Value(e, "model");
        if (!model || model->size() < 2 || (*model)[0] != '*') continue;
        const int m = std::atoi(model->c_str() + 1);
        float o[3], a[3];
        Floats(Value(e, "origin"), 3, o); Floats(Value(e, "angles"), 3, a);
        if (m >= 0 && m < L.n_models) for (int k = 0; k < 3; k++) modelOrigins[3 * m + k] = o[k];
        if (Value(e, "vrad_brush_cast_shadows")) { casterModel.push_back(m); casterOrigin.insert(casterOrigin.end(), o, o + 3); casterAngles.insert(casterAngles.end(), a, a + 3); }
    }
    P.tris = loadbsp::BrushesForRayTrace(L, casterModel, casterOrigin, casterAngles);
    patches::FacePatches fp = patches::MakePatches(L, modelOrigins, maxChop);
    // BaseLightForFace (rad/patches/face.go:208-280): faces whose material is a texlight emit, and are not edge-chopped
    std::vector<float> faceBaseLight(3 * fp.faces.size(), 0.0f);
    if (tex) {
        int nt = 0;
        fatal_on(vrad_texlights_parse(tex->radText.data(), static_cast<int64_t>(tex->radText.size()), tex->hdr, 0, nullptr, &nt, nullptr, 0, nullptr, nullptr, nullptr), "vrad_texlights_parse");
        std::vector<vrad_texlight> table(static_cast<size_t>(nt) + 1);
        fatal_on(vrad_texlights_parse(tex->radText.data(), static_cast<int64_t>(tex->radText.size()), tex->hdr, nt, table.data(), &nt, nullptr, 0, nullptr, nullptr, nullptr), "vrad_texlights_parse");
        fatal_on(vrad_bsp_apply_texlights(&L, tex->stringTable, tex->nStrings, tex->stringData, tex->stringLen, tex->mapName.c_str(), nt, table.data(),
                                          static_cast<int>(fp.faces.size()), fp.faceNumber.data(), fp.faces.data(), faceBaseLight.data()), "vrad_bsp_apply_texlights");
    }
    P.tree = patches::SubdividePatches(fp.faces, fp.points3, minChop);
    const int N = P.tree.size(), nf = L.n_faces;
    P.face_of_patch.resize(N); P.refl3.resize(3 * static_cast<size_t>(N)); P.flags.resize(N);
    for (int p = 0; p < N; p++) {
        const int f = P.tree.face[p];
        P.face_of_patch[p] = fp.faceNumber[f];
        for (int k = 0; k < 3; k++) P.refl3[3 * static_cast<size_t>(p) + k] = fp.reflectivity3[3 * static_cast<size_t>(f) + k];
        P.flags[p] = fp.faces[f].sky;
    }
    // PairEdges -> phong normals of the child patches (CreateChildPatch, subdivide.go:385); roots keep the plane normal
    P.neighbours = lightmap::PairEdges(L, smoothing);
    P.face_origin3.assign(3 * static_cast<size_t>(nf), 0.0f);
    for (int m = 0; m < L.n_models; m++)
        for (int j = 0; j < L.models[m].numfaces; j++) for (int k = 0; k < 3; k++) P.face_origin3[3 * static_cast<size_t>(L.models[m].firstface + j) + k] = modelOrigins[3 * m + k];
    P.centroids3.assign(3 * static_cast<size_t>(nf), 0.0f);
    std::vector<int32_t> kidFace; std::vector<float> kidPoint; std::vector<int> kids;
    for (int p = 0; p < N; p++) {
        const int f = P.face_of_patch[p];
        if (P.tree.parent[p] == -1) { for (int k = 0; k < 3; k++) P.centroids3[3 * static_cast<size_t>(f) + k] = P.tree.origin[3 * static_cast<size_t>(p) + k] - P.face_origin3[3 * static_cast<size_t>(f) + k]; }
        else { kids.push_back(p); kidFace.push_back(f); for (int k = 0; k < 3; k++) kidPoint.push_back(P.tree.origin[3 * static_cast<size_t>(p) + k] - P.face_origin3[3 * static_cast<size_t>(f) + k]); }
    }
    if (!kids.empty()) {
        std::vector<float> nrm(kidPoint.size());
        fatal_on(vrad_bsp_phong_normals(&L, smoothing, P.neighbours.vertexNormals3.data(), P.centroids3.data(), static_cast<int64_t>(kids.size()), kidFace.data(), kidPoint.data(), nrm.data()), "vrad_bsp_phong_normals");
        for (size_t i = 0; i < kids.size(); i++) for (int k = 0; k < 3; k++) P.tree.normal[3 * static_cast<size_t>(kids[i]) + k] = nrm[3 * i + k];
    }
    // Patch.NeedsBumpMap + the bump basis around the patch's (phong) normal
    P.needs_bump.assign(N, 0); P.bump_basis9.assign(9 * static_cast<size_t>(N), 0.0f);
    for (int p = 0; p < N; p++) {
        if (!fp.needsBump[P.tree.face[p]]) continue;
        P.needs_bump[p] = 1;
        const vrad_dface& fc = L.faces[P.face_of_patch[p]];
        const vrad_texinfo& tx = L.texinfo[fc.texinfo];
        fatal_on(vrad_bump_normals(tx.texture_vecs[0], tx.texture_vecs[1], L.planes[fc.planenum].normal, &P.tree.normal[3 * static_cast<size_t>(p)], &P.bump_basis9[9 * static_cast<size_t>(p)]), "vrad_bump_normals");
    }
    // cluster of a face = cluster of the first leaf that lists it; faces of brush models: the leaf their model sits in
    std::vector<int32_t> faceCluster(nf, -1);
    for (int l = 0; l < L.n_leafs; l++)
        for (int k = 0; k < L.leafs[l].numleaffaces; k++) { const int f = L.leaffaces[L.leafs[l].firstleafface + k]; if (faceCluster[f] < 0) faceCluster[f] = L.leafs[l].cluster; }
    for (int m = 1; m < L.n_models; m++) {
        float c[3];
        for (int k = 0; k < 3; k++) c[k] = modelOrigins[3 * m + k] + 0.5f * (L.models[m].mins[k] + L.models[m].maxs[k]);
        const int cl = PointCluster(L, c);
        for (int j = 0; j < L.models[m].numfaces; j++) faceCluster[L.models[m].firstface + j] = cl;
    }
    P.cluster.resize(N);
    for (int p = 0; p < N; p++) P.cluster[p] = faceCluster[P.face_of_patch[p]];
    // the visibility lump -> PVS matrix; the sky lights' merged PVS (BuildVisForLightEnvironment; radial-vis maps need the device)
    P.n_clusters = 0;
    if (L.vis_len >= 4) std::memcpy(&P.n_clusters, L.visdata, 4);
    if (P.n_clusters > 0) {
        P.pvs.resize(static_cast<size_t>(P.n_clusters) * P.n_clusters);
        fatal_on(vrad_pvs_from_vis_lump(P.n_clusters, reinterpret_cast<const int32_t*>(L.visdata + 4), L.visdata, L.vis_len, P.pvs.data()), "vrad_pvs_from_vis_lump");
    }
    bool radial = false;
    for (int l = 0; l < L.n_leafs; l++) radial = radial || ((static_cast<uint16_t>(L.leafs[l].area_flags) >> 9) & VRAD_LEAF_FLAGS_RADIAL);
    if (!radial) {
        std::vector<uint8_t> leafFlags(static_cast<size_t>(L.n_leafs) + 1), pvs((static_cast<size_t>(P.n_clusters) + 7) / 8 + 1);
        int has = 0;
        fatal_on(vrad_bsp_vis_for_light_environment(nullptr, &L, leafFlags.data(), pvs.data(), &has), "vrad_bsp_vis_for_light_environment");
        if (has) P.sky_pvs.assign(pvs.begin(), pvs.begin() + (P.n_clusters + 7) / 8);
    }
    // lightmap geometry: extents, lump layout, samples on their faces, phong normals of the samples
    P.mins2.resize(2 * static_cast<size_t>(nf)); P.size2.resize(2 * static_cast<size_t>(nf));
    int oversize = 0;
    fatal_on(vrad_bsp_face_extents(&L, P.mins2.data(), P.size2.data(), &oversize), "vrad_bsp_face_extents");
    P.lit_faces.resize(nf); P.luxel_first.resize(static_cast<size_t>(nf) + 1);
    fatal_on(vrad_bsp_layout_lighting(&L, P.mins2.data(), P.size2.data(), P.lit_faces.data(), P.luxel_first.data(), &P.lump_bytes), "vrad_bsp_layout_lighting");
    P.lumps = L; P.lumps.faces = P.lit_faces.data();
    const int64_t nl = P.luxel_first[nf];
    P.lux_pos3.resize(3 * static_cast<size_t>(nl) + 3); P.lux_normal3.resize(3 * static_cast<size_t>(nl) + 3); P.lux_face.resize(static_cast<size_t>(nl) + 1);
    fatal_on(vrad_bsp_face_luxels(&P.lumps, P.mins2.data(), P.size2.data(), P.face_origin3.data(), P.luxel_first.data(), P.lux_pos3.data(), P.lux_normal3.data(), P.lux_face.data()), "vrad_bsp_face_luxels");
    fatal_on(vrad_bsp_place_samples(&P.lumps, P.mins2.data(), P.size2.data(), P.face_origin3.data(), P.luxel_first.data(), P.lux_pos3.data(), nullptr), "vrad_bsp_place_samples");
    P.lux_pos3.resize(3 * static_cast<size_t>(nl)); P.lux_normal3.resize(3 * static_cast<size_t>(nl)); P.lux_face.resize(static_cast<size_t>(nl));
    {   // flat blocks only: the extra blocks of a bump-mapped face keep their bump basis
        std::vector<int64_t> idx; std::vector<int32_t> face; std::vector<float> pt;
        for (int64_t l = 0; l < nl; l++) {
            const int f = P.lux_face[l];
            const bool bumped = (L.texinfo[L.faces[f].texinfo].flags & VRAD_SURF_BUMPLIGHT) != 0;
            const int64_t perBlock = static_cast<int64_t>(P.size2[2 * static_cast<size_t>(f)] + 1) * (P.size2[2 * static_cast<size_t>(f) + 1] + 1);
            if (bumped && l - P.luxel_first[f] >= perBlock) continue;
            idx.push_back(l); face.push_back(f);
            for (int k = 0; k < 3; k++) pt.push_back((P.lux_pos3[3 * l + k] - P.lux_normal3[3 * l + k]) - P.face_origin3[3 * static_cast<size_t>(f) + k]);
        }
        if (!idx.empty()) {
            std::vector<float> nrm(pt.size());
            fatal_on(vrad_bsp_phong_normals(&L, smoothing, P.neighbours.vertexNormals3.data(), P.centroids3.data(), static_cast<int64_t>(idx.size()), face.data(), pt.data(), nrm.data()), "vrad_bsp_phong_normals");
            for (size_t i = 0; i < idx.size(); i++) for (int k = 0; k < 3; k++) P.lux_normal3[3 * idx[i] + k] = nrm[3 * i + k];
        }
    }
    // which patches light which face (radial filter)
    P.radial_first.resize(static_cast<size_t>(nf) + 1);
    int64_t ne = 0;
    const int32_t* nb = P.neighbours.neighbours.empty() ? nullptr : P.neighbours.neighbours.data();
    auto entries = [&](int64_t cap, vrad_radial_entry* out) {
        return vrad_bsp_radial_entries(&P.lumps, P.mins2.data(), P.face_origin3.data(), N, P.face_of_patch.data(), P.tree.child1.data(), P.tree.origin.data(),
                                       P.tree.wind_first.data(), P.tree.wind_count.data(), P.tree.wind_points.data(), nb ? P.neighbours.first.data() : nullptr, nb,
                                       cap, P.radial_first.data(), out, &ne);
    };
    fatal_on(entries(0, nullptr), "vrad_bsp_radial_entries");
    P.radial.resize(static_cast<size_t>(ne) + 1);
    fatal_on(entries(ne, P.radial.data()), "vrad_bsp_radial_entries");
    P.radial.resize(static_cast<size_t>(ne));
    // CreateDirectLights, entity part
    const std::vector<vrad_light_entity> le = LightEntities(P.ents);
    P.lights.resize(2 * le.size() + 2);
    int nLights = 0;
    fatal_on(vrad_lights_from_entities(static_cast<int>(le.size()), le.data(), static_cast<int>(P.lights.size()), P.lights.data(), &nLights), "vrad_lights_from_entities");
    P.lights.resize(nLights);
    // ... and the surface part (lights.go:49-82), which comes first in the list
    bool emits = false;
    for (float v : faceBaseLight) emits = emits || v != 0.0f;
    if (emits) {
        std::vector<float> base(3 * static_cast<size_t>(N)), scale(2 * static_cast<size_t>(N)), baseArea(N);
        for (int p = 0; p < N; p++) {
            const int f = P.tree.face[p];
            for (int k = 0; k < 3; k++) base[3 * static_cast<size_t>(p) + k] = faceBaseLight[3 * static_cast<size_t>(f) + k];
            scale[2 * static_cast<size_t>(p)] = fp.scale2[2 * static_cast<size_t>(f)]; scale[2 * static_cast<size_t>(p) + 1] = fp.scale2[2 * static_cast<size_t>(f) + 1];
            baseArea[p] = fp.baseArea[f];
        }
        std::vector<vrad_light> surf(static_cast<size_t>(N) + 1);
        int ns = 0;
        fatal_on(vrad_lights_from_patches(N, P.tree.origin.data(), P.tree.normal.data(), base.data(), P.tree.area.data(), scale.data(), baseArea.data(), P.tree.child1.data(),
                                          0.1f, static_cast<int>(surf.size()), surf.data(), &ns), "vrad_lights_from_patches");
        surf.resize(ns);
        surf.insert(surf.end(), P.lights.begin(), P.lights.end());
        P.lights.swap(surf);
    }
}

struct Lit { int64_t nnz = 0; int bounces = 0; std::vector<float> direct3, emit3, total3, bump9; };

// K3 for a block of points with DirectLight.PVS honoured: one device call per distinct light subset (see vrad_b200/bake.py)
inline void DirectLightCulled(raytracer::Environment& env, const Prepared& P, const std::vector<std::vector<uint8_t>>& lightSees, int64_t n, const float* pos3, const float* nrm3, float* out3) {
    std::fill(out3, out3 + 3 * n, 0.0f);
    const int nLights = static_cast<int>(P.lights.size());
    if (n == 0 || nLights == 0) return;
    if (lightSees.empty()) { fatal_on(vrad_direct_light(env.handle(), n, pos3, nrm3, nLights, P.lights.data(), out3), "vrad_direct_light"); return; }
    std::vector<int32_t> cl(n);
    fatal_on(vrad_cluster_from_point(env.handle(), n, pos3, cl.data()), "vrad_cluster_from_point");
    std::map<std::vector<uint8_t>, std::vector<int64_t>> groups;           // light mask -> the points it applies to
    std::vector<uint8_t> mask(nLights);
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < nLights; k++) mask[k] = (cl[i] < 0 || cl[i] >= P.n_clusters) ? 1 : lightSees[k][cl[i]];
        groups[mask].push_back(i);
    }
    for (const auto& g : groups) {
        std::vector<vrad_light> lights;
        for (int k = 0; k < nLights; k++) if (g.first[k]) lights.push_back(P.lights[k]);
        if (lights.empty()) continue;
        const std::vector<int64_t>& sel = g.second;
        std::vector<float> p(3 * sel.size()), nn(3 * sel.size()), rgb(3 * sel.size());
        for (size_t i = 0; i < sel.size(); i++) for (int k = 0; k < 3; k++) { p[3 * i + k] = pos3[3 * sel[i] + k]; nn[3 * i + k] = nrm3[3 * sel[i] + k]; }
        fatal_on(vrad_direct_light(env.handle(), static_cast<int64_t>(sel.size()), p.data(), nn.data(), static_cast<int>(lights.size()), lights.data(), rgb.data()), "vrad_direct_light");
        for (size_t i = 0; i < sel.size(); i++) for (int k = 0; k < 3; k++) out3[3 * sel[i] + k] = rgb[3 * i + k];
    }
}

// the Nodes / Planes / Leafs lumps trace.PointLeafnum and clustertable.ClusterFromPoint walk -> the environment
inline void UploadBsp(raytracer::Environment& env, const vrad_bsp_lumps& L) {
    std::vector<int32_t> nodePlane(L.n_nodes), nodeChildren(2 * static_cast<size_t>(L.n_nodes)), planeType(L.n_planes), leafCluster(L.n_leafs), leafArea(L.n_leafs);
    std::vector<float> planeNormal(3 * static_cast<size_t>(L.n_planes)), planeDist(L.n_planes);
    for (int i = 0; i < L.n_nodes; i++) { nodePlane[i] = L.nodes[i].planenum; nodeChildren[2 * i] = L.nodes[i].children[0]; nodeChildren[2 * i + 1] = L.nodes[i].children[1]; }
    for (int i = 0; i < L.n_planes; i++) { for (int k = 0; k < 3; k++) planeNormal[3 * i + k] = L.planes[i].normal[k]; planeDist[i] = L.planes[i].dist; planeType[i] = L.planes[i].type; }
    for (int i = 0; i < L.n_leafs; i++) { leafCluster[i] = L.leafs[i].cluster; leafArea[i] = L.leafs[i].area_flags & 0x1ff; }
    fatal_on(vrad_bsp_upload(env.handle(), L.n_nodes, nodePlane.data(), nodeChildren.data(), L.n_planes, planeNormal.data(), planeDist.data(), planeType.data(),
                             L.n_leafs, leafCluster.data(), leafArea.data(), std::max(L.n_areas, 1)), "vrad_bsp_upload");
}

// the device stages: geometry + kd build (K1), transfers (K2), direct light on luxels and patches (K3), bounces (K4)
inline Lit Light(raytracer::Environment& env, const Prepared& P, const std::vector<float>& skyDirs3, int bounces = 8, bool fastTree = false, bool textureShadows = false, bool polygonFormFactor = false) {
    const vrad_bsp_lumps& L = P.lumps;
    const int N = P.tree.size();
    std::vector<uint8_t> triFlags(P.tris.ids.size(), 0);
    fatal_on(vrad_env_add_triangles(env.handle(), static_cast<int>(P.tris.ids.size()), P.tris.ids.data(), P.tris.verts9.data(), triFlags.data()), "vrad_env_add_triangles");
    if (fastTree) fatal_on(vrad_env_build_fast(env.handle(), VRAD_BUILD_AUTO), "vrad_env_build_fast");
    else fatal_on(vrad_env_build(env.handle()), "vrad_env_build");
    if (textureShadows) fatal_on(vrad_set_light_trace_flags(env.handle(), VRAD_TL_TEXTURE_SHADOWS), "vrad_set_light_trace_flags");
    const bool haveVis = !P.pvs.empty();
    if (haveVis) UploadBsp(env, L);
    // Cluster per patch, AFTER subdivision (rad/patches/subdivide.go:92-116): ClusterFromPoint(patch.Origin) -- faces span leaves, so
    // the children of one face can sit in different clusters -- and for an origin in solid space (cluster -1) the first winding
    // point that is not.  A patch that stays at -1 is in no cluster: vrad_build_transfers leaves it out.
    std::vector<int32_t> cluster(N, 0);
    if (haveVis && N > 0) {
        fatal_on(vrad_cluster_from_point(env.handle(), N, P.tree.origin.data(), cluster.data()), "vrad_cluster_from_point");
        std::vector<float> wp; std::vector<int> owner;
        for (int p = 0; p < N; p++) if (cluster[p] < 0)
            for (int j = 0; j < P.tree.wind_count[p]; j++) {
                for (int k = 0; k < 3; k++) wp.push_back(P.tree.wind_points[3 * static_cast<size_t>(P.tree.wind_first[p] + j) + k]);
                owner.push_back(p);
            }
        if (!owner.empty()) {
            std::vector<int32_t> wc(owner.size());
            fatal_on(vrad_cluster_from_point(env.handle(), static_cast<int64_t>(owner.size()), wp.data(), wc.data()), "vrad_cluster_from_point");
            for (size_t q = 0; q < owner.size(); q++) if (cluster[owner[q]] < 0 && wc[q] >= 0) cluster[owner[q]] = wc[q];
        }
    }
    fatal_on(vrad_patches_upload(env.handle(), N, P.tree.origin.data(), P.tree.normal.data(), P.tree.plane_dist.data(), P.tree.area.data(), P.refl3.data(), cluster.data(), P.flags.data()), "vrad_patches_upload");
    fatal_on(vrad_patches_set_hierarchy(env.handle(), N, P.tree.parent.data(), P.tree.child1.data(), P.tree.child2.data(), P.tree.face.data()), "vrad_patches_set_hierarchy");
    if (polygonFormFactor)
        fatal_on(vrad_patches_set_windings(env.handle(), N, P.tree.wind_first.data(), P.tree.wind_count.data(), static_cast<int>(P.tree.wind_points.size() / 3), P.tree.wind_points.data()), "vrad_patches_set_windings");
    Lit out;
    bool bumped = false;
    for (uint8_t b : P.needs_bump) bumped = bumped || b;
    if (bumped) fatal_on(vrad_patches_set_bump(env.handle(), N, P.needs_bump.data(), P.bump_basis9.data()), "vrad_patches_set_bump");
    fatal_on(vrad_build_transfers(env.handle(), P.n_clusters, P.pvs.empty() ? nullptr : P.pvs.data(), &out.nnz), "vrad_build_transfers");
    bool ambient = false;
    for (const vrad_light& l : P.lights) ambient = ambient || l.type == 5;
    if (ambient) fatal_on(vrad_set_sky_dirs(env.handle(), static_cast<int>(skyDirs3.size() / 3), skyDirs3.data()), "vrad_set_sky_dirs");
    // DirectLight.PVS (AllocDLight / SetDLightVis, rad/lightmap/lights.go:118-161)
    std::vector<std::vector<uint8_t>> lightSees;
    if (!P.pvs.empty() && !P.lights.empty()) {
        const int nLights = static_cast<int>(P.lights.size());
        std::vector<float> lo(3 * static_cast<size_t>(nLights));
        for (int k = 0; k < nLights; k++) for (int a = 0; a < 3; a++) lo[3 * k + a] = P.lights[k].origin[a];
        std::vector<int32_t> lcl(nLights);
        fatal_on(vrad_cluster_from_point(env.handle(), nLights, lo.data(), lcl.data()), "vrad_cluster_from_point");
        lightSees.assign(nLights, std::vector<uint8_t>(P.n_clusters, 1));
        for (int k = 0; k < nLights; k++) {
            if (P.lights[k].type == 3 || P.lights[k].type == 5) {
                if (!P.sky_pvs.empty()) for (int c = 0; c < P.n_clusters; c++) lightSees[k][c] = (P.sky_pvs[c >> 3] >> (c & 7)) & 1;
            } else if (lcl[k] >= 0 && lcl[k] < P.n_clusters)
                for (int c = 0; c < P.n_clusters; c++) lightSees[k][c] = P.pvs[static_cast<size_t>(lcl[k]) * P.n_clusters + c] != 0;
        }
    }
    const int64_t nl = static_cast<int64_t>(P.lux_face.size());
    out.direct3.resize(3 * static_cast<size_t>(nl)); out.emit3.resize(3 * static_cast<size_t>(N)); out.total3.resize(3 * static_cast<size_t>(N));
    DirectLightCulled(env, P, lightSees, nl, P.lux_pos3.data(), P.lux_normal3.data(), out.direct3.data());
    std::vector<float> lifted(3 * static_cast<size_t>(N));
    for (size_t i = 0; i < lifted.size(); i++) lifted[i] = P.tree.origin[i] + P.tree.normal[i];
    DirectLightCulled(env, P, lightSees, N, lifted.data(), P.tree.normal.data(), out.emit3.data());
    float added[3];
    fatal_on(vrad_bounce(env.handle(), out.emit3.data(), bounces, 1, out.total3.data(), added, &out.bounces), "vrad_bounce");
    if (bumped) { out.bump9.resize(9 * static_cast<size_t>(N)); fatal_on(vrad_bounce_bump_totals(env.handle(), out.bump9.data()), "vrad_bounce_bump_totals"); }
    return out;
}

// radial filter + K5 + the lighting lump
inline std::vector<uint8_t> Finish(raytracer::Environment& env, const Prepared& P, const Lit& lit) {
    const int64_t nl = static_cast<int64_t>(P.lux_face.size());
    const int nf = P.lumps.n_faces, N = P.tree.size();
    std::vector<float> indirect(3 * static_cast<size_t>(nl) + 3);
    std::vector<vrad_radial_entry> none(1);
    fatal_on(vrad_luxel_radial_light(env.handle(), nl, P.lux_face.data(), nf, P.luxel_first.data(), P.size2.data(), P.radial_first.data(),
                                     P.radial.empty() ? none.data() : P.radial.data(), N, lit.total3.data(), lit.bump9.empty() ? nullptr : lit.bump9.data(), indirect.data()),
             "vrad_luxel_radial_light");
    std::vector<vrad_color_rgbexp32> colors(static_cast<size_t>(nl) + 1);
    fatal_on(vrad_lightmap_finalize(env.handle(), nl, lit.direct3.data(), indirect.data(), colors.data()), "vrad_lightmap_finalize");
    std::vector<uint8_t> lump(static_cast<size_t>(P.lump_bytes) + 1);
    fatal_on(vrad_bsp_pack_lighting(&P.lumps, P.luxel_first.data(), colors.data(), lump.data(), P.lump_bytes), "vrad_bsp_pack_lighting");
    lump.resize(static_cast<size_t>(P.lump_bytes));
    return lump;
}

inline std::vector<float> ReadSkyDirs(const std::string& path) {             // vmath.Anorms as text: 162 lines of x y z
    std::vector<float> d;
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;                      // comment lines, as numpy.loadtxt skips them
        std::istringstream ls(line);
        float v;
        while (ls >> v) d.push_back(v);
    }
    if (d.size() % 3 != 0 || d.empty()) { std::fprintf(stderr, "ReadSkyDirs: %s holds %zu numbers, expected triples\n", path.c_str(), d.size()); std::abort(); }
    return d;
}

// .bsp in -> lit .bsp out
inline std::string ReadText(const std::string& path) { std::ifstream in(path, std::ios::binary); std::stringstream ss; ss << in.rdbuf(); return ss.str(); }
inline std::string MapName(const std::string& path) {
    const size_t slash = path.find_last_of("/\\"), dot = path.rfind('.');
    const size_t a = slash == std::string::npos ? 0 : slash + 1;
    return path.substr(a, (dot == std::string::npos || dot < a) ? std::string::npos : dot - a);
}
// fills `tex` from the file's string lumps and a lights.rad path; the pointers stay valid while `bsp` lives
inline void LoadTexLights(const loadbsp::Bsp& bsp, const std::string& bspPath, const std::string& radPath, TexLights& tex) {
    const void* p = nullptr; int64_t n = 0;
    fatal_on(vrad_bspfile_get_lump(bsp.file, VRAD_LUMP_TEXDATA_STRING_TABLE, &p, &n, nullptr), "vrad_bspfile_get_lump");
    tex.stringTable = static_cast<const int32_t*>(p); tex.nStrings = static_cast<int>(n / 4);
    fatal_on(vrad_bspfile_get_lump(bsp.file, VRAD_LUMP_TEXDATA_STRING_DATA, &p, &n, nullptr), "vrad_bspfile_get_lump");
    tex.stringData = static_cast<const char*>(p); tex.stringLen = n;
    tex.radText = ReadText(radPath); tex.mapName = MapName(bspPath);
}

inline Lit BakeFile(const char* pathIn, const char* pathOut, const std::string& skyDirsPath, int device = 0, const Options& opt = Options()) {
    const char* lightsRadPath = opt.lightsRad.empty() ? nullptr : opt.lightsRad.c_str();
    const int bounces = opt.bounce;
    loadbsp::Bsp bsp(pathIn, opt.hdr);
    const void* ent = nullptr; int64_t entLen = 0;
    fatal_on(vrad_bspfile_get_lump(bsp.file, VRAD_LUMP_ENTITIES, &ent, &entLen, nullptr), "vrad_bspfile_get_lump");
    std::string text(static_cast<const char*>(ent), static_cast<size_t>(entLen));
    while (!text.empty() && text.back() == '\0') text.pop_back();
    Prepared P;
    TexLights tex;
    if (lightsRadPath) { LoadTexLights(bsp, pathIn, lightsRadPath, tex); tex.hdr = opt.hdr; }
    Prepare(bsp.lumps, text, P, lightsRadPath ? &tex : nullptr, opt.chop, opt.maxChop, opt.SmoothingThreshold(), opt.luxelDensity);
    raytracer::Environment env(device);
    const Lit lit = Light(env, P, ReadSkyDirs(skyDirsPath), bounces, opt.fast, opt.textureShadows, opt.polygonFormFactor);
    const std::vector<uint8_t> lump = Finish(env, P, lit);
    fatal_on(vrad_bspfile_set_lump(bsp.file, bsp.lightingLump, lump.data(), static_cast<int64_t>(lump.size()), 1), "vrad_bspfile_set_lump");
    // P.lumps.faces points at P.lit_faces (our copy), so replacing the face lump does not pull the rug from under it
    fatal_on(vrad_bspfile_set_lump(bsp.file, bsp.faceLump, P.lit_faces.data(), static_cast<int64_t>(P.lit_faces.size() * sizeof(vrad_dface)), 1), "vrad_bspfile_set_lump");
    if (!P.texinfo.empty())                                                   // the rescaled lightmap vectors go back into the file (rad/start.go:48-50)
        fatal_on(vrad_bspfile_set_lump(bsp.file, VRAD_LUMP_TEXINFO, P.texinfo.data(), static_cast<int64_t>(P.texinfo.size() * sizeof(vrad_texinfo)), 0), "vrad_bspfile_set_lump");
    {   // lightmap.SaveVertexNormals (rad/start.go:82-85): the vertex-normal lumps, "for use in the engine"
        const int nfv = static_cast<int>(P.neighbours.vertexNormals3.size() / 3) - 1;       // PairEdges leaves one spare entry
        std::vector<float> normals(3 * static_cast<size_t>(std::max(nfv, 1)));
        std::vector<uint16_t> indices(static_cast<size_t>(std::max(nfv, 1)));
        int nn = 0;
        fatal_on(vrad_bsp_save_vertex_normals(nfv, P.neighbours.vertexNormals3.data(), std::max(nfv, 1), normals.data(), indices.data(), &nn), "vrad_bsp_save_vertex_normals");
        fatal_on(vrad_bspfile_set_lump(bsp.file, VRAD_LUMP_VERTNORMALS, normals.data(), static_cast<int64_t>(nn) * 12, 0), "vrad_bspfile_set_lump");
        fatal_on(vrad_bspfile_set_lump(bsp.file, VRAD_LUMP_VERTNORMALINDICES, indices.data(), static_cast<int64_t>(nfv) * 2, 0), "vrad_bspfile_set_lump");
    }
    fatal_on(vrad_bspfile_save(bsp.file, pathOut), "vrad_bspfile_save");
    return lit;
}

// position-weighted 64-bit checksum over the 32-bit words of a buffer (sum of word_i * (2654435761 * i + 1), wrapping): cheap to
// vectorise on the Python side, order-sensitive; lets the tests compare what this header prepares with what vrad_b200/bake.py prepares
inline uint64_t Checksum(const void* data, size_t bytes) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint64_t h = 0;
    const size_t words = (bytes + 3) / 4;
    for (size_t i = 0; i < words; i++) {
        uint32_t w = 0;
        std::memcpy(&w, p + 4 * i, std::min<size_t>(4, bytes - 4 * i));
        h += static_cast<uint64_t>(w) * (2654435761ull * i + 1ull);
    }
    return h;
}
template <class T> inline uint64_t Checksum(const std::vector<T>& v) { return Checksum(v.data(), v.size() * sizeof(T)); }

}  // namespace bake
