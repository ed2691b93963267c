// bsp_file.cpp -- the .bsp container either side of the path: read a Source BSP v20 file
// into 64 opaque lumps, hand typed views of the ones the path consumes to the input code (bsp_input.cpp), take
// replaced lumps (lighting, faces, vertex normals) back and write the file.  Pure host code.
//
// Reference map
//   cmd/tasks/loadbsp/main.go:163-170   loadBSP: bsp.NewReader(file).Read()  (github.com/galaco/bsp, Gopkg.lock:24-28)
//   cache/bsp.go:51-91                   BuildLumpCache: which lumps the program reads, and as what
//   cmd/tasks/finish/main.go:15-18       "Writing %s" -- the bsp.Writer call the reference leaves commented out
// The format itself is not in the reference tree (unvendored dependency): header = int32 ident "VBSP", int32
// version, 64 x {int32 fileofs, int32 filelen, int32 version, char fourCC[4]}, int32 mapRevision = 1036 bytes;
// lump payloads anywhere after it.  Written back in lump-index order, each payload 4-byte aligned.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/vrad_bsp.h"

namespace vrad { void set_error(const char* fmt, ...); }

struct vrad_bspfile {
    int32_t version = VRAD_BSP_VERSION;
    int32_t map_revision = 0;
    std::vector<uint8_t> lump[VRAD_HEADER_LUMPS];
    int32_t lump_version[VRAD_HEADER_LUMPS] = {};
    uint8_t fourcc[VRAD_HEADER_LUMPS][4] = {};
    int face_lump = VRAD_LUMP_FACES;       // cache.SetTargetFaces: LUMP_FACES, or LUMP_FACES_HDR for an HDR compile
};

namespace {

struct LumpDir { int32_t fileofs, filelen, version; uint8_t fourcc[4]; };
static_assert(sizeof(LumpDir) == 16, "lump_t is 16 bytes");
constexpr size_t kHeaderBytes = 8 + VRAD_HEADER_LUMPS * sizeof(LumpDir) + 4;

static_assert(sizeof(vrad_dplane) == 20 && sizeof(vrad_dedge) == 4 && sizeof(vrad_dface) == 56 && sizeof(vrad_texinfo) == 72 &&
              sizeof(vrad_dtexdata) == 32 && sizeof(vrad_dmodel) == 48 && sizeof(vrad_dnode) == 32 && sizeof(vrad_dleaf) == 32 &&
              sizeof(vrad_dbrush) == 12 && sizeof(vrad_dbrushside) == 8 && sizeof(vrad_color_rgbexp32) == 4,
              "lump records must have their on-disk sizes");

template <typename T>
bool view(const vrad_bspfile* f, int lump, const char* name, int32_t* n, const T** p) {
    const std::vector<uint8_t>& b = f->lump[lump];
    if (b.size() % sizeof(T)) { vrad::set_error("bsp: %s lump is %zu bytes, not a multiple of %zu", name, b.size(), sizeof(T)); return false; }
    *n = (int32_t)(b.size() / sizeof(T));
    *p = b.empty() ? nullptr : reinterpret_cast<const T*>(b.data());
    return true;
}

}  // namespace

extern "C" int vrad_bspfile_create(int map_revision, vrad_bspfile** out) {
    if (!out) { vrad::set_error("vrad_bspfile_create: bad arguments"); return VRAD_E_INVALID; }
    vrad_bspfile* f = new vrad_bspfile();
    f->map_revision = map_revision;
    f->lump_version[VRAD_LUMP_LEAFS] = 1;
    *out = f;
    return VRAD_OK;
}

extern "C" int vrad_bspfile_open(const char* path, vrad_bspfile** out) {
    if (!path || !out) { vrad::set_error("vrad_bspfile_open: bad arguments"); return VRAD_E_INVALID; }
    *out = nullptr;
    FILE* fp = std::fopen(path, "rb");
    if (!fp) { vrad::set_error("vrad_bspfile_open: cannot open %s", path); return VRAD_E_INVALID; }
    std::vector<uint8_t> bytes;
    std::fseek(fp, 0, SEEK_END);
    const long size = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    if (size > 0) { bytes.resize((size_t)size); if (std::fread(bytes.data(), 1, bytes.size(), fp) != bytes.size()) bytes.clear(); }
    std::fclose(fp);
    if (bytes.size() < kHeaderBytes) { vrad::set_error("vrad_bspfile_open: %s is %zu bytes, shorter than a BSP header", path, bytes.size()); return VRAD_E_INVALID; }
    int32_t ident, version;
    std::memcpy(&ident, bytes.data(), 4); std::memcpy(&version, bytes.data() + 4, 4);
    if (ident != VRAD_BSP_IDENT) { vrad::set_error("vrad_bspfile_open: %s is not a VBSP file (ident 0x%08x)", path, (unsigned)ident); return VRAD_E_INVALID; }
    if (version < 19 || version > 21) { vrad::set_error("vrad_bspfile_open: BSP version %d not supported (19..21)", version); return VRAD_E_UNSUPPORTED; }
    vrad_bspfile* f = new vrad_bspfile();
    f->version = version;
    std::memcpy(&f->map_revision, bytes.data() + kHeaderBytes - 4, 4);
    for (int i = 0; i < VRAD_HEADER_LUMPS; i++) {
        LumpDir d;
        std::memcpy(&d, bytes.data() + 8 + i * sizeof(LumpDir), sizeof d);
        if (d.filelen < 0 || d.fileofs < 0 || (uint64_t)d.fileofs + (uint64_t)d.filelen > bytes.size()) {
            vrad::set_error("vrad_bspfile_open: lump %d (%d bytes at %d) runs past the end of the file", i, d.filelen, d.fileofs);
            delete f; return VRAD_E_INVALID;
        }
        f->lump[i].assign(bytes.begin() + d.fileofs, bytes.begin() + d.fileofs + d.filelen);
        f->lump_version[i] = d.version;
        std::memcpy(f->fourcc[i], d.fourcc, 4);
    }
    *out = f;
    return VRAD_OK;
}

extern "C" int vrad_bspfile_get_lump(vrad_bspfile* f, int lump, const void** data, int64_t* len, int* lump_version) {
    if (!f || lump < 0 || lump >= VRAD_HEADER_LUMPS) { vrad::set_error("vrad_bspfile_get_lump: bad arguments"); return VRAD_E_INVALID; }
    if (data) *data = f->lump[lump].empty() ? nullptr : f->lump[lump].data();
    if (len) *len = (int64_t)f->lump[lump].size();
    if (lump_version) *lump_version = f->lump_version[lump];
    return VRAD_OK;
}

extern "C" int vrad_bspfile_set_lump(vrad_bspfile* f, int lump, const void* data, int64_t len, int lump_version) {
    if (!f || lump < 0 || lump >= VRAD_HEADER_LUMPS || len < 0 || (len > 0 && !data)) { vrad::set_error("vrad_bspfile_set_lump: bad arguments"); return VRAD_E_INVALID; }
    if (len > INT32_MAX) { vrad::set_error("vrad_bspfile_set_lump: a lump cannot exceed 2 GiB"); return VRAD_E_INVALID; }
    const uint8_t* p = static_cast<const uint8_t*>(data);
    std::vector<uint8_t> fresh(p, p + len);       // copy first: `data` may alias the lump being replaced
    f->lump[lump].swap(fresh);
    f->lump_version[lump] = lump_version;
    return VRAD_OK;
}

extern "C" int vrad_bspfile_save(vrad_bspfile* f, const char* path) {
    if (!f || !path) { vrad::set_error("vrad_bspfile_save: bad arguments"); return VRAD_E_INVALID; }
    std::vector<uint8_t> head(kHeaderBytes, 0);
    const int32_t ident = VRAD_BSP_IDENT;
    std::memcpy(head.data(), &ident, 4); std::memcpy(head.data() + 4, &f->version, 4);
    std::memcpy(head.data() + kHeaderBytes - 4, &f->map_revision, 4);
    uint64_t ofs = kHeaderBytes;
    for (int i = 0; i < VRAD_HEADER_LUMPS; i++) {
        LumpDir d = {0, (int32_t)f->lump[i].size(), f->lump_version[i], {0, 0, 0, 0}};
        std::memcpy(d.fourcc, f->fourcc[i], 4);
        if (!f->lump[i].empty()) {
            ofs = (ofs + 3) & ~uint64_t(3);
            if (ofs + f->lump[i].size() > (uint64_t)INT32_MAX) { vrad::set_error("vrad_bspfile_save: file would exceed 2 GiB"); return VRAD_E_INVALID; }
            d.fileofs = (int32_t)ofs;
            ofs += f->lump[i].size();
        }
        std::memcpy(head.data() + 8 + i * sizeof(LumpDir), &d, sizeof d);
    }
    FILE* fp = std::fopen(path, "wb");
    if (!fp) { vrad::set_error("vrad_bspfile_save: cannot create %s", path); return VRAD_E_INVALID; }
    bool ok = std::fwrite(head.data(), 1, head.size(), fp) == head.size();
    uint64_t pos = kHeaderBytes;
    static const uint8_t zeros[4] = {0, 0, 0, 0};
    for (int i = 0; i < VRAD_HEADER_LUMPS && ok; i++) {
        if (f->lump[i].empty()) continue;
        const uint64_t aligned = (pos + 3) & ~uint64_t(3);
        if (aligned != pos) ok = std::fwrite(zeros, 1, aligned - pos, fp) == aligned - pos;
        ok = ok && std::fwrite(f->lump[i].data(), 1, f->lump[i].size(), fp) == f->lump[i].size();
        pos = aligned + f->lump[i].size();
    }
    ok = (std::fclose(fp) == 0) && ok;
    if (!ok) { vrad::set_error("vrad_bspfile_save: short write to %s", path); return VRAD_E_INVALID; }
    return VRAD_OK;
}

extern "C" void vrad_bspfile_close(vrad_bspfile* f) { delete f; }

// cross-lump indices: everything the input code (bsp_input.cpp, bsp_light.cpp, texlights.cpp) dereferences is checked once, here
extern "C" int vrad_bsp_validate(const vrad_bsp_lumps* L) {
    if (!L) { vrad::set_error("vrad_bsp_validate: bad arguments"); return VRAD_E_INVALID; }
    if ((L->n_planes && !L->planes) || (L->n_vertexes && !L->vertexes3) || (L->n_edges && !L->edges) || (L->n_surfedges && !L->surfedges) ||
        (L->n_faces && !L->faces) || (L->n_texinfo && !L->texinfo) || (L->n_texdata && !L->texdata) || (L->n_models && !L->models) ||
        (L->n_nodes && !L->nodes) || (L->n_leafs && !L->leafs) || (L->n_leaffaces && !L->leaffaces) || (L->n_leafbrushes && !L->leafbrushes) ||
        (L->n_brushes && !L->brushes) || (L->n_brushsides && !L->brushsides) || (L->vis_len && !L->visdata)) {
        vrad::set_error("bsp: a lump has a count but no data"); return VRAD_E_INVALID;
    }
    auto bad = [](const char* what, int i, long v, long n) { vrad::set_error("bsp: %s %d refers to %ld, outside [0,%ld)", what, i, v, n); return VRAD_E_INVALID; };
    for (int i = 0; i < L->n_faces; i++) {
        const vrad_dface& fc = L->faces[i];
        if (fc.planenum >= L->n_planes) return bad("face plane of face", i, fc.planenum, L->n_planes);
        if (fc.texinfo < 0 || fc.texinfo >= L->n_texinfo) return bad("texinfo of face", i, fc.texinfo, L->n_texinfo);
        if (fc.numedges < 0 || fc.firstedge < 0 || (int64_t)fc.firstedge + fc.numedges > L->n_surfedges) return bad("surfedges of face", i, (long)fc.firstedge + fc.numedges, L->n_surfedges + 1);
    }
    for (int i = 0; i < L->n_surfedges; i++) {
        const int64_t e = L->surfedges[i] < 0 ? -(int64_t)L->surfedges[i] : L->surfedges[i];
        if (e >= L->n_edges) return bad("surfedge", i, (long)e, L->n_edges);
    }
    for (int i = 0; i < L->n_edges; i++)
        for (int k = 0; k < 2; k++) if (L->edges[i].v[k] >= L->n_vertexes) return bad("edge", i, L->edges[i].v[k], L->n_vertexes);
    for (int i = 0; i < L->n_texinfo; i++)
        if (L->texinfo[i].texdata < 0 || L->texinfo[i].texdata >= L->n_texdata) return bad("texdata of texinfo", i, L->texinfo[i].texdata, L->n_texdata);
    for (int i = 0; i < L->n_models; i++) {
        const vrad_dmodel& m = L->models[i];
        if (m.numfaces < 0 || m.firstface < 0 || (int64_t)m.firstface + m.numfaces > L->n_faces) return bad("faces of model", i, (long)m.firstface + m.numfaces, L->n_faces + 1);
        if (m.headnode >= L->n_nodes || (m.headnode < 0 && -1 - m.headnode >= L->n_leafs)) return bad("head node of model", i, m.headnode, L->n_nodes);
    }
    for (int i = 0; i < L->n_nodes; i++) {
        const vrad_dnode& nd = L->nodes[i];
        if (nd.planenum < 0 || nd.planenum >= L->n_planes) return bad("plane of node", i, nd.planenum, L->n_planes);
        for (int k = 0; k < 2; k++) {
            const int32_t c = nd.children[k];
            if (c >= L->n_nodes || (c < 0 && -1 - c >= L->n_leafs)) return bad("child of node", i, c, L->n_nodes);
        }
    }
    for (int i = 0; i < L->n_leafs; i++) {
        const vrad_dleaf& lf = L->leafs[i];
        if ((int)lf.firstleafface + lf.numleaffaces > L->n_leaffaces) return bad("leaf faces of leaf", i, (long)lf.firstleafface + lf.numleaffaces, L->n_leaffaces + 1);
        if ((int)lf.firstleafbrush + lf.numleafbrushes > L->n_leafbrushes) return bad("leaf brushes of leaf", i, (long)lf.firstleafbrush + lf.numleafbrushes, L->n_leafbrushes + 1);
    }
    for (int i = 0; i < L->n_leaffaces; i++) if (L->leaffaces[i] >= L->n_faces) return bad("leafface", i, L->leaffaces[i], L->n_faces);
    for (int i = 0; i < L->n_leafbrushes; i++) if (L->leafbrushes[i] >= L->n_brushes) return bad("leafbrush", i, L->leafbrushes[i], L->n_brushes);
    for (int i = 0; i < L->n_brushes; i++) {
        const vrad_dbrush& b = L->brushes[i];
        if (b.numsides < 0 || b.firstside < 0 || (int64_t)b.firstside + b.numsides > L->n_brushsides) return bad("sides of brush", i, (long)b.firstside + b.numsides, L->n_brushsides + 1);
    }
    for (int i = 0; i < L->n_brushsides; i++) {
        const vrad_dbrushside& s = L->brushsides[i];
        if (s.planenum >= L->n_planes || (int)(s.planenum ^ 1) >= L->n_planes) return bad("plane of brush side", i, s.planenum, L->n_planes);
        if (s.texinfo >= L->n_texinfo) return bad("texinfo of brush side", i, s.texinfo, L->n_texinfo);
    }
    if (L->vis_len) {
        if (L->vis_len < 4) { vrad::set_error("bsp: visibility lump is %lld bytes", (long long)L->vis_len); return VRAD_E_INVALID; }
        int32_t nc; std::memcpy(&nc, L->visdata, 4);
        if (nc < 0 || 4 + (int64_t)nc * 8 > L->vis_len) { vrad::set_error("bsp: visibility lump names %d clusters but holds %lld bytes", nc, (long long)L->vis_len); return VRAD_E_INVALID; }
    }
    // the node lump must be a tree under every model's head node: GetBrushRecursive and MakeParents recurse without a visited set
    std::vector<uint8_t> seen((size_t)L->n_nodes, 0);
    for (int m = 0; m < L->n_models; m++) {
        std::vector<int32_t> stack;
        if (L->models[m].headnode >= 0) stack.push_back(L->models[m].headnode);
        while (!stack.empty()) {
            const int32_t n = stack.back(); stack.pop_back();
            if (seen[n]) { vrad::set_error("bsp: node %d is reached twice (the node lump is not a forest)", n); return VRAD_E_INVALID; }
            seen[n] = 1;
            for (int k = 0; k < 2; k++) if (L->nodes[n].children[k] >= 0) stack.push_back(L->nodes[n].children[k]);
        }
    }
    return VRAD_OK;
}

extern "C" int vrad_bspfile_set_target_faces(vrad_bspfile* f, int hdr, int* face_lump_out, int* lighting_lump_out) {
    if (!f) { vrad::set_error("vrad_bspfile_set_target_faces: bad arguments"); return VRAD_E_INVALID; }
    // loadbsp.Main (cmd/tasks/loadbsp/main.go:79-89): HDR compiles light LUMP_FACES_HDR; when that lump is empty the LDR faces are
    // taken over (upstream copies dfaces into dfaces_hdr), so the HDR lump is seeded from LUMP_FACES here.
    if (hdr) {
        if (f->lump[VRAD_LUMP_FACES_HDR].empty()) {
            f->lump[VRAD_LUMP_FACES_HDR] = f->lump[VRAD_LUMP_FACES];
            f->lump_version[VRAD_LUMP_FACES_HDR] = f->lump_version[VRAD_LUMP_FACES];
        }
        f->face_lump = VRAD_LUMP_FACES_HDR;
    } else f->face_lump = VRAD_LUMP_FACES;
    if (face_lump_out) *face_lump_out = f->face_lump;
    if (lighting_lump_out) *lighting_lump_out = hdr ? VRAD_LUMP_LIGHTING_HDR : VRAD_LUMP_LIGHTING;
    return VRAD_OK;
}

extern "C" int vrad_bspfile_lumps(vrad_bspfile* f, vrad_bsp_lumps* L) {
    if (!f || !L) { vrad::set_error("vrad_bspfile_lumps: bad arguments"); return VRAD_E_INVALID; }
    std::memset(L, 0, sizeof *L);
    if (!f->lump[VRAD_LUMP_LEAFS].empty() && f->lump_version[VRAD_LUMP_LEAFS] != 1) {
        vrad::set_error("bsp: leaf lump version %d (only version 1, 32-byte leafs, is read)", f->lump_version[VRAD_LUMP_LEAFS]);
        return VRAD_E_UNSUPPORTED;
    }
    const std::vector<uint8_t>& vx = f->lump[VRAD_LUMP_VERTEXES];
    if (vx.size() % 12) { vrad::set_error("bsp: vertex lump is %zu bytes, not a multiple of 12", vx.size()); return VRAD_E_INVALID; }
    L->n_vertexes = (int32_t)(vx.size() / 12);
    L->vertexes3 = vx.empty() ? nullptr : reinterpret_cast<const float*>(vx.data());
    int32_t n_area_recs = 0; const uint8_t (*areas)[8] = nullptr;
    if (!view(f, VRAD_LUMP_PLANES, "plane", &L->n_planes, &L->planes) || !view(f, VRAD_LUMP_EDGES, "edge", &L->n_edges, &L->edges) ||
        !view(f, VRAD_LUMP_SURFEDGES, "surfedge", &L->n_surfedges, &L->surfedges) || !view(f, f->face_lump, "face", &L->n_faces, &L->faces) ||
        !view(f, VRAD_LUMP_TEXINFO, "texinfo", &L->n_texinfo, &L->texinfo) || !view(f, VRAD_LUMP_TEXDATA, "texdata", &L->n_texdata, &L->texdata) ||
        !view(f, VRAD_LUMP_MODELS, "model", &L->n_models, &L->models) || !view(f, VRAD_LUMP_NODES, "node", &L->n_nodes, &L->nodes) ||
        !view(f, VRAD_LUMP_LEAFS, "leaf", &L->n_leafs, &L->leafs) || !view(f, VRAD_LUMP_LEAFFACES, "leafface", &L->n_leaffaces, &L->leaffaces) ||
        !view(f, VRAD_LUMP_LEAFBRUSHES, "leafbrush", &L->n_leafbrushes, &L->leafbrushes) || !view(f, VRAD_LUMP_BRUSHES, "brush", &L->n_brushes, &L->brushes) ||
        !view(f, VRAD_LUMP_BRUSHSIDES, "brushside", &L->n_brushsides, &L->brushsides) || !view(f, VRAD_LUMP_AREAS, "area", &n_area_recs, &areas))
        return VRAD_E_INVALID;
    L->n_areas = n_area_recs;
    L->vis_len = (int64_t)f->lump[VRAD_LUMP_VISIBILITY].size();
    L->visdata = f->lump[VRAD_LUMP_VISIBILITY].empty() ? nullptr : f->lump[VRAD_LUMP_VISIBILITY].data();

    return vrad_bsp_validate(L);
}
