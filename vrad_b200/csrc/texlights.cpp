// texlights.cpp -- lights.rad: the texture-light table that gives face patches their BaseLight, i.e. the producer of the surface
// lights K3 samples (vrad_lights_from_patches) and of the "no edge chop on surface lights" rule in the subdivision.  Host code.
//
// Reference map
//   common/parser/lights-rad/reader.go:19-120    Reader.Read: one entry per line -- "<material> r g b [scale [r g b scale]]",
//                                                "noshadow <material>", "forcetextureshadow <model>", hdr: / ldr: prefixes
//   common/parser/lights-rad/reader.go:122-186   lightForString (float32 parse, gamma 2.2 to linear, scale / 255; the HDR tuple wins
//                                                when both are given: useHDR := true at :126)
//   common/parser/lights-rad/reader.go:188-192   forceTextureShadowsOnModel
//   common/types/texlight.go:5-9, cache/texlights.go:5-12   TexLight, the cache
//   rad/patches/face.go:208-280                  BaseLightForFace / LightForTexture: texdata name -> texlight value
//   rad/patches/face.go:158-163                  a face with BaseLight gets SURF_LIGHT (which lifts PreventSubdivision's NOLIGHT rule)
// Intent adopted where the literal text is defective: reader.go:37-40 splits the file at '\r' only (an LF file would be one line) and
// :44-46 stops at the first empty line -- lines end at '\n' (a trailing '\r' is dropped) and empty lines are skipped; :49-60 tests
// `Contains("hdr:")` and never strips the prefix -- "hdr:" / "ldr:" are line prefixes selecting the entry for HDR / LDR compiles
// (upstream); :189-190 TrimLeft / TrimRight take cut SETS ("models/" would also eat a leading 'd') -- the prefix "models/" and the
// suffix ".mdl" are removed; face.go:274 `result = &...Value` rebinds a local pointer -- the value is copied; face.go:236-262 cannot
// cut the "_%d_%d_%d" suffix of cubemap-patched names (Go strings are immutable, the code is commented out) -- it is cut.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/vrad_bsp.h"

namespace vrad { void set_error(const char* fmt, ...); }

namespace {

constexpr int kMaxTexlights = 128;             // MAX_TEXLIGHTS (common/parser/lights-rad/rad.go)

// reader.go:122-186
bool light_for_string(const char* light, float out[3]) {
    out[0] = out[1] = out[2] = 0.0f;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n = std::sscanf(light, "%e %e %e %e %e %e %e %e", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
    if (n < 0) n = 0;
    float r = v[0], g = v[1], b = v[2], scaler = v[3];
    if (n == 8) { r = v[4]; g = v[5]; b = v[6]; scaler = v[7]; n = 4; }          // useHDR == true
    if (r < 0.0f || g < 0.0f || b < 0.0f || scaler < 0.0f) return false;
    out[0] = (float)(std::pow((double)(r / 255.0f), 2.2) * 255);
    if (n == 1) { out[2] = out[0]; out[1] = out[2]; }
    else if (n == 3 || n == 4) {
        out[1] = (float)(std::pow((double)(g / 255.0f), 2.2) * 255);
        out[2] = (float)(std::pow((double)(b / 255.0f), 2.2) * 255);
        if (n == 4) for (int k = 0; k < 3; k++) out[k] = out[k] * (scaler / 255.0f);
    } else { out[0] = 0.0f; return false; }                                      // "unknown light specifier type"
    return true;                                                                 // lightScale = 1
}

std::string first_token(const std::string& s, size_t* end) {
    size_t a = s.find_first_not_of(" \t");
    if (a == std::string::npos) { *end = s.size(); return ""; }
    size_t b = s.find_first_of(" \t", a);
    if (b == std::string::npos) b = s.size();
    *end = b;
    return s.substr(a, b - a);
}

bool append_name(char* buf, int64_t cap, int64_t& used, const std::string& name) {
    if (!buf) { used += (int64_t)name.size() + 1; return true; }
    if (used + (int64_t)name.size() + 1 > cap) return false;
    std::memcpy(buf + used, name.c_str(), name.size() + 1);
    used += (int64_t)name.size() + 1;
    return true;
}

}  // namespace

extern "C" int vrad_texlights_parse(const char* text, int64_t len, int hdr, int max_out, vrad_texlight* out, int* n_out,
                                    char* names_out, int64_t names_cap, int* n_noshadow, int* n_forced, int64_t* names_len) {
    if (!text || len < 0 || !n_out || max_out < 0 || (max_out > 0 && !out)) { vrad::set_error("vrad_texlights_parse: bad arguments"); return VRAD_E_INVALID; }
    std::vector<vrad_texlight> table;
    std::vector<std::string> noshadow, forced;
    int64_t pos = 0;
    while (pos < len) {
        int64_t e = pos;
        while (e < len && text[e] != '\n') e++;
        std::string line(text + pos, (size_t)(e - pos));
        pos = e + 1;
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        size_t lead = line.find_first_not_of(" \t");
        if (lead == std::string::npos) continue;
        line = line.substr(lead);
        if (line.compare(0, 4, "hdr:") == 0) { if (!hdr) continue; line = line.substr(4); }
        if (line.compare(0, 4, "ldr:") == 0) { if (hdr) continue; line = line.substr(4); }
        size_t tok_end = 0;
        const std::string tok = first_token(line, &tok_end);
        if (tok.empty()) continue;
        const std::string rest = line.substr(tok_end);
        if (tok == "noshadow") {
            size_t e2 = 0;
            std::string name = first_token(rest, &e2);
            if (name.empty()) continue;
            const size_t dot = name.find('.');                                   // reader.go:66-68: drop the extension
            if (dot != std::string::npos) name = name.substr(0, dot);
            noshadow.push_back(name);
        } else if (tok == "forcetextureshadow") {
            size_t e2 = 0;
            std::string name = first_token(rest, &e2);
            if (name.empty()) continue;
            if (name.compare(0, 7, "models/") == 0) name = name.substr(7);
            if (name.size() >= 4 && name.compare(name.size() - 4, 4, ".mdl") == 0) name = name.substr(0, name.size() - 4);
            forced.push_back(name);
        } else {
            float value[3];
            size_t a = rest.find_first_not_of(" \t");
            if (a == std::string::npos || !light_for_string(rest.c_str() + a, value)) continue;       // "ignoring bad texlight"
            int j = 0;
            for (; j < (int)table.size(); j++) if (tok == table[j].name) break;  // a later definition overrides (reader.go:89-104)
            if (j == (int)table.size()) {
                if ((int)table.size() == kMaxTexlights) { vrad::set_error("Too many texlights, max = %d", kMaxTexlights); return VRAD_E_INVALID; }
                if (tok.size() >= sizeof(table[0].name)) { vrad::set_error("texlight name longer than %zu characters: %s", sizeof(table[0].name) - 1, tok.c_str()); return VRAD_E_INVALID; }
                vrad_texlight t;
                std::memset(&t, 0, sizeof t);
                std::memcpy(t.name, tok.c_str(), tok.size());
                table.push_back(t);
            }
            for (int k = 0; k < 3; k++) table[j].value[k] = value[k];
        }
    }
    *n_out = (int)table.size();
    if (n_noshadow) *n_noshadow = (int)noshadow.size();
    if (n_forced) *n_forced = (int)forced.size();
    int64_t used = 0;
    bool fits = true;
    for (const std::string& s : noshadow) fits = append_name(names_out, names_cap, used, s) && fits;
    for (const std::string& s : forced) fits = append_name(names_out, names_cap, used, s) && fits;
    if (names_len) *names_len = used;
    if (out) {
        if ((int)table.size() > max_out) { vrad::set_error("vrad_texlights_parse: %zu texlights, room for %d", table.size(), max_out); return VRAD_E_NOMEM; }
        if (!table.empty()) std::memcpy(out, table.data(), table.size() * sizeof(vrad_texlight));
    }
    if (!fits) { vrad::set_error("vrad_texlights_parse: name buffer too small"); return VRAD_E_NOMEM; }
    return VRAD_OK;
}

extern "C" int vrad_bsp_apply_texlights(const vrad_bsp_lumps* L, const int32_t* string_table, int n_strings, const char* string_data, int64_t string_len,
                                        const char* map_name, int n_texlights, const vrad_texlight* texlights,
                                        int n_faces, const int32_t* face_number, vrad_face_patch* faces_inout, float* base_light3_out) {
    if (!L || n_faces < 0 || (n_faces && (!face_number || !base_light3_out)) || n_strings < 0 || (n_strings && (!string_table || !string_data)) ||
        n_texlights < 0 || (n_texlights && !texlights)) { vrad::set_error("vrad_bsp_apply_texlights: bad arguments"); return VRAD_E_INVALID; }
    const std::string map = map_name ? map_name : "";
    for (int i = 0; i < n_faces; i++) {
        float* light = base_light3_out + 3 * (size_t)i;
        light[0] = light[1] = light[2] = 0.0f;
        const int fn = face_number[i];
        if (fn < 0 || fn >= L->n_faces) { vrad::set_error("vrad_bsp_apply_texlights: face %d of %d", fn, L->n_faces); return VRAD_E_INVALID; }
        const vrad_texinfo& tx = L->texinfo[L->faces[fn].texinfo];
        const int id = L->texdata[tx.texdata].name_id;
        if (id < 0 || id >= n_strings) continue;
        const int32_t ofs = string_table[id];
        if (ofs < 0 || ofs >= string_len) continue;
        std::string name(string_data + ofs, strnlen(string_data + ofs, (size_t)(string_len - ofs)));
        // LightForTexture (face.go:232-280): "maps/<map>/<original>_%d_%d_%d" (cubemap-patched material) -> <original>
        const std::string prefix = "maps/" + map + "/";
        if (!map.empty() && name.compare(0, prefix.size(), prefix) == 0) {
            std::string base = name.substr(prefix.size());
            bool found = true;
            for (int k = 0; k < 3 && found; k++) {
                const size_t us = base.rfind('_');
                if (us == std::string::npos) found = false; else base = base.substr(0, us);
            }
            if (found) name = base;
        }
        for (int t = 0; t < n_texlights; t++)
            if (name == texlights[t].name) { for (int k = 0; k < 3; k++) light[k] = texlights[t].value[k]; break; }
        if (faces_inout && (light[0] != 0.0f || light[1] != 0.0f || light[2] != 0.0f)) {
            faces_inout[i].has_base_light = 1;
            // face.go:158-163: the texinfo gains SURF_LIGHT, so PreventSubdivision's "NOLIGHT and not LIGHT" no longer holds
            if (!(tx.flags & VRAD_SURF_NOCHOP)) faces_inout[i].no_subdivide = 0;
        }
    }
    return VRAD_OK;
}
