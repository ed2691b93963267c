// light_setup.cpp -- host side of K3's inputs: the `vrad_light` records made from light entities and from
// light-emitting patches, i.e. the part of rad/lightmap/lights.go that sits right before the direct-light kernel
// (SURVEY section 8 a14).  Pure host code, no device needed.
//
// Reference map
//   common/types/entity.go:96-158        LightForKey / LightForString: "r g b [scale]" -> linear RGB intensity
//   rad/lightmap/lights.go:38-116        CreateDirectLights: surface lights from patches (:49-82), entity dispatch (:90-113)
//   rad/lightmap/lights.go:173-213       ParseLightGeneric: intensity, normal from a target or from angles/pitch/angle
//   rad/lightmap/lights.go:216-256       ParseLightSpot: cone angles -> cosines, 180/180 -> point light
//   rad/lightmap/lights.go:259-341       SetLightFalloffParams: 50 % / 0 % distances -> quadratic (vmath/quadratic/solver.go),
//                                        hard falloff / cap distance, or the legacy attenuation + unit-100 intensity scale
//   rad/lightmap/lights.go:343-372       SetupLightNormalFromProps
//   rad/lightmap/lights.go:374-416       ParseLightEnvironment: sky light + sky ambient (first light_environment only)
//   common/types/light.go:38-44          default fade distances
// Intent adopted where the literal text is defective: lights.go:360 `output[1] = cos` -> sin;
// solver.go:65-69 vSwap swaps nothing -> real swap (#21); entity.go:138-139 `case 3:` has no fallthrough in Go, so a
// 3-number "_light" would lose G and B -> cases 3 and 4 share the body as in the C original; lights.go:397-400 halves the
// sun into the ambient when "_ambient" PARSES -> when it is missing/invalid (upstream: `if (!LightForKey(...))`);
// lights.go:73-76 fatal when the normal IS long enough (#18) -> fatal when it is not.
// Output order: entity order (the reference prepends to a linked list; only the summation order in K3 depends on it).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/vrad_cuda.h"

namespace vrad { void set_error(const char* fmt, ...); }

namespace {

constexpr double kPi = 3.14159265358979323846;       // vmath/constants.go:5
constexpr float kEqualEpsilon = 0.001f;              // vmath/constants.go:9
constexpr float kDirectScale = 100.0f * 100.0f;      // lights.go:23

// vmath/quadratic/solver.go:48-63 -- fp32 throughout, Go's left-to-right evaluation
bool solve_inverse_quadratic(float x1, float y1, float x2, float y2, float x3, float y3, float& a, float& b, float& c) {
    const float det = (x1 - x2) * (x1 - x3) * (x2 - x3);
    if (det == 0.0f) return false;
    a = (x3 * (-y1 + y2) + x2 * (y1 - y3) + x1 * (-y2 + y3)) / det;
    b = (x3 * x3 * (y1 - y2) + x1 * x1 * (y2 - y3) + x2 * x2 * (-y1 + y3)) / det;
    c = (x1 * x3 * (-x1 + x3) * y2 + x2 * x2 * (x3 * y1 - x1 * y3) + x2 * (-(x3 * x3 * y1) + x1 * x1 * y3)) / det;
    return true;
}

// solver.go:4-45: pull the middle sample towards the chord until the fitted curve is monotonic at x = 1
bool solve_inverse_quadratic_monotonic(float x1, float y1, float x2, float y2, float x3, float y3, float& a, float& b, float& c) {
    auto order = [](float& xa, float& ya, float& xb, float& yb) { if (xa > xb) { float t = xa; xa = xb; xb = t; t = ya; ya = yb; yb = t; } };
    order(x1, y1, x2, y2); order(x2, y2, x3, y3); order(x1, y1, x2, y2);
    for (double blend = 0.0; blend <= 1.0; blend += 0.05) {
        const double chord = (double)(y1 + (y3 - y1) * (x2 - x1) / (x3 - x1));          // FLerp, lerp.go:14-16 (fp32, then widened)
        const double tempy2 = (1 - blend) * (double)y2 + blend * chord;
        if (!solve_inverse_quadratic(x1, y1, x2, (float)tempy2, x3, y3, a, b, c)) return false;
        const float derivative = 2.0f * a + b;
        if (y1 < y2 && y2 < y3) { if (derivative >= 0.0f) return true; }
        else if (y1 > y2 && y2 > y3) { if (derivative <= 0.0f) return true; }
        else return true;
    }
    return true;
}

void fill_defaults(vrad_light& L) {
    memset(&L, 0, sizeof(L));
    L.start_fade = 0.0f; L.end_fade = -1.0f; L.cap_dist = 1.0e22f;                       // NewDirectLight, light.go:38-44
}

// lights.go:343-372
void normal_from_props(const float angles[3], float angle, float pitch, float out[3]) {
    if (angle == -1.0f) { out[0] = 0; out[1] = 0; out[2] = 1; }                         // ANGLE_UP
    else if (angle == -2.0f) { out[0] = 0; out[1] = 0; out[2] = -1; }                   // ANGLE_DOWN
    else {
        if (angle == 0.0f) angle = angles[1];                                           // YAW
        out[2] = 0;
        out[0] = (float)std::cos((double)angle / 180 * kPi);
        out[1] = (float)std::sin((double)angle / 180 * kPi);
    }
    if (pitch == 0.0f) pitch = angles[0];                                               // PITCH
    out[2] = (float)std::sin((double)pitch / 180 * kPi);
    out[0] *= (float)std::cos((double)pitch / 180 * kPi);
    out[1] *= (float)std::cos((double)pitch / 180 * kPi);
}

// lights.go:173-213 (the HDR keys are off: useHDR = false, :28)
void parse_generic(const vrad_light_entity& E, vrad_light& L) {
    for (int k = 0; k < 3; k++) { L.origin[k] = E.origin[k]; L.intensity[k] = E.light_ok ? E.light[k] : 0.0f; }
    if (E.has_target) {
        float d[3] = {E.target_origin[0] - E.origin[0], E.target_origin[1] - E.origin[1], E.target_origin[2] - E.origin[2]};
        const float len = (float)std::sqrt((double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]));   // mgl32 Normalize: v * (1/Len)
        const float inv = 1.0f / len;
        for (int k = 0; k < 3; k++) L.normal[k] = d[k] * inv;
    } else {
        normal_from_props(E.angles, E.angle, E.pitch, L.normal);
    }
}

// lights.go:259-341
void falloff_params(const vrad_light_entity& E, vrad_light& L) {
    const float d50 = E.fifty_percent_distance;
    L.start_fade = 0.0f; L.end_fade = -1.0f; L.cap_dist = 1.0e22f;
    if (d50 != 0.0f) {
        float d0 = E.zero_percent_distance;
        if (d0 < d50) d0 = 2.0f * d50;
        float a = 0.0f, b = 1.0f, c = 0.0f;
        solve_inverse_quadratic_monotonic(0.0f, 1.0f, d50, 2.0f, d0, 256.0f, a, b, c);  // failure only logs (:274-276)
        const float v50 = c + d50 * (b + d50 * a);
        const float scale = 2.0f / v50;
        a *= scale; b *= scale; c *= scale;
        L.quadratic_attn = a; L.linear_attn = b; L.constant_attn = c;
        if (E.hardfalloff != 0) {
            L.end_fade = d0;
            L.start_fade = 0.75f * d0 + 0.25f * d50;
        } else if (std::fabs((double)a) > 0.0) {
            const float fl_max = b / (-2.0f * a);                                        // where f' = 0
            if (fl_max > 0.0f) { L.cap_dist = fl_max; L.start_fade = fl_max; L.end_fade = 10.0f * fl_max; }
        }
    } else {
        L.constant_attn = E.constant_attn; L.linear_attn = E.linear_attn; L.quadratic_attn = E.quadratic_attn;
        L.radius = E.distance;
        if (L.constant_attn < kEqualEpsilon) L.constant_attn = 0;
        if (L.linear_attn < kEqualEpsilon) L.linear_attn = 0;
        if (L.quadratic_attn < kEqualEpsilon) L.quadratic_attn = 0;
        if (L.constant_attn < kEqualEpsilon && L.linear_attn < kEqualEpsilon && L.quadratic_attn < kEqualEpsilon) L.constant_attn = 1;
        const float ratio = L.constant_attn + 100 * L.linear_attn + 100 * 100 * L.quadratic_attn;   // "scale intensity for unit 100 distance"
        if (ratio > 0) for (int k = 0; k < 3; k++) L.intensity[k] = L.intensity[k] * ratio;
    }
}

} // namespace

extern "C" {

int vrad_light_for_string(const char* value, float rgb_out[3]) {
    if (!value || !rgb_out) { vrad::set_error("vrad_light_for_string: bad arguments"); return VRAD_E_INVALID; }
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n = sscanf(value, "%lf %lf %lf %lf %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
    if (n < 0) n = 0;
    if (n == 8) n = 4;                                                                  // two 4-tuples: LDR first (useHDR = false)
    rgb_out[0] = rgb_out[1] = rgb_out[2] = 0.0f;
    if (v[0] < 0.0 || v[1] < 0.0 || v[2] < 0.0 || v[3] < 0.0) { vrad::set_error("invalid colour for light"); return VRAD_E_INVALID; }
    rgb_out[0] = (float)(std::pow(v[0] / 255.0, 2.2) * 255);                            // convert to linear (entity.go:130)
    if (n == 1) { rgb_out[1] = rgb_out[2] = rgb_out[0]; }
    else if (n == 3 || n == 4) {
        rgb_out[1] = (float)(std::pow(v[1] / 255.0, 2.2) * 255);
        rgb_out[2] = (float)(std::pow(v[2] / 255.0, 2.2) * 255);
        if (n == 4) { const float s = (float)(v[3] / 255.0); for (int k = 0; k < 3; k++) rgb_out[k] = rgb_out[k] * s; }
    } else {
        rgb_out[0] = 0.0f;
        vrad::set_error("unknown light specifier type - %s", value); return VRAD_E_INVALID;
    }
    return VRAD_OK;                                                                     // lightScale = 1 (lights.go:27)
}

int vrad_lights_from_entities(int n, const vrad_light_entity* ents, int max_out, vrad_light* out, int* n_out) {
    if (n < 0 || (n > 0 && !ents) || !n_out || (max_out > 0 && !out)) { vrad::set_error("vrad_lights_from_entities: bad arguments"); return VRAD_E_INVALID; }
    int m = 0;
    bool have_sky = false;
    auto emit = [&](const vrad_light& L) { if (m < max_out) out[m] = L; m++; };
    for (int i = 0; i < n; i++) {
        const vrad_light_entity& E = ents[i];
        vrad_light L;
        fill_defaults(L);
        if (E.classname == 0) {                                                         // "light": ParseLightPoint (:418-426)
            parse_generic(E, L);
            L.type = 1;
            falloff_params(E, L);
            emit(L);
        } else if (E.classname == 1) {                                                  // "light_spot": ParseLightSpot (:216-256)
            parse_generic(E, L);
            L.type = 2;
            L.stopdot = E.inner_cone;
            if (L.stopdot == 0) L.stopdot = 10;
            L.stopdot2 = E.cone;
            if (L.stopdot2 == 0) L.stopdot2 = L.stopdot;
            if (L.stopdot2 < L.stopdot) L.stopdot2 = L.stopdot;
            if (L.stopdot == 180 && L.stopdot2 == 180) {                                // "This is a point light if stop dots are 180"
                L.stopdot = L.stopdot2 = 0; L.type = 1; L.exponent = 0;
            } else {
                if (L.stopdot > 90) L.stopdot = 90;                                     // "Clamp to 90, that's all DX8 can handle!"
                if (L.stopdot2 > 90) L.stopdot2 = 90;
                L.stopdot2 = (float)std::cos((double)(L.stopdot2 / 180 * (float)kPi));
                L.stopdot = (float)std::cos((double)(L.stopdot / 180 * (float)kPi));
                L.exponent = E.exponent;
            }
            falloff_params(E, L);
            emit(L);
        } else if (E.classname == 2) {                                                  // "light_environment" (:374-416)
            if (have_sky) continue;                                                     // only the first one (globalSkyLight == nil)
            have_sky = true;
            parse_generic(E, L);
            L.type = 3;
            emit(L);
            vrad_light A;
            fill_defaults(A);
            A.type = 5;
            for (int k = 0; k < 3; k++) { A.origin[k] = L.origin[k]; A.intensity[k] = E.ambient_ok ? E.ambient[k] : L.intensity[k] * 0.5f; }
            emit(A);
        } else {
            vrad::set_error("unsupported light entity class %d (entity %d)", E.classname, i); return VRAD_E_INVALID;
        }
    }
    *n_out = m;
    if (m > max_out) { vrad::set_error("vrad_lights_from_entities: %d lights, capacity %d", m, max_out); return VRAD_E_NOMEM; }
    return VRAD_OK;
}

int vrad_lights_from_patches(int n, const float* origin3, const float* normal3, const float* base_light3, const float* area,
                             const float* scale2, const float* base_area, const int32_t* child1, float light_threshold,
                             int max_out, vrad_light* out, int* n_out) {
    if (n < 0 || !n_out || (n > 0 && (!origin3 || !normal3 || !base_light3 || !area || !scale2 || !base_area)) || (max_out > 0 && !out)) {
        vrad::set_error("vrad_lights_from_patches: bad arguments"); return VRAD_E_INVALID;
    }
    int m = 0;
    for (int i = 0; i < n; i++) {                                                       // lights.go:49-82
        if (child1 && child1[i] != -1) continue;                                        // skip parent patches
        if (base_area[i] < 1e-6f) continue;
        const float* bl = base_light3 + 3 * (size_t)i;
        if ((double)((bl[0] + bl[1] + bl[2]) / 3) < (double)light_threshold) continue;  // vector.Avg >= lightThreshold
        const float* nr = normal3 + 3 * (size_t)i;
        if (!((float)std::sqrt((double)(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2])) > 1.0e-20f)) {
            vrad::set_error("Patch normal out of bounds during DirectLight creation (patch %d)", i); return VRAD_E_INVALID;
        }
        vrad_light L;
        fill_defaults(L);
        L.type = 0;
        const float s = 1.0f * area[i] * scale2[2 * (size_t)i] * scale2[2 * (size_t)i + 1] / base_area[i];   // lightScale * Area * Scale[0] * Scale[1] / BaseArea
        for (int k = 0; k < 3; k++) {
            L.origin[k] = origin3[3 * (size_t)i + k]; L.normal[k] = nr[k];
            L.intensity[k] = (bl[k] * s) * kDirectScale;                                // "scale to a range that results in actual light"
        }
        if (m < max_out) out[m] = L;
        m++;
    }
    *n_out = m;
    if (m > max_out) { vrad::set_error("vrad_lights_from_patches: %d lights, capacity %d", m, max_out); return VRAD_E_NOMEM; }
    return VRAD_OK;
}

} // extern "C"
