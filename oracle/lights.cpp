// lights.cpp -- ORACLE (test infrastructure): light creation, restated function by function from
//   common/types/entity.go:96-158 (LightForKey / LightForString), rad/lightmap/lights.go:38-116 (CreateDirectLights),
//   :173-213 (ParseLightGeneric), :216-256 (ParseLightSpot), :259-341 (SetLightFalloffParams),
//   :343-372 (SetupLightNormalFromProps), :374-416 (ParseLightEnvironment), :418-426 (ParseLightPoint),
//   vmath/quadratic/solver.go:4-69, vmath/quadratic/lerp.go:14-16, common/types/light.go:38-44.
// Intent adopted for defects in the literal text: App. A #18 (inverted normal assertion), #19 (Y = sin), #21 (vSwap),
// the missing fallthrough of `case 3:` in LightForString, and the inverted `_ambient` test (see light_setup.cpp header).
// Lights come out in entity order.  useHDR = false, lightScale = 1.
#include "oracle_impl.hpp"
#include <cstdio>

namespace orc {

#pragma pack(push, 4)
struct LightEntity {                       // same layout as vrad_light_entity (include/vrad_cuda.h), 124 bytes
    int32_t classname; float origin[3];
    int32_t light_ok; float light[3];
    int32_t has_target; float target_origin[3];
    float angles[3], pitch, angle;
    float inner_cone, cone, exponent;
    float fifty_percent_distance, zero_percent_distance; int32_t hardfalloff;
    float constant_attn, linear_attn, quadratic_attn, distance;
    int32_t ambient_ok; float ambient[3];
};
#pragma pack(pop)
static_assert(sizeof(LightEntity) == 124, "layout");

static const double PI = 3.14159265358979323846;
static const float EQUAL_EPSILON = 0.001f;
static const float DIRECT_SCALE = 100.0f * 100.0f;

static void vSwap(float* a, float* b) { float c = *a; *a = *b; *b = c; }

static double FLerp(float f1, float f2, float i1, float i2, float x) { return (double)(f1 + (f2 - f1) * (x - i1) / (i2 - i1)); }

static bool SolveInverseQuadratic(float x1, float y1, float x2, float y2, float x3, float y3, float* a, float* b, float* c) {
    float det = (x1 - x2) * (x1 - x3) * (x2 - x3);
    if (det == 0.0f) return false;
    *a = (x3 * (-y1 + y2) + x2 * (y1 - y3) + x1 * (-y2 + y3)) / det;
    *b = (x3 * x3 * (y1 - y2) + x1 * x1 * (y2 - y3) + x2 * x2 * (-y1 + y3)) / det;
    *c = (x1 * x3 * (-x1 + x3) * y2 + x2 * x2 * (x3 * y1 - x1 * y3) + x2 * (-(x3 * x3 * y1) + x1 * x1 * y3)) / det;
    return true;
}

static bool SolveInverseQuadraticMonotonic(float x1, float y1, float x2, float y2, float x3, float y3, float* a, float* b, float* c) {
    if (x1 > x2) { vSwap(&x1, &x2); vSwap(&y1, &y2); }
    if (x2 > x3) { vSwap(&x2, &x3); vSwap(&y2, &y3); }
    if (x1 > x2) { vSwap(&x1, &x2); vSwap(&y1, &y2); }
    for (double blend_to_linear_factor = 0.0; blend_to_linear_factor <= 1.0; blend_to_linear_factor += 0.05) {
        double tempy2 = (double)(1 - blend_to_linear_factor) * (double)y2 + blend_to_linear_factor * FLerp(y1, y3, x1, x3, x2);
        if (!SolveInverseQuadratic(x1, y1, x2, (float)tempy2, x3, y3, a, b, c)) return false;
        float derivative = 2.0f * (*a) + (*b);
        if ((y1 < y2) && (y2 < y3)) {
            if (derivative >= 0.0f) return true;
        } else {
            if ((y1 > y2) && (y2 > y3)) {
                if (derivative <= 0.0f) return true;
            } else {
                return true;
            }
        }
    }
    return true;
}

static orc_light NewDirectLight() {
    orc_light dl;
    memset(&dl, 0, sizeof(dl));
    dl.end_fade = -1.0f; dl.start_fade = 0.0f; dl.cap_dist = 1.0e22f;
    return dl;
}

static void SetupLightNormalFromProps(const float angles[3], float angle, float pitch, float output[3]) {
    if (angle == -1) { output[0] = 0; output[1] = 0; output[2] = 1; }
    else if (angle == -2) { output[0] = 0; output[1] = 0; output[2] = -1; }
    else {
        if (0 == angle) angle = angles[1];
        output[2] = 0;
        output[0] = (float)cos((double)angle / 180 * PI);
        output[1] = (float)sin((double)angle / 180 * PI);
    }
    if (0 == pitch) pitch = angles[0];
    output[2] = (float)sin((double)pitch / 180 * PI);
    output[0] *= (float)cos((double)pitch / 180 * PI);
    output[1] *= (float)cos((double)pitch / 180 * PI);
}

static void ParseLightGeneric(const LightEntity& e, orc_light& dl) {
    for (int k = 0; k < 3; k++) dl.intensity[k] = e.light_ok ? e.light[k] : 0.0f;
    if (e.has_target) {
        float n[3] = {e.target_origin[0] - dl.origin[0], e.target_origin[1] - dl.origin[1], e.target_origin[2] - dl.origin[2]};
        float len = (float)sqrt((double)(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]));
        float l = 1.0f / len;
        dl.normal[0] = n[0] * l; dl.normal[1] = n[1] * l; dl.normal[2] = n[2] * l;
    } else {
        SetupLightNormalFromProps(e.angles, e.angle, e.pitch, dl.normal);
    }
}

static void SetLightFalloffParams(const LightEntity& e, orc_light& dl) {
    float d50 = e.fifty_percent_distance;
    dl.start_fade = 0; dl.end_fade = -1; dl.cap_dist = 1.0e22f;
    if (0 != d50) {
        float d0 = e.zero_percent_distance;
        if (d0 < d50) d0 = 2.0f * d50;
        float a = 0.0f, b = 1.0f, c = 0.0f;
        SolveInverseQuadraticMonotonic(0, 1.0f, d50, 2.0f, d0, 256.0f, &a, &b, &c);
        float v50 = c + d50 * (b + d50 * a);
        float scale = 2.0f / v50;
        a *= scale; b *= scale; c *= scale;
        dl.quadratic_attn = a; dl.linear_attn = b; dl.constant_attn = c;
        if (0 != e.hardfalloff) {
            dl.end_fade = d0;
            dl.start_fade = 0.75f * d0 + 0.25f * d50;
        } else {
            if (fabs((double)a) > 0.) {
                float flMax = b / (-2.0f * a);
                if (flMax > 0.0f) { dl.cap_dist = flMax; dl.start_fade = flMax; dl.end_fade = 10.0f * flMax; }
            }
        }
    } else {
        dl.constant_attn = e.constant_attn; dl.linear_attn = e.linear_attn; dl.quadratic_attn = e.quadratic_attn;
        dl.radius = e.distance;
        if (dl.constant_attn < EQUAL_EPSILON) dl.constant_attn = 0;
        if (dl.linear_attn < EQUAL_EPSILON) dl.linear_attn = 0;
        if (dl.quadratic_attn < EQUAL_EPSILON) dl.quadratic_attn = 0;
        if (dl.constant_attn < EQUAL_EPSILON && dl.linear_attn < EQUAL_EPSILON && dl.quadratic_attn < EQUAL_EPSILON) dl.constant_attn = 1;
        float ratio = dl.constant_attn + 100 * dl.linear_attn + 100 * 100 * dl.quadratic_attn;
        if (ratio > 0) { dl.intensity[0] = dl.intensity[0] * ratio; dl.intensity[1] = dl.intensity[1] * ratio; dl.intensity[2] = dl.intensity[2] * ratio; }
    }
}

} // namespace orc

using namespace orc;

extern "C" {

int orc_light_for_string(const char* light, float out[3]) {
    double r = 0, g = 0, b = 0, scaler = 0, r_hdr, g_hdr, b_hdr, scaler_hdr;
    int argCnt = sscanf(light, "%lf %lf %lf %lf %lf %lf %lf %lf", &r, &g, &b, &scaler, &r_hdr, &g_hdr, &b_hdr, &scaler_hdr);
    if (argCnt < 0) argCnt = 0;
    if (argCnt == 8) argCnt = 4;
    out[0] = out[1] = out[2] = 0;
    if (r < 0.0 || g < 0.0 || b < 0.0 || scaler < 0.0) return -1;
    out[0] = (float)(pow(r / 255.0, 2.2) * 255);
    switch (argCnt) {
    case 1:
        out[2] = out[0]; out[1] = out[2];
        break;
    case 3:
    case 4:
        out[1] = (float)(pow((double)(g / 255.0), 2.2) * 255);
        out[2] = (float)(pow((double)(b / 255.0), 2.2) * 255);
        if (argCnt == 4) { float s = (float)(scaler / 255.0); out[0] = out[0] * s; out[1] = out[1] * s; out[2] = out[2] * s; }
        break;
    default:
        out[0] = 0;
        return -1;
    }
    return 0;
}

int orc_lights_from_entities(int n, const void* ents_raw, int max_out, orc_light* out) {
    const LightEntity* ents = (const LightEntity*)ents_raw;
    int m = 0;
    bool globalSkyLight = false;
    for (int i = 0; i < n; i++) {
        const LightEntity& e = ents[i];
        orc_light dl = NewDirectLight();
        for (int k = 0; k < 3; k++) dl.origin[k] = e.origin[k];                       // AllocDLight(&dest, ...)
        if (e.classname == 1) {                                                        // ParseLightSpot
            ParseLightGeneric(e, dl);
            dl.type = 2;
            dl.stopdot = e.inner_cone;
            if (0 == dl.stopdot) dl.stopdot = 10;
            dl.stopdot2 = e.cone;
            if (0 == dl.stopdot2) dl.stopdot2 = dl.stopdot;
            if (dl.stopdot2 < dl.stopdot) dl.stopdot2 = dl.stopdot;
            if ((dl.stopdot == 180) && (dl.stopdot2 == 180)) {
                dl.stopdot2 = 0; dl.stopdot = 0; dl.type = 1; dl.exponent = 0;
            } else {
                if (dl.stopdot > 90) dl.stopdot = 90;
                if (dl.stopdot2 > 90) dl.stopdot2 = 90;
                dl.stopdot2 = (float)cos((double)(dl.stopdot2 / 180 * (float)PI));
                dl.stopdot = (float)cos((double)(dl.stopdot / 180 * (float)PI));
                dl.exponent = e.exponent;
            }
            SetLightFalloffParams(e, dl);
        } else if (e.classname == 2) {                                                 // ParseLightEnvironment
            if (globalSkyLight) continue;
            globalSkyLight = true;
            ParseLightGeneric(e, dl);
            dl.type = 3;
            if (m < max_out) out[m] = dl;
            m++;
            orc_light amb = NewDirectLight();
            amb.type = 5;
            for (int k = 0; k < 3; k++) { amb.origin[k] = dl.origin[k]; amb.intensity[k] = e.ambient_ok ? e.ambient[k] : dl.intensity[k] * 0.5f; }
            if (m < max_out) out[m] = amb;
            m++;
            continue;
        } else if (e.classname == 0) {                                                 // ParseLightPoint
            ParseLightGeneric(e, dl);
            dl.type = 1;
            SetLightFalloffParams(e, dl);
        } else {
            return -1;
        }
        if (m < max_out) out[m] = dl;
        m++;
    }
    return m;
}

int orc_lights_from_patches(int n, const float* origin3, const float* normal3, const float* base_light3, const float* area,
                            const float* scale2, const float* base_area, const int32_t* child1, float light_threshold,
                            int max_out, orc_light* out) {
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (child1 && child1[i] != -1) continue;
        if (base_area[i] < 1e-6f) continue;
        const float* bl = &base_light3[3 * i];
        if ((double)((bl[0] + bl[1] + bl[2]) / 3) >= (double)light_threshold) {
            orc_light dl = NewDirectLight();
            dl.type = 0;
            const float* nn = &normal3[3 * i];
            if (!((float)sqrt((double)(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2])) > 1.0e-20f)) return -1;
            float s = 1.0f * area[i] * scale2[2 * i] * scale2[2 * i + 1] / base_area[i];
            for (int k = 0; k < 3; k++) {
                dl.origin[k] = origin3[3 * i + k]; dl.normal[k] = nn[k];
                dl.intensity[k] = bl[k] * s;
                dl.intensity[k] = dl.intensity[k] * DIRECT_SCALE;
            }
            if (m < max_out) out[m] = dl;
            m++;
        }
    }
    return m;
}

} // extern "C"
