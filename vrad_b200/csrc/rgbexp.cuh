// rgbexp.cuh -- ColorRGBExp32 packing, shared by the host entry points (bsp_light.cpp) and the finalisation kernel
// (k5_finalize.cu): one inline function, compiled for both sides, so the CPU tests of the host entry point pin the
// arithmetic the kernel runs.
//
// Upstream VectorToColorRGBExp32 (UNCITED: Source SDK 2013 bspfile; absent from the reference, whose finish task only
// logs "Writing", cmd/tasks/finish/main.go:15-18): the largest component is brought into [128, 255] by halving /
// doubling, the shared exponent is clamped to a signed byte, the three mantissas are truncated to 8 bits.
// Own rule for the range upstream leaves undefined: components below 2^-120 (scalar 2^-exponent would overflow)
// encode as zero.  Negative components are clamped to zero first (FinalLightFace does the same).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define VRAD_HD __host__ __device__ __forceinline__
#else
#define VRAD_HD inline
#endif

namespace vrad {

struct RgbExp { uint8_t r, g, b; int8_t e; };

VRAD_HD RgbExp pack_rgbexp32(float r, float g, float b) {
    r = r > 0.0f ? r : 0.0f; g = g > 0.0f ? g : 0.0f; b = b > 0.0f ? b : 0.0f;      // also maps NaN to 0
    float mx = r; if (g > mx) mx = g; if (b > mx) mx = b;
    RgbExp o = {0, 0, 0, 0};
    if (!(mx >= 7.5231638e-37f)) return o;                   // 2^-120
    if (mx > 3.0e38f) mx = 3.0e38f;                          // +inf would never leave the halving loop
    int power = 0;
    float in = mx;
    while (in > 255.0f) { power += 1; in *= 0.5f; }
    while (in < 128.0f) { power -= 1; in *= 2.0f; }
    // power is within [-127, 121] here, so the signed-byte clamp of upstream never triggers
    float scalar = 1.0f;                                     // 2^-power, exact
    if (power > 0) for (int k = 0; k < power; k++) scalar *= 0.5f;
    else for (int k = 0; k < -power; k++) scalar *= 2.0f;
    float fr = r * scalar, fg = g * scalar, fb = b * scalar;
    fr = fr > 255.0f ? 255.0f : fr; fg = fg > 255.0f ? 255.0f : fg; fb = fb > 255.0f ? 255.0f : fb;
    o.r = (uint8_t)fr; o.g = (uint8_t)fg; o.b = (uint8_t)fb; o.e = (int8_t)power;
    return o;
}

// upstream ColorRGBExp32ToVector: component * 2^exponent
VRAD_HD void unpack_rgbexp32(RgbExp c, float& r, float& g, float& b) {
    float s = 1.0f;
    if (c.e > 0) for (int k = 0; k < c.e; k++) s *= 2.0f;
    else for (int k = 0; k < -(int)c.e; k++) s *= 0.5f;
    r = (float)c.r * s; g = (float)c.g * s; b = (float)c.b * s;
}

}  // namespace vrad
