/* color.hpp -- colour-space and integer/float helpers needed by textures and post-processing
 * (mirror of the used part of reference libwurblpt/color.hpp:240-305). */
#pragma once

#include <cstdint>

#include "gvm.hpp"

namespace WurblPT {

inline float rgb_to_srgb_helper(float x)
{
    return (x <= 0.0031308f ? (x * 12.92f) : (1.055f * pow(x, 1.0f / 2.4f) - 0.055f));
}

inline vec3 rgb_to_srgb(const vec3& rgb)
{
    return vec3(rgb_to_srgb_helper(rgb.r()), rgb_to_srgb_helper(rgb.g()), rgb_to_srgb_helper(rgb.b()));
}

inline float srgb_to_rgb_helper(float x)
{
    return (x <= 0.04045f ? (x * (1.0f / 12.92f)) : pow((x + 0.055f) * (1.0f / 1.055f), 2.4f));
}

inline vec3 srgb_to_rgb(const vec3& srgb)
{
    return vec3(srgb_to_rgb_helper(srgb.r()), srgb_to_rgb_helper(srgb.g()), srgb_to_rgb_helper(srgb.b()));
}

inline vec3 rgb_to_xyz(const vec3& rgb)
{
    return 100.0f * vec3(
            (0.412453f * rgb.r() + 0.357580f * rgb.g() + 0.180423f * rgb.b()),
            (0.212671f * rgb.r() + 0.715160f * rgb.g() + 0.072169f * rgb.b()),
            (0.019334f * rgb.r() + 0.119193f * rgb.g() + 0.950227f * rgb.b()));
}

/* inverse of rgb_to_xyz (BT.709 primaries, the PBRT / Mitsuba matrix; reference color.hpp:255-261) */
inline vec3 xyz_to_rgb(const vec3& xyz)
{
    return 0.01f * vec3(
            (+3.240479f * xyz.x() - 1.537150f * xyz.y() - 0.498535f * xyz.z()),
            (-0.969256f * xyz.x() + 1.875991f * xyz.y() + 0.041556f * xyz.z()),
            (+0.055648f * xyz.x() - 0.204023f * xyz.y() + 1.057311f * xyz.z()));
}

/* new luminance Y at unchanged chromaticity (reference color.hpp:226-237) */
inline vec3 adjust_y(const vec3& xyz, float new_y)
{
    const float sum = xyz.x() + xyz.y() + xyz.z();
    if (xyz.y() <= 0.0f || sum <= 0.0f)
        return vec3(0.0f);
    const float cx = xyz.x() / sum;
    const float cy = xyz.y() / sum;
    const float ratio = new_y / cy;
    return vec3(ratio * cx, new_y, ratio * (1.0f - cx - cy));
}

inline float byte_to_float(uint8_t x) { return x / 255.0f; }
inline float uint16_to_float(uint16_t x) { return x / 65535.0f; }
inline uint8_t float_to_byte(float x) { return round(x * 255.0f); }
inline uint16_t float_to_uint16(float x) { return round(x * 65535.0f); }

}
