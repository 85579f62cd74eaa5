/* postproc.hpp -- the image post-processing helpers example applications call after mcpt()
 * (mirror of the commonly used part of reference libwurblpt/postproc.hpp:44-110; host code). */
#pragma once

#include <tgd/array.hpp>

#include "gvm.hpp"
#include "color.hpp"

namespace WurblPT {

inline TGD::Array<uint8_t> toSRGB(const TGD::Array<float>& img)
{
    TGD::Array<uint8_t> r(img.dimensions(), 3);
    r.globalTagList() = img.globalTagList();
    for (size_t d = 0; d < img.dimensionCount(); d++)
        r.dimensionTagList(d) = img.dimensionTagList(d);
    r.componentTagList(0).set("INTERPRETATION", "SRGB/R");
    r.componentTagList(1).set("INTERPRETATION", "SRGB/G");
    r.componentTagList(2).set("INTERPRETATION", "SRGB/B");
    for (size_t i = 0; i < r.elementCount(); i++) {
        for (size_t j = 0; j < r.componentCount(); j++) {
            float v = img.get<float>(i, j < img.componentCount() ? j : 0);
            r.set<uint8_t>(i, j, float_to_byte(rgb_to_srgb_helper(clamp(v, 0.0f, 1.0f))));
        }
    }
    return r;
}

inline float maxLuminance(const TGD::Array<float>& img)
{
    float maxLum = 0.0f;
    for (size_t i = 0; i < img.elementCount(); i++) {
        vec3 rgb(img.get<float>(i));
        float lum = rgb_to_xyz(rgb).y();
        if (lum > maxLum)
            maxLum = lum;
    }
    return maxLum;
}

/* Schlick's uniform rational quantisation tone mapping on luminance */
inline TGD::Array<float> uniformRationalQuantization(const TGD::Array<float>& img, float maxVal, float brightness)
{
    TGD::Array<float> r(img.description());
    for (size_t i = 0; i < r.elementCount(); i++) {
        for (size_t c = 0; c < r.componentCount(); c++) {
            float v = img.get<float>(i, c);
            r.set<float>(i, c, brightness * v / ((brightness - 1.0f) * v + maxVal));
        }
    }
    return r;
}

}
