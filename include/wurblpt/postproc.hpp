/* postproc.hpp -- the image post-processing helpers example applications call after mcpt()
 * (mirror of the commonly used part of reference libwurblpt/postproc.hpp:44-110; host code). */
#pragma once

#include <tgd/array.hpp>

#include <cassert>
#include <cstring>
#include <memory>
#include <vector>

#include "gvm.hpp"
#include "color.hpp"
#include "optics.hpp"
#include "texture_image.hpp"

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

/* multiplies the luminance of every pixel, keeping its chromaticity; luminances above 100 * clamp are cut
 * (reference postproc.hpp:93-108) */
inline TGD::Array<float> scaleLuminance(const TGD::Array<float>& img, float factor, float clamp = 1.0f)
{
    TGD::Array<float> r(img.description());
    for (size_t i = 0; i < r.elementCount(); i++) {
        vec3 xyz = rgb_to_xyz(vec3(img.get<float>(i)));
        float y = factor * xyz.y();
        if (clamp > 0.0f && y > 100.0f * clamp)
            y = 100.0f * clamp;
        const vec3 rgb = xyz_to_rgb(adjust_y(xyz, y));
        r.set(i, { rgb.r(), rgb.g(), rgb.b() });
    }
    return r;
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

/* Distance image (e.g. SensorTofAmcw::result()) -> camera space coordinates (reference
 * postproc.hpp:252-287): the distance is looked up at the distorted image position of each pixel
 * centre and put along the undistorted viewing direction; z points away from the camera. */
inline TGD::Array<float> cameraSpaceDistanceToCoords(
        const TGD::Array<float>& distances /* first component must contain distance */,
        const Projection& projection,
        const LensDistortion& distortion)
{
    std::unique_ptr<Texture> lookup(createTextureImage(distances, LinearizeSRGB_Off));
    TGD::Array<float> coords(distances.dimensions(), 3);
    coords.globalTagList() = distances.globalTagList();
    static const char* axes[3] = { "x", "y", "z" };
    for (size_t c = 0; c < 3; c++)
        coords.componentTagList(c).set("INTERPRETATION", axes[c]);
    const unsigned int width = distances.dimension(0);
    const unsigned int height = distances.dimension(1);
    const float invWidth = 1.0f / width;
    const float invHeight = 1.0f / height;
    const LensDistortion::Helper helper = distortion.getHelper(projection, width, height);
    for (size_t y = 0; y < height; y++) {
        for (size_t x = 0; x < width; x++) {
            const float p = (x + 0.5f) * invWidth;
            const float q = (y + 0.5f) * invHeight;
            float dp = p, dq = q;
            distortion.distort(dp, dq, helper);
            const float distance = lookup->value(vec2(dp, dq)).r();
            const vec2 uv = (vec2(p, q) - projection.center()) * projection.inverseFocalLength();
            const vec3 xyz = normalize(vec3(uv, 1.0f)) * distance;
            coords.set({ x, y }, { xyz.x(), xyz.y(), -xyz.z() });
        }
    }
    return coords;
}

/* Camera space (y up, looking along -z) -> PMD sensor space (y down, z forward) (postproc.hpp:289-309) */
inline TGD::Array<float> cameraSpaceToPMDSpace(const TGD::Array<float>& coords /* first 3 components must be x,y,z */)
{
    TGD::Array<float> pmd(coords.dimensions(), 3);
    pmd.globalTagList() = coords.globalTagList();
    static const char* axes[3] = { "x", "y", "z" };
    for (size_t c = 0; c < 3; c++)
        pmd.componentTagList(c).set("INTERPRETATION", axes[c]);
    for (size_t e = 0; e < pmd.elementCount(); e++) {
        pmd[e][0] = coords[e][0];
        pmd[e][1] = -coords[e][1];
        pmd[e][2] = -coords[e][2];
    }
    return pmd;
}

/* A new array holding the listed components of `a`, in the listed order, with their tags
 * (reference postproc.hpp:313-337) */
inline TGD::ArrayContainer extractComponents(const TGD::ArrayContainer& a, const std::vector<size_t>& componentList)
{
    TGD::ArrayContainer picked(a.dimensions(), componentList.size(), a.componentType());
    picked.globalTagList() = a.globalTagList();
    for (size_t d = 0; d < a.dimensionCount(); d++)
        picked.dimensionTagList(d) = a.dimensionTagList(d);
    const size_t bytes = a.componentSize();
    const unsigned char* from = static_cast<const unsigned char*>(a.data());
    unsigned char* to = static_cast<unsigned char*>(picked.data());
    for (size_t c = 0; c < componentList.size(); c++) {
        assert(componentList[c] < a.componentCount());
        picked.componentTagList(c) = a.componentTagList(componentList[c]);
        for (size_t e = 0; e < a.elementCount(); e++)
            std::memcpy(to + e * picked.elementSize() + c * bytes, from + e * a.elementSize() + componentList[c] * bytes, bytes);
    }
    return picked;
}

inline TGD::ArrayContainer extractComponent(const TGD::ArrayContainer& a, size_t component)
{
    return extractComponents(a, std::vector<size_t>(1, component));
}

}
