/* texture_image.hpp -- image textures (mirror of reference libwurblpt/texture_image.hpp:37-366).
 * The reference instantiates one class template per (components, type, sRGB) combination; here one
 * class carries those three as data, createTextureImage() keeps its signatures. */
#pragma once

#include <cstdio>
#include <string>

#include "texture.hpp"

namespace WurblPT {

class TextureImage final : public Texture
{
private:
    const TGD::ArrayContainer _img;
    const bool _linearizeSRGB;
    const vec2 _coordFactor;
    const vec2 _coordOffset;
    const vec4 _valFactor;
    const vec4 _valOffset;

    float component(size_t x, size_t y, size_t c) const
    {
        switch (_img.componentType()) {
        case TGD::uint8: return byte_to_float(_img.get<uint8_t>({ x, y })[c]);
        case TGD::uint16: return uint16_to_float(_img.get<uint16_t>({ x, y })[c]);
        default: return _img.get<float>({ x, y })[c];
        }
    }

    float colorComponent(size_t x, size_t y, size_t c) const
    {
        float v = component(x, y, c);
        return _linearizeSRGB ? srgb_to_rgb_helper(v) : v;
    }

public:
    TextureImage(const TGD::ArrayContainer& img, bool linearizeSRGB,
            const vec2& coordFactor = vec2(1.0f), const vec2& coordOffset = vec2(0.0f),
            const vec4& valFactor = vec4(1.0f), const vec4& valOffset = vec4(0.0f)) :
        _img(img), _linearizeSRGB(linearizeSRGB),
        _coordFactor(coordFactor), _coordOffset(coordOffset), _valFactor(valFactor), _valOffset(valOffset)
    {
    }

    /* one texel as RGBA: grey images replicate, missing alpha is 1, alpha is never sRGB-decoded */
    vec4 value(size_t x, size_t y) const
    {
        if (x >= _img.dimension(0))
            x = _img.dimension(0) - 1;
        if (y >= _img.dimension(1))
            y = _img.dimension(1) - 1;
        switch (_img.componentCount()) {
        case 3:
            return vec4(colorComponent(x, y, 0), colorComponent(x, y, 1), colorComponent(x, y, 2), 1.0f);
        case 4:
            return vec4(colorComponent(x, y, 0), colorComponent(x, y, 1), colorComponent(x, y, 2), component(x, y, 3));
        case 1: {
            float v = colorComponent(x, y, 0);
            return vec4(v, v, v, 1.0f);
        }
        default: {
            float v = colorComponent(x, y, 0);
            return vec4(v, v, v, component(x, y, 1));
        }
        }
    }

    /* bilinear filtering with fract() wrap */
    virtual vec4 value(const vec2& texcoords, float /* t */) const override
    {
        vec2 uv = fract(_coordFactor * texcoords + _coordOffset);
        float uvs = max(0.0f, (uv.s() * _img.dimension(0)) - 0.5f);
        float uvt = max(0.0f, (uv.t() * _img.dimension(1)) - 0.5f);
        size_t x0 = uvs;
        size_t y0 = uvt;
        float alpha = uvs - x0;
        float beta = uvt - y0;
        vec4 a = mix(value(x0, y0), value(x0 + 1, y0), alpha);
        vec4 b = mix(value(x0, y0 + 1), value(x0 + 1, y0 + 1), alpha);
        return _valFactor * mix(a, b, beta) + _valOffset;
    }

    virtual vec2 texelSize() const override { return vec2(1.0f / _img.dimension(0), 1.0f / _img.dimension(1)); }
    virtual unsigned int componentCount() const override { return _img.componentCount(); }
    virtual TGD::Type componentType() const override { return _img.componentType(); }
    virtual enum LinearizeSRGBType linearizeSRGBType() const override { return _linearizeSRGB ? LinearizeSRGB_On : LinearizeSRGB_Off; }

    virtual bool deviceNative() const override { return true; }
    virtual bool flatten(Flattener& F, wpt_texture& out) const override
    {
        out.type = WPT_TEX_IMAGE;
        out.width = _img.dimension(0);
        out.height = _img.dimension(1);
        out.components = _img.componentCount();
        out.texel_type = (_img.componentType() == TGD::uint8 ? WPT_TEXEL_U8
                : _img.componentType() == TGD::uint16 ? WPT_TEXEL_U16 : WPT_TEXEL_F32);
        out.linearize_srgb = _linearizeSRGB;
        while (F.texels.size() % 16 != 0)
            F.texels.push_back(0);
        out.texel_offset = F.texels.size();
        const uint8_t* src = static_cast<const uint8_t*>(_img.data());
        F.texels.insert(F.texels.end(), src, src + _img.dataSize());
        detail::store(out.coord_factor, _coordFactor, 2);
        detail::store(out.coord_offset, _coordOffset, 2);
        detail::store(out.val_factor, _valFactor, 4);
        detail::store(out.val_offset, _valOffset, 4);
        return true;
    }
};

inline Texture* createTextureImage(const TGD::ArrayContainer& img,
        LinearizeSRGBType linearizeSRGBType = LinearizeSRGB_Auto,
        const vec2& coordFactor = vec2(1.0f), const vec2& coordOffset = vec2(0.0f),
        const vec4& valFactor = vec4(1.0f), const vec4& valOffset = vec4(0.0f))
{
    if (img.dimensionCount() != 2 || img.dimension(0) == 0 || img.dimension(1) == 0
            || img.componentCount() < 1 || img.componentCount() > 4
            || (img.componentType() != TGD::uint8 && img.componentType() != TGD::uint16
                && img.componentType() != TGD::float32))
        return nullptr;
    bool linearize = (linearizeSRGBType == LinearizeSRGB_On)
        || (linearizeSRGBType == LinearizeSRGB_Auto
                && (img.componentType() == TGD::uint8 || img.componentType() == TGD::uint16));
    return new TextureImage(img, linearize, coordFactor, coordOffset, valFactor, valOffset);
}

inline Texture* createTextureImage(const std::string& imgFileName,
        LinearizeSRGBType linearizeSRGBType = LinearizeSRGB_Auto,
        const vec2& coordFactor = vec2(1.0f), const vec2& coordOffset = vec2(0.0f),
        const vec4& valFactor = vec4(1.0f), const vec4& valOffset = vec4(0.0f))
{
    TGD::Error e;
    TGD::ArrayContainer img = TGD::load(imgFileName, TGD::TagList(), &e);
    if (img.elementCount() == 0) {
        fprintf(stderr, "%s: %s\n", imgFileName.c_str(), TGD::strerror(e));
        return nullptr;
    }
    Texture* tex = createTextureImage(img, linearizeSRGBType, coordFactor, coordOffset, valFactor, valOffset);
    if (!tex)
        fprintf(stderr, "%s: not a valid texture image\n", imgFileName.c_str());
    return tex;
}

}
