/* texture.hpp -- Texture interface and the procedural textures (mirror of reference
 * libwurblpt/texture.hpp:40-246). value() is evaluated on the host for environment-map importance
 * tables and application use; the device evaluates the same four texture kinds from their flattened
 * description (csrc/device_shade.cuh). A user-defined Texture subclass cannot run on the device:
 * it is sampled into an image texture when the scene is prepared (Flattener::texture). */
#pragma once

#include <tgd/array.hpp>
#include <tgd/io.hpp>

#include "scene_component.hpp"
#include "color.hpp"
#include "animation.hpp"
#include "flatten.hpp"

namespace WurblPT {

enum LinearizeSRGBType
{
    LinearizeSRGB_On,
    LinearizeSRGB_Off,
    LinearizeSRGB_Auto
};

class Texture : public SceneComponent
{
public:
    virtual ~Texture() {}

    virtual vec4 value(const vec2& texcoords, float t = 0.0f) const = 0;
    virtual vec2 texelSize() const { return vec2(0.0f, 0.0f); }
    virtual unsigned int componentCount() const { return 4; }
    virtual TGD::Type componentType() const { return TGD::float32; }
    virtual enum LinearizeSRGBType linearizeSRGBType() const { return LinearizeSRGB_Auto; }

    /* Fill `out` (and append texels) for the C ABI; return false if this texture kind cannot be
     * expressed there. */
    virtual bool flatten(Flattener& /* F */, wpt_texture& /* out */) const { return false; }
    /* true: flatten() expresses this texture exactly; false: it is a function only value() knows (a user-defined
     * class, a noise texture), which the Flattener samples into an image */
    virtual bool deviceNative() const { return false; }
    /* how many times the unit square of this texture's coordinates a lookup in [0,1]^2 spans (see TextureTransformer) */
    virtual vec2 bakeScale() const { return vec2(1.0f, 1.0f); }
};

class TextureConstant final : public Texture
{
private:
    const vec4 _color;

public:
    TextureConstant(const vec4& color) : _color(color) {}

    virtual vec4 value(const vec2& /* texcoords */, float /* t */) const override { return _color; }
    virtual vec2 texelSize() const override { return vec2(1.0f, 1.0f); }

    virtual bool flatten(Flattener&, wpt_texture& out) const override
    {
        out.type = WPT_TEX_CONSTANT;
        detail::store(out.color0, _color, 4);
        return true;
    }
    virtual bool deviceNative() const override { return true; }
};

class TextureChecker final : public Texture
{
private:
    vec4 _color0, _color1;
    int _horiz, _vert;

public:
    TextureChecker(const vec4& color0, const vec4& color1, int horiz = 9, int vert = 9) :
        _color0(color0), _color1(color1), _horiz(horiz), _vert(vert)
    {
    }

    TextureChecker(const vec3& color0, const vec3& color1, int horiz = 9, int vert = 9) :
        TextureChecker(vec4(color0, average(color0)), vec4(color1, average(color1)), horiz, vert)
    {
    }

    virtual vec4 value(const vec2& texcoords, float /* t */) const override
    {
        int row = texcoords.y() * _vert;
        int col = texcoords.x() * _horiz;
        return (row % 2 == col % 2 ? _color0 : _color1);
    }

    virtual bool flatten(Flattener&, wpt_texture& out) const override
    {
        out.type = WPT_TEX_CHECKER;
        detail::store(out.color0, _color0, 4);
        detail::store(out.color1, _color1, 4);
        out.horiz = _horiz;
        out.vert = _vert;
        return true;
    }
    virtual bool deviceNative() const override { return true; }
};

class TextureTransformer final : public Texture
{
private:
    const Texture* _texture;
    const vec2 _coordFactor;
    const vec2 _coordOffset;
    const vec4 _valFactor;
    const vec4 _valOffset;

public:
    TextureTransformer(const Texture* texture,
            const vec2& coordFactor, const vec2& coordOffset = vec2(0.0f),
            const vec4& valFactor = vec4(1.0f), const vec4& valOffset = vec4(0.0f)) :
        _texture(texture), _coordFactor(coordFactor), _coordOffset(coordOffset),
        _valFactor(valFactor), _valOffset(valOffset)
    {
    }

    virtual vec4 value(const vec2& texcoords, float t) const override
    {
        vec4 val = _texture->value(_coordFactor * texcoords + _coordOffset, t);
        return _valFactor * val + _valOffset;
    }

    virtual vec2 texelSize() const override { return _texture->texelSize() / _coordFactor; }
    virtual unsigned int componentCount() const override { return _texture->componentCount(); }
    virtual TGD::Type componentType() const override { return _texture->componentType(); }

    virtual bool deviceNative() const override { return _texture->deviceNative(); }
    virtual vec2 bakeScale() const override { return abs(_coordFactor) * _texture->bakeScale(); }

    virtual bool flatten(Flattener& F, wpt_texture& out) const override
    {
        /* A sampled source is only known on the unit square of ITS coordinates (the image lookup wraps around), but
         * it need not be periodic (e.g. a ramp clamped to [0,1]): then this texture as a whole is sampled instead,
         * over the unit square of the coordinates it is looked up with. */
        if (!_texture->deviceNative())
            return false;
        out.type = WPT_TEX_TRANSFORMER;
        out.src_texture = F.texture(_texture);
        detail::store(out.coord_factor, _coordFactor, 2);
        detail::store(out.coord_offset, _coordOffset, 2);
        detail::store(out.val_factor, _valFactor, 4);
        detail::store(out.val_offset, _valOffset, 4);
        return out.src_texture >= 0;
    }
};

inline int Flattener::texture(const Texture* tex)
{
    if (!tex)
        return -1;
    auto it = textureIndex.find(tex);
    if (it != textureIndex.end())
        return it->second;
    wpt_texture t;
    std::memset(&t, 0, sizeof(t));
    t.src_texture = -1;
    if (!tex->flatten(*this, t)) {
        /* A user-defined Texture: the device cannot call its value(). It is sampled at texel centres into a float
         * RGBA image (bilinear lookup with fract() wrap, like TextureImage) at the time the description is made for;
         * the scene re-bakes when updateBVH() is called for other times. The image has the size the reference itself
         * gives such a texture when it writes it out (Texture::exportToObj, texture.hpp:100-105): 1 / texelSize() per
         * axis where the texture reports one, bakeTextureSize (512) otherwise - times the number of repeats a
         * TextureTransformer in front of a sampled texture asks for, so that its texels do not get coarser. */
        bakedTextures = true;
        const vec2 scale = tex->bakeScale();
        unsigned int w = max(static_cast<unsigned int>(bakeTextureSize * min(scale.x(), 8.0f)), 1u);
        unsigned int h = max(static_cast<unsigned int>(bakeTextureSize * min(scale.y(), 8.0f)), 1u);
        const vec2 ts = tex->texelSize();
        if (ts.x() > 0.0f)
            w = max(static_cast<unsigned int>(1.0f / ts.x()), 1u);
        if (ts.y() > 0.0f)
            h = max(static_cast<unsigned int>(1.0f / ts.y()), 1u);
        w = min(w, 8192u);
        h = min(h, 8192u);
        t.type = WPT_TEX_IMAGE;
        t.width = int(w);
        t.height = int(h);
        t.components = 4;
        t.texel_type = WPT_TEXEL_F32;
        t.linearize_srgb = 0;
        t.coord_factor[0] = t.coord_factor[1] = 1.0f;
        for (int c = 0; c < 4; c++)
            t.val_factor[c] = 1.0f;
        while (texels.size() % 16 != 0)
            texels.push_back(0);
        t.texel_offset = texels.size();
        texels.resize(texels.size() + size_t(w) * h * 4 * sizeof(float));
        float* dst = reinterpret_cast<float*>(texels.data() + t.texel_offset);
        for (unsigned int y = 0; y < h; y++) {
            for (unsigned int x = 0; x < w; x++) {
                const vec4 v = tex->value(vec2((x + 0.5f) / w, (y + 0.5f) / h), bakeTime);
                for (int c = 0; c < 4; c++)
                    dst[(size_t(y) * w + x) * 4 + c] = v[c];
            }
        }
    }
    int index = textures.size();
    textures.push_back(t);
    textureIndex[tex] = index;
    return index;
}

}
