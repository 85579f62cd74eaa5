/* material_mirror.hpp -- perfect mirror (description of reference libwurblpt/material_mirror.hpp:33-62) */
#pragma once

#include "material.hpp"

namespace WurblPT {

class MaterialMirror final : public Material
{
public:
    const vec4 color;
    const Texture* colorTex;
    const bool haveNIR;

    MaterialMirror(const vec4& color_ = vec4(1.0f), const Texture* tex = nullptr) :
        Material(nullptr), color(color_), colorTex(tex), haveNIR(true)
    {
    }
    MaterialMirror(const vec3& color_, const Texture* tex = nullptr) :
        Material(nullptr), color(color_, average(color_)), colorTex(tex), haveNIR(false)
    {
    }

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_MIRROR;
        detail::store(out.v0, color, 4);
        out.tex[0] = F.texture(colorTex);
        if (haveNIR)
            out.flags |= WPT_MATF_HAVE_NIR;
        return true;
    }
};

}
