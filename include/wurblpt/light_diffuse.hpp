/* light_diffuse.hpp -- one-sided diffuse emitter (description of reference libwurblpt/light_diffuse.hpp:32-61) */
#pragma once

#include "material.hpp"

namespace WurblPT {

class LightDiffuse final : public Material
{
private:
    const vec4 _emit;
    const Texture* _tex;

public:
    LightDiffuse(const vec4& emit, const Texture* tex = nullptr) : _emit(emit), _tex(tex) {}
    LightDiffuse(const vec3& emit, const Texture* tex = nullptr) : _emit(emit, average(emit)), _tex(tex) {}

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_LIGHT_DIFFUSE;
        detail::store(out.v0, _emit, 4);
        out.tex[0] = F.texture(_tex);
        return true;
    }
};

}
