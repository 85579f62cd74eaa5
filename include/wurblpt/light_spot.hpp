/* light_spot.hpp -- emitter limited to a cone around its normal (description of reference
 * libwurblpt/light_spot.hpp:32-60) */
#pragma once

#include "material.hpp"

namespace WurblPT {

class LightSpot final : public Material
{
private:
    const float _cosHalfOpeningAngle;
    const vec4 _emit;
    const Texture* _tex;

public:
    LightSpot(float openingAngle, const vec3& emit, const Texture* tex = nullptr) :
        _cosHalfOpeningAngle(cos(0.5f * openingAngle)), _emit(emit, average(emit)), _tex(tex)
    {
    }

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_LIGHT_SPOT;
        detail::store(out.v0, _emit, 4);
        out.f[0] = _cosHalfOpeningAngle;
        out.tex[0] = F.texture(_tex);
        return true;
    }
};

}
