/* light_tof.hpp -- light source of a time-of-flight camera: emits in the NIR channel only and marks
 * its radiance as modulated (description of reference libwurblpt/light_tof.hpp:36-74) */
#pragma once

#include "material.hpp"

namespace WurblPT {

class LightTof final : public Material
{
private:
    float _radiance;
    float _cosHalfOpeningAngle;
    const Texture* _tex;

public:
    LightTof(float radiance, float openingAngle, const Texture* tex = nullptr) :
        _radiance(radiance), _cosHalfOpeningAngle(cos(0.5f * openingAngle)), _tex(tex)
    {
    }

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_LIGHT_TOF;
        out.f[0] = _radiance;
        out.f[1] = _cosHalfOpeningAngle;
        out.tex[0] = F.texture(_tex);
        return true;
    }
};

}
