/* material_phase_function_isotropic.hpp -- isotropic phase function for media (description of
 * reference libwurblpt/material_phase_function_isotropic.hpp:35-59). */
#pragma once

#include "material.hpp"

namespace WurblPT {

class MaterialPhaseFunctionIsotropic final : public Material
{
private:
    const vec4 _albedo;

public:
    MaterialPhaseFunctionIsotropic(const vec4& albedo) : _albedo(albedo) {}
    MaterialPhaseFunctionIsotropic(const vec3& albedo) : _albedo(albedo, 0.0f) {}

    virtual bool flatten(Flattener&, wpt_material& out) const override
    {
        out.type = WPT_MAT_PHASE_ISOTROPIC;
        detail::store(out.v0, _albedo, 4);
        return true;
    }
};

}
