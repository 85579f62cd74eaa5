/* material_lambertian.hpp -- ideal diffuse reflector (description of reference
 * libwurblpt/material_lambertian.hpp:36-102; device code: WPT_MAT_LAMBERTIAN). */
#pragma once

#include "material.hpp"
#include "sampler.hpp"

namespace WurblPT {

class MaterialLambertian final : public Material
{
private:
    const vec4 _albedo;
    const Texture* _tex;
    const bool haveNIR;

public:
    MaterialLambertian(const vec4& albedo, const Texture* tex = nullptr) : _albedo(albedo), _tex(tex), haveNIR(true) {}
    MaterialLambertian(const vec3& albedo, const Texture* tex = nullptr) : _albedo(albedo, 0.0f), _tex(tex), haveNIR(false) {}

    vec4 albedoAt(const vec2& texcoords, float t) const
    {
        vec4 a = _tex ? _tex->value(texcoords, t) : _albedo;
        if (!haveNIR)
            a.a() = average(a.rgb());
        return a;
    }

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_LAMBERTIAN;
        detail::store(out.v0, _albedo, 4);
        out.tex[0] = F.texture(_tex);
        if (haveNIR)
            out.flags |= WPT_MATF_HAVE_NIR;
        return true;
    }
};

}
