/* material_ggx.hpp -- anisotropic GGX microfacet reflector (description of reference
 * libwurblpt/material_ggx.hpp:36-257; device code: WPT_MAT_GGX). */
#pragma once

#include "material.hpp"
#include "fresnel.hpp"

namespace WurblPT {

class MaterialGGX final : public Material
{
public:
    vec4 albedo;
    vec2 roughness;
    const Texture* albedoTex;
    const Texture* roughnessTex;

    MaterialGGX(const vec4& alb, const Texture* albTex, const vec2& r, const Texture* rTex) :
        albedo(alb), roughness(r), albedoTex(albTex), roughnessTex(rTex)
    {
    }
    MaterialGGX() : MaterialGGX(vec4(0.0f), nullptr, vec2(0.0f), nullptr) {}
    MaterialGGX(const vec4& alb, const vec2& r, const Texture* albTex = nullptr, const Texture* rTex = nullptr) :
        MaterialGGX(alb, albTex, r, rTex)
    {
    }
    MaterialGGX(const vec3& alb, const Texture* albTex, const vec2& r, const Texture* rTex) :
        MaterialGGX(vec4(alb, average(alb)), albTex, r, rTex)
    {
    }
    MaterialGGX(const vec3& alb, const vec2& r, const Texture* albTex = nullptr, const Texture* rTex = nullptr) :
        MaterialGGX(vec4(alb, average(alb)), albTex, r, rTex)
    {
    }

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_GGX;
        detail::store(out.v0, albedo, 4);
        out.f[0] = roughness.x();
        out.f[1] = roughness.y();
        out.tex[0] = F.texture(albedoTex);
        out.tex[1] = F.texture(roughnessTex);
        return true;
    }
};

}
