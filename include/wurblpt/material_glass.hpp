/* material_glass.hpp -- dielectric with absorption and optional dispersion (description of reference
 * libwurblpt/material_glass.hpp:37-206; device code: WPT_MAT_GLASS). */
#pragma once

#include "material.hpp"
#include "fresnel.hpp"
#include "constants.hpp"

namespace WurblPT {

class MaterialGlass final : public Material
{
private:
    static vec4 toVec4(const vec3& v) { return vec4(v, average(v)); }
    static bool allComponentsEqual(const vec4& v) { return (v.r() == v.g()) && (v.r() == v.b()) && (v.r() == v.a()); }

public:
    const vec4 absorption;
    const vec4 refractiveIndexOfMaterial;
    const vec4 refractiveIndexOfSurroundingMedium;
    const bool chromaticDispersion;

    MaterialGlass(const vec4& absorption_, const vec4& riMaterial, const vec4& riSurrounding = vec4(refractiveIndexOfVacuum)) :
        Material(nullptr), absorption(absorption_), refractiveIndexOfMaterial(riMaterial),
        refractiveIndexOfSurroundingMedium(riSurrounding),
        chromaticDispersion(!allComponentsEqual(riMaterial) || !allComponentsEqual(riSurrounding))
    {
    }
    MaterialGlass(const vec4& absorption_, float riMaterial, float riSurrounding = refractiveIndexOfVacuum) :
        MaterialGlass(absorption_, vec4(riMaterial), vec4(riSurrounding))
    {
    }
    MaterialGlass(const vec3& absorption_, const vec3& riMaterial, const vec3& riSurrounding = vec3(refractiveIndexOfVacuum)) :
        MaterialGlass(toVec4(absorption_), toVec4(riMaterial), toVec4(riSurrounding))
    {
    }
    MaterialGlass(const vec3& absorption_, float riMaterial, float riSurrounding = refractiveIndexOfVacuum) :
        MaterialGlass(toVec4(absorption_), vec4(riMaterial), vec4(riSurrounding))
    {
    }

    static float transparentColorToAbsorption(float transparentColor, float targetDistance = 0.01f)
    {
        return max(-log(transparentColor) / targetDistance, 0.0f);
    }
    static vec3 transparentColorToAbsorption(const vec3& c, float targetDistance = 0.01f)
    {
        return vec3(transparentColorToAbsorption(c.r(), targetDistance), transparentColorToAbsorption(c.g(), targetDistance),
                transparentColorToAbsorption(c.b(), targetDistance));
    }
    static vec4 transparentColorToAbsorption(const vec4& c, float targetDistance = 0.01f)
    {
        return vec4(transparentColorToAbsorption(c.r(), targetDistance), transparentColorToAbsorption(c.g(), targetDistance),
                transparentColorToAbsorption(c.b(), targetDistance), transparentColorToAbsorption(c.a(), targetDistance));
    }
    static float absorptionToTransparentColor(float a, float targetDistance = 0.01f) { return exp(-a / targetDistance); }
    static vec3 absorptionToTransparentColor(const vec3& a, float targetDistance = 0.01f)
    {
        return vec3(absorptionToTransparentColor(a.r(), targetDistance), absorptionToTransparentColor(a.g(), targetDistance),
                absorptionToTransparentColor(a.b(), targetDistance));
    }

    virtual bool flatten(Flattener&, wpt_material& out) const override
    {
        out.type = WPT_MAT_GLASS;
        detail::store(out.v0, absorption, 4);
        detail::store(out.v1, refractiveIndexOfMaterial, 4);
        detail::store(out.v2, refractiveIndexOfSurroundingMedium, 4);
        if (chromaticDispersion)
            out.flags |= WPT_MATF_CHROMATIC;
        return true;
    }
};

}
