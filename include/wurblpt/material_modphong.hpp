/* material_modphong.hpp -- modified Phong with transmission and emission (description of reference
 * libwurblpt/material_modphong.hpp:36-327; device code: WPT_MAT_MODPHONG). Public data members and
 * constructor overloads as in the reference. */
#pragma once

#include "material.hpp"

namespace WurblPT {

class MaterialModPhong final : public Material
{
public:
    vec4 diffuse;
    const Texture* diffuseTex;
    bool diffuseTexHasAlpha;
    vec4 specular;
    const Texture* specularTex;
    bool specularTexHasAlpha;
    float shininess;
    const Texture* shininessTex;
    float opacity;
    const Texture* opacityTex;
    float indexOfRefraction;
    vec4 transmissive;
    vec4 emissive;
    const Texture* emissiveTex;
    bool haveNIR;

    MaterialModPhong() :
        diffuse(0.0f), diffuseTex(nullptr), diffuseTexHasAlpha(false),
        specular(0.0f), specularTex(nullptr), specularTexHasAlpha(false),
        shininess(0.0f), shininessTex(nullptr), opacity(0.0f), opacityTex(nullptr),
        indexOfRefraction(1.0f), transmissive(0.0f), emissive(0.0f), emissiveTex(nullptr), haveNIR(true)
    {
    }

    MaterialModPhong(const vec4& dif, const Texture* difTex,
            const vec4& spc = vec4(0.0f), const Texture* spcTex = nullptr,
            float shi = 100.0f, const Texture* shiTex = nullptr,
            float opa = 1.0f, const Texture* opaTex = nullptr) : MaterialModPhong()
    {
        diffuse = dif; diffuseTex = difTex;
        specular = spc; specularTex = spcTex;
        shininess = shi; shininessTex = shiTex;
        opacity = opa; opacityTex = opaTex;
    }

    MaterialModPhong(const vec3& dif, const Texture* difTex,
            const vec3& spc = vec3(0.0f), const Texture* spcTex = nullptr,
            float shi = 100.0f, const Texture* shiTex = nullptr,
            float opa = 1.0f, const Texture* opaTex = nullptr) :
        MaterialModPhong(vec4(dif, average(dif)), difTex, vec4(spc, average(spc)), spcTex, shi, shiTex, opa, opaTex)
    {
        haveNIR = false;
    }

    MaterialModPhong(const vec4& dif, const vec4& spc = vec4(0.0f), float shi = 100.0f, float opa = 1.0f,
            const Texture* difTex = nullptr, const Texture* spcTex = nullptr,
            const Texture* shiTex = nullptr, const Texture* opaTex = nullptr) :
        MaterialModPhong(dif, difTex, spc, spcTex, shi, shiTex, opa, opaTex)
    {
    }

    MaterialModPhong(const vec3& dif, const vec3& spc = vec3(0.0f), float shi = 100.0f, float opa = 1.0f,
            const Texture* difTex = nullptr, const Texture* spcTex = nullptr,
            const Texture* shiTex = nullptr, const Texture* opaTex = nullptr) :
        MaterialModPhong(vec4(dif, average(dif)), difTex, vec4(spc, average(spc)), spcTex, shi, shiTex, opa, opaTex)
    {
        haveNIR = false;
    }

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_MODPHONG;
        detail::store(out.v0, diffuse, 4);
        detail::store(out.v1, specular, 4);
        detail::store(out.v2, transmissive, 4);
        detail::store(out.v3, emissive, 4);
        out.f[0] = shininess;
        out.f[1] = opacity;
        out.f[2] = indexOfRefraction;
        out.tex[0] = F.texture(diffuseTex);
        out.tex[1] = F.texture(specularTex);
        out.tex[2] = F.texture(shininessTex);
        out.tex[3] = F.texture(opacityTex);
        out.tex[4] = F.texture(emissiveTex);
        if (haveNIR) out.flags |= WPT_MATF_HAVE_NIR;
        if (diffuseTexHasAlpha) out.flags |= WPT_MATF_DIFFUSE_TEX_ALPHA;
        if (specularTexHasAlpha) out.flags |= WPT_MATF_SPECULAR_TEX_ALPHA;
        return true;
    }
};

}
