/* material.hpp -- Material interface, ScatterRecord, MaterialTwoSided (mirror of the API part of
 * reference libwurblpt/material.hpp:36-334). Scattering itself runs on the device
 * (csrc/device_shade.cuh); a material here is its description plus flatten(). A user-defined
 * Material subclass cannot run on the device and is rejected when the scene is prepared. */
#pragma once

#include <string>

#include "scene_component.hpp"
#include "texture.hpp"
#include "tangentspace.hpp"
#include "ray.hpp"
#include "hitable.hpp"
#include "flatten.hpp"

namespace WurblPT {

typedef enum {
    ScatterNone,
    ScatterExplicit,
    ScatterRandom
} ScatterType;

class ScatterRecord
{
public:
    ScatterType type;
    vec3 direction;
    vec4 attenuation;
    float pdf;
    vec4 refractiveIndex;

    ScatterRecord(ScatterType t) : type(t), direction(0.0f), attenuation(0.0f), pdf(0.0f), refractiveIndex(0.0f) {}
    ScatterRecord(ScatterType t, const vec3& dir, const vec4& att, const vec4& refrInd) :
        type(t), direction(dir), attenuation(att), pdf(0.0f), refractiveIndex(refrInd)
    {
    }
    ScatterRecord(ScatterType t, const vec3& dir, const vec4& att, float p, const vec4& refrInd) :
        type(t), direction(dir), attenuation(att), pdf(p), refractiveIndex(refrInd)
    {
    }
};

class Material : public SceneComponent
{
public:
    const Texture* normalTex;       // public and assignable after construction, as in the reference

    Material(const Texture* nt = nullptr) : normalTex(nt) {}
    virtual ~Material() {}

    /* The base material neither scatters nor emits (reference :158-191). */
    virtual bool flatten(Flattener& /* F */, wpt_material& out) const
    {
        if (typeid(*this) != typeid(Material))
            return false;
        out.type = WPT_MAT_NONE;
        return true;
    }
};

class MaterialTwoSided final : public Material
{
private:
    const Material* _frontSideMaterial;
    const Material* _backSideMaterial;

public:
    MaterialTwoSided(const Material* frontSideMaterial, const Material* backSideMaterial) :
        Material(nullptr), _frontSideMaterial(frontSideMaterial), _backSideMaterial(backSideMaterial)
    {
    }

    virtual bool flatten(Flattener& F, wpt_material& out) const override
    {
        out.type = WPT_MAT_TWOSIDED;
        out.front = F.material(_frontSideMaterial);
        out.back = F.material(_backSideMaterial);
        return true;
    }
};

inline int Flattener::material(const Material* mat)
{
    if (!mat)
        return -1;
    auto it = materialIndex.find(mat);
    if (it != materialIndex.end())
        return it->second;
    wpt_material m;
    std::memset(&m, 0, sizeof(m));
    for (int i = 0; i < 5; i++)
        m.tex[i] = -1;
    m.front = m.back = -1;
    m.normal_tex = texture(mat->normalTex);
    /* reserve the slot first: a two-sided material precedes the sides it pulls in */
    int index = materials.size();
    materials.push_back(m);
    materialIndex[mat] = index;
    if (!mat->flatten(*this, m)) {
        fail("a user-defined Material subclass cannot be evaluated on the device");
        return -1;
    }
    materials[index] = m;
    return index;
}

}
