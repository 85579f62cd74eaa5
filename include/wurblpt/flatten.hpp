/* flatten.hpp -- turns the API mirror's objects into the plain arrays of the C ABI
 * (include/wurblpt_b200.h). Not part of the reference API; used by Scene::updateBVH() and mcpt().
 *
 * Index assignment follows one fixed walk so that the description produced here is byte-identical
 * to the one oracle/ref_harness.cpp derives from the reference's own Scene object: materials in
 * Scene::take() order (others on first use), textures on first use while visiting materials
 * (normal map first, then the material's slots in declaration order) and finally the environment
 * map, meshes on first use by an instance.
 */
#pragma once

#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../wurblpt_b200.h"

namespace WurblPT {

class Texture;
class Material;
class Mesh;

class Flattener
{
public:
    /* user-defined Texture subclasses are sampled ("baked") into float RGBA images of this size at this time */
    float bakeTime = 0.0f;
    unsigned int bakeTextureSize = 512;
    bool bakedTextures = false;
    std::vector<wpt_animation> animations;
    std::vector<wpt_keyframe> keyframes;
    std::vector<wpt_texture> textures;
    std::vector<uint8_t> texels;
    std::vector<wpt_material> materials;
    std::vector<wpt_mesh> meshes;
    std::vector<float> vertices;
    std::vector<uint32_t> indices;
    std::vector<wpt_instance> instances;
    std::vector<wpt_sphere> spheres;
    std::vector<wpt_medium> media;
    std::vector<uint32_t> boundary;
    std::vector<wpt_component> components;
    std::vector<float> envM;
    std::vector<int32_t> envMs;
    std::vector<float> envMcs;
    wpt_envmap envmap;
    std::string error;

    std::unordered_map<const Texture*, int> textureIndex;
    std::unordered_map<const Material*, int> materialIndex;
    std::unordered_map<const Mesh*, int> meshIndex;

    Flattener()
    {
        std::memset(&envmap, 0, sizeof(envmap));
        for (int i = 0; i < 6; i++)
            envmap.tex[i] = -1;
    }

    int texture(const Texture* tex);        // defined in texture.hpp
    int material(const Material* mat);      // defined in material.hpp
    int mesh(const Mesh* m);                // defined in mesh.hpp

    void fail(const std::string& message)
    {
        if (error.empty())
            error = message;
    }

    wpt_scene_desc desc() const
    {
        wpt_scene_desc d;
        std::memset(&d, 0, sizeof(d));
        d.abi_version = WPT_ABI_VERSION;
        d.n_animations = animations.size();     d.animations = animations.data();
        d.n_keyframes = keyframes.size();       d.keyframes = keyframes.data();
        d.n_textures = textures.size();         d.textures = textures.data();
        d.texel_bytes = texels.size();          d.texels = texels.data();
        d.n_materials = materials.size();       d.materials = materials.data();
        d.n_meshes = meshes.size();             d.meshes = meshes.data();
        d.n_vertex_floats = vertices.size();    d.vertices = vertices.data();
        d.n_indices = indices.size();           d.indices = indices.data();
        d.n_instances = instances.size();       d.instances = instances.data();
        d.n_spheres = spheres.size();           d.spheres = spheres.data();
        d.n_media = media.size();               d.media = media.data();
        d.n_boundary = boundary.size();         d.boundary_hitables = boundary.data();
        d.n_components = components.size();     d.components = components.data();
        d.envmap = envmap;
        d.envmap.M = envM.data();
        d.envmap.Ms = envMs.data();
        d.envmap.Mcs = envMcs.data();
        return d;
    }
};

namespace detail {

template<typename V> inline void store(float* dst, const V& v, int n)
{
    for (int i = 0; i < n; i++)
        dst[i] = v[i];
}

}

}
