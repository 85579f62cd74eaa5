/* mesh.hpp -- Mesh (interleaved vertex buffer + indices) and MeshInstance (mirror of reference
 * libwurblpt/mesh.hpp:38-189; same public members). */
#pragma once

#include <vector>

#include "material.hpp"
#include "scene_component.hpp"
#include "geometryproc.hpp"
#include "transformation.hpp"

namespace WurblPT {

class Mesh
{
public:
    /* per vertex: position(3) normal(3) [texcoord(2)] [tangent(3)] */
    constexpr static size_t positionOffset = 0;
    constexpr static size_t normalOffset = 3;
    constexpr static size_t texcoordOffset = 6;
    constexpr static size_t tangentOffset = 8;
    bool haveTexCoords;
    bool haveTangents;
    std::vector<float> vertices;
    std::vector<unsigned int> indices;

    static size_t vertexSize(bool haveTexCoords, bool haveTangents)
    {
        return (haveTexCoords && haveTangents ? 11 : haveTexCoords ? 8 : 6);
    }
    size_t vertexSize() const { return vertexSize(haveTexCoords, haveTangents); }
    size_t vertexCount() const { return vertices.size() / vertexSize(); }
    size_t triangleCount() const { return indices.size() / 3; }
    vec3 position(unsigned int i) const { return vec3(vertices.data() + i * vertexSize() + positionOffset); }
    vec3 normal(unsigned int i) const { return vec3(vertices.data() + i * vertexSize() + normalOffset); }
    vec2 texcoord(unsigned int i) const { return vec2(vertices.data() + i * vertexSize() + texcoordOffset); }
    vec3 tangent(unsigned int i) const { return vec3(vertices.data() + i * vertexSize() + tangentOffset); }

    /* Texture coordinates may be empty or all zero (then there are no tangents either). The
     * geometry is baked through T. */
    Mesh(const std::vector<vec3>& pos, const std::vector<vec3>& nrm, const std::vector<vec2>& tc,
            const std::vector<unsigned int>& ind, const Transformation& T = Transformation(), bool wantTangents = true) :
        indices(ind)
    {
        /* The reference means to drop texture coordinates that are all zero, but its test
         * `if (notEqual(tc[i], vec2(0)))` converts the bvec2 to a (non-null) pointer and is always
         * true (mesh.hpp:92-98 with gvm.hpp:266): any non-empty array counts as "has texture
         * coordinates". Mirrored, because it decides the vertex layout. */
        bool anyTexCoord = tc.size() > 0;
        if (!anyTexCoord)
            wantTangents = false;
        std::vector<vec3> tng;
        if (wantTangents)
            tng = computeTangents(pos, nrm, tc, ind);
        haveTexCoords = anyTexCoord;
        haveTangents = wantTangents;
        const size_t stride = vertexSize();
        vertices.resize(pos.size() * stride);

        const bool transform = !T.isIdentity();
        mat4 M(1.0f);
        mat3 N(1.0f);
        if (transform) {
            M = T.toMat4();
            N = T.toNormalMatrix();
        }
        for (size_t i = 0; i < pos.size(); i++) {
            float* v = vertices.data() + i * stride;
            vec3 p = pos[i];
            vec3 n = nrm[i];
            if (transform) {
                p = (M * vec4(p, 1.0f)).xyz();
                n = N * n;
            }
            v[0] = p.x(); v[1] = p.y(); v[2] = p.z();
            v[3] = n.x(); v[4] = n.y(); v[5] = n.z();
            if (haveTexCoords) {
                v[6] = tc[i][0];
                v[7] = tc[i][1];
            }
            if (haveTangents) {
                vec3 t = tng[i];
                if (transform)
                    t = N * t;
                v[8] = t.x(); v[9] = t.y(); v[10] = t.z();
            }
        }
    }
};

class MeshInstance : public SceneComponent
{
public:
    const Mesh* mesh;
    const Material* material;
    const Transformation transformation;
    const int animationIndex;
    const mat4 transformationM;
    const mat3 transformationN;

    MeshInstance(const Mesh* mesh_, const Material* m, const Transformation& t, int ai = -1) :
        mesh(mesh_), material(m), transformation(t), animationIndex(ai),
        transformationM(transformation.toMat4()), transformationN(transformation.toNormalMatrix())
    {
    }

    MeshInstance(const Mesh* mesh_, const Material* m, int ai = -1) : MeshInstance(mesh_, m, Transformation(), ai) {}

    /* one hitable handle per triangle, like the reference's one HitableTriangle per triangle */
    virtual std::vector<Hitable*> createHitables() const override
    {
        std::vector<Hitable*> h(mesh->triangleCount());
        for (size_t i = 0; i < h.size(); i++)
            h[i] = new Hitable(Hitable::KindTriangle, this, i, material, animationIndex);
        return h;
    }
};

inline int Flattener::mesh(const Mesh* m)
{
    auto it = meshIndex.find(m);
    if (it != meshIndex.end())
        return it->second;
    wpt_mesh out;
    std::memset(&out, 0, sizeof(out));
    out.vertex_offset = vertices.size();
    out.index_offset = indices.size();
    out.vertex_count = m->vertexCount();
    out.vertex_stride = m->vertexSize();
    out.triangle_count = m->triangleCount();
    out.flags = (m->haveTexCoords ? WPT_MESHF_TEXCOORDS : 0) | (m->haveTangents ? WPT_MESHF_TANGENTS : 0);
    vertices.insert(vertices.end(), m->vertices.begin(), m->vertices.end());
    indices.insert(indices.end(), m->indices.begin(), m->indices.end());
    int index = meshes.size();
    meshes.push_back(out);
    meshIndex[m] = index;
    return index;
}

}
