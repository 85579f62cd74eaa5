/* geometryproc.hpp -- vertex normals and tangents from indexed triangles (mirror of reference
 * libwurblpt/geometryproc.hpp:36-226; normals after Thuermer & Wuethrich 1998, tangents by the
 * usual per-triangle uv-gradient accumulation followed by Gram-Schmidt). Results feed the vertex
 * buffer, so the arithmetic order follows the reference. */
#pragma once

#include <vector>

#include "gvm.hpp"
#include "constants.hpp"

namespace WurblPT {

typedef enum {
    NormalsFromFirstFace,
    NormalsFromFaceAverage,
    NormalsFromWeightedFaceAverage
} NormalSource;

inline std::vector<vec3> computeNormals(
        const std::vector<vec3>& positions,
        const std::vector<unsigned int> indices,
        NormalSource normalSource = NormalsFromWeightedFaceAverage)
{
    const size_t nv = positions.size();
    const size_t nt = indices.size() / 3;
    std::vector<vec3> faceNormals(nt);
    std::vector<std::vector<unsigned int>> facesOf(nv);
    for (size_t f = 0; f < nt; f++) {
        const unsigned int* ix = indices.data() + 3 * f;
        vec3 e0 = positions[ix[1]] - positions[ix[0]];
        vec3 e1 = positions[ix[2]] - positions[ix[0]];
        vec3 e2 = e1 - e0;
        vec3 fn(0.0f, 0.0f, 1.0f);
        if (dot(e0, e0) > 0.0f && dot(e1, e1) > 0.0f && dot(e2, e2) > 0.0f) {
            vec3 c = cross(e0, e1);
            if (dot(c, c) > 0.0f)
                fn = normalize(c);
        }
        faceNormals[f] = fn;
        for (int k = 0; k < 3; k++)
            facesOf[ix[k]].push_back(f);
    }
    std::vector<vec3> normals(nv);
    for (size_t v = 0; v < nv; v++) {
        const std::vector<unsigned int>& faces = facesOf[v];
        vec3 n(0.0f);
        if (faces.size() == 0) {
            n = vec3(0.0f, 0.0f, 1.0f);
        } else if (faces.size() == 1) {
            n = faceNormals[faces[0]];
        } else {
            if (normalSource == NormalsFromWeightedFaceAverage) {
                /* weight each face by the angle it spans at this vertex */
                for (unsigned int f : faces) {
                    const unsigned int* ix = indices.data() + 3 * f;
                    vec3 e0, e1;
                    if (v == ix[0]) {
                        e0 = positions[ix[1]]; e1 = positions[ix[2]];
                    } else if (v == ix[1]) {
                        e0 = positions[ix[2]]; e1 = positions[ix[0]];
                    } else {
                        e0 = positions[ix[0]]; e1 = positions[ix[1]];
                    }
                    e0 = e0 - positions[v];
                    e1 = e1 - positions[v];
                    if (dot(e0, e0) <= 0.0f || dot(e1, e1) <= 0.0f)
                        continue;
                    float x = dot(normalize(e0), normalize(e1));
                    float alpha = acos(clamp(x, -1.0f, +1.0f));
                    n += alpha * faceNormals[f];
                }
            }
            if (normalSource == NormalsFromFaceAverage
                    || (normalSource == NormalsFromWeightedFaceAverage && dot(n, n) <= 0.0f)) {
                n = vec3(0.0f);
                for (unsigned int f : faces)
                    n += faceNormals[f];
            }
            if (normalSource == NormalsFromFirstFace || dot(n, n) <= 0.0f)
                n = faceNormals[faces[0]];
            n = normalize(n);
        }
        normals[v] = n;
    }
    return normals;
}

inline std::vector<vec3> computeTangents(
        const std::vector<vec3>& positions,
        const std::vector<vec3>& normals,
        const std::vector<vec2>& texcoords,
        const std::vector<unsigned int> indices)
{
    const size_t nv = positions.size();
    const size_t nt = indices.size() / 3;
    std::vector<vec3> tangents(nv, vec3(0.0f));
    for (size_t f = 0; f < nt; f++) {
        const unsigned int i0 = indices[3 * f + 0], i1 = indices[3 * f + 1], i2 = indices[3 * f + 2];
        vec3 e1 = positions[i1] - positions[i0];
        vec3 e2 = positions[i2] - positions[i0];
        float s1 = texcoords[i1].x() - texcoords[i0].x();
        float t1 = texcoords[i1].y() - texcoords[i0].y();
        float s2 = texcoords[i2].x() - texcoords[i0].x();
        float t2 = texcoords[i2].y() - texcoords[i0].y();
        float det = (s1 * t2 - s2 * t1);
        if (abs(det) > epsilon) {
            vec3 tp = 1.0f / det * (t2 * e1 - t1 * e2);
            tangents[i0] += tp;
            tangents[i1] += tp;
            tangents[i2] += tp;
        }
    }
    for (size_t v = 0; v < nv; v++) {
        const vec3 tp = tangents[v];
        vec3 t(1.0f, 0.0f, 0.0f);
        if (dot(tp, tp) > 0.0f)
            t = normalize(tp - dot(normals[v], tp) * normals[v]);
        tangents[v] = t;
    }
    return tangents;
}

}
