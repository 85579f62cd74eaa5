/* generator.hpp -- procedural meshes (mirror of reference libwurblpt/generator.hpp:39-733: same
 * function names, parameters, vertex order, index order and per-vertex arithmetic, so a generated
 * mesh is byte-identical to the reference's). All shapes fill [-1,+1]^3 before T is applied.
 *
 * Most shapes are a (rows+1) x (cols+1) vertex grid with two triangles per cell; GridBuilder holds
 * that common part and each generator only supplies the vertex function. */
#pragma once

#include "mesh.hpp"

namespace WurblPT {

namespace detail {

struct GridBuilder {
    std::vector<vec3> positions, normals;
    std::vector<vec2> texcoords;
    std::vector<unsigned int> indices;

    /* Cell triangulations used by the reference: "A" splits (i,j)-(i,j+1)-(i+1,j) / (i,j+1)-(i+1,j+1)-(i+1,j),
     * "B" splits (i,j)-(i+1,j)-(i,j+1) / (i,j+1)-(i+1,j)-(i+1,j+1). */
    template<typename VertexFn>
    void grid(int rows, int cols, bool windingB, VertexFn vertex)
    {
        const unsigned int start = positions.size();
        for (int i = 0; i <= rows; i++) {
            for (int j = 0; j <= cols; j++) {
                vec3 p, n;
                vec2 tc;
                vertex(i, j, p, n, tc);
                positions.push_back(p);
                normals.push_back(n);
                texcoords.push_back(tc);
                if (i < rows && j < cols) {
                    const unsigned int a = start + (i + 0) * (cols + 1) + (j + 0);
                    const unsigned int b = start + (i + 0) * (cols + 1) + (j + 1);
                    const unsigned int c = start + (i + 1) * (cols + 1) + (j + 0);
                    const unsigned int d = start + (i + 1) * (cols + 1) + (j + 1);
                    if (!windingB) {
                        indices.insert(indices.end(), { a, b, c, b, d, c });
                    } else {
                        indices.insert(indices.end(), { a, c, b, b, c, d });
                    }
                }
            }
        }
    }

    void triangle(const vec3& v0, const vec3& v1, const vec3& v2, const vec2& tc0, const vec2& tc1, const vec2& tc2)
    {
        const unsigned int start = positions.size();
        positions.insert(positions.end(), { v0, v1, v2 });
        vec3 n = normalize(cross(v1 - v0, v2 - v0));
        normals.insert(normals.end(), { n, n, n });
        texcoords.insert(texcoords.end(), { tc0, tc1, tc2 });
        indices.insert(indices.end(), { start, start + 1, start + 2 });
    }

    Mesh* finish(const Transformation& T) const { return new Mesh(positions, normals, texcoords, indices, T); }
};

/* one face of the [-1,1]^3 cube in the reference's face numbering: 0 front(+z) 1 back(-z) 2 left(-x)
 * 3 right(+x) 4 top(+y) 5 bottom(-y); tx runs along columns, ty along rows */
inline void cubeFaceVertex(int face, float tx, float ty, vec3& p, vec3& n)
{
    switch (face) {
    case 0: p = vec3(-1.0f + tx, -1.0f + ty, 1.0f); n = vec3(0.0f, 0.0f, +1.0f); break;
    case 1: p = vec3(1.0f - tx, -1.0f + ty, -1.0f); n = vec3(0.0f, 0.0f, -1.0f); break;
    case 2: p = vec3(-1.0f, -1.0f + ty, -1.0f + tx); n = vec3(-1.0f, 0.0f, 0.0f); break;
    case 3: p = vec3(1.0f, -1.0f + ty, 1.0f - tx); n = vec3(+1.0f, 0.0f, 0.0f); break;
    case 4: p = vec3(-1.0f + ty, 1.0f, -1.0f + tx); n = vec3(0.0f, +1.0f, 0.0f); break;
    default: p = vec3(1.0f - ty, -1.0f, -1.0f + tx); n = vec3(0.0f, -1.0f, 0.0f); break;
    }
}

/* a disk of radius innerRadius..1 in the plane `height` of axis y (caps) */
inline void capVertex(int i, int j, int loops, int slices, float innerRadius, float y, vec3& p, vec3& n, vec2& tc)
{
    float ty = static_cast<float>(i) / loops;
    float r = innerRadius + ty * (1.0f - innerRadius);
    float tx = static_cast<float>(j) / slices;
    float alpha = tx * (2.0f * pi) + pi_2;
    p = vec3(r * cos(alpha), y, r * sin(alpha));
    n = vec3(0.0f, y, 0.0f);
    tc = vec2(1.0f - tx, ty);
}

}

/* corners (-1,-1,0) (+1,-1,0) (+1,+1,0) (-1,+1,0) */
inline Mesh* generateQuad(const Transformation& T = Transformation(), int slices = 1)
{
    detail::GridBuilder g;
    g.grid(slices, slices, false, [slices](int i, int j, vec3& p, vec3& n, vec2& tc) {
        float ty = i / (slices / 2.0f);
        float tx = j / (slices / 2.0f);
        p = vec3(-1.0f + tx, -1.0f + ty, 0.0f);
        n = vec3(0.0f, 0.0f, 1.0f);
        tc = 0.5f * vec2(tx, ty);
    });
    return g.finish(T);
}

inline Mesh* generateCube(const Transformation& T = Transformation(), int slices = 1)
{
    detail::GridBuilder g;
    for (int face = 0; face < 6; face++) {
        g.grid(slices, slices, false, [slices, face](int i, int j, vec3& p, vec3& n, vec2& tc) {
            float ty = i / (slices / 2.0f);
            float tx = j / (slices / 2.0f);
            detail::cubeFaceVertex(face, tx, ty, p, n);
            tc = 0.5f * vec2(tx, ty);
        });
    }
    return g.finish(T);
}

/* side: 0 posx, 1 negx, 2 posy, 3 negy, 4 posz, 5 negz */
inline Mesh* generateCubeSide(int side, const Transformation& T = Transformation(), int slices = 1)
{
    static const int faceOfSide[6] = { 3, 2, 4, 5, 0, 1 };
    const int face = faceOfSide[side >= 0 && side < 6 ? side : 0];
    detail::GridBuilder g;
    g.grid(slices, slices, false, [slices, face](int i, int j, vec3& p, vec3& n, vec2& tc) {
        float ty = i / (slices / 2.0f);
        float tx = j / (slices / 2.0f);
        detail::cubeFaceVertex(face, tx, ty, p, n);
        tc = 0.5f * vec2(tx, ty);
    });
    return g.finish(T);
}

inline Mesh* generateDisk(const Transformation& T = Transformation(), float innerRadius = 0.0f, int slices = 40)
{
    const int loops = 1;
    detail::GridBuilder g;
    g.grid(loops, slices, true, [loops, slices, innerRadius](int i, int j, vec3& p, vec3& n, vec2& tc) {
        float ty = static_cast<float>(i) / loops;
        float r = innerRadius + ty * (1.0f - innerRadius);
        float tx = static_cast<float>(j) / slices;
        float alpha = tx * (2.0f * pi) + pi_2;
        p = vec3(r * cos(alpha), r * sin(alpha), 0.0f);
        n = vec3(0.0f, 0.0f, 1.0f);
        tc = vec2(1.0f - tx, ty);
    });
    return g.finish(T);
}

inline Mesh* generateSphere(const Transformation& T = Transformation(), int slices = 40, int stacks = 20)
{
    detail::GridBuilder g;
    g.grid(stacks, slices, false, [slices, stacks](int i, int j, vec3& p, vec3& n, vec2& tc) {
        float ty = static_cast<float>(i) / stacks;
        float lat = ty * pi;
        float tx = static_cast<float>(j) / slices;
        float lon = tx * (2.0f * pi) - pi_2;
        float sinlat = sin(lat);
        float coslat = cos(lat);
        float sinlon = sin(lon);
        float coslon = cos(lon);
        p = vec3(sinlat * coslon, coslat, sinlat * sinlon);
        n = p;
        tc = vec2(1.0f - tx, 1.0f - ty);
    });
    return g.finish(T);
}

inline Mesh* generateCylinder(bool closed, const Transformation& T, int slices)
{
    const int stacks = 1;
    detail::GridBuilder g;
    g.grid(stacks, slices, false, [slices, stacks](int i, int j, vec3& p, vec3& n, vec2& tc) {
        float ty = static_cast<float>(i) / stacks;
        float tx = static_cast<float>(j) / slices;
        float alpha = tx * (2.0f * pi) - pi_2;
        float x = cos(alpha);
        float y = -(ty * 2.0f - 1.0f);
        float z = sin(alpha);
        p = vec3(x, y, z);
        n = vec3(x, 0.0f, z);
        tc = vec2(1.0f - tx, 1.0f - ty);
    });
    if (closed) {
        for (int side = 0; side < 2; side++) {
            const float y = (side == 0 ? +1.0f : -1.0f);
            g.grid(1, slices, side == 1, [slices, y](int i, int j, vec3& p, vec3& n, vec2& tc) {
                detail::capVertex(i, j, 1, slices, 0.0f, y, p, n, tc);
            });
        }
    }
    return g.finish(T);
}

inline Mesh* generateCylinder(const Transformation& T = Transformation(), int slices = 40) { return generateCylinder(false, T, slices); }
inline Mesh* generateClosedCylinder(const Transformation& T = Transformation(), int slices = 40) { return generateCylinder(true, T, slices); }

inline Mesh* generateCone(bool closed, const Transformation& T, int slices = 40, int stacks = 20)
{
    detail::GridBuilder g;
    g.grid(stacks, slices, false, [slices, stacks](int i, int j, vec3& p, vec3& n, vec2& tc) {
        float ty = static_cast<float>(i) / stacks;
        float tx = static_cast<float>(j) / slices;
        float alpha = tx * (2.0f * pi) - pi_2;
        float x = ty * cos(alpha);
        float y = -(ty * 2.0f - 1.0f);
        float z = ty * sin(alpha);
        float nx = x, ny = 0.5f, nz = z;
        float nl = sqrt(nx * nx + ny * ny + nz * nz);
        p = vec3(x, y, z);
        n = vec3(nx, ny, nz) / nl;
        tc = vec2(1.0f - tx, 1.0f - ty);
    });
    if (closed) {
        g.grid(1, slices, true, [slices](int i, int j, vec3& p, vec3& n, vec2& tc) {
            detail::capVertex(i, j, 1, slices, 0.0f, -1.0f, p, n, tc);
        });
    }
    return g.finish(T);
}

inline Mesh* generateCone(const Transformation& T = Transformation(), int slices = 40, int stacks = 20) { return generateCone(false, T, slices, stacks); }
inline Mesh* generateClosedCone(const Transformation& T = Transformation(), int slices = 40, int stacks = 20) { return generateCone(true, T, slices, stacks); }

inline Mesh* generateTorus(const Transformation& T = Transformation(), float innerRadius = 0.4f, int sides = 40, int rings = 40)
{
    const float ringradius = (1.0f - innerRadius) / 2.0f;
    const float ringcenter = innerRadius + ringradius;
    detail::GridBuilder g;
    g.grid(sides, rings, false, [sides, rings, ringradius, ringcenter](int i, int j, vec3& p, vec3& n, vec2& tc) {
        float ty = static_cast<float>(i) / sides;
        float alpha = ty * (2.0f * pi) - pi_2;
        float c = cos(alpha);
        float s = sin(alpha);
        float tx = static_cast<float>(j) / rings;
        float beta = tx * (2.0f * pi) - pi;
        float x = ringcenter + ringradius * cos(beta);
        float y = 0.0f;
        float z = ringradius * sin(beta);
        float rx = c * x + s * y;
        float ry = c * y - s * x;
        float rz = z;
        p = vec3(rx, ry, rz);
        float rcx = c * ringcenter;
        float rcy = -s * ringcenter;
        float rcz = 0.0f;
        float nx = rx - rcx, ny = ry - rcy, nz = rz - rcz;
        float nl = sqrt(nx * nx + ny * ny + nz * nz);
        n = vec3(nx, ny, nz) / nl;
        tc = vec2(1.0f - tx, 1.0f - ty);
    });
    return g.finish(T);
}

inline Mesh* generateTetrahedron(const Transformation& T = Transformation())
{
    const float a = 1.0f / 3.0f;
    const float b = sqrt(8.0f / 9.0f);
    const float c = sqrt(2.0f / 9.0f);
    const float d = sqrt(2.0f / 3.0f);
    const vec3 v[4] = { vec3(-c, -a, d), vec3(b, -a, 0.0f), vec3(-d, -a, -d), vec3(0.0f, 1.0f, 0.0f) };
    const vec2 t0(0.0f, 0.0f), t1(1.0f, 0.0f), t2(0.5f, 1.0f);
    static const int faces[4][3] = { { 0, 1, 3 }, { 1, 2, 3 }, { 2, 0, 3 }, { 2, 1, 0 } };
    detail::GridBuilder g;
    for (const auto& f : faces)
        g.triangle(v[f[0]], v[f[1]], v[f[2]], t0, t1, t2);
    return g.finish(T);
}

inline Mesh* generateOctahedron(const Transformation& T = Transformation())
{
    const vec3 v[6] = { vec3(0.0f, -1.0f, 0.0f), vec3(0.0f, 0.0f, -1.0f), vec3(+1.0f, 0.0f, 0.0f),
        vec3(0.0f, 0.0f, +1.0f), vec3(-1.0f, 0.0f, 0.0f), vec3(0.0f, +1.0f, 0.0f) };
    const vec2 t0(0.0f, 0.0f), t1(1.0f, 0.0f), t2(0.5f, 1.0f);
    static const int lower[4][3] = { { 1, 2, 0 }, { 2, 3, 0 }, { 3, 4, 0 }, { 4, 1, 0 } };
    static const int upper[4][3] = { { 1, 5, 2 }, { 2, 5, 3 }, { 3, 5, 4 }, { 4, 5, 1 } };
    detail::GridBuilder g;
    for (const auto& f : lower)
        g.triangle(v[f[0]], v[f[1]], v[f[2]], t0, t1, t2);
    for (const auto& f : upper)
        g.triangle(v[f[0]], v[f[1]], v[f[2]], t1, t2, t0);
    return g.finish(T);
}

inline Mesh* generateIcosahedron(const Transformation& T = Transformation())
{
    const float r = 2.0f / (1 + sqrt(5.0f));   // 1 / golden ratio
    const vec3 v[12] = {
        vec3(0.0f, +r, -1.0f), vec3(+r, +1.0f, 0.0f), vec3(-r, +1.0f, 0.0f), vec3(0.0f, +r, +1.0f),
        vec3(0.0f, -r, +1.0f), vec3(-1.0f, 0.0f, +r), vec3(0.0f, -r, -1.0f), vec3(+1.0f, 0.0f, -r),
        vec3(+1.0f, 0.0f, +r), vec3(-1.0f, 0.0f, -r), vec3(+r, -1.0f, 0.0f), vec3(-r, -1.0f, 0.0f)
    };
    const vec2 t0(0.0f, 0.0f), t1(1.0f, 0.0f), t2(0.5f, 1.0f);
    static const int faces[20][3] = {
        { 2, 1, 0 }, { 1, 2, 3 }, { 5, 4, 3 }, { 4, 8, 3 }, { 7, 6, 0 }, { 6, 9, 0 }, { 11, 10, 4 }, { 10, 11, 6 },
        { 9, 5, 2 }, { 5, 9, 11 }, { 8, 7, 1 }, { 7, 8, 10 }, { 2, 5, 3 }, { 8, 1, 3 }, { 9, 2, 0 }, { 1, 7, 0 },
        { 11, 9, 6 }, { 7, 10, 6 }, { 5, 11, 4 }, { 10, 8, 4 }
    };
    detail::GridBuilder g;
    for (const auto& f : faces)
        g.triangle(v[f[0]], v[f[1]], v[f[2]], t0, t1, t2);
    return g.finish(T);
}

}
