/*
 * oracle/port.cpp -- TEST INFRASTRUCTURE: a CPU restatement of the reference's primary-ray path.
 *
 * This is the checker's own statement of the algorithm, written from the reference's headers and
 * independent of both the product code (wurblpt_b200/csrc) and the compiled reference
 * (oracle/_ref). It covers the integer/index part of the hot path, where bit-exact parity is
 * demanded: Prng, camera rays without randomness, Scene::updateBVH's SAH tree, and BVH::hit with
 * the slab / watertight-triangle / sphere tests, for static scenes of triangles and spheres.
 * Radiance (materials, lights, media) is checked against the compiled reference only.
 *
 * Parity status: pinned. tests/test_oracle_port.py checks this file bit for bit against
 * oracle/_ref/libwpt_ref_strict.so and against tests/golden/*.npz (generated from the reference).
 * Build: -O2 -ffp-contract=off (IEEE-strict, like the strict reference build).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this library.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "wurblpt_b200.h"

namespace {

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(V3 a, V3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline V3 operator*(float s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline V3 operator/(V3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline float at(V3 v, int k) { return k == 0 ? v.x : k == 1 ? v.y : v.z; }
/* gvm.hpp:1183-1189: d = 0; d += a[i] * b[i] */
inline float dot(V3 a, V3 b) { float d = 0.0f; d += a.x * b.x; d += a.y * b.y; d += a.z * b.z; return d; }
/* gvm.hpp:1225-1231 */
inline V3 cross(V3 v, V3 w) { return { v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x }; }
/* gvm.hpp:1201-1204 */
inline V3 normalize(V3 v) { return v / std::sqrt(dot(v, v)); }
/* gvm.hpp:90-96 */
inline float gmin(float x, float y) { return x < y ? x : y; }
inline float gmax(float x, float y) { return x > y ? x : y; }

/* ---- prng.hpp:47-101 ---- */
struct Prng {
    uint32_t s[4];
    static uint64_t splitmix64(uint64_t x)
    {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    explicit Prng(unsigned int pixelIndex)
    {
        uint64_t seed = pixelIndex;
        seed += 42;
        uint64_t s01 = splitmix64(seed), s23 = splitmix64(s01);
        s[0] = uint32_t(s01 >> 32); s[1] = uint32_t(s01); s[2] = uint32_t(s23 >> 32); s[3] = uint32_t(s23);
    }
    uint32_t next()
    {
        const uint32_t result = s[0] + s[3];
        const uint32_t t = s[1] << 9;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t;
        s[3] = (s[3] << 11) | (s[3] >> 21);
        return result;
    }
    float in01() { return float(next() >> 8) * 0x1.0p-24f; }
};

struct Ray { V3 o, d; };

/* ---- hitable.hpp:66-113 ---- */
struct Helper {
    V3 inv;
    int kx, ky, kz;
    float Sx, Sy, Sz;
    explicit Helper(const Ray& r)
    {
        inv = { 1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z };
        float ax = std::fabs(r.d.x), ay = std::fabs(r.d.y), az = std::fabs(r.d.z);
        if (az >= ay && az >= ax) kz = 2; else if (ay >= ax) kz = 1; else kz = 0;
        kx = kz + 1; if (kx == 3) kx = 0;
        ky = kx + 1; if (ky == 3) ky = 0;
        if (at(r.d, kz) < 0.0f) std::swap(kx, ky);
        Sx = at(r.d, kx) * at(inv, kz);
        Sy = at(r.d, ky) * at(inv, kz);
        Sz = at(inv, kz);
    }
};

struct Box { V3 lo, hi; };

/* aabb.hpp:70-86 with gvm's vector min/max (gvm.hpp:1060-1090) and reductions (gvm.hpp:1275-1291) */
inline bool mayHit(const Box& b, const Ray& r, float amin, float amax, V3 inv)
{
    V3 t0 = (b.lo - r.o) * inv, t1 = (b.hi - r.o) * inv;
    float tmin[4] = { amin, gmin(t0.x, t1.x), gmin(t0.y, t1.y), gmin(t0.z, t1.z) };
    float tmax[4] = { amax, gmax(t0.x, t1.x), gmax(t0.y, t1.y), gmax(t0.z, t1.z) };
    float mx = tmin[0], mn = tmax[0];
    for (int i = 1; i < 4; i++) { if (tmin[i] > mx) mx = tmin[i]; if (tmax[i] < mn) mn = tmax[i]; }
    return mx <= mn;
}

struct Prim {
    int type;           /* 0 triangle, 1 sphere */
    V3 v0, v1, v2;      /* triangle corners in world space */
    V3 center; float radius;
    Box box;
    int32_t id;         /* global hitable id */
};

/* hitable_triangle.hpp:189-262: the watertight test on the ray's whole line. On a hit: Ts = T with the sign of det
 * folded in, ds = |det| (:263 compares these with the bounds), a = the distance (:266-267). */
inline bool triangleOnLine(const Prim& p, const Ray& r, const Helper& h, float& Ts, float& ds, float& a)
{
    const V3 A = p.v0 - r.o, B = p.v1 - r.o, C = p.v2 - r.o;
    const float Ax = at(A, h.kx) - h.Sx * at(A, h.kz), Ay = at(A, h.ky) - h.Sy * at(A, h.kz);
    const float Bx = at(B, h.kx) - h.Sx * at(B, h.kz), By = at(B, h.ky) - h.Sy * at(B, h.kz);
    const float Cx = at(C, h.kx) - h.Sx * at(C, h.kz), Cy = at(C, h.ky) - h.Sy * at(C, h.kz);
    float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    const float eps = float(std::numeric_limits<long double>::epsilon());
    if (std::fabs(U) < eps || std::fabs(V) < eps || std::fabs(W) < eps) {
        U = float(double(Cx) * double(By) - double(Cy) * double(Bx));
        V = float(double(Ax) * double(Cy) - double(Ay) * double(Cx));
        W = float(double(Bx) * double(Ay) - double(By) * double(Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
        return false;
    const float det = U + V + W;
    if (det == 0.0f)
        return false;
    const float Az = h.Sz * at(A, h.kz), Bz = h.Sz * at(B, h.kz), Cz = h.Sz * at(C, h.kz);
    const float T = U * Az + V * Bz + W * Cz;
    Ts = std::signbit(det) ? -T : T;        /* xorf with the sign mask */
    ds = std::signbit(det) ? -det : det;
    a = (1.0f / det) * T;
    return true;
}

/* hitable_triangle.hpp:189-271 (distance only) */
inline bool hitTriangle(const Prim& p, const Ray& r, const Helper& h, float amin, float amax, float& a)
{
    float Ts, ds, dist;
    if (!triangleOnLine(p, r, h, Ts, ds, dist))
        return false;
    if (Ts < amin * ds || Ts > amax * ds)
        return false;
    a = dist;
    return true;
}

/* hitable_sphere.hpp:111-153 (distance only) */
inline bool hitSphere(const Prim& p, const Ray& r, float amin, float amax, float& a)
{
    const V3 oc = r.o - p.center;
    const float aq = -dot(oc, r.d);
    const V3 tmp = oc - dot(oc, r.d) * r.d;
    const float disc = p.radius * p.radius - dot(tmp, tmp);
    if (!(disc > 0.0f))
        return false;
    float a1, a2;
    if (aq < 0.0f) { a2 = aq - std::sqrt(disc); a1 = 2.0f * aq - a2; }
    else { a1 = aq + std::sqrt(disc); a2 = 2.0f * aq - a1; }
    if (a2 > amin && a2 < amax) { a = a2; return true; }
    if (a1 > amin && a1 < amax) { a = a1; return true; }
    return false;
}

struct Node { Box box; uint32_t link; int32_t prim; };      /* prim < 0: inner, child 1 = next node */

struct Tree {
    std::vector<Prim> prims;
    std::vector<Node> nodes;
    std::vector<unsigned int> subset;
    std::vector<float> areas0, areas1;

    static Box merge(const Box& a, const Box& b)
    {
        return { { gmin(a.lo.x, b.lo.x), gmin(a.lo.y, b.lo.y), gmin(a.lo.z, b.lo.z) },
                 { gmax(a.hi.x, b.hi.x), gmax(a.hi.y, b.hi.y), gmax(a.hi.z, b.hi.z) } };
    }
    static float area(const Box& b) { V3 l = b.hi - b.lo; return 2.0f * (l.x * l.y + l.y * l.z + l.x * l.z); }   /* aabb.hpp:64-68 */
    static int longestAxis(const Box& b)            /* aabb.hpp:52-62 */
    {
        V3 l = b.hi - b.lo;
        if (l.x > l.y && l.x > l.z) return 0;
        return l.y > l.z ? 1 : 2;
    }
    float center(unsigned int i, int axis) const    /* aabb.hpp:47-50 */
    {
        const Box& b = prims[i].box;
        return at(0.5f * (b.lo + b.hi), axis);
    }

    /* BVHNode::build (bvh.hpp:93-164) + BVH::flatten (bvh.hpp:233-246) */
    void build(size_t off, size_t n, uint32_t node)
    {
        if (n == 1) {
            nodes[node] = { prims[subset[off]].box, 0u, int32_t(subset[off]) };
            return;
        }
        Box box = prims[subset[off]].box;
        for (size_t i = 1; i < n; i++) box = merge(box, prims[subset[off + i]].box);
        const int axis = longestAxis(box);
        std::sort(subset.begin() + off, subset.begin() + off + n,
                [this, axis](unsigned int i, unsigned int j) { return center(i, axis) < center(j, axis); });
        Box b0 = prims[subset[off]].box;
        areas0[off] = area(b0);
        for (size_t i = 1; i + 1 < n; i++) { b0 = merge(b0, prims[subset[off + i]].box); areas0[off + i] = area(b0); }
        Box b1 = prims[subset[off + n - 1]].box;
        areas1[off + n - 1] = area(b1);
        for (size_t i = n - 2; i > 0; i--) { b1 = merge(b1, prims[subset[off + i]].box); areas1[off + i] = area(b1); }
        size_t split = 1;
        float best = 1 * areas0[off] + float(n - 1) * areas1[off + 1];
        for (size_t i = 2; i < n; i++) {
            float c = float(int(i)) * areas0[off + i - 1] + float(n - i) * areas1[off + i];
            if (c < best) { best = c; split = i; }
        }
        const uint32_t child1 = node + uint32_t(2 * split);
        nodes[node] = { box, child1, -1 };
        build(off, split, node + 1);
        build(off + split, n - split, child1);
    }

    void finish()
    {
        const size_t n = prims.size();
        nodes.assign(n ? 2 * n - 1 : 0, Node());
        subset.resize(n); areas0.resize(n); areas1.resize(n);
        for (size_t i = 0; i < n; i++) subset[i] = i;
        if (n) build(0, n, 0);
    }

    /* BVH::hit (bvh.hpp:277-311) */
    bool hit(const Ray& r, float amin, float amax, int32_t& id, float& dist) const
    {
        if (nodes.empty())
            return false;
        const Helper h(r);
        bool have = false;
        size_t stack[128], sp = 0, cur = 0;
        for (;;) {
            const Node& n = nodes[cur];
            if (mayHit(n.box, r, amin, amax, h.inv)) {
                if (n.prim >= 0) {
                    const Prim& p = prims[n.prim];
                    float a;
                    bool ok = (p.type == 0 ? hitTriangle(p, r, h, amin, amax, a) : hitSphere(p, r, amin, amax, a));
                    if (ok) { have = true; id = p.id; dist = a; amax = a; }
                    if (sp == 0) break;
                    cur = stack[--sp];
                } else {
                    stack[sp++] = n.link;
                    cur++;
                }
            } else {
                if (sp == 0) break;
                cur = stack[--sp];
            }
        }
        return have;
    }
};

/* (M * vec4(v,1)).xyz, gvm.hpp:1481-1491 */
inline V3 xform(const float* M, V3 v)
{
    V3 r;
    float* o[3] = { &r.x, &r.y, &r.z };
    const float w[4] = { v.x, v.y, v.z, 1.0f };
    for (int i = 0; i < 3; i++) { float s = 0.0f; for (int j = 0; j < 4; j++) s += M[j * 4 + i] * w[j]; *o[i] = s; }
    return r;
}

/* Expand the components of a static scene into primitives (hitable_triangle.hpp:194-206,336-350;
 * hitable_sphere.hpp:97-101). Returns 0, or -4 for what this restatement does not cover. */
int buildTree(const wpt_scene_desc* d, Tree& t)
{
    if (d->n_media > 0)
        return WPT_E_UNSUPPORTED;
    int32_t id = 0;
    for (uint32_t c = 0; c < d->n_components; c++) {
        const wpt_component& comp = d->components[c];
        const bool hitable = (comp.flags & WPT_COMPF_HITABLE) != 0;
        if (comp.type == WPT_COMP_MESH_INSTANCE) {
            const wpt_instance& in = d->instances[comp.ref];
            const wpt_mesh& m = d->meshes[in.mesh];
            if (in.animation >= 0)
                return WPT_E_UNSUPPORTED;
            for (uint32_t k = 0; k < m.triangle_count; k++, id++) {
                if (!hitable) continue;
                Prim p; std::memset(&p, 0, sizeof(p));
                p.type = 0; p.id = id;
                V3* corner[3] = { &p.v0, &p.v1, &p.v2 };
                for (int j = 0; j < 3; j++) {
                    const float* v = d->vertices + m.vertex_offset + size_t(d->indices[m.index_offset + 3 * k + j]) * m.vertex_stride;
                    V3 q = { v[0], v[1], v[2] };
                    if (in.has_transform) q = xform(in.M, q);
                    *corner[j] = q;
                }
                p.box.lo = { gmin(gmin(p.v0.x, p.v1.x), p.v2.x), gmin(gmin(p.v0.y, p.v1.y), p.v2.y), gmin(gmin(p.v0.z, p.v1.z), p.v2.z) };
                p.box.hi = { gmax(gmax(p.v0.x, p.v1.x), p.v2.x), gmax(gmax(p.v0.y, p.v1.y), p.v2.y), gmax(gmax(p.v0.z, p.v1.z), p.v2.z) };
                t.prims.push_back(p);
            }
        } else if (comp.type == WPT_COMP_SPHERE) {
            const wpt_sphere& s = d->spheres[comp.ref];
            if (s.animation >= 0)
                return WPT_E_UNSUPPORTED;
            if (hitable) {
                Prim p; std::memset(&p, 0, sizeof(p));
                p.type = 1; p.id = id;
                p.center = { s.center[0], s.center[1], s.center[2] };
                p.radius = s.radius;
                p.box.lo = p.center - V3{ s.radius, s.radius, s.radius };
                p.box.hi = p.center + V3{ s.radius, s.radius, s.radius };
                t.prims.push_back(p);
            }
            id++;
        } else {
            return WPT_E_UNSUPPORTED;
        }
    }
    t.finish();
    return 0;
}

/* quaternion * vector (gvm.hpp:1713-1720) and Transformation * vector (transformation.hpp:83-86) */
inline V3 qrot(const float* q, V3 v)
{
    V3 s = { q[0], q[1], q[2] };
    V3 t = 2.0f * cross(s, v);
    return (v + q[3] * t) + cross(s, t);
}

/* Camera::getRay without randomness (camera.hpp:123-185), perspective and surround, no lens distortion */
int cameraRay(const wpt_camera_desc* c, float p, float q, Ray& r)
{
    if (c->distortion_type != WPT_DISTORTION_NONE || c->n_keyframes != 0)
        return WPT_E_UNSUPPORTED;
    float shift = 0.0f;
    if (c->stereoscopic_distance > 0.0f) {
        q *= 2.0f;
        if (q < 1.0f) shift = -0.5f * c->stereoscopic_distance;
        else { q -= 1.0f; shift = +0.5f * c->stereoscopic_distance; }
    }
    V3 O, D;
    const float pi = 3.14159274f;
    if (c->surround_mode == WPT_SURROUND_OFF) {
        V3 P = { c->l + p * (c->r - c->l), c->b + q * (c->t - c->b), -1.0f };   /* mix(), gvm.hpp:166 */
        O = { 0.0f, 0.0f, 0.0f };
        D = P - O;
        O = O + V3{ shift, 0.0f, 0.0f };
    } else {
        float lon = (2.0f * p - 1.0f) * pi;
        if (c->surround_mode == WPT_SURROUND_180) lon *= 0.5f;
        float lat = (q - 0.5f) * pi;
        D = { std::cos(lat) * std::sin(lon), std::sin(lat), -std::cos(lat) * std::cos(lon) };
        V3 e = { -std::cos(lon), 0.0f, -std::sin(lon) };
        O = { e.x * shift, e.y * shift, e.z * shift };
    }
    const V3 T = { c->pose_translation[0], c->pose_translation[1], c->pose_translation[2] };
    const V3 Sc = { c->pose_scaling[0], c->pose_scaling[1], c->pose_scaling[2] };
    r.o = T + qrot(c->pose_rotation, O * Sc);
    r.d = normalize(qrot(c->pose_rotation, D));
    return 0;
}

}

extern "C" {

int port_prng_kat(const uint32_t* seeds, uint32_t nSeeds, uint32_t nDraws, float* out)
{
    for (uint32_t i = 0; i < nSeeds; i++) {
        Prng p(seeds[i]);
        for (uint32_t k = 0; k < nDraws; k++)
            out[size_t(i) * nDraws + k] = p.in01();
    }
    return 0;
}

/* the loop of getGroundTruth() (wurblpt.hpp:659-668) */
int port_primary_hits(const wpt_scene_desc* desc, const wpt_camera_desc* cam, uint32_t width, uint32_t height,
        float minHit, int32_t* ids, float* a)
{
    Tree tree;
    int rc = buildTree(desc, tree);
    if (rc != 0)
        return rc;
    int err = 0;
    #pragma omp parallel for schedule(dynamic, 64)
    for (long long pixel = 0; pixel < (long long)width * height; pixel++) {
        unsigned int y = pixel / width, x = pixel % width;
        Ray r;
        if (cameraRay(cam, (x + 0.5f) / width, (y + 0.5f) / height, r) != 0) { err = 1; continue; }
        int32_t id = -1;
        float dist = 0.0f;
        if (!tree.hit(r, minHit, std::numeric_limits<float>::max(), id, dist)) { id = -1; dist = 0.0f; }
        ids[pixel] = id;
        a[pixel] = dist;
    }
    return err ? WPT_E_UNSUPPORTED : 0;
}

/* BVH::hit for caller-supplied rays */
int port_trace_rays(const wpt_scene_desc* desc, const float* origins, const float* directions, uint64_t n,
        float minHit, int32_t* ids, float* a)
{
    Tree tree;
    int rc = buildTree(desc, tree);
    if (rc != 0)
        return rc;
    #pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; i++) {
        Ray r = { { origins[3 * i], origins[3 * i + 1], origins[3 * i + 2] }, { directions[3 * i], directions[3 * i + 1], directions[3 * i + 2] } };
        int32_t id = -1;
        float dist = 0.0f;
        if (!tree.hit(r, minHit, std::numeric_limits<float>::max(), id, dist)) { id = -1; dist = 0.0f; }
        ids[i] = id;
        a[i] = dist;
    }
    return 0;
}

/* The hitables that bound medium `medium` (static triangles only) as a Tree: HitableMedium::updateBVH
 * (hitable_medium.hpp:59-62) builds the same kind of BVH over them as the scene does over its hitables. */
int buildBoundaryTree(const wpt_scene_desc* d, uint32_t medium, Tree& t)
{
    if (medium >= d->n_media)
        return WPT_E_INVALID;
    std::vector<Prim> all;      /* every hitable of the scene by id (component order); type -1: not covered here */
    for (uint32_t c = 0; c < d->n_components; c++) {
        const wpt_component& comp = d->components[c];
        if (comp.type == WPT_COMP_MESH_INSTANCE) {
            const wpt_instance& in = d->instances[comp.ref];
            const wpt_mesh& m = d->meshes[in.mesh];
            for (uint32_t k = 0; k < m.triangle_count; k++) {
                Prim p; std::memset(&p, 0, sizeof(p));
                p.type = (in.animation >= 0 ? -1 : 0);
                p.id = int32_t(all.size());
                V3* corner[3] = { &p.v0, &p.v1, &p.v2 };
                for (int j = 0; j < 3; j++) {
                    const float* v = d->vertices + m.vertex_offset + size_t(d->indices[m.index_offset + 3 * k + j]) * m.vertex_stride;
                    V3 q = { v[0], v[1], v[2] };
                    if (in.has_transform) q = xform(in.M, q);
                    *corner[j] = q;
                }
                p.box.lo = { gmin(gmin(p.v0.x, p.v1.x), p.v2.x), gmin(gmin(p.v0.y, p.v1.y), p.v2.y), gmin(gmin(p.v0.z, p.v1.z), p.v2.z) };
                p.box.hi = { gmax(gmax(p.v0.x, p.v1.x), p.v2.x), gmax(gmax(p.v0.y, p.v1.y), p.v2.y), gmax(gmax(p.v0.z, p.v1.z), p.v2.z) };
                all.push_back(p);
            }
        } else {
            Prim p; std::memset(&p, 0, sizeof(p));
            p.type = -1;
            p.id = int32_t(all.size());
            all.push_back(p);
        }
    }
    const wpt_medium& m = d->media[medium];
    for (uint32_t k = 0; k < m.boundary_count; k++) {
        const uint32_t id = d->boundary_hitables[m.first_boundary + k];
        if (id >= all.size() || all[id].type != 0)
            return WPT_E_UNSUPPORTED;
        t.prims.push_back(all[id]);
    }
    t.finish();
    return 0;
}

/* The two boundary queries of HitableMedium::hit (hitable_medium.hpp:75-79) for caller-supplied rays, computed twice:
 *   nested[2i], nested[2i+1]   as the reference does: two BVH::hit calls on the boundary's tree, (-maxval, maxval) and
 *                              (a1 + minHitDistance, maxval);
 *   replay[2i], replay[2i+1]   as the device does for a small triangle boundary (mediumRoundCoop / mediumTest in
 *                              wurblpt_b200/csrc/device_trace.cuh): every boundary triangle once on the whole line, then
 *                              both queries replayed over the triangles that were hit, in the tree's storage order, with
 *                              the comparisons of the bounded test against the shrinking upper bound.
 * A query without a hit is reported as +infinity. corners (optional): the boundary triangles in storage order, 9 floats each.
 * Returns the number of boundary triangles, or a negative error. */
long long port_medium_boundary(const wpt_scene_desc* desc, uint32_t medium, const float* origins, const float* directions,
        uint64_t n, float minHit, float* nested, float* replay, float* corners)
{
    Tree tree;
    int rc = buildBoundaryTree(desc, medium, tree);
    if (rc != 0)
        return rc;
    std::vector<int32_t> order;     /* boundary triangles in the order the tree stores (= its walk meets) them */
    for (const Node& node : tree.nodes)
        if (node.prim >= 0)
            order.push_back(node.prim);
    if (corners)
        for (size_t k = 0; k < order.size(); k++) {
            const Prim& p = tree.prims[order[k]];
            const float c[9] = { p.v0.x, p.v0.y, p.v0.z, p.v1.x, p.v1.y, p.v1.z, p.v2.x, p.v2.y, p.v2.z };
            std::memcpy(corners + 9 * k, c, sizeof(c));
        }
    const float inf = std::numeric_limits<float>::infinity();
    const float maxval = std::numeric_limits<float>::max();
    #pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; i++) {
        Ray r = { { origins[3 * i], origins[3 * i + 1], origins[3 * i + 2] }, { directions[3 * i], directions[3 * i + 1], directions[3 * i + 2] } };
        /* the reference's way */
        int32_t id;
        float a1 = inf, a2 = inf, dist = 0.0f;
        if (tree.hit(r, -maxval, maxval, id, dist)) {
            a1 = dist;
            if (tree.hit(r, a1 + minHit, maxval, id, dist))
                a2 = dist;
        }
        nested[2 * i] = a1;
        nested[2 * i + 1] = a2;
        /* the device's way */
        const Helper h(r);
        std::vector<float> Ts(order.size()), ds(order.size()), as(order.size());
        std::vector<char> found(order.size());
        for (size_t k = 0; k < order.size(); k++)
            found[k] = triangleOnLine(tree.prims[order[k]], r, h, Ts[k], ds[k], as[k]);
        float b1 = maxval, b2 = maxval;
        bool have1 = false, have2 = false;
        for (size_t k = 0; k < order.size(); k++)
            if (found[k] && !(Ts[k] < -maxval * ds[k] || Ts[k] > b1 * ds[k])) {
                b1 = as[k];
                have1 = true;
            }
        if (have1) {
            const float lo = b1 + minHit;
            for (size_t k = 0; k < order.size(); k++)
                if (found[k] && !(Ts[k] < lo * ds[k] || Ts[k] > b2 * ds[k])) {
                    b2 = as[k];
                    have2 = true;
                }
        }
        replay[2 * i] = have1 ? b1 : inf;
        replay[2 * i + 1] = (have1 && have2) ? b2 : inf;
    }
    return (long long)order.size();
}

/* node count of the restated tree and, optionally, its arrays in the layout of wpt_scene_get_binary_nodes() */
long long port_bvh_nodes(const wpt_scene_desc* desc, uint64_t capacity, float* boxes, uint32_t* link, int32_t* leaf)
{
    Tree tree;
    if (buildTree(desc, tree) != 0)
        return -1;
    if (!boxes)
        return (long long)tree.nodes.size();
    if (capacity < tree.nodes.size())
        return -1;
    for (size_t i = 0; i < tree.nodes.size(); i++) {
        const Node& n = tree.nodes[i];
        boxes[i * 6 + 0] = n.box.lo.x; boxes[i * 6 + 1] = n.box.lo.y; boxes[i * 6 + 2] = n.box.lo.z;
        boxes[i * 6 + 3] = n.box.hi.x; boxes[i * 6 + 4] = n.box.hi.y; boxes[i * 6 + 5] = n.box.hi.z;
        link[i] = n.prim < 0 ? n.link : 0;
        leaf[i] = n.prim < 0 ? -1 : tree.prims[n.prim].id;
    }
    return (long long)tree.nodes.size();
}

}
