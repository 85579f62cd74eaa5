/*
 * scene_host.cpp -- scene ingest and the host side of Scene::updateBVH().
 *
 * Compiled with -ffp-contract=off: bounds and the SAH sweep must round exactly like the
 * reference's strict build so that both produce the same tree (same std::sort, same
 * comparator, same input order => same permutation, ties included).
 */
#include "scene_host.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <limits>

#include <omp.h>

namespace wpt {

/* ------------------------------------------------------------------ ingest */

namespace {

template<typename T>
void copyArray(std::vector<T>& dst, const T* src, uint64_t n)
{
    dst.clear();
    if (n > 0 && src)
        dst.assign(src, src + n);
}

bool texOk(int32_t t, size_t n) { return t >= -1 && (t < 0 || size_t(t) < n); }

}

int HostScene::ingest(const wpt_scene_desc* d)
{
    if (!d) {
        error = "null scene description";
        return WPT_E_INVALID;
    }
    if (d->abi_version != WPT_ABI_VERSION) {
        error = "scene description has ABI version " + std::to_string(d->abi_version)
            + ", library expects " + std::to_string(WPT_ABI_VERSION);
        return WPT_E_INVALID;
    }
    copyArray(animations, d->animations, d->n_animations);
    copyArray(keyframes, d->keyframes, d->n_keyframes);
    copyArray(textures, d->textures, d->n_textures);
    copyArray(texels, d->texels, d->texel_bytes);
    copyArray(materials, d->materials, d->n_materials);
    copyArray(meshes, d->meshes, d->n_meshes);
    copyArray(vertices, d->vertices, d->n_vertex_floats);
    copyArray(indices, d->indices, d->n_indices);
    copyArray(instances, d->instances, d->n_instances);
    copyArray(spheres, d->spheres, d->n_spheres);
    copyArray(media, d->media, d->n_media);
    copyArray(boundary, d->boundary_hitables, d->n_boundary);
    copyArray(components, d->components, d->n_components);
    env = d->envmap;
    if (env.type != WPT_ENV_NONE && env.is_n > 0) {
        uint64_t n2 = uint64_t(env.is_n) * env.is_n;
        if (!env.M || !env.Ms || !env.Mcs) {
            error = "environment map importance tables missing";
            return WPT_E_INVALID;
        }
        copyArray(envM, env.M, n2);
        copyArray(envMs, env.Ms, n2);
        copyArray(envMcs, env.Mcs, n2);
        for (int32_t b : envMs)
            if (b < 0 || uint64_t(b) >= n2) {
                error = "environment map bin index out of range";
                return WPT_E_INVALID;
            }
    } else {
        env.is_n = 0;
    }
    env.M = nullptr; env.Ms = nullptr; env.Mcs = nullptr;

    /* validation: every index must point inside its array */
    for (const wpt_animation& a : animations)
        if (uint64_t(a.first_keyframe) + a.keyframe_count > keyframes.size()) {
            error = "animation key frame range out of bounds";
            return WPT_E_INVALID;
        }
    for (size_t i = 0; i < textures.size(); i++) {
        const wpt_texture& t = textures[i];
        if (t.type == WPT_TEX_IMAGE) {
            size_t comp = (t.texel_type == WPT_TEXEL_U8 ? 1 : t.texel_type == WPT_TEXEL_U16 ? 2 : 4);
            if (t.width <= 0 || t.height <= 0 || t.components < 1 || t.components > 4
                    || t.texel_type < WPT_TEXEL_U8 || t.texel_type > WPT_TEXEL_F32
                    || t.texel_offset + uint64_t(t.width) * t.height * t.components * comp > texels.size()) {
                error = "image texture " + std::to_string(i) + " is malformed";
                return WPT_E_INVALID;
            }
        } else if (t.type == WPT_TEX_TRANSFORMER) {
            /* the wrapped texture must come earlier, which also rules out cycles */
            if (t.src_texture < 0 || size_t(t.src_texture) >= i) {
                error = "texture transformer " + std::to_string(i) + " wraps an invalid texture";
                return WPT_E_INVALID;
            }
        } else if (t.type != WPT_TEX_CONSTANT && t.type != WPT_TEX_CHECKER) {
            error = "texture " + std::to_string(i) + " has unknown type " + std::to_string(t.type);
            return WPT_E_UNSUPPORTED;
        }
    }
    for (size_t i = 0; i < materials.size(); i++) {
        const wpt_material& m = materials[i];
        if (m.type < WPT_MAT_NONE || m.type > WPT_MAT_TWOSIDED) {
            error = "material " + std::to_string(i) + " has unknown type " + std::to_string(m.type);
            return WPT_E_UNSUPPORTED;
        }
        bool ok = texOk(m.normal_tex, textures.size());
        for (int k = 0; k < 5; k++)
            ok = ok && texOk(m.tex[k], textures.size());
        if (m.type == WPT_MAT_TWOSIDED) {
            ok = ok && m.front >= -1 && m.back >= -1
                && (m.front < 0 || size_t(m.front) < materials.size())
                && (m.back < 0 || size_t(m.back) < materials.size());
            /* nesting two-sided materials would need recursion on the device */
            if (ok && ((m.front >= 0 && materials[m.front].type == WPT_MAT_TWOSIDED)
                        || (m.back >= 0 && materials[m.back].type == WPT_MAT_TWOSIDED))) {
                error = "nested two-sided materials are not supported";
                return WPT_E_UNSUPPORTED;
            }
        }
        if (!ok) {
            error = "material " + std::to_string(i) + " refers to an invalid texture or material";
            return WPT_E_INVALID;
        }
    }
    for (size_t i = 0; i < meshes.size(); i++) {
        const wpt_mesh& m = meshes[i];
        uint32_t expectStride = (m.flags & WPT_MESHF_TEXCOORDS) ? ((m.flags & WPT_MESHF_TANGENTS) ? 11 : 8) : 6;
        if (m.vertex_stride != expectStride
                || m.vertex_offset + uint64_t(m.vertex_count) * m.vertex_stride > vertices.size()
                || m.index_offset + uint64_t(m.triangle_count) * 3 > indices.size()) {
            error = "mesh " + std::to_string(i) + " is malformed";
            return WPT_E_INVALID;
        }
        for (uint64_t k = 0; k < uint64_t(m.triangle_count) * 3; k++)
            if (indices[m.index_offset + k] >= m.vertex_count) {
                error = "mesh " + std::to_string(i) + " has a vertex index out of range";
                return WPT_E_INVALID;
            }
    }
    auto matOk = [this](int32_t m) { return m >= -1 && (m < 0 || size_t(m) < materials.size()); };
    auto animOk = [this](int32_t a) { return a >= -1 && (a < 0 || size_t(a) < animations.size()); };
    for (size_t i = 0; i < instances.size(); i++) {
        const wpt_instance& in = instances[i];
        if (in.mesh < 0 || size_t(in.mesh) >= meshes.size() || !matOk(in.material) || !animOk(in.animation)) {
            error = "mesh instance " + std::to_string(i) + " is malformed";
            return WPT_E_INVALID;
        }
    }
    for (size_t i = 0; i < spheres.size(); i++)
        if (!matOk(spheres[i].material) || !animOk(spheres[i].animation)) {
            error = "sphere " + std::to_string(i) + " is malformed";
            return WPT_E_INVALID;
        }
    if (env.type != WPT_ENV_NONE) {
        int nt = (env.type == WPT_ENV_EQUIRECT ? 1 : env.type == WPT_ENV_CUBE ? 6 : -1);
        if (nt < 0) {
            error = "unknown environment map type";
            return WPT_E_UNSUPPORTED;
        }
        for (int k = 0; k < nt; k++)
            if (env.tex[k] < 0 || size_t(env.tex[k]) >= textures.size()) {
                error = "environment map texture missing";
                return WPT_E_INVALID;
            }
    }

    /* expand components into hitables */
    prims.clear();
    topLevel.clear();
    hotSpotIds.clear();
    for (size_t c = 0; c < components.size(); c++) {
        const wpt_component& comp = components[c];
        uint32_t count = 1;
        uint32_t type;
        if (comp.type == WPT_COMP_MESH_INSTANCE) {
            if (comp.ref < 0 || size_t(comp.ref) >= instances.size()) {
                error = "component refers to an invalid mesh instance";
                return WPT_E_INVALID;
            }
            type = PRIM_TRIANGLE;
            count = meshes[instances[comp.ref].mesh].triangle_count;
        } else if (comp.type == WPT_COMP_SPHERE) {
            if (comp.ref < 0 || size_t(comp.ref) >= spheres.size()) {
                error = "component refers to an invalid sphere";
                return WPT_E_INVALID;
            }
            type = PRIM_SPHERE;
        } else if (comp.type == WPT_COMP_MEDIUM) {
            if (comp.ref < 0 || size_t(comp.ref) >= media.size()) {
                error = "component refers to an invalid medium";
                return WPT_E_INVALID;
            }
            type = PRIM_MEDIUM;
        } else {
            error = "unknown component type";
            return WPT_E_UNSUPPORTED;
        }
        for (uint32_t k = 0; k < count; k++) {
            uint32_t id = prims.size();
            prims.push_back(Prim{ type, uint32_t(comp.ref), k, comp.flags });
            if (comp.flags & WPT_COMPF_HITABLE) {
                topLevel.push_back(id);
                if (comp.flags & WPT_COMPF_HOTSPOT)
                    hotSpotIds.push_back(id);
            }
        }
    }
    if (prims.size() >= REF_SLOT_MASK) {
        error = "too many hitables";
        return WPT_E_UNSUPPORTED;
    }
    for (size_t i = 0; i < media.size(); i++) {
        const wpt_medium& m = media[i];
        if (uint64_t(m.first_boundary) + m.boundary_count > boundary.size() || !matOk(m.material)) {
            error = "medium " + std::to_string(i) + " is malformed";
            return WPT_E_INVALID;
        }
        for (uint32_t k = 0; k < m.boundary_count; k++) {
            uint32_t id = boundary[m.first_boundary + k];
            if (id >= prims.size() || prims[id].type == PRIM_MEDIUM) {
                error = "medium " + std::to_string(i) + " has an invalid boundary hitable";
                return id >= prims.size() ? WPT_E_INVALID : WPT_E_UNSUPPORTED;
            }
        }
    }
    haveBvh = false;
    return WPT_OK;
}

/* -------------------------------------------------------------- animations */

static Trs toTrs(const wpt_keyframe& k)
{
    return Trs{ v3(k.translation[0], k.translation[1], k.translation[2]),
        Quat{ k.rotation[0], k.rotation[1], k.rotation[2], k.rotation[3] },
        v3(k.scaling[0], k.scaling[1], k.scaling[2]) };
}

/* AnimationKeyframes::at (animation_keyframes.hpp:191-215) with its binary search (:82-103) */
Trs keyframesAt(const wpt_keyframe* kf, uint32_t count, float t)
{
    if (count == 0)
        return Trs{ v3(0.0f, 0.0f, 0.0f), Quat{ 0.0f, 0.0f, 0.0f, 1.0f }, v3(1.0f, 1.0f, 1.0f) };
    if (t <= kf[0].t)
        return toTrs(kf[0]);
    if (t >= kf[count - 1].t)
        return toTrs(kf[count - 1]);
    int a = 0, b = int(count) - 1;
    int lower = -1, higher = -1;
    while (b >= a) {
        int c = (a + b) / 2;
        if (kf[c].t < t) {
            a = c + 1;
        } else if (kf[c].t > t) {
            b = c - 1;
        } else {
            lower = higher = c;
            break;
        }
    }
    if (lower < 0) {
        lower = b;
        higher = a;
    }
    if (lower == higher)
        return toTrs(kf[lower]);
    float alpha = 1.0f - (kf[higher].t - t) / (kf[higher].t - kf[lower].t);
    return mix(toTrs(kf[lower]), toTrs(kf[higher]), alpha);
}

Trs HostScene::animationAt(int index, float t) const
{
    const wpt_animation& a = animations[index];
    return keyframesAt(keyframes.data() + a.first_keyframe, a.keyframe_count, t);
}

/* ------------------------------------------------------------------ bounds */

namespace {

struct BoundsContext {
    const HostScene& s;
    std::vector<Trs> T0, T1;        /* AnimationCache at t0 / t1 (animation.hpp:52-125) */
    std::vector<Mat34> M0;

    BoundsContext(const HostScene& scene, float t0, float t1) : s(scene)
    {
        for (size_t i = 0; i < s.animations.size(); i++) {
            T0.push_back(s.animationAt(i, t0));
            T1.push_back(s.animationAt(i, t1));
            M0.push_back(toMat34(T0.back()));
        }
    }

    V3 vertex(const wpt_mesh& m, uint32_t vi) const
    {
        const float* p = s.vertices.data() + m.vertex_offset + uint64_t(vi) * m.vertex_stride;
        return v3(p[0], p[1], p[2]);
    }

    /* world-space, un-animated corner positions of a triangle (hitable_triangle.hpp:194-206) */
    void triangleCorners(const Prim& p, V3 out[3]) const
    {
        const wpt_instance& in = s.instances[p.ref];
        const wpt_mesh& m = s.meshes[in.mesh];
        const uint32_t* idx = s.indices.data() + m.index_offset + uint64_t(p.tri) * 3;
        for (int k = 0; k < 3; k++) {
            V3 v = vertex(m, idx[k]);
            if (in.has_transform)
                v = transformPoint(in.M, 4, v);
            out[k] = v;
        }
    }

    /* HitableTriangle::aabb (hitable_triangle.hpp:336-395) */
    Box triangleBox(const Prim& p) const
    {
        V3 v[3];
        triangleCorners(p, v);
        const wpt_instance& in = s.instances[p.ref];
        if (in.animation < 0)
            return Box{ vmin3(v[0], v[1], v[2]), vmax3(v[0], v[1], v[2]) };
        int ai = in.animation;
        const Mat34& M0m = M0[ai];
        V3 a0 = transformPoint(M0m.m, 3, v[0]);
        V3 a1 = transformPoint(M0m.m, 3, v[1]);
        V3 a2 = transformPoint(M0m.m, 3, v[2]);
        Box box0{ vmin3(a0, a1, a2), vmax3(a0, a1, a2) };
        if (T0[ai] == T1[ai])
            return box0;
        /* The reference takes the matrix for t1 from the t0 cache as well
         * (hitable_triangle.hpp:362); reproduce that, do not "fix" it. */
        Box box = merge(box0, box0);
        if (!(T0[ai].r == T1[ai].r)) {
            unsigned int samples = 4;
            const Quat& q0 = T0[ai].r;
            const Quat& q1 = T1[ai].r;
            float cosHalfAngle = (((0.0f + q0.x * q1.x) + q0.y * q1.y) + q0.z * q1.z) + q0.w * q1.w;
            if (std::fabs(cosHalfAngle) < 1.0f)
                samples += (std::acos(cosHalfAngle) * (180.0f / 3.14159274f) /* degrees(), gvm.hpp:112-115 */) * 2.0f;
            for (unsigned int i = 1; i < samples - 1; i++) {
                float alpha = i / (samples - 1.0f);
                Trs T = mix(T0[ai], T1[ai], alpha);
                V3 b0 = apply(T, v[0]);
                V3 b1 = apply(T, v[1]);
                V3 b2 = apply(T, v[2]);
                box = merge(box, Box{ vmin3(b0, b1, b2), vmax3(b0, b1, b2) });
            }
        }
        return box;
    }

    /* HitableSphere::aabb (hitable_sphere.hpp:97-110) */
    Box sphereBox(const Prim& p) const
    {
        const wpt_sphere& sp = s.spheres[p.ref];
        V3 c = v3(sp.center[0], sp.center[1], sp.center[2]);
        float r = sp.radius;
        if (sp.animation < 0)
            return Box{ c - v3(r, r, r), c + v3(r, r, r) };
        const Trs& A = T0[sp.animation];
        const Trs& B = T1[sp.animation];
        V3 c0 = apply(A, c);
        V3 c1 = apply(B, c);
        float r0 = fmaxr(fmaxr(A.s.x, A.s.y), A.s.z) * r;
        float r1 = fmaxr(fmaxr(B.s.x, B.s.y), B.s.z) * r;
        return merge(Box{ c0 - v3(r0, r0, r0), c0 + v3(r0, r0, r0) },
                Box{ c1 - v3(r1, r1, r1), c1 + v3(r1, r1, r1) });
    }
};

/* ------------------------------------------------------------- SAH builder */

/* BVHNode::build (bvh.hpp:93-164) writing straight into the linear layout of BVH::flatten
 * (bvh.hpp:233-246): a subtree over n hitables occupies 2n-1 consecutive nodes, child 1 follows
 * its parent, child 2 follows child 1's subtree. */
struct TreeBuilder {
    const std::vector<Box>& boxes;      /* per local primitive */
    const std::vector<uint32_t>& leafIds;
    std::vector<V3> centers;
    std::vector<uint32_t> subset;
    std::vector<float> areas0, areas1;
    BNode* out;
    uint32_t base;
    uint32_t maxDepth = 0;

    TreeBuilder(const std::vector<Box>& b, const std::vector<uint32_t>& ids, BNode* o, uint32_t nodeBase)
        : boxes(b), leafIds(ids), out(o), base(nodeBase)
    {
        size_t n = boxes.size();
        centers.resize(n);
        subset.resize(n);
        areas0.resize(n);
        areas1.resize(n);
        for (size_t i = 0; i < n; i++) {
            centers[i] = center(boxes[i]);
            subset[i] = i;
        }
    }

    static void setBox(BNode& node, const Box& b)
    {
        node.lo[0] = b.lo.x; node.lo[1] = b.lo.y; node.lo[2] = b.lo.z;
        node.hi[0] = b.hi.x; node.hi[1] = b.hi.y; node.hi[2] = b.hi.z;
    }

    void build(size_t offset, size_t size, uint32_t nodeIndex, uint32_t level, uint32_t* depthOut)
    {
        BNode& node = out[nodeIndex];
        if (size == 1) {
            setBox(node, boxes[subset[offset]]);
            node.link = 0;
            node.leaf = leafIds[subset[offset]];
            *depthOut = level;
            return;
        }
        int n = size;
        Box box = boxes[subset[offset]];
        for (int i = 1; i < n; i++)
            box = merge(box, boxes[subset[offset + i]]);
        const int axis = longestAxis(box);
        const std::vector<V3>& c = centers;
        std::sort(subset.begin() + offset, subset.begin() + offset + size,
                [&c, axis](unsigned int i, unsigned int j) { return c[i][axis] < c[j][axis]; });
        Box box0 = boxes[subset[offset]];
        areas0[offset] = surfaceArea(box0);
        for (int i = 1; i < n - 1; i++) {
            box0 = merge(box0, boxes[subset[offset + i]]);
            areas0[offset + i] = surfaceArea(box0);
        }
        Box box1 = boxes[subset[offset + n - 1]];
        areas1[offset + n - 1] = surfaceArea(box1);
        for (int i = n - 2; i > 0; i--) {
            box1 = merge(box1, boxes[subset[offset + i]]);
            areas1[offset + i] = surfaceArea(box1);
        }
        /* sah(i) = i * areas0[i-1] + (n - i) * areas1[i] (bvh.hpp:73-77): int and size_t factors
         * convert to float before the multiplication */
        auto sah = [this, offset, size](int i) {
            return float(i) * areas0[offset + i - 1] + float(size - size_t(i)) * areas1[offset + i];
        };
        int split = 1;
        float best = sah(1);
        for (int i = 2; i < n; i++) {
            float v = sah(i);
            if (v < best) {
                best = v;
                split = i;
            }
        }
        setBox(node, box);
        const size_t n0 = split, n1 = size - n0;
        const uint32_t child0 = nodeIndex + 1;
        const uint32_t child1 = nodeIndex + uint32_t(2 * n0);
        node.link = child1;
        node.leaf = INNER;
        uint32_t d0 = 0, d1 = 0;
        constexpr size_t parallelThreshold = 16384;     /* bvh.hpp:149 */
        if (n0 >= parallelThreshold && n1 >= parallelThreshold) {
            #pragma omp task shared(d0) firstprivate(offset, n0, child0, level)
            build(offset, n0, child0, level + 1, &d0);
            #pragma omp task shared(d1) firstprivate(offset, n0, n1, child1, level)
            build(offset + n0, n1, child1, level + 1, &d1);
            #pragma omp taskwait
        } else {
            build(offset, n0, child0, level + 1, &d0);
            build(offset + n0, n1, child1, level + 1, &d1);
        }
        *depthOut = d0 > d1 ? d0 : d1;
    }

    uint32_t run()
    {
        uint32_t depth = 0;
        if (boxes.empty())
            return 0;
        #pragma omp parallel
        #pragma omp single
        build(0, boxes.size(), base, 1, &depth);
        return depth;
    }
};

/* --------------------------------------------------------- 4-wide collapse */

struct Collapser {
    const std::vector<BNode>& b;
    std::vector<Node4>& out;
    uint32_t maxRun = 1;            /* triangles per leaf of the wide tree (1 = like the reference) */
    uint32_t maxDepth = 0;

    static float area(const BNode& n)
    {
        float lx = n.hi[0] - n.lo[0], ly = n.hi[1] - n.lo[1], lz = n.hi[2] - n.lo[2];
        return lx * ly + ly * lz + lx * lz;
    }

    /* Returns the encoded child for binary node `bi`. */
    /* Number of leaves below binary node `bi` if they are all triangles and at most `limit`, else 0.
     * `first` receives the lowest leaf reference (= lowest slot: slots follow storage order). */
    uint32_t triangleRun(uint32_t bi, uint32_t limit, uint32_t& first) const
    {
        const BNode& n = b[bi];
        if (n.leaf != INNER) {
            if ((n.leaf >> REF_TYPE_SHIFT) != PRIM_TRIANGLE)
                return 0;
            first = n.leaf;
            return 1;
        }
        uint32_t f0 = 0, f1 = 0;
        const uint32_t c0 = triangleRun(bi + 1, limit, f0);
        if (c0 == 0 || c0 >= limit)
            return 0;
        const uint32_t c1 = triangleRun(n.link, limit - c0, f1);
        if (c1 == 0 || c0 + c1 > limit || f1 != f0 + c0)
            return 0;
        first = f0;
        return c0 + c1;
    }

    int32_t collapse(uint32_t bi, uint32_t depth)
    {
        const BNode& n = b[bi];
        if (n.leaf != INNER)
            return ~int32_t(n.leaf);
        if (maxRun > 1) {
            /* a small all-triangle subtree becomes one leaf holding a run of consecutive triangle slots */
            uint32_t first = 0;
            const uint32_t run = triangleRun(bi, maxRun, first);
            if (run >= 2)
                return ~int32_t((PRIM_TRIANGLE_RUN << REF_TYPE_SHIFT) | ((run - 1) << RUN_COUNT_SHIFT) | (first & RUN_SLOT_MASK));
        }
        if (depth > maxDepth)
            maxDepth = depth;
        uint32_t kids[4] = { bi + 1, n.link, 0, 0 };
        int count = 2;
        /* open the inner child with the largest surface area until four children are held */
        while (count < 4) {
            int pick = -1;
            float bestArea = -1.0f;
            for (int k = 0; k < count; k++) {
                if (b[kids[k]].leaf == INNER) {
                    float a = area(b[kids[k]]);
                    if (a > bestArea) {
                        bestArea = a;
                        pick = k;
                    }
                }
            }
            if (pick < 0)
                break;
            uint32_t open = kids[pick];
            /* keep the reference's child order: child 1 stays in place, child 2 goes right after it */
            for (int k = count; k > pick + 1; k--)
                kids[k] = kids[k - 1];
            kids[pick] = open + 1;
            kids[pick + 1] = b[open].link;
            count++;
        }
        uint32_t myIndex = out.size();
        out.emplace_back();
        const float far = std::numeric_limits<float>::max();
        for (int k = 0; k < 4; k++) {
            Node4& w = out[myIndex];
            if (k < count) {
                const BNode& c = b[kids[k]];
                w.lox[k] = c.lo[0]; w.hix[k] = c.hi[0];
                w.loy[k] = c.lo[1]; w.hiy[k] = c.hi[1];
                w.loz[k] = c.lo[2]; w.hiz[k] = c.hi[2];
            } else {
                w.lox[k] = w.hix[k] = w.loy[k] = w.hiy[k] = w.loz[k] = w.hiz[k] = far;
                w.child[k] = EMPTY_CHILD;
            }
            w.pad[k] = 0;
        }
        for (int k = 0; k < count; k++) {
            int32_t code = collapse(kids[k], depth + 1);
            out[myIndex].child[k] = code;   /* `out` may have been reallocated: index, don't hold a reference */
        }
        return int32_t(myIndex);
    }
};

}

/* Shading queue of a material (device_scene.cuh MC_*): the scattering types with a kernel of their own; a
 * two-sided material belongs to its sides' class if they agree (a null side scatters nothing and fits any). */
uint8_t HostScene::materialClass(int32_t material) const
{
    auto direct = [this](int32_t m) -> int {
        if (m < 0 || size_t(m) >= materials.size())
            return -1;      /* null: fits any class */
        switch (materials[m].type) {
        case WPT_MAT_LAMBERTIAN: return 1;
        case WPT_MAT_GGX: return 2;
        case WPT_MAT_GLASS: return 3;
        case WPT_MAT_MODPHONG: return 4;
        default: return 0;
        }
    };
    if (material < 0 || size_t(material) >= materials.size())
        return 0;
    if (materials[material].type != WPT_MAT_TWOSIDED)
        return uint8_t(direct(material));
    const int f = direct(materials[material].front), b = direct(materials[material].back);
    if (f < 0 && b < 0)
        return 0;
    if (f < 0)
        return uint8_t(b);
    if (b < 0 || b == f)
        return uint8_t(f);
    return 0;
}

/* --------------------------------------------------------------- updateBvh */

int HostScene::updateBvh(float t0, float t1)
{
    auto start = std::chrono::steady_clock::now();
    BoundsContext ctx(*this, t0, t1);

    /* bounds of every triangle and sphere (media need their nested tree first) */
    std::vector<Box> primBox(prims.size());
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < prims.size(); i++) {
        if (prims[i].type == PRIM_TRIANGLE)
            primBox[i] = ctx.triangleBox(prims[i]);
        else if (prims[i].type == PRIM_SPHERE)
            primBox[i] = ctx.sphereBox(prims[i]);
    }

    /* node budget: top-level tree + one nested tree per medium */
    uint64_t totalNodes = topLevel.empty() ? 1 : 2 * uint64_t(topLevel.size()) - 1;
    topLevelNodes = totalNodes;
    std::vector<uint32_t> nestedBase(media.size(), NO_NODE);
    for (size_t m = 0; m < media.size(); m++) {
        if (media[m].boundary_count > 0) {
            nestedBase[m] = totalNodes;
            totalNodes += 2 * uint64_t(media[m].boundary_count) - 1;
        }
    }
    if (totalNodes >= 0xFFFFFFF0ull) {
        error = "scene too large for 32-bit node indices";
        return WPT_E_UNSUPPORTED;
    }
    bnodes.assign(totalNodes, BNode{ { 0, 0, 0 }, { 0, 0, 0 }, 0, INNER });
    mediumRoot = nestedBase;

    /* nested boundary trees (HitableMedium::updateBVH, hitable_medium.hpp:59-62) and medium
     * bounds (:64-67: the nested tree's root box; an empty tree has the box (0,0,0)-(0,0,0),
     * bvh.hpp:186-189) */
    uint32_t depth = 0;
    for (size_t m = 0; m < media.size(); m++) {
        Box mb{ v3(0.0f, 0.0f, 0.0f), v3(0.0f, 0.0f, 0.0f) };
        if (media[m].boundary_count > 0) {
            std::vector<Box> boxes;
            std::vector<uint32_t> ids;
            for (uint32_t k = 0; k < media[m].boundary_count; k++) {
                uint32_t id = boundary[media[m].first_boundary + k];
                boxes.push_back(primBox[id]);
                ids.push_back(id);
            }
            TreeBuilder tb(boxes, ids, bnodes.data(), nestedBase[m]);
            uint32_t d = tb.run();
            if (d > depth)
                depth = d;
            const BNode& root = bnodes[nestedBase[m]];
            mb = Box{ v3(root.lo[0], root.lo[1], root.lo[2]), v3(root.hi[0], root.hi[1], root.hi[2]) };
        }
        for (size_t i = 0; i < prims.size(); i++)
            if (prims[i].type == PRIM_MEDIUM && prims[i].ref == m)
                primBox[i] = mb;
    }

    /* top-level tree (Scene::updateBVH -> BVH::build, scene.hpp:151-169, bvh.hpp:254-270) */
    if (topLevel.empty()) {
        /* BVHNode::buildBVH with no hitables: a single leaf with a null hitable (bvh.hpp:186-189).
         * Give it a box no ray can enter. */
        const float far = std::numeric_limits<float>::max();
        bnodes[0] = BNode{ { far, far, far }, { far, far, far }, 0, INNER };
        binaryDepth = 1;
    } else {
        std::vector<Box> boxes(topLevel.size());
        for (size_t i = 0; i < topLevel.size(); i++)
            boxes[i] = primBox[topLevel[i]];
        TreeBuilder tb(boxes, topLevel, bnodes.data(), 0);
        binaryDepth = tb.run();
    }
    if (binaryDepth > 128 || depth > 128) {      /* BVH::maxTreeDepth, bvh.hpp:230,260-263 */
        error = "Too many levels";
        return WPT_E_UNSUPPORTED;
    }

    /* Slots: triangles / spheres / media are numbered in the order the reference's depth-first
     * storage meets them, so "later in the reference's walk" == "higher slot". The device uses
     * that to settle exact distance ties the way the reference's visiting order does
     * (SURVEY.md A.3). Hitables that only bound media follow in id order. */
    refOfGlobal.assign(prims.size(), 0xFFFFFFFFu);
    triGlobal.clear(); sphGlobal.clear(); medGlobal.clear();
    auto assignSlot = [this](uint32_t id) {
        if (refOfGlobal[id] != 0xFFFFFFFFu)
            return;
        std::vector<uint32_t>& table = (prims[id].type == PRIM_TRIANGLE ? triGlobal
                : prims[id].type == PRIM_SPHERE ? sphGlobal : medGlobal);
        refOfGlobal[id] = (prims[id].type << REF_TYPE_SHIFT) | uint32_t(table.size());
        table.push_back(id);
    };
    if (!topLevel.empty())
        for (uint64_t i = 0; i < topLevelNodes; i++)
            if (bnodes[i].leaf != INNER)
                assignSlot(bnodes[i].leaf);
    /* the hitables that bound a medium, in the order the nested tree stores them: a medium whose boundary is a small
     * triangle mesh then owns a run of consecutive triangle slots in the order the nested walk meets them (medData) */
    for (size_t m = 0; m < media.size(); m++)
        if (nestedBase[m] != NO_NODE)
            for (uint64_t i = nestedBase[m]; i < uint64_t(nestedBase[m]) + 2 * uint64_t(media[m].boundary_count) - 1; i++)
                if (bnodes[i].leaf != INNER)
                    assignSlot(bnodes[i].leaf);
    for (uint32_t id = 0; id < prims.size(); id++)
        assignSlot(id);
    if (!topLevel.empty() || !media.empty()) {
        uint64_t first = topLevel.empty() ? topLevelNodes : 0;
        for (uint64_t i = first; i < bnodes.size(); i++)
            if (bnodes[i].leaf != INNER)
                bnodes[i].leaf = refOfGlobal[bnodes[i].leaf];
    }

    /* 4-wide tree over the top level */
    nodes4.clear();
    nodes4.reserve(topLevel.size() / 2 + 16);
    Collapser col{ bnodes, nodes4 };
    col.maxRun = (triGlobal.size() <= RUN_SLOT_MASK ? wideLeafTriangles : 1);
    root4 = topLevel.empty() ? EMPTY_CHILD : col.collapse(0, 1);
    wideDepth = col.maxDepth;

    /* primitive arrays in slot order */
    auto bits = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
    triData.resize(triGlobal.size() * 12);
    #pragma omp parallel for schedule(static)
    for (size_t s = 0; s < triGlobal.size(); s++) {
        const Prim& p = prims[triGlobal[s]];
        V3 v[3];
        ctx.triangleCorners(p, v);
        const wpt_instance& in = instances[p.ref];
        float* d = triData.data() + s * 12;
        d[0] = v[0].x; d[1] = v[0].y; d[2] = v[0].z; d[3] = bits(p.ref);
        d[4] = v[1].x; d[5] = v[1].y; d[6] = v[1].z; d[7] = bits(p.tri);
        d[8] = v[2].x; d[9] = v[2].y; d[10] = v[2].z; d[11] = bits(uint32_t(in.animation + 1));
    }
    /* what the hit record needs per triangle, resolved once: mesh -> index buffer -> vertex blocks */
    if (vertices.size() >= (size_t(1) << 32)) {
        error = "more than 2^32 vertex attribute floats";
        return WPT_E_UNSUPPORTED;
    }
    triShade.resize(triGlobal.size() * 4);
    #pragma omp parallel for schedule(static)
    for (size_t s = 0; s < triGlobal.size(); s++) {
        const Prim& p = prims[triGlobal[s]];
        const wpt_mesh& mesh = meshes[instances[p.ref].mesh];
        const uint32_t* ix = indices.data() + mesh.index_offset + size_t(p.tri) * 3;
        uint32_t* d = triShade.data() + s * 4;
        for (int k = 0; k < 3; k++)
            d[k] = uint32_t(mesh.vertex_offset + size_t(ix[k]) * mesh.vertex_stride);
        d[3] = mesh.flags;
    }
    sphData.resize(sphGlobal.size() * 12);
    for (size_t s = 0; s < sphGlobal.size(); s++) {
        const wpt_sphere& sp = spheres[prims[sphGlobal[s]].ref];
        float* d = sphData.data() + s * 12;
        d[0] = sp.center[0]; d[1] = sp.center[1]; d[2] = sp.center[2]; d[3] = sp.radius;
        d[4] = sp.rotation[0]; d[5] = sp.rotation[1]; d[6] = sp.rotation[2]; d[7] = sp.rotation[3];
        d[8] = bits(uint32_t(sp.material)); d[9] = bits(uint32_t(sp.animation + 1)); d[10] = 0.0f; d[11] = 0.0f;
    }
    medData.resize(medGlobal.size() * 4);
    for (size_t s = 0; s < medGlobal.size(); s++) {
        uint32_t mi = prims[medGlobal[s]].ref;
        float* d = medData.data() + s * 4;
        d[0] = media[mi].density;
        d[1] = bits(uint32_t(media[mi].material));
        d[2] = bits(mediumRoot[mi]);
        /* boundary = up to 32 triangles in consecutive slots (the nested tree's storage order, see above): first slot |
         * count << 26 for the trace kernel's warp-wide boundary test; 0 = walk the nested tree */
        uint32_t run = 0;
        const uint32_t count = media[mi].boundary_count;
        if (mediumRoot[mi] != NO_NODE && count >= 1 && count <= 32) {
            bool ok = true;
            uint32_t first = 0xFFFFFFFFu, next = 0;
            for (uint64_t i = mediumRoot[mi]; ok && i < uint64_t(mediumRoot[mi]) + 2 * uint64_t(count) - 1; i++) {
                if (bnodes[i].leaf == INNER)
                    continue;
                const uint32_t ref = bnodes[i].leaf;        /* already a slot reference */
                if ((ref >> REF_TYPE_SHIFT) != PRIM_TRIANGLE) {
                    ok = false;
                } else if (first == 0xFFFFFFFFu) {
                    first = ref & REF_SLOT_MASK;
                    next = first + 1;
                } else if ((ref & REF_SLOT_MASK) == next) {
                    next++;
                } else {
                    ok = false;
                }
            }
            if (ok && first != 0xFFFFFFFFu && next - first == count && first < (1u << 26))
                run = first | (count << 26);
        }
        d[3] = bits(run);
    }
    hotSpotRefs.clear();
    for (uint32_t id : hotSpotIds)
        hotSpotRefs.push_back(refOfGlobal[id]);

    /* material class of every slot: which shading queue a hit on it goes to */
    triClass.resize(triGlobal.size());
    sphClass.resize(sphGlobal.size());
    medClass.resize(medGlobal.size());
    std::vector<uint8_t> classOfMaterial(materials.size());
    for (size_t m = 0; m < materials.size(); m++)
        classOfMaterial[m] = materialClass(int32_t(m));
    auto cls = [&](int32_t material) { return material >= 0 && size_t(material) < classOfMaterial.size() ? classOfMaterial[material] : uint8_t(0); };
    #pragma omp parallel for schedule(static)
    for (size_t s = 0; s < triGlobal.size(); s++)
        triClass[s] = cls(instances[prims[triGlobal[s]].ref].material);
    for (size_t s = 0; s < sphGlobal.size(); s++)
        sphClass[s] = cls(spheres[prims[sphGlobal[s]].ref].material);
    for (size_t s = 0; s < medGlobal.size(); s++)
        medClass[s] = cls(media[prims[medGlobal[s]].ref].material);
    classMask = 0;
    for (uint8_t c : triClass) classMask |= 1u << c;
    for (uint8_t c : sphClass) classMask |= 1u << c;
    for (uint8_t c : medClass) classMask |= 1u << c;

    haveBvh = true;
    deviceBuilt = false;
    for (int k = 0; k < 3; k++) {
        sceneLo[k] = topLevel.empty() ? 0.0f : bnodes[0].lo[k];
        sceneHi[k] = topLevel.empty() ? 0.0f : bnodes[0].hi[k];
    }
    nTriSlots = triGlobal.size();
    nSphSlots = sphGlobal.size();
    nWideNodes = nodes4.size();
    bvhT0 = t0;
    bvhT1 = t1;
    buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    return WPT_OK;
}

/* ------------------------------------------------------ device builder, host part */

int HostScene::prepareDeviceBuild(float t0, float t1)
{
    if (!media.empty() || prims.size() < 2 || topLevel.size() != prims.size() || prims.size() >= (1u << RUN_COUNT_SHIFT))
        return WPT_E_UNSUPPORTED;
    BoundsContext ctx(*this, t0, t1);
    const size_t n = prims.size();
    /* index of every hitable among the hitables of its type */
    std::vector<uint32_t> typeIndex(n);
    uint32_t nTri = 0, nSph = 0;
    for (size_t i = 0; i < n; i++)
        typeIndex[i] = (prims[i].type == PRIM_TRIANGLE ? nTri++ : nSph++);
    if (vertices.size() >= (size_t(1) << 32)) {
        error = "more than 2^32 vertex attribute floats";
        return WPT_E_INVALID;
    }
    auto bits = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
    devBox.resize(n * 8);
    devTriPrim.resize(size_t(nTri) * 12);
    devTriShadePrim.resize(size_t(nTri) * 4);
    devTriClassPrim.resize(nTri);
    devSphPrim.resize(size_t(nSph) * 12);
    devSphClassPrim.resize(nSph);
    std::vector<uint8_t> classOfMaterial(materials.size());
    for (size_t m = 0; m < materials.size(); m++)
        classOfMaterial[m] = materialClass(int32_t(m));
    auto cls = [&](int32_t material) { return material >= 0 && size_t(material) < classOfMaterial.size() ? classOfMaterial[material] : uint8_t(0); };
    uint32_t mask = 0;
    #pragma omp parallel for schedule(static) reduction(|:mask)
    for (size_t i = 0; i < n; i++) {
        const Prim& p = prims[i];
        const uint32_t ti = typeIndex[i];
        Box b;
        if (p.type == PRIM_TRIANGLE) {
            b = ctx.triangleBox(p);
            V3 v[3];
            ctx.triangleCorners(p, v);
            const wpt_instance& in = instances[p.ref];
            float* d = devTriPrim.data() + size_t(ti) * 12;
            d[0] = v[0].x; d[1] = v[0].y; d[2] = v[0].z; d[3] = bits(p.ref);
            d[4] = v[1].x; d[5] = v[1].y; d[6] = v[1].z; d[7] = bits(p.tri);
            d[8] = v[2].x; d[9] = v[2].y; d[10] = v[2].z; d[11] = bits(uint32_t(in.animation + 1));
            const wpt_mesh& mesh = meshes[in.mesh];
            const uint32_t* ix = indices.data() + mesh.index_offset + size_t(p.tri) * 3;
            uint32_t* sh = devTriShadePrim.data() + size_t(ti) * 4;
            for (int k = 0; k < 3; k++)
                sh[k] = uint32_t(mesh.vertex_offset + size_t(ix[k]) * mesh.vertex_stride);
            sh[3] = mesh.flags;
            devTriClassPrim[ti] = cls(in.material);
            mask |= 1u << devTriClassPrim[ti];
        } else {
            b = ctx.sphereBox(p);
            const wpt_sphere& sp = spheres[p.ref];
            float* d = devSphPrim.data() + size_t(ti) * 12;
            d[0] = sp.center[0]; d[1] = sp.center[1]; d[2] = sp.center[2]; d[3] = sp.radius;
            d[4] = sp.rotation[0]; d[5] = sp.rotation[1]; d[6] = sp.rotation[2]; d[7] = sp.rotation[3];
            d[8] = bits(uint32_t(sp.material)); d[9] = bits(uint32_t(sp.animation + 1)); d[10] = 0.0f; d[11] = 0.0f;
            devSphClassPrim[ti] = cls(sp.material);
            mask |= 1u << devSphClassPrim[ti];
        }
        float* o = devBox.data() + i * 8;
        o[0] = b.lo.x; o[1] = b.lo.y; o[2] = b.lo.z; o[3] = bits(p.type);
        o[4] = b.hi.x; o[5] = b.hi.y; o[6] = b.hi.z; o[7] = bits(ti);
    }
    classMask = mask;
    nTriSlots = nTri;
    nSphSlots = nSph;
    for (int k = 0; k < 3; k++) {
        float lo = devBox[k], hi = devBox[4 + k];
        for (size_t i = 1; i < n; i++) {
            lo = std::min(lo, devBox[i * 8 + k]);
            hi = std::max(hi, devBox[i * 8 + 4 + k]);
        }
        sceneLo[k] = lo;
        sceneHi[k] = hi;
    }
    return WPT_OK;
}

void HostScene::finishDeviceBuild(float t0, float t1, uint32_t wideNodes, uint32_t depth, uint32_t triangles, double seconds)
{
    (void)triangles;
    haveBvh = true;
    deviceBuilt = true;
    bvhT0 = t0;
    bvhT1 = t1;
    root4 = 0;
    wideDepth = depth;
    nWideNodes = wideNodes;
    binaryDepth = 0;
    topLevelNodes = 0;
    bnodes.clear();
    nodes4.clear();
    buildSeconds = seconds;
}

}
