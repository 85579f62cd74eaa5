/*
 * device_trace.cuh -- closest-hit queries on the device: ray/box, ray/triangle, ray/sphere,
 * medium free-flight sampling, and the two tree walks.
 *
 *   traverseReference  the reference's walk (bvh.hpp:277-311): binary tree in storage order,
 *                      child 1 before child 2, no near/far ordering, last-accepted-wins.
 *                      Same tests in the same order => same answer as the reference, ties
 *                      included. Used for parity runs and for the nested medium boundaries.
 *   traverseWide       the production walk: 4-wide nodes (one 128-byte line, eight 128-bit
 *                      loads), children visited near to far, stack in shared memory with a
 *                      local-memory spill. Exact distance ties between triangles are settled
 *                      by slot number, which the host assigned in the reference's visiting
 *                      order, so the reference's winner is reproduced.
 *
 * Roofline bound of both: node/primitive fetch bandwidth (L1/L2 for small scenes, HBM when
 * the tree exceeds the 126 MB L2). Algorithmic bytes per query as defined by SURVEY.md 8(d):
 * 32 B per reference node visit + 68 B per reference leaf test.
 */
#pragma once

#include "device_math.cuh"
#include "device_scene.cuh"

namespace wptd {

constexpr uint32_t NO_HIT = 0xFFFFFFFFu;
constexpr uint32_t REF_TYPE_SHIFT = 29;
constexpr uint32_t REF_SLOT_MASK = (1u << REF_TYPE_SHIFT) - 1u;
constexpr uint32_t PRIM_TRIANGLE = 0, PRIM_SPHERE = 1, PRIM_MEDIUM = 2, PRIM_TRIANGLE_RUN = 3;
constexpr uint32_t RUN_COUNT_SHIFT = 27;
constexpr uint32_t RUN_SLOT_MASK = (1u << RUN_COUNT_SHIFT) - 1u;
constexpr int32_t EMPTY_CHILD = 0x7FFFFFFF;
constexpr uint32_t NO_NODE = 0xFFFFFFFFu;

constexpr int F_MEDIA = 1;      /* scene contains media (nested walks, PRNG draws inside the walk) */
constexpr int F_ANIM = 2;       /* scene contains animated hitables */
constexpr int F_OPL = 4;        /* track optical path length / refractive index (ToF sensor, gated RGB) */
constexpr int F_COUNT = 8;      /* count node visits and leaf tests (query kernels only) */

#ifndef WPT_SM_STACK
#define WPT_SM_STACK 12
#endif
#ifndef WPT_STACK_STRIDE
#define WPT_STACK_STRIDE 128    /* threads per block of the kernels that walk the tree (kernels.cu BLOCK) */
#endif
constexpr int SM_STACK = WPT_SM_STACK;          /* shared-memory stack entries per thread */
constexpr int LOCAL_STACK = 96 - SM_STACK;      /* spill entries per thread (local memory, touched only when deep) */
constexpr int BINARY_STACK = 128;   /* BVH::maxTreeDepth, bvh.hpp:230 */

struct Ray {
    float3 o, d;
    float time;
};

/* RayIntersectionHelper (hitable.hpp:66-113) */
struct RayHelper {
    float3 inv;
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

WPT_DEV RayHelper makeRayHelper(const Ray& ray)
{
    RayHelper h;
    h.inv = 1.0f / ray.d;
    float ax = fabsf(ray.d.x), ay = fabsf(ray.d.y), az = fabsf(ray.d.z);
    if (az >= ay && az >= ax)
        h.kz = 2;
    else if (ay >= ax)
        h.kz = 1;
    else
        h.kz = 0;
    h.kx = h.kz + 1;
    if (h.kx == 3)
        h.kx = 0;
    h.ky = h.kx + 1;
    if (h.ky == 3)
        h.ky = 0;
    if (sel(ray.d, h.kz) < 0.0f) {
        int tmp = h.kx;
        h.kx = h.ky;
        h.ky = tmp;
    }
    h.Sx = sel(ray.d, h.kx) * sel(h.inv, h.kz);
    h.Sy = sel(ray.d, h.ky) * sel(h.inv, h.kz);
    h.Sz = sel(h.inv, h.kz);
    return h;
}

struct Hit {
    float a;
    uint32_t ref;
};

/* Per-thread state that a walk may touch: the pixel's Prng (media draw inside the walk,
 * hitable_medium.hpp:89) and the animation cache (animation.hpp:52-125). */
template<int F>
struct Ctx {
    Prng prng;
    unsigned int nested;
    unsigned int nodes, leaves;
    /* AnimationCache (animation.hpp:52-125): filled lazily, entry by entry, for the time `animTime` */
    mutable float animTime;
    mutable uint32_t animValid;     /* bit e: cache entry e holds an animation at animTime */
    mutable uint32_t animShared;    /* which animation the last (shared) entry holds */
    mutable Trs animT[(F & F_ANIM) ? MAX_CACHED_ANIMATIONS : 1];
    mutable Mat34 animM[(F & F_ANIM) ? MAX_CACHED_ANIMATIONS : 1];
    mutable Mat33 animN[(F & F_ANIM) ? MAX_CACHED_ANIMATIONS : 1];
};

/* AnimationKeyframes::at (animation_keyframes.hpp:191-215) */
WPT_DEV Trs keyframeTrs(const wpt_keyframe& k)
{
    Trs T;
    T.t = f3(k.translation[0], k.translation[1], k.translation[2]);
    T.r = f4(k.rotation[0], k.rotation[1], k.rotation[2], k.rotation[3]);
    T.s = f3(k.scaling[0], k.scaling[1], k.scaling[2]);
    return T;
}

static __device__ __noinline__ Trs keyframesAt(const wpt_keyframe* kf, uint32_t count, float t)
{
    if (count == 0) {
        Trs T;
        T.t = f3(0.0f);
        T.r = f4(0.0f, 0.0f, 0.0f, 1.0f);
        T.s = f3(1.0f);
        return T;
    }
    if (t <= kf[0].t)
        return keyframeTrs(kf[0]);
    if (t >= kf[count - 1].t)
        return keyframeTrs(kf[count - 1]);
    int a = 0, b = int(count) - 1;
    int lower = -1, higher = -1;
    while (b >= a) {
        int c = (a + b) / 2;
        float tc = kf[c].t;
        if (tc < t) {
            a = c + 1;
        } else if (tc > t) {
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
        return keyframeTrs(kf[lower]);
    float alpha = 1.0f - (kf[higher].t - t) / (kf[higher].t - kf[lower].t);
    return trsMix(keyframeTrs(kf[lower]), keyframeTrs(kf[higher]), alpha);
}

/* AnimationCache::init(t) (animation.hpp:62-70): a new time invalidates the entries; they are evaluated on
 * first use (animation.hpp:96-101). */
template<int F>
WPT_DEV void animationCacheInit(const DScene& S, Ctx<F>& ctx, float t)
{
    if (F & F_ANIM) {
        ctx.animTime = t;
        ctx.animValid = 0u;
    }
}

/* Cache entry of animation i: the first MAX_CACHED_ANIMATIONS - 1 animations have an entry of their own, all the others
 * share the last one (it then holds whichever of them was used last: correct for any number of animations, slower when
 * several of the shared ones alternate). */
WPT_DEV uint32_t animationCacheEntry(uint32_t i)
{
    return i < uint32_t(MAX_CACHED_ANIMATIONS - 1) ? i : uint32_t(MAX_CACHED_ANIMATIONS - 1);
}

template<int F>
WPT_DEV uint32_t animationCacheEnsure(const DScene& S, const Ctx<F>& ctx, uint32_t i)
{
    const uint32_t e = animationCacheEntry(i);
    const bool shared = (e == uint32_t(MAX_CACHED_ANIMATIONS - 1));
    if (!(ctx.animValid & (1u << e)) || (shared && ctx.animShared != i)) {
        const wpt_animation an = S.animations[i];
        const Trs T = keyframesAt(S.keyframes + an.first_keyframe, an.keyframe_count, ctx.animTime);
        const Mat33 N = toMat33(T.r);
        ctx.animT[e] = T;
        ctx.animN[e] = N;
        ctx.animM[e] = toMat34(T, N);
        ctx.animValid |= (1u << e);
        if (shared)
            ctx.animShared = i;
    }
    return e;
}

/* Transformation of animation `anim` (1-based, as stored with the primitives) at the thread's current time:
 * from the launch's table when all rays share one time, else from the thread's cache. */
template<int F>
WPT_DEV Trs animTrs(const DScene& S, const Ctx<F>& ctx, uint32_t anim)
{
    if (S.animTable) {
        const float* e = S.animTable + size_t(anim - 1u) * 32u;
        Trs T;
        T.t = f3(e[0], e[1], e[2]);
        T.r = f4(e[3], e[4], e[5], e[6]);
        T.s = f3(e[7], e[8], e[9]);
        return T;
    }
    return ctx.animT[animationCacheEnsure<F>(S, ctx, anim - 1u)];
}
template<int F>
WPT_DEV const float* animMatrix(const DScene& S, const Ctx<F>& ctx, uint32_t anim)
{
    if (S.animTable)
        return S.animTable + size_t(anim - 1u) * 32u + 10u;
    return ctx.animM[animationCacheEnsure<F>(S, ctx, anim - 1u)].m;
}
template<int F>
WPT_DEV const float* animNormalMatrix(const DScene& S, const Ctx<F>& ctx, uint32_t anim)
{
    if (S.animTable)
        return S.animTable + size_t(anim - 1u) * 32u + 22u;
    return ctx.animN[animationCacheEnsure<F>(S, ctx, anim - 1u)].m;
}

/* ------------------------------------------------------------ ray / triangle */

/* Corners of the triangle in slot `slot`, animated to the current cache time. */
template<int F>
WPT_DEV void triangleCorners(const DScene& S, uint32_t slot, const Ctx<F>& ctx,
        float3& v0, float3& v1, float3& v2, uint32_t& instance, uint32_t& triIndex, uint32_t& anim)
{
    const float4* p = S.tri + size_t(slot) * 3;
    const float4 a = __ldg(p + 0), b = __ldg(p + 1), c = __ldg(p + 2);
    v0 = xyz(a); v1 = xyz(b); v2 = xyz(c);
    instance = __float_as_uint(a.w);
    triIndex = __float_as_uint(b.w);
    anim = __float_as_uint(c.w);
    if (F & F_ANIM) {
        if (anim != 0) {
            const float* M = animMatrix<F>(S, ctx, anim);
            v0 = transformPoint(M, 3, v0);
            v1 = transformPoint(M, 3, v1);
            v2 = transformPoint(M, 3, v2);
        }
    }
}

/* Woop/Benthin/Wald watertight test as the reference runs it (hitable_triangle.hpp:189-271).
 * Returns false on a miss; on a hit `a` is the distance, (U,V,W)/det the barycentrics. */
WPT_DEV bool triangleTest(float3 v0, float3 v1, float3 v2, const Ray& ray, const RayHelper& h,
        float amin, float amax, bool bounded, float& a, float& U, float& V, float& W, float& det,
        float& Tsigned, float& detAbs)
{
    const float3 A = v0 - ray.o;
    const float3 B = v1 - ray.o;
    const float3 C = v2 - ray.o;
    const float Akz = sel(A, h.kz), Bkz = sel(B, h.kz), Ckz = sel(C, h.kz);
    const float Ax = sel(A, h.kx) - h.Sx * Akz;
    const float Ay = sel(A, h.ky) - h.Sy * Akz;
    const float Bx = sel(B, h.kx) - h.Sx * Bkz;
    const float By = sel(B, h.ky) - h.Sy * Bkz;
    const float Cx = sel(C, h.kx) - h.Sx * Ckz;
    const float Cy = sel(C, h.ky) - h.Sy * Ckz;
    U = Cx * By - Cy * Bx;
    V = Ax * Cy - Ay * Cx;
    W = Bx * Ay - By * Ax;
    if (fabsf(U) < LDBL_EPS_F || fabsf(V) < LDBL_EPS_F || fabsf(W) < LDBL_EPS_F) {
        /* FP64 edge fallback (hitable_triangle.hpp:240-250); __dmul_rn/__dsub_rn keep the
         * products and the difference un-contracted like the strict host build */
        U = float(__dsub_rn(__dmul_rn(double(Cx), double(By)), __dmul_rn(double(Cy), double(Bx))));
        V = float(__dsub_rn(__dmul_rn(double(Ax), double(Cy)), __dmul_rn(double(Ay), double(Cx))));
        W = float(__dsub_rn(__dmul_rn(double(Bx), double(Ay)), __dmul_rn(double(By), double(Ax))));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
        return false;
    det = U + V + W;
    if (det == 0.0f)
        return false;
    const float Az = h.Sz * Akz;
    const float Bz = h.Sz * Bkz;
    const float Cz = h.Sz * Ckz;
    const float T = U * Az + V * Bz + W * Cz;
    const uint32_t detSign = __float_as_uint(det) & 0x80000000u;
    Tsigned = __uint_as_float(__float_as_uint(T) ^ detSign);
    detAbs = __uint_as_float(__float_as_uint(det) ^ detSign);
    if (bounded) {
        if (Tsigned < amin * detAbs || Tsigned > amax * detAbs)
            return false;
    }
    const float invDet = 1.0f / det;
    a = invDet * T;
    return true;
}

WPT_DEV bool triangleTest(float3 v0, float3 v1, float3 v2, const Ray& ray, const RayHelper& h,
        float amin, float amax, bool bounded, float& a, float& U, float& V, float& W, float& det)
{
    float Ts, da;
    return triangleTest(v0, v1, v2, ray, h, amin, amax, bounded, a, U, V, W, det, Ts, da);
}

/* ------------------------------------------------------------- ray / sphere */

struct SphereGeom {
    float3 center;
    float radius;
    float4 rotation;
    uint32_t material;
};

/* HitableSphere::hit prologue (hitable_sphere.hpp:114-126) */
template<int F>
WPT_DEV SphereGeom sphereGeom(const DScene& S, uint32_t slot, const Ctx<F>& ctx)
{
    const float4* p = S.sph + size_t(slot) * 3;
    const float4 a = __ldg(p + 0), b = __ldg(p + 1), c = __ldg(p + 2);
    SphereGeom g;
    g.center = xyz(a);
    g.radius = a.w;
    g.rotation = b;
    g.material = __float_as_uint(c.x);
    if (F & F_ANIM) {
        uint32_t anim = __float_as_uint(c.y);
        if (anim != 0) {
            const Trs T = animTrs<F>(S, ctx, anim);
            g.center = g.center + T.t;
            g.radius = g.radius * rmax(rmax(T.s.x, T.s.y), T.s.z);
            g.rotation = qmul(g.rotation, T.r);
        }
    }
    return g;
}

/* HitableSphere::hit (hitable_sphere.hpp:128-152): distance only */
WPT_DEV bool sphereTest(float3 center, float radius, const Ray& ray, float amin, float amax, float& a)
{
    const float3 oc = ray.o - center;
    const float ocd = dot(oc, ray.d);
    const float aq = -ocd;
    const float3 tmp = oc - ocd * ray.d;
    const float discriminant = radius * radius - dot(tmp, tmp);
    if (discriminant > 0.0f) {
        float a1, a2;
        if (aq < 0.0f) {
            a2 = aq - sqrtf(discriminant);
            a1 = 2.0f * aq - a2;
        } else {
            a1 = aq + sqrtf(discriminant);
            a2 = 2.0f * aq - a1;
        }
        if (a2 > amin && a2 < amax) {
            a = a2;
            return true;
        } else if (a1 > amin && a1 < amax) {
            a = a1;
            return true;
        }
    }
    return false;
}

/* ------------------------------------------------------------ ray / box */

/* AABB::mayHit (aabb.hpp:70-86) with gvm's select-based min/max and reductions */
WPT_DEV bool boxTestReference(float3 lo, float3 hi, const Ray& ray, float3 inv, float amin, float amax)
{
    const float3 t0 = (lo - ray.o) * inv;
    const float3 t1 = (hi - ray.o) * inv;
    const float3 mn = vmin(t0, t1);
    const float3 mx = vmax(t0, t1);
    float tmin = amin;
    if (mn.x > tmin) tmin = mn.x;
    if (mn.y > tmin) tmin = mn.y;
    if (mn.z > tmin) tmin = mn.z;
    float tmax = amax;
    if (mx.x < tmax) tmax = mx.x;
    if (mx.y < tmax) tmax = mx.y;
    if (mx.z < tmax) tmax = mx.z;
    return tmin <= tmax;
}

/* ------------------------------------------------- reference-order tree walk */

template<int F, bool TOP>
__device__ Hit traverseReference(const DScene& S, uint32_t root, const Ray& ray, const RayHelper& h,
        float amin, float amax, float minHit, Ctx<F>& ctx);

/* HitableMedium::hit (hitable_medium.hpp:69-98) */
template<int F>
WPT_DEV bool mediumTest(const DScene& S, uint32_t slot, const Ray& ray, const RayHelper& h,
        float amin, float amax, float minHit, Ctx<F>& ctx, float& a)
{
    const float4 m = __ldg(S.med + slot);
    const uint32_t root = __float_as_uint(m.z);
    if (root == NO_NODE)
        return false;
    float a1, a2;
    const uint32_t run = (F & F_ANIM) ? 0u : __float_as_uint(m.w);     /* (animated scenes: nested walks in every kernel) */
    if (run != 0u) {
        /* The boundary is a small triangle mesh in consecutive slots, in the nested tree's storage order (scene_host.cpp).
         * The persistent trace kernel tests such a boundary with a whole warp (mediumRoundCoop): every triangle once
         * without bounds, then the reference's two queries replayed over the hit triangles with the bounded test's own
         * comparisons. This is the same computation for one thread - same tests, same comparisons, same order - so that a
         * query gets the same answer bit for bit whichever kernel serves it (frames do not depend on which slots the drain
         * kernel finishes, nor on the number of GPUs). The box tests of the nested walk are left out: what they cull
         * fails the comparisons too (see mediumRoundCoop for the exception: rays exactly through a vertex or an edge). */
        const uint32_t first = run & ((1u << 26) - 1u), count = run >> 26;
        bool have1 = false, have2 = false;
        a1 = a2 = MAXVAL;
        ctx.nested++;
        for (uint32_t j = 0; j < count; j++) {
            float3 v0, v1, v2;
            uint32_t inst, tri, anim;
            triangleCorners<F>(S, first + j, ctx, v0, v1, v2, inst, tri, anim);
            float ta, U, V, W, det, Ts, da;
            if (triangleTest(v0, v1, v2, ray, h, 0.0f, 0.0f, false, ta, U, V, W, det, Ts, da)
                    && !(Ts < -MAXVAL * da || Ts > a1 * da)) {
                a1 = ta;
                have1 = true;
            }
        }
        if (!have1)
            return false;
        ctx.nested++;
        const float lo = a1 + minHit;
        for (uint32_t j = 0; j < count; j++) {
            float3 v0, v1, v2;
            uint32_t inst, tri, anim;
            triangleCorners<F>(S, first + j, ctx, v0, v1, v2, inst, tri, anim);
            float ta, U, V, W, det, Ts, da;
            if (triangleTest(v0, v1, v2, ray, h, 0.0f, 0.0f, false, ta, U, V, W, det, Ts, da)
                    && !(Ts < lo * da || Ts > a2 * da)) {
                a2 = ta;
                have2 = true;
            }
        }
        if (!have2)
            return false;
    } else {
        ctx.nested++;
        Hit hr1 = traverseReference<F, false>(S, root, ray, h, -MAXVAL, MAXVAL, minHit, ctx);
        if (hr1.ref == NO_HIT)
            return false;
        ctx.nested++;
        Hit hr2 = traverseReference<F, false>(S, root, ray, h, hr1.a + minHit, MAXVAL, minHit, ctx);
        if (hr2.ref == NO_HIT)
            return false;
        a1 = hr1.a;
        a2 = hr2.a;
    }
    if (a1 < amin)
        a1 = amin;
    if (a2 > amax)
        a2 = amax;
    if (a1 < a2) {
        if (a1 < 0.0f)
            a1 = 0.0f;
        const float inside = a2 - a1;
        const float freeFlight = -(1.0f / m.x) * logf(in01(ctx.prng));
        if (freeFlight < inside) {
            a = a1 + freeFlight;
            return true;
        }
    }
    return false;
}

/* One leaf of either walk. Returns true if the hitable reports a hit under (amin, amax). */
template<int F, bool TOP>
WPT_DEV bool leafTest(const DScene& S, uint32_t ref, const Ray& ray, const RayHelper& h,
        float amin, float amax, float minHit, Ctx<F>& ctx, float& a, float& Tsigned, float& detAbs)
{
    Tsigned = 0.0f;
    detAbs = 0.0f;
    const uint32_t type = ref >> REF_TYPE_SHIFT;
    const uint32_t slot = ref & REF_SLOT_MASK;
    if (F & F_COUNT)
        ctx.leaves++;
    if (type == PRIM_TRIANGLE) {
        float3 v0, v1, v2;
        uint32_t inst, tri, anim;
        triangleCorners<F>(S, slot, ctx, v0, v1, v2, inst, tri, anim);
        float U, V, W, det;
        return triangleTest(v0, v1, v2, ray, h, amin, amax, true, a, U, V, W, det, Tsigned, detAbs);
    } else if (type == PRIM_SPHERE) {
        SphereGeom g = sphereGeom<F>(S, slot, ctx);
        return sphereTest(g.center, g.radius, ray, amin, amax, a);
    } else if (TOP && (F & F_MEDIA)) {
        return mediumTest<F>(S, slot, ray, h, amin, amax, minHit, ctx, a);
    }
    return false;
}

/* BVH::hit (bvh.hpp:277-311). The 32-byte node is two 128-bit loads:
 *   (lo.x lo.y lo.z hi.x) (hi.y hi.z link leaf) */
template<int F, bool TOP>
__device__ Hit traverseReference(const DScene& S, uint32_t root, const Ray& ray, const RayHelper& h,
        float amin, float amax, float minHit, Ctx<F>& ctx)
{
    Hit best;
    best.a = amax;
    best.ref = NO_HIT;
    uint32_t stack[BINARY_STACK];
    int sp = 0;
    uint32_t node = root;
    for (;;) {
        const float4 n0 = __ldg(S.bnode + size_t(node) * 2);
        const float4 n1 = __ldg(S.bnode + size_t(node) * 2 + 1);
        if (F & F_COUNT)
            ctx.nodes++;
        if (boxTestReference(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), ray, h.inv, amin, best.a)) {
            const uint32_t leaf = __float_as_uint(n1.w);
            if (leaf != 0xFFFFFFFFu) {
                float a, Ts, da;
                if (leafTest<F, TOP>(S, leaf, ray, h, amin, best.a, minHit, ctx, a, Ts, da)) {
                    best.a = a;
                    best.ref = leaf;
                }
                if (sp == 0)
                    break;
                node = stack[--sp];
            } else {
                stack[sp++] = __float_as_uint(n1.z);
                node++;
            }
        } else {
            if (sp == 0)
                break;
            node = stack[--sp];
        }
    }
    return best;
}

/* ------------------------------------------------------------ 4-wide walk */

#ifndef WPT_SHARED_STACK_PTX
#define WPT_SHARED_STACK_PTX 1      /* ld.shared / st.shared instead of generic accesses (same frame, +1 % on C3) */
#endif

/* The wide walk's stack: SM_STACK entries per thread in shared memory (a column of the block's array, WPT_STACK_STRIDE
 * entries apart, addressed as shared memory: ld.shared / st.shared instead of the generic LD.E / ST.E a pointer gives),
 * deeper entries in `spill` (local memory). Base address, spill array and stack pointer are separate variables of the
 * walk: in a struct next to the dynamically indexed spill array the stack pointer lives in local memory (an STL / LDL
 * per push and pop). */
WPT_DEV void stackPush(uint32_t smBase, uint2* spill, int& sp, uint32_t code, uint32_t tminBits)
{
    if (sp < SM_STACK)
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(smBase + uint32_t(sp) * uint32_t(WPT_STACK_STRIDE * sizeof(uint2))),
                "r"(code), "r"(tminBits) : "memory");
    else
        spill[sp - SM_STACK] = make_uint2(code, tminBits);
    sp++;
}

WPT_DEV int32_t stackPopNext(uint32_t smBase, const uint2* spill, int& sp, float limit)
{
    while (sp > 0) {
        --sp;
        uint2 e;
        if (sp < SM_STACK)
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y)
                    : "r"(smBase + uint32_t(sp) * uint32_t(WPT_STACK_STRIDE * sizeof(uint2))) : "memory");
        else
            e = spill[sp - SM_STACK];
        if (__uint_as_float(e.y) <= limit)
            return int32_t(e.x);
    }
    return EMPTY_CHILD;
}

#ifndef WPT_PUSH_PREDICATED
#define WPT_PUSH_PREDICATED 0   /* 1: the far children of a node are pushed by predicated stores (no branch per child) while the
                                   stack is in shared memory. Most nodes have one or two children hit and the branches skip
                                   the work of the others: -1.5 % on C3 (profiles/r2y_trace_micro_ab.txt) */
#endif

/* st.shared of one stack entry under the predicate key != 0xFFFFFFFF (a child the ray enters) */
WPT_DEV void stackStoreIfHit(uint32_t addr, uint32_t code, uint32_t key)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "setp.ne.u32 p, %2, 0xffffffff;\n\t"
        "and.b32 t, %2, 0xfffffffc;\n\t"
        "@p st.shared.v2.u32 [%0], {%1, t};\n\t}"
        :: "r"(addr), "r"(code), "r"(key) : "memory");
}

WPT_DEV void sort2(uint32_t& a, uint32_t& b)
{
    uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

/* Slab test of one child with hardware min/max. NaNs (ray parallel to and inside a slab plane)
 * drop out of fminf/fmaxf, which only ever widens the interval: conservative. */
WPT_DEV uint32_t childKey(float lox, float hix, float loy, float hiy, float loz, float hiz,
        float3 o, float3 inv, float amin, float amax, uint32_t k)
{
    const float t0x = (lox - o.x) * inv.x, t1x = (hix - o.x) * inv.x;
    const float t0y = (loy - o.y) * inv.y, t1y = (hiy - o.y) * inv.y;
    const float t0z = (loz - o.z) * inv.z, t1z = (hiz - o.z) * inv.z;
    const float tmin = fmaxf(fmaxf(fmaxf(amin, fminf(t0x, t1x)), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float tmax = fminf(fminf(fminf(amax, fmaxf(t0x, t1x)), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    return tmin <= tmax ? ((__float_as_uint(tmin) & ~3u) | k) : 0xFFFFFFFFu;
}

#ifndef WPT_OCTANT_ROWS
#define WPT_OCTANT_ROWS 1   /* (+3.5 % on C3) pick each slab's near / far plane row by the sign of the ray direction (no min/max per plane) */
#endif

/* Slab test of one child when the near and the far plane of every axis are already known (picked by the sign of
 * the ray direction): the products are the very (plane - o) * inv of AABB::mayHit (aabb.hpp:70-86), min(t0, t1) is
 * the near plane's and max(t0, t1) the far plane's because rounding is monotonic. NaNs drop out of fmaxf / fminf. */
WPT_DEV uint32_t childKeyNearFar(float nx, float fx, float ny, float fy, float nz, float fz,
        float3 o, float3 inv, float amin, float amax, uint32_t k)
{
    const float tnx = (nx - o.x) * inv.x, tfx = (fx - o.x) * inv.x;
    const float tny = (ny - o.y) * inv.y, tfy = (fy - o.y) * inv.y;
    const float tnz = (nz - o.z) * inv.z, tfz = (fz - o.z) * inv.z;
    const float tmin = fmaxf(fmaxf(amin, tnx), fmaxf(tny, tnz));
    const float tmax = fminf(fminf(amax, tfx), fminf(tfy, tfz));
    return tmin <= tmax ? ((__float_as_uint(tmin) & ~3u) | k) : 0xFFFFFFFFu;
}

/* The four child keys of a wide node, sorted ascending (nearest entry distance first; 0xFFFFFFFF = missed). */
#ifndef WPT_NODE_FAST
#define WPT_NODE_FAST 0     /* 1: the persistent trace kernel forms the node's row addresses and child keys with nodeKeysFast();
                               2: and marks missed children by the sign bit of the key */
#endif
WPT_DEV bool keyMissed(uint32_t k)
{
#if WPT_NODE_FAST >= 2
    return int32_t(k) < 0;      /* 0xFFFFFFFF or any key with the sign bit set (hit keys are bits of a positive distance) */
#else
    return k == 0xFFFFFFFFu;
#endif
}
struct NodeKeys {
    uint32_t k0, k1, k2, k3;
    int4 ch;
};

WPT_DEV NodeKeys nodeKeys(const DScene& S, int32_t cur, float3 o, float3 inv, float amin, float amax, uint32_t rows)
{
    NodeKeys r;
    const float4* n = S.node4 + size_t(cur) * 8;
#if WPT_OCTANT_ROWS
    /* rows: bit 0/1/2 set = direction x/y/z negative: the near plane is the hi row */
    const uint32_t sx = rows & 1u, sy = (rows >> 1) & 1u, sz = (rows >> 2) & 1u;
    const float4 nx = __ldg(n + sx), fx = __ldg(n + (1u - sx));
    const float4 ny = __ldg(n + 2u + sy), fy = __ldg(n + (3u - sy));
    const float4 nz = __ldg(n + 4u + sz), fz = __ldg(n + (5u - sz));
    r.ch = __ldg(reinterpret_cast<const int4*>(n + 6));
    r.k0 = childKeyNearFar(nx.x, fx.x, ny.x, fy.x, nz.x, fz.x, o, inv, amin, amax, 0);
    r.k1 = childKeyNearFar(nx.y, fx.y, ny.y, fy.y, nz.y, fz.y, o, inv, amin, amax, 1);
    r.k2 = childKeyNearFar(nx.z, fx.z, ny.z, fy.z, nz.z, fz.z, o, inv, amin, amax, 2);
    r.k3 = childKeyNearFar(nx.w, fx.w, ny.w, fy.w, nz.w, fz.w, o, inv, amin, amax, 3);
#else
    const float4 lox = __ldg(n + 0), hix = __ldg(n + 1);
    const float4 loy = __ldg(n + 2), hiy = __ldg(n + 3);
    const float4 loz = __ldg(n + 4), hiz = __ldg(n + 5);
    r.ch = __ldg(reinterpret_cast<const int4*>(n + 6));
    r.k0 = childKey(lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, o, inv, amin, amax, 0);
    r.k1 = childKey(lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, o, inv, amin, amax, 1);
    r.k2 = childKey(lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, o, inv, amin, amax, 2);
    r.k3 = childKey(lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, o, inv, amin, amax, 3);
#endif
    /* 5-comparator sorting network, ascending */
    sort2(r.k0, r.k1); sort2(r.k2, r.k3); sort2(r.k0, r.k2); sort2(r.k1, r.k3); sort2(r.k1, r.k2);
    return r;
}

/* Index of the near row inside each axis' pair of rows (lo row, hi row): 1 when the ray direction's component is
 * negative (sign bit; -0 counts, its inverse is -inf), else 0. Per ray, kept in registers by the persistent trace kernel. */
struct RowOffsets {
    uint32_t x, y, z;
};

WPT_DEV RowOffsets rowOffsets(float3 d)
{
    RowOffsets r;
    r.x = __float_as_uint(d.x) >> 31;
    r.y = __float_as_uint(d.y) >> 31;
    r.z = __float_as_uint(d.z) >> 31;
    return r;
}

/* node + row * 16 bytes as ONE 64-bit multiply-add (IMAD.WIDE.U32, FMA pipe) instead of the shift / mask /
 * add-with-carry sequence the compiler makes of an index expression (ALU pipe, the busier one in the trace kernel). */
WPT_DEV const float4* rowAddress(const float4* node, uint32_t row)
{
    const float4* p;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(p) : "r"(row), "l"(node));
    return p;
}

/* The other row of a pair: the address with bit 4 flipped (nodes are 128-byte aligned, a pair of rows 32-byte aligned),
 * one logic operation on the low word. */
WPT_DEV const float4* otherRow(const float4* row)
{
    const float4* p;
    asm("{\n\t.reg .b32 lo, hi;\n\t"
        "mov.b64 {lo, hi}, %1;\n\t"
        "xor.b32 lo, lo, 16;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "=l"(p) : "l"(row));
    return p;
}

/* tmin <= tmax ? (bits of tmin with the child number in the two low bits) : 0xFFFFFFFF, as compare + predicated
 * three-input logic operation (the select of the plain form is one more ALU instruction per child). */
template<uint32_t K>
WPT_DEV uint32_t childKeyBits(float tmin, float tmax)
{
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.le.f32 p, %1, %2;\n\t"
        "mov.b32 %0, 0xffffffff;\n\t"
        "@p lop3.b32 %0, %1, 0xfffffffc, %3, 0xEA;\n\t}"
        : "=r"(r) : "f"(tmin), "f"(tmax), "n"(K));
    return r;
}

template<uint32_t K>
WPT_DEV uint32_t childKeyFast(float nx, float fx, float ny, float fy, float nz, float fz, float3 o, float3 inv, float amin, float amax)
{
    /* the products and the min / max of childKeyNearFar() */
    const float tnx = (nx - o.x) * inv.x, tfx = (fx - o.x) * inv.x;
    const float tny = (ny - o.y) * inv.y, tfy = (fy - o.y) * inv.y;
    const float tnz = (nz - o.z) * inv.z, tfz = (fz - o.z) * inv.z;
    const float tmin = fmaxf(fmaxf(amin, tnx), fmaxf(tny, tnz));
    const float tmax = fminf(fminf(amax, tfx), fminf(tfy, tfz));
#if WPT_NODE_FAST >= 2
    /* A missed child is marked by the SIGN BIT of its key instead of the value 0xFFFFFFFF (keyMissed() accepts both):
     * tmax - tmin is negative exactly when tmin > tmax (no flush to zero: a difference of two different floats is not
     * zero; equal values give +0), tmin >= amin > 0 has a clear sign bit. One subtraction on the FMA pipe and one logic
     * operation replace a compare and a select on the ALU pipe. */
    const float d = tmax - tmin;
    uint32_t t, key;
    asm("lop3.b32 %0, %1, 0xfffffffc, %2, 0xEA;" : "=r"(t) : "f"(tmin), "n"(K));
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0xF8;" : "=r"(key) : "r"(t), "f"(d));
    return key;
#else
    return childKeyBits<K>(tmin, tmax);
#endif
}

/* nodeKeys() for a walk that keeps the ray's row offsets: same loads, same arithmetic, same keys. */
WPT_DEV NodeKeys nodeKeysFast(const DScene& S, int32_t cur, float3 o, float3 inv, float amin, float amax, RowOffsets off)
{
    NodeKeys r;
    const float4* n = S.node4 + size_t(cur) * 8;
    const float4* px = rowAddress(n, off.x);
    const float4* py = rowAddress(n, off.y);
    const float4* pz = rowAddress(n, off.z);
    const float4 nx = __ldg(px), fx = __ldg(otherRow(px));
    const float4 ny = __ldg(py + 2), fy = __ldg(otherRow(py) + 2);
    const float4 nz = __ldg(pz + 4), fz = __ldg(otherRow(pz) + 4);
    r.ch = __ldg(reinterpret_cast<const int4*>(n + 6));
    r.k0 = childKeyFast<0>(nx.x, fx.x, ny.x, fy.x, nz.x, fz.x, o, inv, amin, amax);
    r.k1 = childKeyFast<1>(nx.y, fx.y, ny.y, fy.y, nz.y, fz.y, o, inv, amin, amax);
    r.k2 = childKeyFast<2>(nx.z, fx.z, ny.z, fy.z, nz.z, fz.z, o, inv, amin, amax);
    r.k3 = childKeyFast<3>(nx.w, fx.w, ny.w, fy.w, nz.w, fz.w, o, inv, amin, amax);
    sort2(r.k0, r.k1); sort2(r.k2, r.k3); sort2(r.k0, r.k2); sort2(r.k1, r.k3); sort2(r.k1, r.k2);
    return r;
}

/* The far children of a node whose nearest child is visited next: k1..k3 of the sorted keys, farthest first (the hit
 * ones are a prefix of the sorted order: k3 hit implies k2 and k1 hit). */
WPT_DEV void pushFarChildren(uint32_t smBase, uint2* spill, int& sp, const NodeKeys& nk)
{
#if WPT_PUSH_PREDICATED
    if (sp <= SM_STACK - 3) {
        constexpr uint32_t STRIDE = uint32_t(WPT_STACK_STRIDE * sizeof(uint2));
        const uint32_t p3 = (nk.k3 != 0xFFFFFFFFu), p2 = (nk.k2 != 0xFFFFFFFFu), p1 = (nk.k1 != 0xFFFFFFFFu);
        const uint32_t a3 = smBase + uint32_t(sp) * STRIDE;
        const uint32_t a2 = a3 + p3 * STRIDE;
        const uint32_t a1 = a2 + p2 * STRIDE;
        stackStoreIfHit(a3, uint32_t(sel4i(nk.ch, nk.k3 & 3u)), nk.k3);
        stackStoreIfHit(a2, uint32_t(sel4i(nk.ch, nk.k2 & 3u)), nk.k2);
        stackStoreIfHit(a1, uint32_t(sel4i(nk.ch, nk.k1 & 3u)), nk.k1);
        sp += int(p3 + p2 + p1);
        return;
    }
#endif
    if (!keyMissed(nk.k3))
        stackPush(smBase, spill, sp, uint32_t(sel4i(nk.ch, nk.k3 & 3u)), nk.k3 & ~3u);
    if (!keyMissed(nk.k2))
        stackPush(smBase, spill, sp, uint32_t(sel4i(nk.ch, nk.k2 & 3u)), nk.k2 & ~3u);
    if (!keyMissed(nk.k1))
        stackPush(smBase, spill, sp, uint32_t(sel4i(nk.ch, nk.k1 & 3u)), nk.k1 & ~3u);
}

#ifndef WPT_STACK_TOS
#define WPT_STACK_TOS 0     /* 1: the persistent trace kernel keeps the youngest stack entry in registers */
#endif

/* The youngest entry of the walk's stack in registers (code == EMPTY_CHILD: the stack is empty). A pop hands it out and
 * asks shared memory for the entry below it, which is not needed before the NEXT pop: the load's latency (the wait at
 * `tmin <= limit`, 6 % of the trace kernel's stall samples) is off the critical path. */
struct StackTop {
    uint32_t code, tmin;
};

WPT_DEV void stackPushTop(uint32_t smBase, uint2* spill, int& sp, StackTop& top, uint32_t code, uint32_t tminBits)
{
    if (top.code != uint32_t(EMPTY_CHILD))
        stackPush(smBase, spill, sp, top.code, top.tmin);
    top.code = code;
    top.tmin = tminBits;
}

WPT_DEV int32_t stackPopNextTop(uint32_t smBase, const uint2* spill, int& sp, StackTop& top, float limit)
{
    while (top.code != uint32_t(EMPTY_CHILD)) {
        const uint32_t code = top.code;
        const float tmin = __uint_as_float(top.tmin);
        if (sp > 0) {
            --sp;
            uint2 e;
            if (sp < SM_STACK)
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y)
                        : "r"(smBase + uint32_t(sp) * uint32_t(WPT_STACK_STRIDE * sizeof(uint2))) : "memory");
            else
                e = spill[sp - SM_STACK];
            top.code = e.x;
            top.tmin = e.y;
        } else {
            top.code = uint32_t(EMPTY_CHILD);
        }
        if (tmin <= limit)
            return int32_t(code);
    }
    return EMPTY_CHILD;
}

WPT_DEV void pushFarChildrenTop(uint32_t smBase, uint2* spill, int& sp, StackTop& top, const NodeKeys& nk)
{
    if (!keyMissed(nk.k3))
        stackPushTop(smBase, spill, sp, top, uint32_t(sel4i(nk.ch, nk.k3 & 3u)), nk.k3 & ~3u);
    if (!keyMissed(nk.k2))
        stackPushTop(smBase, spill, sp, top, uint32_t(sel4i(nk.ch, nk.k2 & 3u)), nk.k2 & ~3u);
    if (!keyMissed(nk.k1))
        stackPushTop(smBase, spill, sp, top, uint32_t(sel4i(nk.ch, nk.k1 & 3u)), nk.k1 & ~3u);
}

/* bit 0/1/2: the ray direction's x/y/z is negative (sign bit; -0 counts, its inverse is -inf) */
WPT_DEV uint32_t octantRows(float3 d)
{
    return (__float_as_uint(d.x) >> 31) | ((__float_as_uint(d.y) >> 31) << 1) | ((__float_as_uint(d.z) >> 31) << 2);
}

/* State of the best hit so far in the wide walk, with what the near-tie rule needs. */
struct WideBest {
    Hit hit;
    float Ts, det;      /* T and |det| of the best hit if it is a triangle */
};

/* One candidate of the wide walk against the current best.
 * (Near-)ties between two triangles: the reference meets triangles in storage order, i.e. in slot
 * order, and a later triangle Y replaces an earlier X iff T_Y <= a_X * det_Y
 * (hitable_triangle.hpp:263-264, bvh.hpp:292-295). When this walk meets them the other way round
 * (best = Y, candidate = X with the lower slot), that very comparison is redone with Y's stored T and
 * det, so the reference's decision is reproduced. */
WPT_DEV void considerHit(WideBest& best, uint32_t ref, float a, float Ts, float da)
{
    const bool keepBest = (best.hit.ref != NO_HIT)
        && ((best.hit.ref >> REF_TYPE_SHIFT) == PRIM_TRIANGLE) && ((ref >> REF_TYPE_SHIFT) == PRIM_TRIANGLE)
        && (ref < best.hit.ref) && (best.Ts <= a * best.det);
    if (!keepBest) {
        best.hit.a = a;
        best.hit.ref = ref;
        best.Ts = Ts;
        best.det = da;
    }
}

#ifndef WPT_MEDIUM_COOP
#define WPT_MEDIUM_COOP 1   /* the persistent trace kernel tests small triangle boundaries of media with the whole warp */
#endif

/* HitableMedium::hit (hitable_medium.hpp:69-98) for the lanes of a warp whose walk waits at a medium whose boundary is a
 * small triangle mesh (`run` != 0: first triangle slot | count << 26, count <= 32; scene_host.cpp). Called by ALL lanes
 * of the warp. The per-lane form (mediumTest) walks the nested tree twice per medium - some thousand instructions at
 * the handful of lanes that wait at a medium, while the rest of the warp idles. Here the lanes that wait are served one
 * after the other: the owner's ray goes to all lanes by shuffle, lane j runs the watertight test of boundary triangle j
 * ONCE without bounds (a triangle's distance along the line does not depend on them), and the two closest-hit queries of
 * the reference (first crossing on the whole line, next crossing behind it) are replayed over the triangles that were
 * hit, in slot order = the order the nested walk meets them, with the very comparisons of the bounded test
 * (hitable_triangle.hpp:263-264) against the shrinking upper bound (bvh.hpp:292-295). The box tests of the nested walk
 * only cull what fails those comparisons, so the two distances are the nested walk's: bit for bit on 500 000 random
 * rays against the restated reference (oracle/port.cpp, tests/test_oracle_port.py); only rays aimed exactly at a vertex
 * or an edge can differ - by which of the coincident triangles is accepted (a few ulp), or because a box test of the
 * reference culls a crossing the watertight triangle test finds. The owner then samples the free flight from its own
 * Prng exactly as mediumTest does. Scenes without animated hitables only (the
 * corners are read for the owner's time). */
template<int F>
WPT_DEV void mediumRoundCoop(const DScene& S, uint32_t run, uint32_t mediumRef, const Ray& ray, const RayHelper& h,
        float amin, float minHit, Ctx<F>& ctx, WideBest& best)
{
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    unsigned waiting = __ballot_sync(FULL, run != 0u);
    float myA1 = 0.0f, myA2 = 0.0f;     /* this lane's two boundary distances, if it is an owner and both queries hit */
    bool myBoth = false;
    const uint32_t kpack = uint32_t(h.kx) | (uint32_t(h.ky) << 2) | (uint32_t(h.kz) << 4);
    while (waiting != 0u) {
        const int owner = __ffs(waiting) - 1;
        waiting &= waiting - 1u;
        Ray oray;
        oray.o = f3(__shfl_sync(FULL, ray.o.x, owner), __shfl_sync(FULL, ray.o.y, owner), __shfl_sync(FULL, ray.o.z, owner));
        oray.d = f3(0.0f);
        oray.time = 0.0f;
        RayHelper oh;
        oh.inv = f3(0.0f);
        oh.Sx = __shfl_sync(FULL, h.Sx, owner);
        oh.Sy = __shfl_sync(FULL, h.Sy, owner);
        oh.Sz = __shfl_sync(FULL, h.Sz, owner);
        const uint32_t ok = __shfl_sync(FULL, kpack, owner);
        oh.kx = int(ok & 3u); oh.ky = int((ok >> 2) & 3u); oh.kz = int(ok >> 4);
        const uint32_t orun = __shfl_sync(FULL, run, owner);
        const uint32_t first = orun & ((1u << 26) - 1u), count = orun >> 26;
        float ta = 0.0f, Ts = 0.0f, da = 0.0f;
        bool found = false;
        if (lane < count) {
            const float4* tp = S.tri + size_t(first + lane) * 3;
            const float4 t0 = __ldg(tp + 0), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            float U, V, W, det;
            found = triangleTest(xyz(t0), xyz(t1), xyz(t2), oray, oh, 0.0f, 0.0f, false, ta, U, V, W, det, Ts, da);
        }
        unsigned hits = __ballot_sync(FULL, found);
        /* first query: (-maxval, maxval); second query: (a1 + minHitDistance, maxval) */
        float a1 = MAXVAL, a2 = MAXVAL;
        bool have1 = false, have2 = false;
        for (unsigned m = hits; m != 0u; m &= m - 1u) {
            const int j = __ffs(m) - 1;
            const float Tj = __shfl_sync(FULL, Ts, j), dj = __shfl_sync(FULL, da, j), aj = __shfl_sync(FULL, ta, j);
            if (!(Tj < -MAXVAL * dj || Tj > a1 * dj)) {
                a1 = aj;
                have1 = true;
            }
        }
        if (have1) {
            const float lo = a1 + minHit;
            for (unsigned m = hits; m != 0u; m &= m - 1u) {
                const int j = __ffs(m) - 1;
                const float Tj = __shfl_sync(FULL, Ts, j), dj = __shfl_sync(FULL, da, j), aj = __shfl_sync(FULL, ta, j);
                if (!(Tj < lo * dj || Tj > a2 * dj)) {
                    a2 = aj;
                    have2 = true;
                }
            }
        }
        if (int(lane) == owner) {
            ctx.nested += have1 ? 2u : 1u;
            myA1 = a1;
            myA2 = a2;
            myBoth = have1 && have2;
        }
    }
    /* the owners sample their free flights together (hitable_medium.hpp:83-95) */
    if (myBoth) {
        float a1 = myA1, a2 = myA2;
        const float amax = best.hit.a;
        if (a1 < amin)
            a1 = amin;
        if (a2 > amax)
            a2 = amax;
        if (a1 < a2) {
            if (a1 < 0.0f)
                a1 = 0.0f;
            const float inside = a2 - a1;
            const float density = __ldg(S.med + (mediumRef & REF_SLOT_MASK)).x;
            const float freeFlight = -(1.0f / density) * logf(in01(ctx.prng));
            if (freeFlight < inside)
                considerHit(best, mediumRef, a1 + freeFlight, 0.0f, 0.0f);
        }
    }
}

/* A leaf of the wide tree: one hitable, or a run of up to four consecutive triangle slots. */
template<int F>
WPT_DEV void wideLeaf(const DScene& S, uint32_t ref, const Ray& ray, const RayHelper& h,
        float amin, float minHit, Ctx<F>& ctx, WideBest& best)
{
    float a, Ts, da;
    if ((ref >> REF_TYPE_SHIFT) == PRIM_TRIANGLE_RUN) {
        const uint32_t first = ref & RUN_SLOT_MASK;
        const uint32_t count = ((ref >> RUN_COUNT_SHIFT) & 3u) + 1u;
        for (uint32_t i = 0; i < count; i++) {
            const uint32_t tref = (PRIM_TRIANGLE << REF_TYPE_SHIFT) | (first + i);
            if (leafTest<F, true>(S, tref, ray, h, amin, best.hit.a, minHit, ctx, a, Ts, da))
                considerHit(best, tref, a, Ts, da);
        }
    } else if (leafTest<F, true>(S, ref, ray, h, amin, best.hit.a, minHit, ctx, a, Ts, da)) {
        considerHit(best, ref, a, Ts, da);
    }
}

/* One step of a leaf: tests ONE primitive of the leaf `ref` and returns what is left of the leaf (a shorter run
 * of triangles), or NO_HIT when the leaf is done. The persistent trace kernel runs its leaf rounds with this, so
 * that every lane of a leaf round does exactly one primitive test whatever the length of its run. */
template<int F>
WPT_DEV uint32_t wideLeafStep(const DScene& S, uint32_t ref, const Ray& ray, const RayHelper& h,
        float amin, float minHit, Ctx<F>& ctx, WideBest& best)
{
    float a, Ts, da;
    uint32_t tref = ref, rest = NO_HIT;
    if ((ref >> REF_TYPE_SHIFT) == PRIM_TRIANGLE_RUN) {
        const uint32_t first = ref & RUN_SLOT_MASK;
        const uint32_t more = (ref >> RUN_COUNT_SHIFT) & 3u;       /* triangles after this one */
        tref = (PRIM_TRIANGLE << REF_TYPE_SHIFT) | first;
        if (more != 0u)
            rest = (PRIM_TRIANGLE_RUN << REF_TYPE_SHIFT) | ((more - 1u) << RUN_COUNT_SHIFT) | (first + 1u);
    }
    if (leafTest<F, true>(S, tref, ray, h, amin, best.hit.a, minHit, ctx, a, Ts, da))
        considerHit(best, tref, a, Ts, da);
    return rest;
}

/* Asks for the triangle data of a leaf the walk has just arrived at; the lanes of the persistent trace
 * kernel then wait a few iterations for their leaf round (see there), which covers the fetch. */
WPT_DEV void prefetchLeaf(const DScene& S, uint32_t ref)
{
    const uint32_t type = ref >> REF_TYPE_SHIFT;
    if (type == PRIM_TRIANGLE_RUN || type == PRIM_TRIANGLE) {
        const uint32_t first = (type == PRIM_TRIANGLE_RUN ? (ref & RUN_SLOT_MASK) : (ref & REF_SLOT_MASK));
        const uint32_t count = (type == PRIM_TRIANGLE_RUN ? ((ref >> RUN_COUNT_SHIFT) & 3u) + 1u : 1u);
        const float4* p = S.tri + size_t(first) * 3;
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 3u * count - 1u));
    }
}

/* Closest hit under (amin, amax), amin > 0.
 *
 * "while-while" form (Aila & Laine 2009): an inner loop that only descends through inner nodes,
 * then one leaf test. A lane that reaches a leaf early waits at the end of the inner loop until
 * the other lanes of the warp have found theirs, so that the (long) node code and the (long) leaf
 * code each run with as many lanes as possible instead of interleaving lane by lane. */
template<int F>
__device__ Hit traverseWide(const DScene& S, const Ray& ray, const RayHelper& h,
        float amin, float amax, float minHit, Ctx<F>& ctx, uint2* smStack, int smStride)
{
    WideBest best;
    best.hit.a = amax;
    best.hit.ref = NO_HIT;
    best.Ts = best.det = 0.0f;
    /* (smStride is the block size of the calling kernel: WPT_STACK_STRIDE) */
    const uint32_t smBase = uint32_t(__cvta_generic_to_shared(smStack));
    uint2 spill[LOCAL_STACK];
    int sp = 0;
    (void)smStride;
    int32_t cur = S.root4;                  /* EMPTY_CHILD doubles as "walk finished" */
    const uint32_t rows = octantRows(ray.d);
    for (;;) {
        while (cur >= 0 && cur != EMPTY_CHILD) {
            if (F & F_COUNT)
                ctx.nodes++;
            const NodeKeys nk = nodeKeys(S, cur, ray.o, h.inv, amin, best.hit.a, rows);
            if (nk.k0 != 0xFFFFFFFFu) {
                /* far children go on the stack farthest first, the nearest is visited next */
                pushFarChildren(smBase, spill, sp, nk);
                cur = sel4i(nk.ch, nk.k0 & 3u);
            } else {
                cur = stackPopNext(smBase, spill, sp, best.hit.a);
            }
        }
        if (cur == EMPTY_CHILD)
            return best.hit;
        wideLeaf<F>(S, uint32_t(~cur), ray, h, amin, minHit, ctx, best);
        cur = stackPopNext(smBase, spill, sp, best.hit.a);
    }
}

}
