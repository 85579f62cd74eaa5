/*
 * device_shade.cuh -- everything tracePath() calls apart from the tree walk: hit records,
 * textures, materials and lights, environment maps, hot-spot sampling, camera rays, sensors.
 * Each function names the reference code it restates; arithmetic is kept in the reference's
 * order (see device_math.cuh).
 */
#pragma once

#include "device_trace.cuh"

namespace wptd {

/* HitRecord (hitable.hpp:39-64) */
struct HitRec {
    float a;
    float3 p, n, t;
    float2 tc;
    bool backside;
    int32_t material;       /* material index of the hitable, -1 = null */
};

enum { SCATTER_NONE = 0, SCATTER_EXPLICIT = 1, SCATTER_RANDOM = 2 };

/* ScatterRecord (material.hpp:43-80) */
struct ScatterRec {
    int type;
    float3 dir;
    float4 att;
    float pdf;
    float4 ri;
};

WPT_DEV ScatterRec scatterNone()
{
    ScatterRec s;
    s.type = SCATTER_NONE;
    s.dir = f3(0.0f);
    s.att = f4(0.0f);
    s.pdf = 0.0f;
    s.ri = f4(0.0f);
    return s;
}

/* ------------------------------------------------------------ hit records */

/* HitableTriangle::hit with fullInfo (hitable_triangle.hpp:273-324) */
template<int F>
WPT_DEV HitRec triangleHitRecord(const DScene& S, uint32_t slot, const Ray& ray, const RayHelper& h, const Ctx<F>& ctx)
{
    float3 v0, v1, v2;
    uint32_t instIdx, triIdx, anim;
    triangleCorners<F>(S, slot, ctx, v0, v1, v2, instIdx, triIdx, anim);
    float a = 0.0f, U = 0.0f, V = 0.0f, W = 0.0f, det = 1.0f;
    triangleTest(v0, v1, v2, ray, h, 0.0f, 0.0f, false, a, U, V, W, det);
    const float invDet = 1.0f / det;
    const bool backfacing = (det < 0.0f);
    const float b0 = invDet * U, b1 = invDet * V, b2 = invDet * W;

    /* vertex attribute blocks and mesh flags were resolved per slot at updateBVH time (scene_host.cpp) */
    const uint4 sh = __ldg(S.triShade + slot);
    const wpt_instance* inst = S.instances + instIdx;
    const float* p0 = S.vertices + sh.x;
    const float* p1 = S.vertices + sh.y;
    const float* p2 = S.vertices + sh.z;
    const uint32_t meshFlags = sh.w;
    const bool transform = inst->has_transform != 0;

    HitRec hr;
    hr.a = a;
    hr.p = ray.o + a * ray.d;
    hr.material = inst->material;
    hr.backside = backfacing;

    float3 n = (b0 * f3(p0[3], p0[4], p0[5]) + b1 * f3(p1[3], p1[4], p1[5])) + b2 * f3(p2[3], p2[4], p2[5]);
    if (transform)
        n = mul33(inst->N, n);
    if ((F & F_ANIM) && anim != 0)
        n = mul33(animNormalMatrix<F>(S, ctx, anim), n);
    n = normalize(n);
    if (backfacing)
        n = -n;
    hr.n = n;

    hr.tc = make_float2(0.0f, 0.0f);
    if (meshFlags & WPT_MESHF_TEXCOORDS) {
        hr.tc.x = (b0 * p0[6] + b1 * p1[6]) + b2 * p2[6];
        hr.tc.y = (b0 * p0[7] + b1 * p1[7]) + b2 * p2[7];
    }
    hr.t = f3(0.0f);
    if (meshFlags & WPT_MESHF_TANGENTS) {
        float3 t = (b0 * f3(p0[8], p0[9], p0[10]) + b1 * f3(p1[8], p1[9], p1[10])) + b2 * f3(p2[8], p2[9], p2[10]);
        if (dot(t, t) > 0.0f) {
            if (transform)
                t = mul33(inst->N, t);
            if ((F & F_ANIM) && anim != 0)
                t = mul33(animNormalMatrix<F>(S, ctx, anim), t);
            t = normalize(t - dot(n, t) * n);
        }
        hr.t = t;
    }
    return hr;
}

/* HitableSphere::constructHitRecord (hitable_sphere.hpp:43-75) */
WPT_DEV HitRec sphereHitRecord(const SphereGeom& g, const Ray& ray, float a)
{
    HitRec hr;
    hr.a = a;
    hr.p = ray.o + a * ray.d;
    float3 n = normalize(hr.p - g.center);
    const float3 rn = qrotate(g.rotation, n);
    const float alpha = atan2f(rn.x, rn.z);
    const float beta = asinf(rclamp(rn.y, -1.0f, 1.0f));
    hr.tc = make_float2(0.5f * INV_PI * (alpha + PI), INV_PI * (beta + 0.5f * PI));
    float3 t = f3(cosf(alpha), 0.0f, -sinf(alpha));
    if (fabsf(dot(rn, t)) >= COS_ORTHO_TOL)
        t = f3(0.0f);
    hr.t = t;
    hr.backside = false;
    if (dot(n, -ray.d) < 0.0f) {
        hr.backside = true;
        n = -n;
    }
    hr.n = n;
    hr.material = int32_t(g.material);
    return hr;
}

template<int F>
WPT_DEV HitRec buildHitRecord(const DScene& S, const Hit& hit, const Ray& ray, const RayHelper& h, const Ctx<F>& ctx)
{
    const uint32_t type = hit.ref >> REF_TYPE_SHIFT;
    const uint32_t slot = hit.ref & REF_SLOT_MASK;
    if (type == PRIM_TRIANGLE)
        return triangleHitRecord<F>(S, slot, ray, h, ctx);
    if (type == PRIM_SPHERE)
        return sphereHitRecord(sphereGeom<F>(S, slot, ctx), ray, hit.a);
    /* medium (hitable_medium.hpp:91-92) */
    HitRec hr;
    hr.a = hit.a;
    hr.p = ray.o + hit.a * ray.d;
    hr.n = f3(0.0f);
    hr.t = f3(0.0f);
    hr.tc = make_float2(0.0f, 0.0f);
    hr.backside = false;
    hr.material = int32_t(__float_as_uint(__ldg(S.med + slot).y));
    return hr;
}

/* --------------------------------------------------------------- textures */

/* TextureImage::value (texture_image.hpp:85-212) on texels decoded to RGBA float at upload */
WPT_DEV float4 imageTexel(const DScene& S, const DTexture& t, uint32_t x, uint32_t y)
{
    if (x >= uint32_t(t.width))
        x = t.width - 1;
    if (y >= uint32_t(t.height))
        y = t.height - 1;
    return __ldg(S.texels + t.texelBase + size_t(y) * t.width + x);
}

static __device__ __noinline__ float4 textureValue(const DScene& S, int32_t index, float2 uv)
{
    /* unwind TextureTransformer chains (texture.hpp:207-246) iteratively */
    float4 factors[4], offsets[4];
    int depth = 0;
    const DTexture* t = S.textures + index;
    while (t->type == WPT_TEX_TRANSFORMER && depth < 4) {
        uv = make_float2(t->coordFactor.x * uv.x + t->coordOffset.x, t->coordFactor.y * uv.y + t->coordOffset.y);
        factors[depth] = t->valFactor;
        offsets[depth] = t->valOffset;
        depth++;
        t = S.textures + t->src;
    }
    float4 v;
    if (t->type == WPT_TEX_CONSTANT) {
        v = t->color0;
    } else if (t->type == WPT_TEX_CHECKER) {
        int row = int(uv.y * t->vert);
        int col = int(uv.x * t->horiz);
        v = (row % 2 == col % 2 ? t->color0 : t->color1);
    } else {
        const float u = fractf(t->coordFactor.x * uv.x + t->coordOffset.x);
        const float w = fractf(t->coordFactor.y * uv.y + t->coordOffset.y);
        const float uvs = rmax(0.0f, (u * t->width) - 0.5f);
        const float uvt = rmax(0.0f, (w * t->height) - 0.5f);
        const uint32_t x0 = uint32_t(uvs), y0 = uint32_t(uvt);
        const float alpha = uvs - float(x0);
        const float beta = uvt - float(y0);
        const float4 v00 = imageTexel(S, *t, x0, y0);
        const float4 v10 = imageTexel(S, *t, x0 + 1, y0);
        const float4 v01 = imageTexel(S, *t, x0, y0 + 1);
        const float4 v11 = imageTexel(S, *t, x0 + 1, y0 + 1);
        const float4 a = vmix(v00, v10, alpha);
        const float4 b = vmix(v01, v11, alpha);
        v = t->valFactor * vmix(a, b, beta) + t->valOffset;
    }
    while (depth > 0) {
        depth--;
        v = factors[depth] * v + offsets[depth];
    }
    return v;
}

/* -------------------------------------------------------------- materials */

/* Material::normalAt (material.hpp:195-205) */
WPT_DEV float3 normalAt(const DScene& S, const wpt_material* m, const HitRec& hr)
{
    float3 n = hr.n;
    if (m->normal_tex >= 0) {
        float3 c = xyz(textureValue(S, m->normal_tex, hr.tc));
        c = 2.0f * c - f3(1.0f);
        n = normalize(toWorld(tangentSpace(hr.n, hr.t), c));
    }
    return n;
}

/* Material::tangentSpaceAt (material.hpp:210-228) */
WPT_DEV TangentSpace tangentSpaceAt(const DScene& S, const wpt_material* m, const HitRec& hr)
{
    TangentSpace ts;
    if (dot(hr.t, hr.t) > EPSILON) {
        ts = tangentSpace(hr.n, hr.t);
        if (m->normal_tex >= 0) {
            float3 n = xyz(textureValue(S, m->normal_tex, hr.tc));
            n = 2.0f * n - f3(1.0f);
            n = normalize(toWorld(ts, n));
            float3 t = normalize(hr.t - dot(n, hr.t) * n);
            ts = tangentSpace(n, t);
        }
    } else {
        ts = tangentSpace(hr.n);
    }
    return ts;
}

WPT_DEV float4 ld4(const float* p) { return f4(p[0], p[1], p[2], p[3]); }

/* value-or-texture with the "no NIR channel given" rule (e.g. material_lambertian.hpp:53-59) */
WPT_DEV float4 colorAt(const DScene& S, const wpt_material* m, const float* constant, int32_t tex, float2 tc)
{
    float4 a = (tex >= 0 ? textureValue(S, tex, tc) : ld4(constant));
    if (!(m->flags & WPT_MATF_HAVE_NIR))
        a.w = average(xyz(a));
    return a;
}

/* MaterialTwoSided (material.hpp:273-334): pick the side's material; the hit is then treated as
 * a front-side hit (toFrontSide). A null side behaves like the base Material. */
WPT_DEV const wpt_material* resolveMaterial(const DScene& S, int32_t index, HitRec& hr)
{
    if (index < 0)
        return nullptr;
    const wpt_material* m = S.materials + index;
    if (m->type == WPT_MAT_TWOSIDED) {
        const int32_t side = hr.backside ? m->back : m->front;
        hr.backside = false;
        return side < 0 ? nullptr : S.materials + side;
    }
    return m;
}

/* ---- GGX helpers (material_ggx.hpp:89-171) ---- */
WPT_DEV float ggxLambda(float3 v, float2 r)
{
    const float ax = r.x * r.x, ay = r.y * r.y;
    const float3 v2 = v * v;
    const float disc = 1.0f + (ax * v2.x + ay * v2.y) / v2.z;
    return 0.5f * (-1.0f + sqrtf(disc));
}
WPT_DEV float ggxG1(float3 v, float2 r) { return 1.0f / (1.0f + ggxLambda(v, r)); }
WPT_DEV float ggxG2(float3 v, float3 l, float2 r) { return 1.0f / ((1.0f + ggxLambda(v, r)) + ggxLambda(l, r)); }
WPT_DEV float ggxD(float3 hh, float2 r)
{
    const float ax = r.x * r.x, ay = r.y * r.y;
    const float3 h2 = hh * hh;
    const float t = (h2.x / ax + h2.y / ay) + h2.z;
    return 1.0f / ((((PI * r.x) * r.y) * t) * t);
}
WPT_DEV float ggxDV(float3 hh, float3 v, float dotVH, float2 r)
{
    return ((ggxG1(v, r) * dotVH) * ggxD(hh, r)) / v.z;
}
WPT_DEV float3 ggxSampleVNDF(float3 Ve, float2 r, Prng& prng)
{
    const float U1 = in01(prng);
    const float U2 = in01(prng);
    const float3 Vh = normalize(f3(r.x * Ve.x, r.y * Ve.y, Ve.z));
    const float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    const float3 T1 = lensq > 0.0f ? f3(-Vh.y, Vh.x, 0.0f) * (1.0f / sqrtf(lensq)) : f3(1.0f, 0.0f, 0.0f);
    const float3 T2 = cross(Vh, T1);
    const float rr = sqrtf(U1);
    const float phi = 2.0f * PI * U2;
    const float t1 = rr * cosf(phi);
    float t2 = rr * sinf(phi);
    const float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
    const float3 Nh = (t1 * T1 + t2 * T2) + sqrtf(rmax(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(f3(r.x * Nh.x, r.y * Nh.y, rmax(0.0f, Nh.z)));
}
WPT_DEV float2 ggxRoughnessAt(const DScene& S, const wpt_material* m, float2 tc)
{
    if (m->tex[1] >= 0) {
        float4 v = textureValue(S, m->tex[1], tc);
        return make_float2(v.x, v.y);
    }
    return make_float2(m->f[0], m->f[1]);
}
WPT_DEV float4 ggxAlbedoAt(const DScene& S, const wpt_material* m, float2 tc)
{
    return m->tex[0] >= 0 ? textureValue(S, m->tex[0], tc) : ld4(m->v0);
}

/* ---- ModPhong helpers (material_modphong.hpp:133-239) ---- */
WPT_DEV float4 phongDiffuseAt(const DScene& S, const wpt_material* m, float2 tc) { return colorAt(S, m, m->v0, m->tex[0], tc); }
WPT_DEV float4 phongSpecularAt(const DScene& S, const wpt_material* m, float2 tc)
{
    float4 ks = (m->tex[1] >= 0 ? textureValue(S, m->tex[1], tc) : ld4(m->v1));
    if (m->flags & WPT_MATF_SPECULAR_TEX_ALPHA)
        ks = f4(vmix(xyz(ks), f3(m->v1[0], m->v1[1], m->v1[2]), ks.w), ks.w);
    if (!(m->flags & WPT_MATF_HAVE_NIR))
        ks.w = average(xyz(ks));
    return ks;
}
WPT_DEV float phongShininessAt(const DScene& S, const wpt_material* m, float2 tc)
{
    float s = m->f[0];
    if (m->tex[2] >= 0)
        s *= textureValue(S, m->tex[2], tc).x;
    return s;
}
WPT_DEV float phongOpacityAt(const DScene& S, const wpt_material* m, float2 tc)
{
    if (m->tex[3] >= 0)
        return textureValue(S, m->tex[3], tc).x;
    if (m->flags & WPT_MATF_DIFFUSE_TEX_ALPHA)
        return textureValue(S, m->tex[0], tc).w;
    return m->f[1];
}
WPT_DEV float4 phongAttenuation(float3 n, float3 v, float3 l, float4 kd, float4 ks, float s, float cosTheta)
{
    const float3 r = reflect(-l, n);
    const float cosRV = rmax(dot(r, v), 0.0f);
    return ((kd + ((0.5f * ks) * (s + 2.0f)) * powf(cosRV, s)) * INV_PI) * rmin(cosTheta, 1.0f);
}
WPT_DEV float phongSpecularProbability(float4 kd, float4 ks)
{
    const float skd = ((kd.x + kd.y) + kd.z) + kd.w;
    const float sks = ((ks.x + ks.y) + ks.z) + ks.w;
    const float sum = (skd + sks) + 1e-4f;
    return rclamp(sks / sum, 0.1f, 0.9f);
}
WPT_DEV float phongPdf(float3 n, float3 v, float3 l, float s, float cosTheta, float specProb)
{
    const float diffusePdf = cosTheta * INV_PI;
    const float3 r = reflect(-v, n);
    const float cosRL = rmax(dot(r, l), 0.0f);
    const float specularPdf = ((0.5f * INV_PI) * (s + 1.0f)) * powf(cosRL, s);
    return rmix(diffusePdf, specularPdf, specProb);
}

/* Material::scatter of the resolved material. MT >= 0: the material's type is known at compile time (the
 * shading queues are sorted by material class, device_scene.cuh MC_*), the switch collapses to one case. */
template<int MT>
WPT_DEV ScatterRec materialScatterT(const DScene& S, const wpt_material* m, const Ray& ray, float4 rayRi,
        const HitRec& hr, Prng& prng)
{
    if (!m)
        return scatterNone();
    ScatterRec sr;
    switch (MT >= 0 ? MT : m->type) {
    case WPT_MAT_LAMBERTIAN: {      /* material_lambertian.hpp:61-84 */
        if (hr.backside)
            return scatterNone();
        const float3 cd = cosineDirection(in01x2(prng));
        const float cosTheta = cd.z;
        const TangentSpace ts = tangentSpaceAt(S, m, hr);
        sr.type = SCATTER_RANDOM;
        sr.dir = normalize(toWorld(ts, cd));
        sr.pdf = cosTheta * INV_PI;
        sr.att = colorAt(S, m, m->v0, m->tex[0], hr.tc) * sr.pdf;
        sr.ri = rayRi;
        return sr;
    }
    case WPT_MAT_MIRROR: {          /* material_mirror.hpp:53-62 */
        if (hr.backside)
            return scatterNone();
        const float3 reflected = reflect(ray.d, normalAt(S, m, hr));
        sr.type = SCATTER_EXPLICIT;
        sr.dir = normalize(reflected);
        sr.att = colorAt(S, m, m->v0, m->tex[0], hr.tc);
        sr.pdf = 0.0f;
        sr.ri = rayRi;
        return sr;
    }
    case WPT_MAT_GLASS: {           /* material_glass.hpp:91-152 */
        float4 att = f4(1.0f);
        float4 ourRI = ld4(m->v1);
        float4 theirRI = ld4(m->v2);
        int riIndex = 0;
        if (m->flags & WPT_MATF_CHROMATIC) {
            riIndex = int(in01(prng) * 4);
            att = f4(riIndex == 0 ? 4.0f : 0.0f, riIndex == 1 ? 4.0f : 0.0f,
                    riIndex == 2 ? 4.0f : 0.0f, riIndex == 3 ? 4.0f : 0.0f);
        }
        if (hr.backside) {
            float4 tmp = ourRI;
            ourRI = theirRI;
            theirRI = tmp;
            const float dist = hr.a;
            att = att * f4(expf(-m->v0[0] * dist), expf(-m->v0[1] * dist), expf(-m->v0[2] * dist), expf(-m->v0[3] * dist));
        }
        const float3 n = normalAt(S, m, hr);
        const float theirs = sel(theirRI, riIndex), ours = sel(ourRI, riIndex);
        const float3 refracted = refract(ray.d, n, theirs / ours);
        bool doReflection = true;
        if (dot(refracted, refracted) > 0.0f) {
            const float cosIncident = dot(-ray.d, n);
            const float cosTransmitted = -dot(refracted, n);
            const float fresnel = fresnelUnpolarized(cosIncident, cosTransmitted, theirs, ours);
            doReflection = in01(prng) < fresnel;
        }
        sr.type = SCATTER_EXPLICIT;
        sr.att = att;
        sr.pdf = 0.0f;
        if (doReflection) {
            sr.dir = normalize(reflect(ray.d, n));
            sr.ri = theirRI;
        } else {
            sr.dir = normalize(refracted);
            sr.ri = ourRI;
        }
        return sr;
    }
    case WPT_MAT_GGX: {             /* material_ggx.hpp:173-225 */
        if (hr.backside)
            return scatterNone();
        const float3 view = -ray.d;
        const float2 rough = ggxRoughnessAt(S, m, hr.tc);
        const TangentSpace ts = tangentSpaceAt(S, m, hr);
        const float3 tsV = toTangent(ts, view);
        const float3 tsH = ggxSampleVNDF(tsV, rough, prng);
        const float3 tsL = reflect(-tsV, tsH);
        const float3 light = toWorld(ts, tsL);
        const float l = dot(light, light);
        if (l < EPSILON)
            return scatterNone();
        const float3 dir = light / sqrtf(l);
        const float dotVH = dot(tsV, tsH);
        const float p = ggxDV(tsH, tsV, dotVH, rough) / (4.0f * dotVH);
        if (!isfinite(p) || p < 0.0f)
            return scatterNone();
        float4 att = f4(0.0f);
        const float dotNL = dot(ts.normal, light);
        const float dotNV = dot(ts.normal, view);
        if (dotNL > 0.0f && dotNV > 0.0f) {
            const float Dval = ggxD(tsH, rough);
            const float4 Fval = fresnelSchlick(ggxAlbedoAt(S, m, hr.tc), dotVH);
            const float Gval = ggxG2(tsV, tsL, rough);
            att = ((Dval * Fval) * Gval) / (4.0f * dotNV);
        }
        sr.type = SCATTER_RANDOM;
        sr.dir = dir;
        sr.att = att;
        sr.pdf = p;
        sr.ri = rayRi;
        return sr;
    }
    case WPT_MAT_MODPHONG: {        /* material_modphong.hpp:241-308 */
        const float opa = phongOpacityAt(S, m, hr.tc);
        const bool transparent = (opa < 1.0f && opa < in01(prng));
        if (transparent) {
            float ourRI = m->f[2];
            float theirRI = 1.0f;
            if (hr.backside) {
                float tmp = ourRI;
                ourRI = theirRI;
                theirRI = tmp;
            }
            const float3 n = normalize(normalAt(S, m, hr));
            float3 refracted = refract(ray.d, n, theirRI / ourRI);
            const float l = dot(refracted, refracted);
            if (l < EPSILON)
                return scatterNone();
            refracted = refracted / sqrtf(l);
            float4 att = ld4(m->v2);
            if (!(m->flags & WPT_MATF_HAVE_NIR))
                att.w = average(xyz(att));
            sr.type = SCATTER_EXPLICIT;
            sr.dir = refracted;
            sr.att = att;
            sr.pdf = 0.0f;
            sr.ri = f4(ourRI);
            return sr;
        }
        if (hr.backside)
            return scatterNone();
        const float4 kd = phongDiffuseAt(S, m, hr.tc);
        const float4 ks = phongSpecularAt(S, m, hr.tc);
        const float s = phongShininessAt(S, m, hr.tc);
        const float specProb = phongSpecularProbability(kd, ks);
        float3 dir, n;
        float cosTheta;
        if (in01(prng) < specProb) {
            const float r1 = in01(prng);
            const float r2 = in01(prng);
            const float cosThetaSpec = powf(1.0f - r1, 1.0f / (1.0f + s));
            const float disc = rmax(1.0f - cosThetaSpec * cosThetaSpec, 0.0f);
            const float sinThetaSpec = sqrtf(disc);
            const float phi = 2.0f * PI * r2;
            const float x = cosf(phi) * sinThetaSpec;
            const float y = sinf(phi) * sinThetaSpec;
            const float z = cosThetaSpec;
            n = normalAt(S, m, hr);
            const TangentSpace specTS = tangentSpace(reflect(ray.d, n));
            dir = normalize(toWorld(specTS, f3(x, y, z)));
            cosTheta = rmax(dot(dir, n), 0.0f);
        } else {
            const TangentSpace ts = tangentSpaceAt(S, m, hr);
            n = ts.normal;
            const float3 cd = cosineDirection(in01x2(prng));
            cosTheta = cd.z;
            dir = normalize(toWorld(ts, cd));
        }
        sr.type = SCATTER_RANDOM;
        sr.dir = dir;
        sr.att = phongAttenuation(n, -ray.d, dir, kd, ks, s, cosTheta);
        sr.pdf = phongPdf(n, -ray.d, dir, s, cosTheta, specProb);
        sr.ri = rayRi;
        return sr;
    }
    case WPT_MAT_PHASE_ISOTROPIC: { /* material_phase_function_isotropic.hpp:51-59 */
        sr.type = SCATTER_EXPLICIT;
        sr.dir = onUnitSphere(in01x2(prng));
        sr.att = ld4(m->v0);
        sr.pdf = 0.0f;
        sr.ri = rayRi;
        return sr;
    }
    default:                        /* lights and the base Material do not scatter */
        return scatterNone();
    }
}

static __device__ __noinline__ ScatterRec materialScatter(const DScene& S, const wpt_material* m, const Ray& ray, float4 rayRi,
        const HitRec& hr, Prng& prng)
{
    return materialScatterT<-1>(S, m, ray, rayRi, hr, prng);
}

/* Material::scatterToDirection of the resolved material. Only materials that return
 * ScatterRandom from scatter() are ever asked. */
template<int MT>
WPT_DEV ScatterRec materialScatterToDirectionT(const DScene& S, const wpt_material* m, const Ray& ray,
        float4 rayRi, const HitRec& hr, float3 direction)
{
    ScatterRec sr;
    sr.type = SCATTER_RANDOM;
    sr.dir = direction;
    sr.att = f4(0.0f);
    sr.pdf = 0.0f;
    sr.ri = rayRi;
    if (!m)
        return scatterNone();
    switch (MT >= 0 ? MT : m->type) {
    case WPT_MAT_LAMBERTIAN: {      /* material_lambertian.hpp:86-102 */
        const float cosTheta = dot(normalAt(S, m, hr), direction);
        if (cosTheta > 0.0f) {
            sr.pdf = cosTheta * INV_PI;
            sr.att = colorAt(S, m, m->v0, m->tex[0], hr.tc) * sr.pdf;
        }
        return sr;
    }
    case WPT_MAT_GGX: {             /* material_ggx.hpp:227-257 */
        const float3 view = -ray.d;
        const float3 light = direction;
        const TangentSpace ts = tangentSpaceAt(S, m, hr);
        const float dotNL = dot(ts.normal, light);
        const float dotNV = dot(ts.normal, view);
        if (dotNL > 0.0f && dotNV > 0.0f) {
            const float2 rough = ggxRoughnessAt(S, m, hr.tc);
            const float3 tsV = toTangent(ts, view);
            const float3 tsL = toTangent(ts, light);
            const float3 tsH = normalize(tsV + tsL);
            const float dotVH = dot(tsV, tsH);
            if (dotVH > 0.0f) {
                sr.pdf = ggxDV(tsH, tsV, dotVH, rough) / (4.0f * dotVH);
                const float Dval = ggxD(tsH, rough);
                const float4 Fval = fresnelSchlick(ggxAlbedoAt(S, m, hr.tc), dotVH);
                const float Gval = ggxG2(tsV, tsL, rough);
                sr.att = ((Dval * Fval) * Gval) / (4.0f * dotNV);
            }
        }
        return sr;
    }
    case WPT_MAT_MODPHONG: {        /* material_modphong.hpp:310-327 */
        const float3 n = normalAt(S, m, hr);
        const float cosTheta = dot(n, direction);
        if (cosTheta > 0.0f) {
            const float4 kd = phongDiffuseAt(S, m, hr.tc);
            const float4 ks = phongSpecularAt(S, m, hr.tc);
            const float s = phongShininessAt(S, m, hr.tc);
            const float specProb = phongSpecularProbability(kd, ks);
            sr.att = phongAttenuation(n, -ray.d, direction, kd, ks, s, cosTheta);
            sr.pdf = phongPdf(n, -ray.d, direction, s, cosTheta, specProb);
        }
        return sr;
    }
    default:                        /* Material::scatterToDirection (material.hpp:177-183) */
        return scatterNone();
    }
}

static __device__ __noinline__ ScatterRec materialScatterToDirection(const DScene& S, const wpt_material* m, const Ray& ray,
        float4 rayRi, const HitRec& hr, float3 direction)
{
    return materialScatterToDirectionT<-1>(S, m, ray, rayRi, hr, direction);
}

/* Material::emitted of the resolved material */
WPT_DEV float4 materialEmitted(const DScene& S, const wpt_material* m, const Ray& ray, const HitRec& hr)
{
    float4 e = f4(0.0f);
    if (!m)
        return e;
    if (m->type == WPT_MAT_LIGHT_DIFFUSE) {         /* light_diffuse.hpp:50-61 */
        if (!hr.backside) {
            e = ld4(m->v0);
            if (m->tex[0] >= 0) {
                const float3 rgb = xyz(textureValue(S, m->tex[0], hr.tc));
                e = e * f4(rgb, average(rgb));
            }
        }
    } else if (m->type == WPT_MAT_LIGHT_SPOT) {     /* light_spot.hpp:45-60 */
        if (!hr.backside) {
            const float cosine = dot(hr.n, -ray.d);
            if (cosine >= m->f[0]) {
                e = ld4(m->v0);
                if (m->tex[0] >= 0) {
                    const float3 rgb = xyz(textureValue(S, m->tex[0], hr.tc));
                    e = e * f4(rgb, average(rgb));
                }
            }
        }
    } else if (m->type == WPT_MAT_LIGHT_TOF) {      /* light_tof.hpp:55-69 */
        if (!hr.backside) {
            const float cosine = dot(hr.n, -ray.d);
            if (cosine >= m->f[1]) {
                e.w = m->f[0];
                if (m->tex[0] >= 0)
                    e.w *= textureValue(S, m->tex[0], hr.tc).x;
            }
        }
    } else if (m->type == WPT_MAT_MODPHONG) {       /* material_modphong.hpp:183-190 */
        if (!hr.backside)
            e = colorAt(S, m, m->v3, m->tex[4], hr.tc);
    }
    return e;
}

/* ------------------------------------------------------- environment map */

/* EnvironmentMap::m (envmap.hpp:55-81) */
WPT_DEV float2 envSquareFromDirection(float3 d)
{
    const float lat = asinf(rclamp(d.y, -1.0f, 1.0f));
    const float lon = atan2f(-d.x, d.z);
    const float r = sinf(0.5f * (PI_2 - lat));
    float alpha = lon - PI_2;
    float u, v;
    if (alpha < -PI_4)
        alpha += 2.0f * PI;
    if (alpha < PI_4) {
        u = r;
        v = alpha * u / PI_4;
    } else if (alpha < PI_2 + PI_4) {
        v = r;
        u = -(alpha - PI_2) * v / PI_4;
    } else if (alpha < PI + PI_4) {
        u = -r;
        v = (alpha - PI) * u / PI_4;
    } else {
        v = -r;
        u = -(alpha - (PI + PI_2)) * v / PI_4;
    }
    return make_float2(0.5f * (u + 1.0f), 0.5f * (v + 1.0f));
}

/* EnvironmentMap::invm (envmap.hpp:84-109) */
WPT_DEV float3 envDirectionFromSquare(float2 uv)
{
    const float u = 2.0f * uv.x - 1.0f;
    const float v = 2.0f * uv.y - 1.0f;
    float r, alpha;
    if (u * u > v * v) {
        r = u;
        alpha = PI_4 * v / u;
    } else {
        r = v;
        if (fabsf(v) > 0.0f)
            alpha = PI_2 - PI_4 * u / v;
        else
            alpha = 0.0f;
    }
    const float lat = PI_2 - 2.0f * asinf(r);
    const float lon = alpha + PI_2;
    const float3 d = f3(-cosf(lat) * sinf(lon), sinf(lat), cosf(lat) * cosf(lon));
    return normalize(d);
}

/* EnvironmentMapEquiRect::L (envmap.hpp:231-247) / EnvironmentMapCube::L (:264-285) */
static __device__ __noinline__ float4 envRadiance(const DScene& S, float3 direction)
{
    if (S.envType == WPT_ENV_EQUIRECT) {
        float y = asinf(rclamp(direction.y, -1.0f, 1.0f));
        float x = atan2f(-direction.x, direction.z);
        if (S.envCompat == WPT_ENV_COMPAT_MITSUBA) {
            x -= PI;
            if (x < 0.0f)
                x += 2.0f * PI;
        }
        x *= 0.5f * INV_PI;
        y = y * INV_PI + 0.5f;
        return textureValue(S, S.envTex[0], make_float2(x, y));
    }
    const float ax = fabsf(direction.x), ay = fabsf(direction.y), az = fabsf(direction.z);
    int side;
    float u, v;
    if (ax > ay && ax > az) {
        u = 0.5f * (direction.z / -direction.x + 1.0f);
        v = 0.5f * (direction.y / ax + 1.0f);
        side = 0 + (signbit(direction.x) ? 1 : 0);
    } else if (ay > az) {
        u = 0.5f * (direction.x / ay + 1.0f);
        v = 0.5f * (direction.z / -direction.y + 1.0f);
        side = 2 + (signbit(direction.y) ? 1 : 0);
    } else {
        u = 0.5f * (direction.x / direction.z + 1.0f);
        v = 0.5f * (direction.y / az + 1.0f);
        side = 4 + (signbit(direction.z) ? 1 : 0);
    }
    const int32_t tex = (side == 0 ? S.envTex[0] : side == 1 ? S.envTex[1] : side == 2 ? S.envTex[2]
            : side == 3 ? S.envTex[3] : side == 4 ? S.envTex[4] : S.envTex[5]);
    return textureValue(S, tex, make_float2(u, v));
}

/* EnvironmentMap::p (envmap.hpp:170-183) */
WPT_DEV float envPdf(const DScene& S, float3 direction)
{
    const int N = S.envN;
    const float2 uv = envSquareFromDirection(direction);
    int x = int(uv.x * N);
    int y = int(uv.y * N);
    if (x >= N)
        x = N - 1;
    if (y >= N)
        y = N - 1;
    const float q = __ldg(S.envM + y * N + x);
    const float invBinSizeOnSphere = (float(N * N) * 0.25f) * INV_PI;
    return q * invBinSizeOnSphere;
}

/* EnvironmentMap::d (envmap.hpp:186-210) */
WPT_DEV float3 envSample(const DScene& S, Prng& prng)
{
    const int N = S.envN;
    const float r = in01(prng);
    int bin;
    if (S.envGuideBits > 0) {
        /* the first bin with Mcs >= r (the last bin if there is none), looked for between two entries of
         * the guide table: the bin the reference's bisection over all N*N bins ends at */
        const uint32_t j = uint32_t(r * float(1u << S.envGuideBits));
        int a = int(__ldg(S.envGuide + j));
        int b = int(__ldg(S.envGuide + j + 1));
        while (a < b) {
            const int c = (a + b) / 2;
            if (__ldg(S.envMcs + c) < r)
                a = c + 1;
            else
                b = c;
        }
        bin = a;
    } else {
        int a = 0;
        int b = N * N - 1;
        while (b > a + 1) {
            const int c = (a + b) / 2;
            if (__ldg(S.envMcs + c) < r)
                a = c;
            else
                b = c;
        }
        bin = (__ldg(S.envMcs + a) >= r ? a : b);
    }
    bin = __ldg(S.envMs + bin);
    const int x = bin % N;
    const int y = bin / N;
    const float u = (float(x) + in01(prng)) / float(N);
    const float v = (float(y) + in01(prng)) / float(N);
    return envDirectionFromSquare(make_float2(u, v));
}

/* ------------------------------------------------------------- hot spots */

/* Untransformed corner `k` of a hot-spot triangle and its instance (for Hitable::direction) */
template<int F>
WPT_DEV float hotSpotPdf(const DScene& S, uint32_t ref, float3 origin, float3 direction, float time, const Ctx<F>& ctx)
{
    const uint32_t type = ref >> REF_TYPE_SHIFT;
    const uint32_t slot = ref & REF_SLOT_MASK;
    Ray ray;
    ray.o = origin;
    ray.d = direction;
    ray.time = time;
    if (type == PRIM_TRIANGLE) {        /* HitableTriangle::pdfValue (hitable_triangle.hpp:405-423) */
        float3 v0, v1, v2;
        uint32_t inst, tri, anim;
        triangleCorners<F>(S, slot, ctx, v0, v1, v2, inst, tri, anim);
        const RayHelper h = makeRayHelper(ray);
        float a, U, V, W, det;
        if (!triangleTest(v0, v1, v2, ray, h, 0.0f, MAXVAL, true, a, U, V, W, det))
            return 0.0f;
        const float3 edgeCross = cross(v1 - v0, v2 - v0);
        const float edgeCrossLength = sqrtf(dot(edgeCross, edgeCross));
        const float3 faceNormal = edgeCross / edgeCrossLength;
        const float faceArea = 0.5f * edgeCrossLength;
        const float cosine = fabsf(dot(faceNormal, -direction));
        return (a * a) / (cosine * faceArea);
    }
    if (type == PRIM_SPHERE) {          /* HitableSphere::pdfValue (hitable_sphere.hpp:155-190) */
        const float4* p = S.sph + size_t(slot) * 3;
        const float4 s0 = __ldg(p + 0), s2 = __ldg(p + 2);
        float3 center = xyz(s0);
        float radius = s0.w;
        /* the hit() below animates with `center += translation` while this function uses T * center;
         * the reference has the same inconsistency */
        float3 hitCenter = center;
        float hitRadius = radius;
        if (F & F_ANIM) {
            const uint32_t anim = __float_as_uint(s2.y);
            if (anim != 0) {
                const Trs T = animTrs<F>(S, ctx, anim);
                const float scale = rmax(rmax(T.s.x, T.s.y), T.s.z);
                hitCenter = center + T.t;
                hitRadius = radius * scale;
                center = trsApply(T, center);
                radius = radius * scale;
            }
        }
        const float3 cmo = center - origin;
        const float distanceSquared = dot(cmo, cmo);
        const float radiusSquared = radius * radius;
        if (distanceSquared <= radiusSquared)
            return 0.25f * INV_PI;
        float a;
        if (!sphereTest(hitCenter, hitRadius, ray, 0.0f, MAXVAL, a))
            return 0.0f;
        const float discriminant = 1.0f - radiusSquared / distanceSquared;
        const float cosThetaMax = (discriminant > 0.0f ? sqrtf(discriminant) : 0.0f);
        const float solidAngle = (2.0f * PI) * (1.0f - cosThetaMax);
        return 1.0f / solidAngle;
    }
    return 0.0f;                        /* Hitable::pdfValue default (hitable.hpp:150-154) */
}

template<int F>
WPT_DEV float hotSpotsPdf(const DScene& S, float3 origin, float3 direction, float time, const Ctx<F>& ctx)
{
    float pdf = 0.0f;
    for (uint32_t i = 0; i < S.nHot; i++)
        pdf += hotSpotPdf<F>(S, __ldg(S.hotRefs + i), origin, direction, time, ctx);
    return pdf * (1.0f / float(S.nHot));    /* invHotSpotsSize (wurblpt.hpp:309) */
}

template<int F>
WPT_DEV float3 hotSpotDirection(const DScene& S, uint32_t ref, float3 origin, Ctx<F>& ctx)
{
    const uint32_t type = ref >> REF_TYPE_SHIFT;
    const uint32_t slot = ref & REF_SLOT_MASK;
    if (type == PRIM_TRIANGLE) {        /* HitableTriangle::direction (hitable_triangle.hpp:425-443) */
        const float4* tp = S.tri + size_t(slot) * 3;
        const uint32_t instIdx = __float_as_uint(__ldg(tp + 0).w);
        const uint32_t triIdx = __float_as_uint(__ldg(tp + 1).w);
        const uint32_t anim = __float_as_uint(__ldg(tp + 2).w);
        const wpt_instance* inst = S.instances + instIdx;
        const wpt_mesh* mesh = S.meshes + inst->mesh;
        const uint32_t* ix = S.indices + mesh->index_offset + size_t(triIdx) * 3;
        const float* base = S.vertices + mesh->vertex_offset;
        const float* p0 = base + size_t(ix[0]) * mesh->vertex_stride;
        const float* p1 = base + size_t(ix[1]) * mesh->vertex_stride;
        const float* p2 = base + size_t(ix[2]) * mesh->vertex_stride;
        const float3 bary = inTriangle(in01x2(ctx.prng));
        float3 p = (bary.x * f3(p0[0], p0[1], p0[2]) + bary.y * f3(p1[0], p1[1], p1[2])) + bary.z * f3(p2[0], p2[1], p2[2]);
        if (inst->has_transform)
            p = transformPoint(inst->M, 4, p);
        if ((F & F_ANIM) && anim != 0)
            p = transformPoint(animMatrix<F>(S, ctx, anim), 3, p);
        return normalize(p - origin);
    }
    if (type == PRIM_SPHERE) {          /* HitableSphere::direction (hitable_sphere.hpp:192-219) */
        const SphereGeom g = sphereGeom<F>(S, slot, ctx);
        const float3 cmo = g.center - origin;
        const float distanceSquared = dot(cmo, cmo);
        const float radiusSquared = g.radius * g.radius;
        if (distanceSquared <= radiusSquared)
            return onUnitSphere(in01x2(ctx.prng));
        const float discriminant = 1.0f - radiusSquared / distanceSquared;
        const float cosThetaMax = (discriminant > 0.0f ? sqrtf(discriminant) : 0.0f);
        return toSphere(normalize(cmo), cosThetaMax, in01x2(ctx.prng));
    }
    return f3(0.0f);                    /* Hitable::direction default (hitable.hpp:156-161) */
}

/* ------------------------------------------------------------------ camera */

/* LensDistortion::distort (optics.hpp:226-246) */
WPT_DEV void lensDistort(const RenderParams& P, float& p, float& q)
{
    if (P.distortionType != WPT_DISTORTION_NONE) {
        const float s = (p - P.distCenter.x) * P.distInvFocal.x;
        const float t = (q - P.distCenter.y) * P.distInvFocal.y;
        const float r2 = s * s + t * t;
        const float r4 = r2 * r2;
        const float r6 = r4 * r2;
        const float rd = 1.0f + P.k1 * r2 + P.k2 * r4 + P.k3 * r6;
        const float term1 = 2.0f * s * t;
        const float term2 = r2 + 2.0f * s * s;
        const float term3 = r2 + 2.0f * t * t;
        const float ns = s * rd + P.p1 * term1 + P.p2 * term2;
        const float nt = t * rd + P.p1 * term3 + P.p2 * term1;
        p = ns * P.distFocal.x + P.distCenter.x;
        q = nt * P.distFocal.y + P.distCenter.y;
    }
}

/* LensDistortion::undistort (optics.hpp:249-309) */
static __device__ __noinline__ void lensUndistort(const RenderParams& P, float& p, float& q)
{
    if (P.distortionType == WPT_DISTORTION_RADIAL_AND_PLANAR) {
        const float s = (p - P.distCenter.x) * P.distInvFocal.x;
        const float t = (q - P.distCenter.y) * P.distInvFocal.y;
        const float r2 = s * s + t * t;
        const float r4 = r2 * r2;
        const float d1 = P.k1 * r2 + P.k2 * r4;
        const float d2 = 1.0f / (4.0f * P.k1 * r2 + 6.0f * P.k2 * r4 + 8.0f * P.p1 * t + 8.0f * P.p2 * s + 1.0f);
        p = (s - d2 * (d1 * s + 2.0f * P.p1 * s * t + P.p2 * (r2 + 2.0f * s * s))) * P.distFocal.x + P.distCenter.x;
        q = (t - d2 * (d1 * t + P.p1 * (r2 + 2.0f * t * t) + 2.0f * P.p2 * s * t)) * P.distFocal.y + P.distCenter.y;
    } else if (P.distortionType == WPT_DISTORTION_RADIAL_ONLY) {
        const float s = (p - P.distCenter.x) * P.distInvFocal.x;
        const float t = (q - P.distCenter.y) * P.distInvFocal.y;
        const float r2 = s * s + t * t;
        const float r4 = r2 * r2;
        const float r6 = r4 * r2;
        const float r8 = r4 * r4;
        const float d = 1.0f + P.b1 * r2 + P.b2 * r4 + P.b3 * r6 + P.b4 * r8;
        p = s * d * P.distFocal.x + P.distCenter.x;
        q = t * d * P.distFocal.y + P.distCenter.y;
    } else if (P.distortionType == WPT_DISTORTION_OPENCV) {
        float s = (p - P.distCenter.x) * P.distInvFocal.x;
        float t = (q - P.distCenter.y) * P.distInvFocal.y;
        const float s0 = s, t0 = t;
        float squaredError = MAXVAL;
        for (int i = 0; i < 256 && squaredError >= 0.001f * 0.001f; i++) {
            const float r2 = s * s + t * t;
            const float r4 = r2 * r2;
            const float r6 = r4 * r2;
            const float rd = 1.0f + P.k1 * r2 + P.k2 * r4 + P.k3 * r6;
            const float invrd = 1.0f / rd;
            const float ds = 2.0f * P.p1 * s * t + P.p2 * (r2 + 2.0f * s * s);
            const float dt = 2.0f * P.p2 * s * t + P.p1 * (r2 + 2.0f * t * t);
            s = (s0 - ds) * invrd;
            t = (t0 - dt) * invrd;
            float ep = s * P.distFocal.x + P.distCenter.x;
            float eq = t * P.distFocal.y + P.distCenter.y;
            lensDistort(P, ep, eq);
            const float ox = (ep - p) * float(P.width);
            const float oy = (eq - q) * float(P.height);
            squaredError = (0.0f + ox * ox) + oy * oy;
        }
        p = s * P.distFocal.x + P.distCenter.x;
        q = t * P.distFocal.y + P.distCenter.y;
    }
}

/* Camera::getRay (camera.hpp:123-185) */
WPT_DEV Ray cameraRay(const RenderParams& P, float p, float q, Prng& prng, bool withRandomness)
{
    float stereoShift = 0.0f;
    if (P.stereo > 0.0f) {
        q *= 2.0f;
        if (q < 1.0f) {
            stereoShift = -0.5f * P.stereo;
        } else {
            q -= 1.0f;
            stereoShift = 0.5f * P.stereo;
        }
    }
    float3 O, D;
    if (P.surround == WPT_SURROUND_OFF) {
        if (P.distortionType != WPT_DISTORTION_NONE)
            lensUndistort(P, p, q);
        float3 Pt = f3(rmix(P.l, P.r, p), rmix(P.b, P.t, q), -1.0f);
        O = f3(0.0f);
        if (withRandomness && P.lensRadius > 0.0f) {    /* LensDepthOfField::offset (optics.hpp:326-334) */
            Pt = Pt * P.focusDist;
            const float2 dsk = inUnitDisk(in01x2(prng));
            O = f3(P.lensRadius * dsk.x, P.lensRadius * dsk.y, 0.0f);
        }
        D = Pt - O;
        O = O + f3(stereoShift, 0.0f, 0.0f);
    } else {
        float lon = (2.0f * p - 1.0f) * PI;
        if (P.surround == WPT_SURROUND_180)
            lon *= 0.5f;
        const float lat = (q - 0.5f) * PI;
        D = f3(cosf(lat) * sinf(lon), sinf(lat), -cosf(lat) * cosf(lon));
        O = f3(-cosf(lon), 0.0f, -sinf(lon)) * stereoShift;
    }
    float t = P.t0;
    Trs T;
    T.t = P.camT;
    T.r = P.camR;
    T.s = P.camS;
    if (P.t0 != P.t1 && withRandomness) {
        t += in01(prng) * (P.t1 - P.t0);
        if (P.nCamKeys > 0)
            T = keyframesAt(P.camKeys, P.nCamKeys, t);      /* Camera::at(t) (camera.hpp:109-112) */
    }
    Ray ray;
    ray.o = trsApply(T, O);
    ray.d = normalize(qrotate(T.r, D));
    ray.time = t;
    return ray;
}

/* ------------------------------------------------------------------ sensor */

/* Sensor::accumulateRadiance: SensorRGB (sensor_rgb.hpp:63-80), SensorTofAmcw (sensor_tof_amcw.hpp:227-252).
 * `lightIsTof` is hitRecord.haveHit && material()->isTofLight(hitRecord) of the hit that carried the radiance. */
template<int F>
WPT_DEV void accumulate(const RenderParams& P, float* acc, float4 opl, float distanceToLight, float4 radiance, bool lightIsTof)
{
    if (!(F & F_OPL)) {
        acc[0] += radiance.x;
        acc[1] += radiance.y;
        acc[2] += radiance.z;
        return;
    }
    if (P.sensorType == WPT_SENSOR_RGB) {
        const bool distOk = distanceToLight >= P.minDistToLight && distanceToLight <= P.maxDistToLight;
        if (distOk && opl.x >= P.minPathLen && opl.x <= P.maxPathLen) acc[0] += radiance.x;
        if (distOk && opl.y >= P.minPathLen && opl.y <= P.maxPathLen) acc[1] += radiance.y;
        if (distOk && opl.z >= P.minPathLen && opl.z <= P.maxPathLen) acc[2] += radiance.z;
    } else {
        const float irradiance = radiance.w * 1000.0f;
        const float power = irradiance * P.tofPixelArea;
        const float energy = power * 0.5f * P.tofExposureTime;      /* dutyCycle = 0.5 (sensor_tof_amcw.hpp:62) */
        const float halfEnergy = 0.5f * energy;
        float t = 0.0f;
        if (lightIsTof) {
            const float phaseShift = 2.0f * PI * opl.w * P.tofFracModfreqC;
            t = P.tofContrast * cosf(P.tofTau + phaseShift);
        }
        acc[0] += halfEnergy * (1.0f + t);
        acc[1] += halfEnergy * (1.0f - t);
        acc[2] += energy;
    }
}

}
