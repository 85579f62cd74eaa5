/*
 * device_math.cuh -- FP32 vector arithmetic for the device path, written so that every
 * expression evaluates in the same order as the reference's gvm.hpp (GLSL-like vectors):
 * the translation unit is compiled with --fmad=false and IEEE division / square root, so
 * what is written here is what executes. Bit-exact primary-ray parity depends on this.
 *
 *   dot       gvm.hpp:1183-1189   d = 0; d += a[i]*b[i] left to right
 *   normalize gvm.hpp:1201-1204   v / sqrt(dot(v,v))
 *   reflect   gvm.hpp:1213-1216   i - 2*dot(n,i)*n
 *   refract   gvm.hpp:1218-1223
 *   cross     gvm.hpp:1225-1231
 *   min/max   gvm.hpp:90-96       (x < y ? x : y), (x > y ? x : y)  -- NOT fminf/fmaxf (NaN rules differ)
 *   mix       gvm.hpp:166         x + alpha*(y - x)
 *   quat*vec  gvm.hpp:1713-1720
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define WPT_DEV __device__ __forceinline__

namespace wptd {

constexpr float PI = 3.14159274f;           /* float(pi), gvm.hpp:64-65 */
constexpr float PI_2 = 1.57079637f;
constexpr float PI_4 = 0.785398185f;
constexpr float INV_PI = 0.318309873f;
constexpr float MAXVAL = 3.40282347e+38f;   /* gvm.hpp:77-78 */
constexpr float EPSILON = 1.1920929e-07f;   /* gvm.hpp:81-82 */
constexpr float LDBL_EPS_F = 1.08420217e-19f;   /* float(epsilon_v<long double>) = 2^-63 on x86-64 */
constexpr float COS_ORTHO_TOL = 0.0003f;    /* constants.hpp:40 */

WPT_DEV float rmin(float x, float y) { return x < y ? x : y; }
WPT_DEV float rmax(float x, float y) { return x > y ? x : y; }
WPT_DEV float rclamp(float x, float lo, float hi) { return rmin(hi, rmax(lo, x)); }     /* gvm.hpp:98 */
WPT_DEV float rmix(float x, float y, float a) { return x + a * (y - x); }
WPT_DEV float fractf(float x) { return x - floorf(x); }                                 /* gvm.hpp:160 */

/* ---- float3 ---- */
WPT_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
WPT_DEV float3 f3(float s) { return make_float3(s, s, s); }
WPT_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
WPT_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
WPT_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
WPT_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
WPT_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
WPT_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
WPT_DEV float3 operator/(float s, float3 a) { return f3(s / a.x, s / a.y, s / a.z); }
WPT_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
WPT_DEV float dot(float3 a, float3 b) { return ((0.0f + a.x * b.x) + a.y * b.y) + a.z * b.z; }
WPT_DEV float3 normalize(float3 v) { return v / sqrtf(dot(v, v)); }
WPT_DEV float3 cross(float3 v, float3 w)
{
    return f3(v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x);
}
WPT_DEV float3 reflect(float3 i, float3 n) { return i - (2.0f * dot(n, i)) * n; }
WPT_DEV float3 refract(float3 i, float3 n, float eta)
{
    const float d = dot(n, i);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k <= 0.0f)
        return f3(0.0f);
    return i * eta - n * (eta * d + sqrtf(k));
}
WPT_DEV float3 vmin(float3 a, float3 b) { return f3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z)); }
WPT_DEV float3 vmax(float3 a, float3 b) { return f3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z)); }
WPT_DEV float3 vmix(float3 a, float3 b, float t) { return f3(rmix(a.x, b.x, t), rmix(a.y, b.y, t), rmix(a.z, b.z, t)); }
/* component k (0..2) of v. Written as two selp: left to itself the compiler turns the ternaries into divergent branches
 * (six instructions and a reconvergence point per use, nine uses per triangle test). */
WPT_DEV float sel(float3 v, int k)
{
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\t"
        "setp.eq.s32 p0, %4, 0;\n\t"
        "setp.eq.s32 p1, %4, 1;\n\t"
        "selp.f32 %0, %2, %3, p1;\n\t"
        "selp.f32 %0, %1, %0, p0;\n\t}"
        : "=f"(r) : "f"(v.x), "f"(v.y), "f"(v.z), "r"(k));
    return r;
}
WPT_DEV float average(float3 v) { return (1.0f / 3.0f) * (((0.0f + v.x) + v.y) + v.z); }  /* gvm.hpp:1293-1300 */

/* ---- float4 ---- */
WPT_DEV float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
WPT_DEV float4 f4(float s) { return make_float4(s, s, s, s); }
WPT_DEV float4 f4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
WPT_DEV float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }
WPT_DEV float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
WPT_DEV float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
WPT_DEV float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
WPT_DEV float4 operator*(float s, float4 a) { return f4(s * a.x, s * a.y, s * a.z, s * a.w); }
WPT_DEV float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
WPT_DEV float4 operator/(float4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
WPT_DEV float4 operator-(float4 a) { return f4(-a.x, -a.y, -a.z, -a.w); }
WPT_DEV float4 vmix(float4 a, float4 b, float t)
{
    return f4(rmix(a.x, b.x, t), rmix(a.y, b.y, t), rmix(a.z, b.z, t), rmix(a.w, b.w, t));
}
/* max(vec4): r = a[0]; if (a[i] > r) r = a[i] (gvm.hpp:1284-1291) */
WPT_DEV float vmaxcomp(float4 v)
{
    float r = v.x;
    if (v.y > r) r = v.y;
    if (v.z > r) r = v.z;
    if (v.w > r) r = v.w;
    return r;
}
WPT_DEV float sel(float4 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
/* component k (0..3) of v with three selects (written as selp so that no branches are generated) */
WPT_DEV int sel4i(int4 v, uint32_t k)
{
    int r;
    asm("{\n\t.reg .pred p0, p1;\n\t.reg .b32 lo, hi;\n\t"
        "and.b32 lo, %5, 1;\n\t"
        "setp.ne.u32 p0, lo, 0;\n\t"
        "and.b32 hi, %5, 2;\n\t"
        "setp.ne.u32 p1, hi, 0;\n\t"
        "selp.b32 lo, %2, %1, p0;\n\t"
        "selp.b32 hi, %4, %3, p0;\n\t"
        "selp.b32 %0, hi, lo, p1;\n\t}"
        : "=r"(r) : "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(k));
    return r;
}

/* ---- quaternion (x, y, z, w) ---- */
WPT_DEV float3 qrotate(float4 q, float3 v)
{
    float3 s = f3(q.x, q.y, q.z);
    float3 t = 2.0f * cross(s, v);
    return (v + q.w * t) + cross(s, t);
}
/* quaternion product, gvm.hpp:1682-1690 */
WPT_DEV float4 qmul(float4 a, float4 q)
{
    float4 p;
    p.x = a.w * q.x + a.x * q.w + a.y * q.z - a.z * q.y;
    p.y = a.w * q.y + a.y * q.w + a.z * q.x - a.x * q.z;
    p.z = a.w * q.z + a.z * q.w + a.x * q.y - a.y * q.x;
    p.w = a.w * q.w - a.x * q.x - a.y * q.y - a.z * q.z;
    return p;
}
/* slerp, gvm.hpp:1765-1796 */
WPT_DEV float4 qslerp(float4 q, float4 r, float alpha)
{
    float4 w = r;
    float c = q.x * r.x + q.y * r.y + q.z * r.z + q.w * r.w;
    if (c < 0.0f) {
        w = -w;
        c = -c;
    }
    float tq, tw;
    if (c >= 1.0f) {
        tq = 1.0f;
        tw = 0.0f;
    } else {
        float halfAngle = acosf(c);
        float s = sqrtf(1.0f - c * c);
        if (fabsf(s) < EPSILON) {
            tq = 0.5f;
            tw = 0.5f;
        } else {
            tq = sinf((1.0f - alpha) * halfAngle) / s;
            tw = sinf(alpha * halfAngle) / s;
        }
    }
    return f4(q.x * tq + w.x * tw, q.y * tq + w.y * tw, q.z * tq + w.z * tw, q.w * tq + w.w * tw);
}

/* ---- Transformation (translation, rotation, scaling), transformation.hpp:47-122 ---- */
struct Trs {
    float3 t;
    float4 r;
    float3 s;
};
WPT_DEV float3 trsApply(const Trs& T, float3 v) { return T.t + qrotate(T.r, v * T.s); }
WPT_DEV Trs trsMix(const Trs& a, const Trs& b, float alpha)
{
    Trs r;
    r.t = vmix(a.t, b.t, alpha);
    r.r = qslerp(a.r, b.r, alpha);
    r.s = vmix(a.s, b.s, alpha);
    return r;
}

/* 3 rows x 4 columns of a column-major mat4: m[c*3 + r] */
struct Mat34 { float m[12]; };
struct Mat33 { float m[9]; };

/* toMat3(quat), gvm.hpp:1979-2001 */
WPT_DEV Mat33 toMat33(float4 q)
{
    Mat33 M;
    float xx = q.x * q.x, xy = q.x * q.y, xz = q.x * q.z, xw = q.x * q.w;
    float yy = q.y * q.y, yz = q.y * q.z, yw = q.y * q.w;
    float zz = q.z * q.z, zw = q.z * q.w;
    M.m[0] = 1.0f - 2.0f * (yy + zz);
    M.m[1] = 2.0f * (xy + zw);
    M.m[2] = 2.0f * (xz - yw);
    M.m[3] = 2.0f * (xy - zw);
    M.m[4] = 1.0f - 2.0f * (xx + zz);
    M.m[5] = 2.0f * (yz + xw);
    M.m[6] = 2.0f * (xz + yw);
    M.m[7] = 2.0f * (yz - xw);
    M.m[8] = 1.0f - 2.0f * (xx + yy);
    return M;
}
/* Transformation::toMat4 (transformation.hpp:105-113); see host_math.hpp for why this is exact */
WPT_DEV Mat34 toMat34(const Trs& T, const Mat33& R)
{
    Mat34 M;
    M.m[0] = R.m[0] * T.s.x; M.m[1] = R.m[1] * T.s.x; M.m[2] = R.m[2] * T.s.x;
    M.m[3] = R.m[3] * T.s.y; M.m[4] = R.m[4] * T.s.y; M.m[5] = R.m[5] * T.s.y;
    M.m[6] = R.m[6] * T.s.z; M.m[7] = R.m[7] * T.s.z; M.m[8] = R.m[8] * T.s.z;
    M.m[9] = T.t.x; M.m[10] = T.t.y; M.m[11] = T.t.z;
    return M;
}
/* (M * vec4(v,1)).xyz, gvm.hpp:1481-1491 */
WPT_DEV float3 transformPoint(const float* m, int cs, float3 v)
{
    return f3((((0.0f + m[0 * cs + 0] * v.x) + m[1 * cs + 0] * v.y) + m[2 * cs + 0] * v.z) + m[3 * cs + 0] * 1.0f,
              (((0.0f + m[0 * cs + 1] * v.x) + m[1 * cs + 1] * v.y) + m[2 * cs + 1] * v.z) + m[3 * cs + 1] * 1.0f,
              (((0.0f + m[0 * cs + 2] * v.x) + m[1 * cs + 2] * v.y) + m[2 * cs + 2] * v.z) + m[3 * cs + 2] * 1.0f);
}
/* mat3 * vec3 */
WPT_DEV float3 mul33(const float* m, float3 v)
{
    return f3(((0.0f + m[0] * v.x) + m[3] * v.y) + m[6] * v.z,
              ((0.0f + m[1] * v.x) + m[4] * v.y) + m[7] * v.z,
              ((0.0f + m[2] * v.x) + m[5] * v.y) + m[8] * v.z);
}

/* ---- TangentSpace (tangentspace.hpp:43-136) ---- */
struct TangentSpace {
    float3 normal, tangent, bitangent;
};
/* from a normal only: Duff et al. (tangentspace.hpp:57-74) */
WPT_DEV TangentSpace tangentSpace(float3 n)
{
    TangentSpace ts;
    ts.normal = n;
    float sign = copysignf(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    ts.tangent = f3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    ts.bitangent = f3(b, sign + n.y * n.y * a, -n.y);
    return ts;
}
WPT_DEV TangentSpace tangentSpace(float3 n, float3 t)
{
    TangentSpace ts;
    ts.normal = n;
    ts.tangent = t;
    ts.bitangent = cross(n, t);
    return ts;
}
/* matrixFromTangentSpace() * v = mat3(tangent, bitangent, normal) * v (tangentspace.hpp:122-135) */
WPT_DEV float3 toWorld(const TangentSpace& ts, float3 v)
{
    return f3(((0.0f + ts.tangent.x * v.x) + ts.bitangent.x * v.y) + ts.normal.x * v.z,
              ((0.0f + ts.tangent.y * v.x) + ts.bitangent.y * v.y) + ts.normal.y * v.z,
              ((0.0f + ts.tangent.z * v.x) + ts.bitangent.z * v.y) + ts.normal.z * v.z);
}
/* matrixToTangentSpace() * v (tangentspace.hpp:99-119): rows are tangent, bitangent, normal */
WPT_DEV float3 toTangent(const TangentSpace& ts, float3 v)
{
    return f3(((0.0f + ts.tangent.x * v.x) + ts.tangent.y * v.y) + ts.tangent.z * v.z,
              ((0.0f + ts.bitangent.x * v.x) + ts.bitangent.y * v.y) + ts.bitangent.z * v.z,
              ((0.0f + ts.normal.x * v.x) + ts.normal.y * v.y) + ts.normal.z * v.z);
}

/* ---- Prng (prng.hpp:47-101): xoshiro128+ seeded through splitmix64 ---- */
struct Prng {
    uint32_t s0, s1, s2, s3;
};
WPT_DEV uint64_t splitmix64(uint64_t x)
{
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
WPT_DEV Prng prngInit(uint32_t pixelIndex)
{
    uint64_t seed = pixelIndex;
    seed += 42;
    uint64_t s01 = splitmix64(seed);
    uint64_t s23 = splitmix64(s01);
    Prng p;
    p.s0 = uint32_t(s01 >> 32);
    p.s1 = uint32_t(s01 & 0xffffffffull);
    p.s2 = uint32_t(s23 >> 32);
    p.s3 = uint32_t(s23 & 0xffffffffull);
    return p;
}
WPT_DEV uint32_t prngNext(Prng& p)
{
    const uint32_t result = p.s0 + p.s3;
    const uint32_t t = p.s1 << 9;
    p.s2 ^= p.s0;
    p.s3 ^= p.s1;
    p.s1 ^= p.s2;
    p.s0 ^= p.s3;
    p.s2 ^= t;
    p.s3 = (p.s3 << 11) | (p.s3 >> 21);
    return result;
}
WPT_DEV float in01(Prng& p) { return float(prngNext(p) >> 8) * 5.9604644775390625e-08f; }
/* Sub-streams: when a pixel's samples are split over several work items, item `sub` of the pixel
 * needs its own stream. sub == 0 is exactly the reference's Prng(pixel); sub > 0 seeds the same
 * splitmix64 -> xoshiro128+ construction with (pixel + 42) + (sub << 32), which no other
 * (pixel, sub) pair produces (pixel indices are below 2^31). (The generator's jump() function would
 * also do, but costs ~1500 instructions per item start, executed by one lane at a time.) */
WPT_DEV Prng prngInitSub(uint32_t pixelIndex, uint32_t sub)
{
    uint64_t seed = pixelIndex;
    seed += 42;
    seed += uint64_t(sub) << 32;
    uint64_t s01 = splitmix64(seed);
    uint64_t s23 = splitmix64(s01);
    Prng p;
    p.s0 = uint32_t(s01 >> 32);
    p.s1 = uint32_t(s01 & 0xffffffffull);
    p.s2 = uint32_t(s23 >> 32);
    p.s3 = uint32_t(s23 & 0xffffffffull);
    return p;
}
/* in01x2() returns vec2(in01(), in01()) (prng.hpp:97-100). The oracle is built with g++, which
 * evaluates those two calls right to left: the FIRST draw lands in .y.
 * The same holds for the pixel jitter (wurblpt.hpp:355). Mirrored here so that a lane with
 * substreams == 1 consumes its stream exactly like the oracle does. */
WPT_DEV float2 in01x2(Prng& p)
{
    float y = in01(p);
    float x = in01(p);
    return make_float2(x, y);
}

/* ---- Sampler (sampler.hpp:36-121) ---- */
WPT_DEV float2 inUnitDisk(float2 u)
{
    float ox = 2.0f * u.x - 1.0f;
    float oy = 2.0f * u.y - 1.0f;
    if (ox == 0.0f && oy == 0.0f)
        return make_float2(0.0f, 0.0f);
    float theta, r;
    if (fabsf(ox) > fabsf(oy)) {
        r = ox;
        theta = PI_4 * (oy / ox);
    } else {
        r = oy;
        theta = PI_2 - PI_4 * (ox / oy);
    }
    return make_float2(r * cosf(theta), r * sinf(theta));
}
WPT_DEV float3 onUnitSphere(float2 u)
{
    float z = 1.0f - 2.0f * u.x;
    float r = sqrtf(rmax(0.0f, 1.0f - z * z));
    float phi = 2.0f * PI * u.y;
    return f3(r * cosf(phi), r * sinf(phi), z);
}
WPT_DEV float3 inTriangle(float2 u)
{
    float su0 = sqrtf(u.x);
    float b0 = 1.0f - su0;
    float b1 = u.y * su0;
    return f3(b0, b1, 1.0f - b0 - b1);
}
WPT_DEV float3 cosineDirection(float2 u)
{
    float2 d = inUnitDisk(u);
    float dd = (0.0f + d.x * d.x) + d.y * d.y;
    float z = sqrtf(rmax(0.0f, 1.0f - dd));
    return f3(d.x, d.y, z);
}
WPT_DEV float3 toSphere(float3 direction, float cosThetaMax, float2 u)
{
    float cosTheta = (1.0f - u.x) + u.x * cosThetaMax;
    float sinTheta = sqrtf(rmax(0.0f, 1.0f - cosTheta * cosTheta));
    float phi = u.y * 2.0f * PI;
    float3 v = f3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
    return normalize(toWorld(tangentSpace(direction), v));
}

/* ---- Fresnel (fresnel.hpp:32-72) ---- */
WPT_DEV float4 fresnelSchlick(float4 r0, float cosTheta)
{
    float t = 1.0f - cosTheta;
    float t2 = t * t;
    /* r0 + (vec4(1) - r0) * t_squared * t_squared * t evaluates as (((1-r0)*t2)*t2)*t */
    return f4(r0.x + (((1.0f - r0.x) * t2) * t2) * t, r0.y + (((1.0f - r0.y) * t2) * t2) * t,
              r0.z + (((1.0f - r0.z) * t2) * t2) * t, r0.w + (((1.0f - r0.w) * t2) * t2) * t);
}
WPT_DEV float fresnelUnpolarized(float cosI, float cosT, float n1, float n2)
{
    float Fs = (n1 * cosI - n2 * cosT) / (n1 * cosI + n2 * cosT);
    Fs *= Fs;
    float Fp = (n1 * cosT - n2 * cosI) / (n1 * cosT + n2 * cosI);
    Fp *= Fp;
    return 0.5f * (Fs + Fp);
}

/* powerHeuristicWeight (wurblpt.hpp:101-106) */
WPT_DEV float powerHeuristic(float f, float g)
{
    f *= f;
    g *= g;
    return (f + g > 0.0f ? f / (f + g) : 0.0f);
}

}
