/*
 * host_math.hpp -- the little FP32 vector / quaternion / TRS arithmetic the host side needs to
 * bound hitables (Scene::updateBVH). Every function states the reference expression it must
 * agree with bit for bit; this translation unit is compiled with -ffp-contract=off so the
 * operation order written here is the operation order executed.
 */
#pragma once

#include <cmath>
#include <cstdint>

namespace wpt {

struct V3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
};

inline V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }

/* gvm.hpp min/max are (x < y ? x : y) / (x > y ? x : y) (gvm.hpp:90-96) */
inline float fminr(float x, float y) { return x < y ? x : y; }
inline float fmaxr(float x, float y) { return x > y ? x : y; }
inline V3 vmin(V3 a, V3 b) { return v3(fminr(a.x, b.x), fminr(a.y, b.y), fminr(a.z, b.z)); }
inline V3 vmax(V3 a, V3 b) { return v3(fmaxr(a.x, b.x), fmaxr(a.y, b.y), fmaxr(a.z, b.z)); }
inline V3 vmin3(V3 a, V3 b, V3 c) { return vmin(vmin(a, b), c); }   /* gvm.hpp:92 */
inline V3 vmax3(V3 a, V3 b, V3 c) { return vmax(vmax(a, b), c); }

/* gvm.hpp:1225-1231 */
inline V3 cross(V3 v, V3 w)
{
    return v3(v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x);
}

struct Quat { float x, y, z, w; };

/* quaternion * vector, gvm.hpp:1713-1720: v + q.w * t + cross(s, t), t = 2 * cross(s, v) */
inline V3 rotate(Quat q, V3 v)
{
    V3 s = v3(q.x, q.y, q.z);
    V3 t = 2.0f * cross(s, v);
    return (v + q.w * t) + cross(s, t);
}

/* gvm.hpp:1765-1796 */
inline Quat slerp(Quat q, Quat r, float alpha)
{
    Quat w = r;
    float cosHalfAngle = q.x * r.x + q.y * r.y + q.z * r.z + q.w * r.w;
    if (cosHalfAngle < 0.0f) {
        w.x = -w.x; w.y = -w.y; w.z = -w.z; w.w = -w.w;
        cosHalfAngle = -cosHalfAngle;
    }
    float tq, tw;
    if (cosHalfAngle >= 1.0f) {
        tq = 1.0f;
        tw = 0.0f;
    } else {
        float halfAngle = std::acos(cosHalfAngle);
        float sinHalfAngle = std::sqrt(1.0f - cosHalfAngle * cosHalfAngle);
        if (std::fabs(sinHalfAngle) < 1.1920929e-07f) {
            tq = 0.5f;
            tw = 0.5f;
        } else {
            tq = std::sin((1.0f - alpha) * halfAngle) / sinHalfAngle;
            tw = std::sin(alpha * halfAngle) / sinHalfAngle;
        }
    }
    return Quat{ q.x * tq + w.x * tw, q.y * tq + w.y * tw, q.z * tq + w.z * tw, q.w * tq + w.w * tw };
}

/* Transformation (transformation.hpp:47-122) */
struct Trs {
    V3 t;
    Quat r;
    V3 s;
};

inline bool operator==(const Quat& a, const Quat& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
inline bool operator==(const V3& a, const V3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator==(const Trs& a, const Trs& b) { return a.t == b.t && a.r == b.r && a.s == b.s; }

/* transformation.hpp:83-86: translation + (rotation * (v * scaling)) */
inline V3 apply(const Trs& T, V3 v) { return T.t + rotate(T.r, v * T.s); }

inline float mixf(float x, float y, float a) { return x + a * (y - x); }   /* gvm.hpp:166 */
inline V3 mix(V3 a, V3 b, float alpha) { return v3(mixf(a.x, b.x, alpha), mixf(a.y, b.y, alpha), mixf(a.z, b.z, alpha)); }

/* transformation.hpp:199-205 */
inline Trs mix(const Trs& a, const Trs& b, float alpha)
{
    return Trs{ mix(a.t, b.t, alpha), slerp(a.r, b.r, alpha), mix(a.s, b.s, alpha) };
}

/* Column-major 4x4 restricted to its three rows: m[c][r] = M[c * 3 + r], c = 0..3.
 * Transformation::toMat4() (transformation.hpp:105-113) = translate * toMat4(rotation), then
 * columns 0..2 scaled. With the translation matrix being identity apart from its last column,
 * the matrix product (gvm.hpp:1505-1517) reproduces the rotation entries exactly, so the
 * result is R[c][r] * s[c] and the translation column. */
struct Mat34 { float m[12]; };

inline void toMat3(Quat q, float* m /* 9, column-major */)
{
    /* gvm.hpp:1979-2001 */
    float xx = q.x * q.x, xy = q.x * q.y, xz = q.x * q.z, xw = q.x * q.w;
    float yy = q.y * q.y, yz = q.y * q.z, yw = q.y * q.w;
    float zz = q.z * q.z, zw = q.z * q.w;
    m[0] = 1.0f - 2.0f * (yy + zz);
    m[1] = 2.0f * (xy + zw);
    m[2] = 2.0f * (xz - yw);
    m[3] = 2.0f * (xy - zw);
    m[4] = 1.0f - 2.0f * (xx + zz);
    m[5] = 2.0f * (yz + xw);
    m[6] = 2.0f * (xz + yw);
    m[7] = 2.0f * (yz - xw);
    m[8] = 1.0f - 2.0f * (xx + yy);
}

inline Mat34 toMat34(const Trs& T)
{
    float r[9];
    toMat3(T.r, r);
    Mat34 M;
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < 3; i++)
            M.m[c * 3 + i] = r[c * 3 + i] * T.s[c];
    M.m[9] = T.t.x; M.m[10] = T.t.y; M.m[11] = T.t.z;
    return M;
}

/* (M * vec4(v, 1)).xyz, gvm.hpp:1481-1491: r[i] = 0; r[i] += m[j][i] * w[j] for j = 0..3 */
inline V3 transformPoint(const float* m /* column-major, column stride cs */, int cs, V3 v)
{
    V3 r;
    r.x = (((0.0f + m[0 * cs + 0] * v.x) + m[1 * cs + 0] * v.y) + m[2 * cs + 0] * v.z) + m[3 * cs + 0] * 1.0f;
    r.y = (((0.0f + m[0 * cs + 1] * v.x) + m[1 * cs + 1] * v.y) + m[2 * cs + 1] * v.z) + m[3 * cs + 1] * 1.0f;
    r.z = (((0.0f + m[0 * cs + 2] * v.x) + m[1 * cs + 2] * v.y) + m[2 * cs + 2] * v.z) + m[3 * cs + 2] * 1.0f;
    return r;
}

struct Box {
    V3 lo, hi;
};

inline Box merge(const Box& a, const Box& b) { return Box{ vmin(a.lo, b.lo), vmax(a.hi, b.hi) }; }  /* aabb.hpp:89-92 */

/* aabb.hpp:64-68 */
inline float surfaceArea(const Box& b)
{
    V3 l = b.hi - b.lo;
    return 2.0f * (l.x * l.y + l.y * l.z + l.x * l.z);
}

/* aabb.hpp:52-62 */
inline int longestAxis(const Box& b)
{
    int axis = 2;
    V3 l = b.hi - b.lo;
    if (l.x > l.y && l.x > l.z)
        axis = 0;
    else if (l.y > l.z)
        axis = 1;
    return axis;
}

/* aabb.hpp:47-50 */
inline V3 center(const Box& b) { return 0.5f * (b.lo + b.hi); }

}
