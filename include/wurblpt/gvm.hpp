/*
 * gvm.hpp -- GLSL-like vectors, matrices and quaternions for the WurblPT API mirror.
 *
 * Source-compatible with the subset of the reference's libwurblpt/gvm.hpp that scene
 * construction code uses (vec2/3/4, ivec/uvec, mat3/mat4, quat, the GLSL function set,
 * the rotation conversions). Written from scratch; where a result ends up in scene data
 * (vertices, transformation matrices, camera poses) the operation order follows the
 * reference so that both produce bit-identical numbers under IEEE-strict compilation:
 *   dot / length / normalize   gvm.hpp:1183-1204
 *   reflect / refract / cross  gvm.hpp:1213-1231
 *   matrix * vector, matrix * matrix   gvm.hpp:1481-1517
 *   quaternion product, rotation of a vector, slerp   gvm.hpp:1682-1796
 *   toQuat / toMat3 / toMat4   gvm.hpp:1848-2062
 */
#pragma once

#include <cassert>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <type_traits>

namespace WurblPT {

/* ---- constants (gvm.hpp:54-83) ---- */
template<typename T> constexpr T e_v = T(2.7182818284590452353602874713526625L);
inline constexpr float e = e_v<float>;
template<typename T> constexpr T pi_v = T(3.1415926535897932384626433832795029L);
inline constexpr float pi = pi_v<float>;
template<typename T> constexpr T pi_2_v = T(1.5707963267948966192313216916397514L);
inline constexpr float pi_2 = pi_2_v<float>;
template<typename T> constexpr T pi_4_v = T(0.7853981633974483096156608458198757L);
inline constexpr float pi_4 = pi_4_v<float>;
template<typename T> constexpr T inv_pi_v = T(0.3183098861837906715377675267450287L);
inline constexpr float inv_pi = inv_pi_v<float>;
template<typename T> constexpr T sqrt2_v = T(1.4142135623730950488016887242096981L);
inline constexpr float sqrt2 = sqrt2_v<float>;
template<typename T> constexpr T inv_sqrt2_v = T(0.7071067811865475244008443621048490L);
inline constexpr float inv_sqrt2 = inv_sqrt2_v<float>;
template<typename T> constexpr T maxval_v = std::numeric_limits<T>::max();
inline constexpr float maxval = maxval_v<float>;
template<typename T> constexpr T minval_v = std::numeric_limits<T>::lowest();
inline constexpr float minval = minval_v<float>;
template<typename T> constexpr T epsilon_v = std::numeric_limits<T>::epsilon();
inline constexpr float epsilon = epsilon_v<float>;

/* ---- scalar functions ---- */
template<typename T> constexpr T min(T x, T y) requires (std::is_arithmetic_v<T>) { return (x < y ? x : y); }
template<typename T> constexpr T min(T x, T y, T z) { return min(min(x, y), z); }
template<typename T> constexpr T min(T x, T y, T z, T w) { return min(min(min(x, y), z), w); }
template<typename T> constexpr T max(T x, T y) requires (std::is_arithmetic_v<T>) { return (x > y ? x : y); }
template<typename T> constexpr T max(T x, T y, T z) { return max(max(x, y), z); }
template<typename T> constexpr T max(T x, T y, T z, T w) { return max(max(max(x, y), z), w); }
template<typename T> constexpr T clamp(T x, T lo, T hi) requires (std::is_arithmetic_v<T>) { return min(hi, max(lo, x)); }
template<typename T> constexpr T step(T edge, T x) requires (std::is_arithmetic_v<T>) { return (x < edge ? T(0) : T(1)); }
template<typename T> constexpr T mod(T x, T y) requires (std::is_floating_point_v<T>) { return x - std::floor(x / y) * y; }
template<typename T> constexpr T sign(T x) requires (std::is_arithmetic_v<T>) { return (x < T(0) ? T(-1) : x > T(0) ? T(1) : T(0)); }
template<typename T> constexpr bool signbit(T x) requires (std::is_floating_point_v<T>) { return std::signbit(x); }
template<typename T> constexpr T deg2rad_v = pi_v<T> / T(180.0L);
template<typename T> constexpr T rad2deg_v = T(180.0L) / pi_v<T>;
template<typename T> constexpr T radians(T x) requires (std::is_floating_point_v<T>) { return x * deg2rad_v<T>; }
template<typename T> constexpr T degrees(T x) requires (std::is_floating_point_v<T>) { return x * rad2deg_v<T>; }
using std::abs; using std::sin; using std::cos; using std::tan; using std::asin; using std::acos; using std::atan;
using std::pow; using std::exp; using std::exp2; using std::log; using std::log2; using std::log10; using std::sqrt;
using std::cbrt; using std::round; using std::floor; using std::ceil;
using std::isfinite; using std::isnan; using std::isinf; using std::isnormal;
template<typename T> constexpr T atan(T y, T x) requires (std::is_floating_point_v<T>) { return std::atan2(y, x); }
template<typename T> constexpr T sqr(T x) requires (std::is_arithmetic_v<T>) { return x * x; }
template<typename T> constexpr T inversesqrt(T x) requires (std::is_floating_point_v<T>) { return T(1) / std::sqrt(x); }
template<typename T> constexpr T fract(T x) requires (std::is_floating_point_v<T>) { return x - std::floor(x); }
template<typename T> constexpr T mix(T x, T y, T alpha) requires (std::is_floating_point_v<T>) { return x + alpha * (y - x); }
template<typename T> constexpr T smoothstep(T e0, T e1, T x) requires (std::is_floating_point_v<T>)
{
    T t = clamp((x - e0) / (e1 - e0), T(0), T(1));
    return t * t * (T(3) - T(2) * t);
}
template<typename T> constexpr bool greaterThan(T x, T y) requires (std::is_arithmetic_v<T>) { return x > y; }
template<typename T> constexpr bool greaterThanEqual(T x, T y) requires (std::is_arithmetic_v<T>) { return x >= y; }
template<typename T> constexpr bool lessThan(T x, T y) requires (std::is_arithmetic_v<T>) { return x < y; }
template<typename T> constexpr bool lessThanEqual(T x, T y) requires (std::is_arithmetic_v<T>) { return x <= y; }
template<typename T> constexpr bool equal(T x, T y) requires (std::is_arithmetic_v<T>) { return x == y; }
template<typename T> constexpr bool notEqual(T x, T y) requires (std::is_arithmetic_v<T>) { return x != y; }
constexpr bool any(bool a) { return a; }
constexpr bool all(bool a) { return a; }

/* ---- vector ---- */
template<typename T, int N>
class vector
{
public:
    T values[N];

    vector() {}
    vector(T s) { for (int i = 0; i < N; i++) values[i] = s; }   /* implicit, as in the reference (gvm.hpp:230) */
    vector(T x, T y) requires (N == 2) : values { x, y } {}
    vector(T x, T y, T z) requires (N == 3) : values { x, y, z } {}
    vector(T x, T y, T z, T w) requires (N == 4) : values { x, y, z, w } {}
    vector(const vector<T, 2>& v, T z) requires (N == 3) : values { v[0], v[1], z } {}
    vector(T x, const vector<T, 2>& v) requires (N == 3) : values { x, v[0], v[1] } {}
    vector(const vector<T, 3>& v, T w) requires (N == 4) : values { v[0], v[1], v[2], w } {}
    vector(T x, const vector<T, 3>& v) requires (N == 4) : values { x, v[0], v[1], v[2] } {}
    vector(const vector<T, 2>& v, T z, T w) requires (N == 4) : values { v[0], v[1], z, w } {}
    vector(const vector<T, 2>& a, const vector<T, 2>& b) requires (N == 4) : values { a[0], a[1], b[0], b[1] } {}
    explicit vector(const T* p) { for (int i = 0; i < N; i++) values[i] = p[i]; }
    template<typename U> explicit vector(const vector<U, N>& v) { for (int i = 0; i < N; i++) values[i] = T(v[i]); }

    /* like the reference, a vector converts to a pointer to its components; the index operators
     * take ptrdiff_t so that v[i] stays unambiguous next to that conversion */
    operator const T*() const { return values; }
    const T& operator[](std::ptrdiff_t i) const { return values[i]; }
    T& operator[](std::ptrdiff_t i) { return values[i]; }

    /* named components */
    const T& x() const { return values[0]; }
    T& x() { return values[0]; }
    const T& y() const requires (N >= 2) { return values[1]; }
    T& y() requires (N >= 2) { return values[1]; }
    const T& z() const requires (N >= 3) { return values[2]; }
    T& z() requires (N >= 3) { return values[2]; }
    const T& w() const requires (N >= 4) { return values[3]; }
    T& w() requires (N >= 4) { return values[3]; }
    const T& r() const { return values[0]; }
    T& r() { return values[0]; }
    const T& g() const requires (N >= 2) { return values[1]; }
    T& g() requires (N >= 2) { return values[1]; }
    const T& b() const requires (N >= 3) { return values[2]; }
    T& b() requires (N >= 3) { return values[2]; }
    const T& a() const requires (N >= 4) { return values[3]; }
    T& a() requires (N >= 4) { return values[3]; }
    const T& s() const { return values[0]; }
    T& s() { return values[0]; }
    const T& t() const requires (N >= 2) { return values[1]; }
    T& t() requires (N >= 2) { return values[1]; }

    /* the swizzles scene code uses */
    vector<T, 2> xy() const requires (N >= 2) { return vector<T, 2>(values[0], values[1]); }
    vector<T, 2> rg() const requires (N >= 2) { return vector<T, 2>(values[0], values[1]); }
    vector<T, 2> st() const requires (N >= 2) { return vector<T, 2>(values[0], values[1]); }
    vector<T, 3> xyz() const requires (N >= 3) { return vector<T, 3>(values[0], values[1], values[2]); }
    vector<T, 3> rgb() const requires (N >= 3) { return vector<T, 3>(values[0], values[1], values[2]); }

    /* arithmetic */
    vector operator+() const { return *this; }
    vector operator-() const { vector r; for (int i = 0; i < N; i++) r[i] = -values[i]; return r; }
    friend vector operator+(const vector& a, const vector& b) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] + b[i]; return r; }
    friend vector operator-(const vector& a, const vector& b) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] - b[i]; return r; }
    friend vector operator*(const vector& a, const vector& b) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] * b[i]; return r; }
    friend vector operator/(const vector& a, const vector& b) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] / b[i]; return r; }
    friend vector operator*(T s, const vector& a) { vector r; for (int i = 0; i < N; i++) r[i] = s * a[i]; return r; }
    friend vector operator*(const vector& a, T s) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] * s; return r; }
    friend vector operator/(const vector& a, T s) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] / s; return r; }
    friend vector operator/(T s, const vector& a) { vector r; for (int i = 0; i < N; i++) r[i] = s / a[i]; return r; }
    friend vector operator+(const vector& a, T s) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] + s; return r; }
    friend vector operator-(const vector& a, T s) { vector r; for (int i = 0; i < N; i++) r[i] = a[i] - s; return r; }
    vector& operator+=(const vector& b) { for (int i = 0; i < N; i++) values[i] += b[i]; return *this; }
    vector& operator-=(const vector& b) { for (int i = 0; i < N; i++) values[i] -= b[i]; return *this; }
    vector& operator*=(const vector& b) { for (int i = 0; i < N; i++) values[i] *= b[i]; return *this; }
    vector& operator/=(const vector& b) { for (int i = 0; i < N; i++) values[i] /= b[i]; return *this; }
    vector& operator*=(T s) { for (int i = 0; i < N; i++) values[i] *= s; return *this; }
    vector& operator/=(T s) { for (int i = 0; i < N; i++) values[i] /= s; return *this; }
    friend bool operator==(const vector& a, const vector& b) { for (int i = 0; i < N; i++) if (!(a[i] == b[i])) return false; return true; }
    friend bool operator!=(const vector& a, const vector& b) { return !(a == b); }

    /* component-wise functions */
#define WURBLPT_VEC_FN1(name) friend vector name(const vector& v) { vector r; for (int i = 0; i < N; i++) r[i] = name(v[i]); return r; }
    WURBLPT_VEC_FN1(abs) WURBLPT_VEC_FN1(sin) WURBLPT_VEC_FN1(cos) WURBLPT_VEC_FN1(tan) WURBLPT_VEC_FN1(asin)
    WURBLPT_VEC_FN1(acos) WURBLPT_VEC_FN1(exp) WURBLPT_VEC_FN1(log) WURBLPT_VEC_FN1(sqrt) WURBLPT_VEC_FN1(floor)
    WURBLPT_VEC_FN1(ceil) WURBLPT_VEC_FN1(round) WURBLPT_VEC_FN1(fract) WURBLPT_VEC_FN1(sign) WURBLPT_VEC_FN1(radians)
    WURBLPT_VEC_FN1(degrees) WURBLPT_VEC_FN1(inversesqrt) WURBLPT_VEC_FN1(sqr)
#undef WURBLPT_VEC_FN1
    friend vector pow(const vector& v, T p) { vector r; for (int i = 0; i < N; i++) r[i] = pow(v[i], p); return r; }
    friend vector pow(const vector& v, const vector& p) { vector r; for (int i = 0; i < N; i++) r[i] = pow(v[i], p[i]); return r; }
    friend vector<bool, N> isfinite(const vector& v) { vector<bool, N> r; for (int i = 0; i < N; i++) r[i] = isfinite(v[i]); return r; }
    friend vector<bool, N> isnan(const vector& v) { vector<bool, N> r; for (int i = 0; i < N; i++) r[i] = isnan(v[i]); return r; }
    friend vector min(const vector& a, const vector& b) { vector r; for (int i = 0; i < N; i++) r[i] = min(a[i], b[i]); return r; }
    friend vector min(const vector& a, T x) { vector r; for (int i = 0; i < N; i++) r[i] = min(a[i], x); return r; }
    friend vector max(const vector& a, const vector& b) { vector r; for (int i = 0; i < N; i++) r[i] = max(a[i], b[i]); return r; }
    friend vector max(const vector& a, T x) { vector r; for (int i = 0; i < N; i++) r[i] = max(a[i], x); return r; }
    friend vector min(const vector& a, const vector& b, const vector& c) { return min(min(a, b), c); }
    friend vector max(const vector& a, const vector& b, const vector& c) { return max(max(a, b), c); }
    friend vector clamp(const vector& v, T lo, T hi) { vector r; for (int i = 0; i < N; i++) r[i] = clamp(v[i], lo, hi); return r; }
    friend vector clamp(const vector& v, const vector& lo, const vector& hi) { vector r; for (int i = 0; i < N; i++) r[i] = clamp(v[i], lo[i], hi[i]); return r; }
    friend vector mix(const vector& a, const vector& b, T alpha) { vector r; for (int i = 0; i < N; i++) r[i] = mix(a[i], b[i], alpha); return r; }
    friend vector mix(const vector& a, const vector& b, const vector& alpha) { vector r; for (int i = 0; i < N; i++) r[i] = mix(a[i], b[i], alpha[i]); return r; }
    friend vector step(T edge, const vector& v) { vector r; for (int i = 0; i < N; i++) r[i] = step(edge, v[i]); return r; }
    friend vector smoothstep(T e0, T e1, const vector& v) { vector r; for (int i = 0; i < N; i++) r[i] = smoothstep(e0, e1, v[i]); return r; }
    friend vector mod(const vector& v, T y) { vector r; for (int i = 0; i < N; i++) r[i] = mod(v[i], y); return r; }
#define WURBLPT_VEC_CMP(name) friend vector<bool, N> name(const vector& a, const vector& b) { vector<bool, N> r; for (int i = 0; i < N; i++) r[i] = name(a[i], b[i]); return r; }
    WURBLPT_VEC_CMP(greaterThan) WURBLPT_VEC_CMP(greaterThanEqual) WURBLPT_VEC_CMP(lessThan)
    WURBLPT_VEC_CMP(lessThanEqual) WURBLPT_VEC_CMP(equal) WURBLPT_VEC_CMP(notEqual)
#undef WURBLPT_VEC_CMP

    /* reductions */
    friend T min(const vector& a) { T r = a[0]; for (int i = 1; i < N; i++) if (a[i] < r) r = a[i]; return r; }
    friend T max(const vector& a) { T r = a[0]; for (int i = 1; i < N; i++) if (a[i] > r) r = a[i]; return r; }
    friend T average(const vector& a)
    {
        constexpr T invN = T(1) / T(N);
        T sum = 0;
        for (int i = 0; i < N; i++)
            sum += a[i];
        return invN * sum;
    }

    /* geometry */
    friend T dot(const vector& a, const vector& b) { T d = T(0); for (int i = 0; i < N; i++) d += a[i] * b[i]; return d; }
    friend T length(const vector& a) { return sqrt(dot(a, a)); }
    friend T distance(const vector& a, const vector& b) { return length(a - b); }
    friend vector normalize(const vector& v) { return v / length(v); }
    friend vector faceforward(const vector& v, const vector& I, const vector& Nref) requires (N == 3) { return dot(Nref, I) < T(0) ? v : -v; }
    friend vector reflect(const vector& i, const vector& n) requires (N == 3) { return i - T(2) * dot(n, i) * n; }
    friend vector refract(const vector& i, const vector& n, T eta) requires (N == 3)
    {
        const T d = dot(n, i);
        const T k = T(1) - eta * eta * (T(1) - d * d);
        return k <= T(0) ? vector(T(0)) : i * eta - n * (eta * d + sqrt(k));
    }
    friend vector cross(const vector& v, const vector& w) requires (N == 3)
    {
        return vector(v[1] * w[2] - v[2] * w[1], v[2] * w[0] - v[0] * w[2], v[0] * w[1] - v[1] * w[0]);
    }
};

template<int N> bool any(const vector<bool, N>& v) { bool r = false; for (int i = 0; i < N; i++) r = r || v[i]; return r; }
template<int N> bool all(const vector<bool, N>& v) { bool r = true; for (int i = 0; i < N; i++) r = r && v[i]; return r; }

typedef vector<bool, 2> bvec2;  typedef vector<bool, 3> bvec3;  typedef vector<bool, 4> bvec4;
typedef vector<int, 2> ivec2;   typedef vector<int, 3> ivec3;   typedef vector<int, 4> ivec4;
typedef vector<unsigned int, 2> uvec2; typedef vector<unsigned int, 3> uvec3; typedef vector<unsigned int, 4> uvec4;
typedef vector<float, 2> vec2;  typedef vector<float, 3> vec3;  typedef vector<float, 4> vec4;
typedef vector<double, 2> dvec2; typedef vector<double, 3> dvec3; typedef vector<double, 4> dvec4;

/* ---- matrix: C columns of R rows, column-major, m[c][r] ---- */
template<typename T, int C, int R>
class matrix
{
public:
    T values[C * R];

    matrix() {}
    explicit matrix(T diagonal)
    {
        for (int c = 0; c < C; c++)
            for (int r = 0; r < R; r++)
                values[c * R + r] = (c == r ? diagonal : T(0));
    }
    matrix(const vector<T, R>& c0, const vector<T, R>& c1, const vector<T, R>& c2) requires (C == 3)
    {
        for (int r = 0; r < R; r++) { values[r] = c0[r]; values[R + r] = c1[r]; values[2 * R + r] = c2[r]; }
    }
    matrix(const vector<T, R>& c0, const vector<T, R>& c1, const vector<T, R>& c2, const vector<T, R>& c3) requires (C == 4)
    {
        for (int r = 0; r < R; r++) { values[r] = c0[r]; values[R + r] = c1[r]; values[2 * R + r] = c2[r]; values[3 * R + r] = c3[r]; }
    }

    const T* operator[](int c) const { return values + c * R; }
    T* operator[](int c) { return values + c * R; }

    friend vector<T, R> operator*(const matrix& m, const vector<T, C>& w)
    {
        vector<T, R> r;
        for (int i = 0; i < R; i++) {
            r[i] = T(0);
            for (int j = 0; j < C; j++)
                r[i] += m[j][i] * w[j];
        }
        return r;
    }
    friend matrix<T, R, R> operator*(const matrix<T, C, R>& m, const matrix<T, R, C>& n)
    {
        matrix<T, R, R> r;
        for (int i = 0; i < R; i++) {
            for (int j = 0; j < R; j++) {
                r[i][j] = T(0);
                for (int k = 0; k < C; k++)
                    r[i][j] += m[k][j] * n[i][k];
            }
        }
        return r;
    }
    matrix& operator*=(const matrix& b) { matrix r = *this * b; *this = r; return *this; }
    friend bool operator==(const matrix& a, const matrix& b) { for (int i = 0; i < C * R; i++) if (!(a.values[i] == b.values[i])) return false; return true; }
    friend bool operator!=(const matrix& a, const matrix& b) { return !(a == b); }

    void translate(const vector<T, 3>& t) requires (C == 4 && R == 4)
    {
        for (int r = 0; r < 4; r++)
            values[3 * R + r] = dot(vector<T, 3>(values[0 * R + r], values[1 * R + r], values[2 * R + r]), t) + values[3 * R + r];
    }
    void scale(const vector<T, 3>& v) requires (C == 4 && R == 4)
    {
        for (int j = 0; j < C - 1; j++)
            for (int i = 0; i < R; i++)
                values[j * R + i] *= v[j];
    }
    friend matrix transpose(const matrix& m) requires (C == R)
    {
        matrix r;
        for (int c = 0; c < C; c++)
            for (int i = 0; i < R; i++)
                r[c][i] = m[i][c];
        return r;
    }
};

typedef matrix<float, 2, 2> mat2;
typedef matrix<float, 3, 3> mat3;
typedef matrix<float, 4, 4> mat4;
typedef matrix<double, 3, 3> dmat3;
typedef matrix<double, 4, 4> dmat4;

/* ---- quaternion (gvm.hpp:1657-1797) ---- */
template<typename T>
class quaternion
{
public:
    T x, y, z, w;

    quaternion() {}
    quaternion(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
    constexpr static quaternion null() { quaternion q; q.x = T(0); q.y = T(0); q.z = T(0); q.w = T(1); return q; }

    quaternion operator+() const { return *this; }
    quaternion operator-() const { return conjugate(*this); }
    quaternion operator*(const quaternion& q) const
    {
        quaternion p;
        p.x = w * q.x + x * q.w + y * q.z - z * q.y;
        p.y = w * q.y + y * q.w + z * q.x - x * q.z;
        p.z = w * q.z + z * q.w + x * q.y - y * q.x;
        p.w = w * q.w - x * q.x - y * q.y - z * q.z;
        return p;
    }
    const quaternion& operator*=(const quaternion& q) { *this = *this * q; return *this; }
    friend bool operator==(const quaternion& a, const quaternion& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
    friend bool operator!=(const quaternion& a, const quaternion& b) { return !(a == b); }
    friend vector<T, 3> operator*(const quaternion& q, const vector<T, 3>& v)
    {
        vector<T, 3> s(q.x, q.y, q.z);
        vector<T, 3> t = T(2) * cross(s, v);
        return v + q.w * t + cross(s, t);
    }
    friend vector<T, 3> operator*(const vector<T, 3>& v, const quaternion& q) { return conjugate(q) * v; }

    vector<T, 3> axis() const
    {
        T t = sqrt(x * x + y * y + z * z);
        if (t < epsilon_v<T>)
            return vector<T, 3>(T(0));
        return vector<T, 3>(x / t, y / t, z / t);
    }
    T angle() const { return T(2) * atan(sqrt(x * x + y * y + z * z), w); }
    friend T magnitude(const quaternion& q) { return sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); }
    friend quaternion normalize(const quaternion q)
    {
        T m = magnitude(q);
        return quaternion(q.x / m, q.y / m, q.z / m, q.w / m);
    }
    friend quaternion conjugate(const quaternion& q) { return quaternion(-q.x, -q.y, -q.z, q.w); }
    friend quaternion inverse(const quaternion& q) { return conjugate(q); }
    friend quaternion slerp(const quaternion& q, const quaternion& r, T alpha)
    {
        quaternion w = r;
        T cosHalfAngle = q.x * r.x + q.y * r.y + q.z * r.z + q.w * r.w;
        if (cosHalfAngle < T(0)) {
            w.x = -w.x; w.y = -w.y; w.z = -w.z; w.w = -w.w;
            cosHalfAngle = -cosHalfAngle;
        }
        T tq, tw;
        if (cosHalfAngle >= T(1)) {
            tq = T(1);
            tw = T(0);
        } else {
            T halfAngle = acos(cosHalfAngle);
            T sinHalfAngle = sqrt(T(1) - cosHalfAngle * cosHalfAngle);
            if (abs(sinHalfAngle) < epsilon_v<T>) {
                tq = T(0.5);
                tw = T(0.5);
            } else {
                tq = sin((T(1) - alpha) * halfAngle) / sinHalfAngle;
                tw = sin(alpha * halfAngle) / sinHalfAngle;
            }
        }
        return quaternion(q.x * tq + w.x * tw, q.y * tq + w.y * tw, q.z * tq + w.z * tw, q.w * tq + w.w * tw);
    }
};

typedef quaternion<float> quat;
typedef quaternion<double> dquat;

/* ---- rotation conversions (gvm.hpp:1822-2062) ---- */

/* some unit vector orthogonal to v (Duff et al. 2017) */
template<typename T> vector<T, 3> someTangentTo(const vector<T, 3>& v)
{
    T sign = std::copysign(T(1), v.z());
    T a = T(-1) / (sign + v.z());
    T b = v.x() * v.y() * a;
    return vector<T, 3>(T(1) + sign * v.x() * v.x() * a, sign * b, -sign * v.x());
}

template<typename T> quaternion<T> toQuat(T angle, const vector<T, 3>& axis)
{
    vector<T, 3> n = normalize(axis);
    T sinA = sin(T(0.5l) * angle);
    T cosA = cos(T(0.5l) * angle);
    return quaternion<T>(n.x() * sinA, n.y() * sinA, n.z() * sinA, cosA);
}

template<typename T> quaternion<T> toQuat(const vector<T, 3>& dir1Normalized, const vector<T, 3>& dir2Normalized)
{
    T cosAngle = dot(dir1Normalized, dir2Normalized);
    quaternion<T> q;
    if (cosAngle >= T(1) - epsilon_v<T>) {
        q = quaternion<T>::null();
    } else if (cosAngle <= T(-1) + epsilon_v<T>) {
        vector<T, 3> axis = someTangentTo(dir1Normalized);
        q = quaternion<T>(axis.x(), axis.y(), axis.z(), T(0));
    } else {
        q = toQuat(acos(cosAngle), cross(dir1Normalized, dir2Normalized));
    }
    return q;
}

template<typename T> quaternion<T> toQuat(const vector<T, 3>& eulerRot)
{
    T x2 = T(0.5l) * eulerRot.x(), y2 = T(0.5l) * eulerRot.y(), z2 = T(0.5l) * eulerRot.z();
    T cx = cos(x2), sx = sin(x2), cy = cos(y2), sy = sin(y2), cz = cos(z2), sz = sin(z2);
    quaternion<T> q;
    q.x = sx * cy * cz - cx * sy * sz;
    q.y = cx * sy * cz + sx * cy * sz;
    q.z = cx * cy * sz - sx * sy * cz;
    q.w = cx * cy * cz + sx * sy * sz;
    return q;
}

template<typename T> matrix<T, 3, 3> toMat3(const quaternion<T>& q)
{
    matrix<T, 3, 3> m;
    T xx = q.x * q.x, xy = q.x * q.y, xz = q.x * q.z, xw = q.x * q.w;
    T yy = q.y * q.y, yz = q.y * q.z, yw = q.y * q.w;
    T zz = q.z * q.z, zw = q.z * q.w;
    m[0][0] = T(1) - T(2) * (yy + zz);
    m[0][1] = T(2) * (xy + zw);
    m[0][2] = T(2) * (xz - yw);
    m[1][0] = T(2) * (xy - zw);
    m[1][1] = T(1) - T(2) * (xx + zz);
    m[1][2] = T(2) * (yz + xw);
    m[2][0] = T(2) * (xz + yw);
    m[2][1] = T(2) * (yz - xw);
    m[2][2] = T(1) - T(2) * (xx + yy);
    return m;
}

template<typename T> matrix<T, 4, 4> toMat4(const quaternion<T>& q)
{
    matrix<T, 3, 3> r = toMat3(q);
    matrix<T, 4, 4> m(T(1));
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < 3; i++)
            m[c][i] = r[c][i];
    return m;
}

template<typename T> matrix<T, 3, 3> toMat3(T angle, const vector<T, 3>& axisNormalized)
{
    return toMat3(toQuat(angle, axisNormalized));
}

}
