/* tangentspace.hpp -- orthonormal frames around a normal (mirror of reference
 * libwurblpt/tangentspace.hpp:43-136; Duff et al. 2017 for the normal-only case). */
#pragma once

#include <cmath>

#include "gvm.hpp"
#include "constants.hpp"

namespace WurblPT {

class TangentSpace
{
public:
    vec3 normal, tangent, bitangent;

    TangentSpace() {}

    TangentSpace(const vec3& n) : normal(n)
    {
        float sign = std::copysignf(1.0f, normal.z());
        float a = -1.0f / (sign + normal.z());
        float b = normal.x() * normal.y() * a;
        tangent = vec3(1.0f + sign * normal.x() * normal.x() * a, sign * b, -sign * normal.x());
        bitangent = vec3(b, sign + normal.y() * normal.y() * a, -normal.y());
    }

    TangentSpace(const vec3& n, const vec3& t) : normal(n), tangent(t), bitangent(cross(n, t)) {}
    TangentSpace(const vec3& n, const vec3& t, const vec3& b) : normal(n), tangent(t), bitangent(b) {}

    mat3 matrixToTangentSpace() const
    {
        return mat3(vec3(tangent.x(), bitangent.x(), normal.x()), vec3(tangent.y(), bitangent.y(), normal.y()),
                vec3(tangent.z(), bitangent.z(), normal.z()));
    }
    mat3 matrixFromTangentSpace() const { return mat3(tangent, bitangent, normal); }
    vec3 toTangentSpace(const vec3& v) const { return matrixToTangentSpace() * v; }
    vec3 fromWorldSpace(const vec3& v) const { return toTangentSpace(v); }
    vec3 fromTangentSpace(const vec3& v) const { return matrixFromTangentSpace() * v; }
    vec3 toWorldSpace(const vec3& v) const { return fromTangentSpace(v); }
};

}
