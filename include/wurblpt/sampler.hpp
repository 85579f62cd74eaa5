/* sampler.hpp -- the sample warps of the WurblPT API (reference libwurblpt/sampler.hpp:36-121; the maps are
 * the ones of PBR3 section 13.6). Host side: applications use them to place objects, the noise textures and the
 * scene recipes to draw directions; the arithmetic is kept operation for operation so that a scene built from a
 * given Prng is the same scene as in the reference. The device has its own copies (csrc/device_math.cuh). */
#pragma once

#include "gvm.hpp"
#include "tangentspace.hpp"

namespace WurblPT {

class Sampler
{
private:
    /* point at height z on the unit sphere, at angle phi around the z axis */
    static vec3 atHeight(float z, float phi)
    {
        const float ringRadius = sqrt(max(0.0f, 1.0f - z * z));
        return vec3(ringRadius * cos(phi), ringRadius * sin(phi), z);
    }

public:
    /* concentric map of the unit square onto the unit disk */
    static vec2 inUnitDisk(const vec2& u)
    {
        const vec2 p = 2.0f * u - vec2(1.0f, 1.0f);
        if (p == vec2(0.0f, 0.0f))
            return p;
        const bool widerThanHigh = abs(p.x()) > abs(p.y());
        const float radius = widerThanHigh ? p.x() : p.y();
        const float angle = widerThanHigh ? pi_4 * (p.y() / p.x()) : pi_2 - pi_4 * (p.x() / p.y());
        return radius * vec2(cos(angle), sin(angle));
    }

    /* point on the unit circle */
    static vec2 onUnitDisk(float u0)
    {
        const float angle = u0 * 2.0f * pi;
        return vec2(cos(angle), sin(angle));
    }

    static vec3 onUnitSphere(const vec2& u) { return atHeight(1.0f - 2.0f * u.x(), 2.0f * pi * u.y()); }

    static vec3 onUnitHemisphere(const vec2& u) { return atHeight(u.x(), 2.0f * pi * u.y()); }

    /* barycentric coordinates, uniform over a triangle */
    static vec3 inTriangle(const vec2& u)
    {
        const float root = sqrt(u.x());
        const float first = 1.0f - root;
        const float second = u.y() * root;
        return vec3(first, second, 1.0f - first - second);
    }

    /* cosine-weighted direction around +z (Malley's method on the concentric disk) */
    static vec3 cosineDirection(const vec2& u)
    {
        const vec2 onDisk = inUnitDisk(u);
        return vec3(onDisk, sqrt(max(0.0f, 1.0f - dot(onDisk, onDisk))));
    }

    /* direction inside the cone of half-angle acos(cosThetaMax) around `direction` */
    static vec3 toSphere(const vec3& direction, float cosThetaMax, const vec2& u)
    {
        const float cosTheta = (1.0f - u.x()) + u.x() * cosThetaMax;
        const float sinTheta = sqrt(max(0.0f, 1.0f - cosTheta * cosTheta));
        const float phi = u.y() * 2.0f * pi;
        const vec3 inConeFrame(cos(phi) * sinTheta, sin(phi) * sinTheta, cosTheta);
        return normalize(TangentSpace(direction).toWorldSpace(inConeFrame));
    }
};

}
