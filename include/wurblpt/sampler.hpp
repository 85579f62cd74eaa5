/* sampler.hpp -- sample warps (mirror of reference libwurblpt/sampler.hpp:36-121; PBR3 13.6). */
#pragma once

#include "gvm.hpp"
#include "tangentspace.hpp"

namespace WurblPT {

class Sampler
{
public:
    static vec2 inUnitDisk(const vec2& u)
    {
        vec2 offset = 2.0f * u - vec2(1.0f, 1.0f);
        if (offset == vec2(0.0f, 0.0f))
            return vec2(0.0f, 0.0f);
        float theta, r;
        if (abs(offset.x()) > abs(offset.y())) {
            r = offset.x();
            theta = pi_4 * (offset.y() / offset.x());
        } else {
            r = offset.y();
            theta = pi_2 - pi_4 * (offset.x() / offset.y());
        }
        return r * vec2(cos(theta), sin(theta));
    }

    static vec2 onUnitDisk(float u0)
    {
        float phi = u0 * 2.0f * pi;
        return vec2(cos(phi), sin(phi));
    }

    static vec3 onUnitSphere(const vec2& u)
    {
        float z = 1.0f - 2.0f * u.x();
        float r = sqrt(max(0.0f, 1.0f - z * z));
        float phi = 2.0f * pi * u.y();
        return vec3(r * cos(phi), r * sin(phi), z);
    }

    static vec3 onUnitHemisphere(const vec2& u)
    {
        float z = u.x();
        float r = sqrt(max(0.0f, 1.0f - z * z));
        float phi = 2.0f * pi * u.y();
        return vec3(r * cos(phi), r * sin(phi), z);
    }

    static vec3 inTriangle(const vec2& u)
    {
        float su0 = sqrt(u.x());
        float b0 = 1.0f - su0;
        float b1 = u.y() * su0;
        return vec3(b0, b1, 1.0f - b0 - b1);
    }

    static vec3 cosineDirection(const vec2& u)
    {
        vec2 d = inUnitDisk(u);
        float z = sqrt(max(0.0f, 1.0f - dot(d, d)));
        return vec3(d, z);
    }

    static vec3 toSphere(const vec3& direction, float cosThetaMax, const vec2& u)
    {
        float cosTheta = (1.0f - u.x()) + u.x() * cosThetaMax;
        float sinTheta = sqrt(max(0.0f, 1.0f - cosTheta * cosTheta));
        float phi = u.y() * 2.0f * pi;
        vec3 aroundZ = vec3(cos(phi) * sinTheta, sin(phi) * sinTheta, cosTheta);
        return normalize(TangentSpace(direction).toWorldSpace(aroundZ));
    }
};

}
