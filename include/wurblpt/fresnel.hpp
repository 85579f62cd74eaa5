/* fresnel.hpp -- Fresnel terms (mirror of reference libwurblpt/fresnel.hpp:32-72) */
#pragma once

#include "gvm.hpp"

namespace WurblPT {

inline float fresnelSchlickR0(float ourIndexOfRefraction, float theirIndexOfRefraction)
{
    float r0 = (theirIndexOfRefraction - ourIndexOfRefraction) / (theirIndexOfRefraction + ourIndexOfRefraction);
    r0 *= r0;
    return r0;
}

inline float fresnelSchlick(float r0, float cosTheta)
{
    float t = 1.0f - cosTheta;
    float t2 = t * t;
    return r0 + (1.0f - r0) * t2 * t2 * t;
}

inline vec4 fresnelSchlick(const vec4& r0, float cosTheta)
{
    float t = 1.0f - cosTheta;
    float t2 = t * t;
    return r0 + (vec4(1.0f) - r0) * t2 * t2 * t;
}

inline float fresnelUnpolarized(float cosI, float cosT, float n1, float n2)
{
    float Fs = (n1 * cosI - n2 * cosT) / (n1 * cosI + n2 * cosT);
    Fs *= Fs;
    float Fp = (n1 * cosT - n2 * cosI) / (n1 * cosT + n2 * cosI);
    Fp *= Fp;
    return 0.5f * (Fs + Fp);
}

}
