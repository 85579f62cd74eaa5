/* fresnel.hpp -- reflectance at a dielectric interface (reference libwurblpt/fresnel.hpp:32-72): Schlick's
 * approximation from the reflectance at normal incidence, and the exact unpolarised Fresnel equations. */
#pragma once

#include "gvm.hpp"

namespace WurblPT {

/* reflectance at normal incidence between two refractive indices */
inline float fresnelSchlickR0(float ourIndexOfRefraction, float theirIndexOfRefraction)
{
    const float ratio = (theirIndexOfRefraction - ourIndexOfRefraction) / (theirIndexOfRefraction + ourIndexOfRefraction);
    return ratio * ratio;
}

namespace detail {

/* (1 - cos)^5, multiplied up the way the reference does: ((x^2) * (x^2)) * x is applied to (1 - r0) first */
template<typename T>
inline T schlick(const T& r0, const T& one, float cosTheta)
{
    const float x = 1.0f - cosTheta;
    const float xx = x * x;
    return r0 + (one - r0) * xx * xx * x;
}

}

inline float fresnelSchlick(float r0, float cosTheta) { return detail::schlick<float>(r0, 1.0f, cosTheta); }

inline vec4 fresnelSchlick(const vec4& r0, float cosTheta) { return detail::schlick<vec4>(r0, vec4(1.0f), cosTheta); }

/* mean of the s- and p-polarised reflectances; cosI / cosT are the cosines on the n1 / n2 side */
inline float fresnelUnpolarized(float cosI, float cosT, float n1, float n2)
{
    const float sAmplitude = (n1 * cosI - n2 * cosT) / (n1 * cosI + n2 * cosT);
    const float pAmplitude = (n1 * cosT - n2 * cosI) / (n1 * cosT + n2 * cosI);
    return 0.5f * (sAmplitude * sAmplitude + pAmplitude * pAmplitude);
}

}
