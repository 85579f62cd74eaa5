/* ray.hpp -- Ray (mirror of reference libwurblpt/ray.hpp:36-56) */
#pragma once

#include "gvm.hpp"
#include "constants.hpp"

namespace WurblPT {

class Ray
{
public:
    vec3 origin;
    vec3 direction;     // unit length
    float time;
    vec4 refractiveIndex;

    Ray(const vec3& o, const vec3& d, float t, const vec4& refrInd) :
        origin(o), direction(d), time(t), refractiveIndex(refrInd)
    {
    }

    Ray(const vec3& o, const vec3& d, float t, float refrInd) : Ray(o, d, t, vec4(refrInd)) {}

    vec3 at(float a) const { return origin + a * direction; }
};

}
