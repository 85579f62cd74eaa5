/* aabb.hpp -- axis-aligned box (mirror of reference libwurblpt/aabb.hpp:36-92) */
#pragma once

#include "gvm.hpp"
#include "ray.hpp"

namespace WurblPT {

class AABB
{
public:
    vec3 lo, hi;

    AABB() {}
    AABB(const vec3& lo_, const vec3& hi_) : lo(lo_), hi(hi_) {}

    vec3 center() const { return 0.5f * (lo + hi); }

    int longestAxis() const
    {
        vec3 l = hi - lo;
        if (l[0] > l[1] && l[0] > l[2])
            return 0;
        return l[1] > l[2] ? 1 : 2;
    }

    float surfaceArea() const
    {
        vec3 l = hi - lo;
        return 2.0f * (l.x() * l.y() + l.y() * l.z() + l.x() * l.z());
    }

    bool mayHit(const Ray& ray, float amin, float amax, const vec3& invRayDirection) const
    {
        vec3 t0 = (lo - ray.origin) * invRayDirection;
        vec3 t1 = (hi - ray.origin) * invRayDirection;
        vec4 tmin = vec4(amin, min(t0, t1));
        vec4 tmax = vec4(amax, max(t0, t1));
        return max(tmin) <= min(tmax);
    }
};

inline AABB merge(const AABB& a, const AABB& b) { return AABB(min(a.lo, b.lo), max(a.hi, b.hi)); }

}
