/* aabb.hpp -- the axis-aligned box of the WurblPT API (reference libwurblpt/aabb.hpp:36-92). Host side only:
 * applications and scene recipes use it; the bounds the tree is built from are computed in
 * wurblpt_b200/csrc/scene_host.cpp and the slab test of the walk lives in csrc/device_trace.cuh. */
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

    /* 0, 1 or 2; on ties the later axis wins */
    int longestAxis() const
    {
        const vec3 extent = hi - lo;
        const bool xIsLongest = extent[0] > extent[1] && extent[0] > extent[2];
        return xIsLongest ? 0 : (extent[1] > extent[2] ? 1 : 2);
    }

    float surfaceArea() const
    {
        const vec3 extent = hi - lo;
        return 2.0f * (extent.x() * extent.y() + extent.y() * extent.z() + extent.x() * extent.z());
    }

    /* slab test: does the ray's parameter interval [amin, amax] overlap the box? The interval bounds are
     * folded into the per-axis entry and exit distances before the comparison (aabb.hpp:70-86). */
    bool mayHit(const Ray& ray, float amin, float amax, const vec3& invRayDirection) const
    {
        const vec3 towardsLo = (lo - ray.origin) * invRayDirection;
        const vec3 towardsHi = (hi - ray.origin) * invRayDirection;
        const float entry = max(vec4(amin, min(towardsLo, towardsHi)));
        const float exit = min(vec4(amax, max(towardsLo, towardsHi)));
        return entry <= exit;
    }
};

inline AABB merge(const AABB& a, const AABB& b)
{
    return AABB(min(a.lo, b.lo), max(a.hi, b.hi));
}

}
