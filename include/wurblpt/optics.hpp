/* optics.hpp -- Projection, LensDistortion, LensDepthOfField, Optics (mirror of reference
 * libwurblpt/optics.hpp:37-351: same constructors and public members; distort()/undistort() are
 * kept on the host for applications, ray generation runs them on the device). */
#pragma once

#include <limits>

#include "gvm.hpp"
#include "sampler.hpp"
#include "prng.hpp"

namespace WurblPT {

class Projection
{
public:
    float t, b, r, l;    // frustum at near=1

    Projection(float l_, float r_, float b_, float t_) : t(t_), b(b_), r(r_), l(l_) {}
    Projection(float vfov, float aspectRatio) : t(tan(vfov * 0.5f)), b(-t), r(t * aspectRatio), l(-r) {}
    Projection(unsigned int width, unsigned int height, const vec2& centerPixel, const vec2& focalLength) :
        t((height - centerPixel.y()) / focalLength.y()),
        b((0.0f - centerPixel.y()) / focalLength.y()),
        r((width - centerPixel.x()) / focalLength.x()),
        l((0.0f - centerPixel.x()) / focalLength.x())
    {
    }
    Projection() : Projection(radians(60.0f), 4.0f / 3.0f) {}

    float vFov() const { return atan(t) - atan(b); }
    float hFov() const { return atan(r) - atan(l); }
    float aspectRatio() const { return (r - l) / (t - b); }
    vec2 center() const { return vec2(l / (l - r), b / (b - t)); }
    vec2 focalLength() const { return vec2(1.0f / (r - l), 1.0f / (t - b)); }
    vec2 inverseFocalLength() const { return vec2(r - l, t - b); }
    vec2 centerPixel(unsigned int w, unsigned int h) const { return center() * vec2(w, h); }
    vec2 focalLengthInPixels(unsigned int w, unsigned int h) const { return focalLength() * vec2(w, h); }
};

class LensDistortion
{
private:
    static constexpr int maxIterations = 256;
    static constexpr float epsilon = 0.001f;    // in pixels

public:
    typedef enum { None, RadialAndPlanar, RadialOnly, OpenCV } Type;

    class Helper {
    public:
        vec2 center;
        vec2 focalLength;
        vec2 inverseFocalLength;
        unsigned int width;
        unsigned int height;
    };

    Type type;
    float k1, k2, k3, p1, p2;
    float b1, b2, b3, b4;

    LensDistortion() : type(None), k1(0.0f), k2(0.0f), k3(0.0f), p1(0.0f), p2(0.0f), b1(0.0f), b2(0.0f), b3(0.0f), b4(0.0f) {}

    /* RadialAndPlanar (Lambers, Sommerhoff, Kolb 2018) */
    LensDistortion(float k1_, float k2_, float p1_, float p2_) :
        type((k1_ == 0.0f && k2_ == 0.0f && p1_ == 0.0f && p2_ == 0.0f) ? None : RadialAndPlanar),
        k1(k1_), k2(k2_), k3(0.0f), p1(p1_), p2(p2_), b1(0.0f), b2(0.0f), b3(0.0f), b4(0.0f)
    {
    }

    /* RadialOnly (Drap, Lefevre 2016) */
    LensDistortion(float k1_, float k2_, float k3_) :
        type((k1_ == 0.0f && k2_ == 0.0f && k3_ == 0.0f) ? None : RadialOnly),
        k1(k1_), k2(k2_), k3(k3_), p1(0.0f), p2(0.0f),
        b1(-k1),
        b2(3.0f * k1 * k1 - k2),
        b3(-12.0f * k1 * k1 * k1 + 8.0f * k1 * k2 - k3),
        b4(55.0f * k1 * k1 * k1 * k1 - 55.0f * k1 * k1 * k2 + 5.0f * k2 * k2 + 10.0f * k1 * k3)
    {
    }

    /* OpenCV model */
    LensDistortion(float k1_, float k2_, float k3_, float p1_, float p2_) :
        type((k1_ == 0.0f && k2_ == 0.0f && k3_ == 0.0f && p1_ == 0.0f && p2_ == 0.0f) ? None : OpenCV),
        k1(k1_), k2(k2_), k3(k3_), p1(p1_), p2(p2_), b1(0.0f), b2(0.0f), b3(0.0f), b4(0.0f)
    {
    }

    Helper getHelper(const Projection& projection, unsigned int width, unsigned int height) const
    {
        Helper h;
        h.center = projection.center();
        h.focalLength = projection.focalLength();
        h.inverseFocalLength = projection.inverseFocalLength();
        h.width = width;
        h.height = height;
        return h;
    }

    void distort(float& p, float& q, const Helper& helper) const
    {
        if (type == None)
            return;
        float s = (p - helper.center.x()) * helper.inverseFocalLength.x();
        float t = (q - helper.center.y()) * helper.inverseFocalLength.y();
        float r2 = s * s + t * t;
        float r4 = r2 * r2;
        float r6 = r4 * r2;
        float rd = 1.0f + k1 * r2 + k2 * r4 + k3 * r6;
        float term1 = 2.0f * s * t;
        float term2 = r2 + 2.0f * s * s;
        float term3 = r2 + 2.0f * t * t;
        float ns = s * rd + p1 * term1 + p2 * term2;
        float nt = t * rd + p1 * term3 + p2 * term1;
        p = ns * helper.focalLength.x() + helper.center.x();
        q = nt * helper.focalLength.y() + helper.center.y();
    }

    void undistort(float& p, float& q, const Helper& helper) const
    {
        if (type == None)
            return;
        float s = (p - helper.center.x()) * helper.inverseFocalLength.x();
        float t = (q - helper.center.y()) * helper.inverseFocalLength.y();
        if (type == RadialAndPlanar) {
            float r2 = s * s + t * t;
            float r4 = r2 * r2;
            float d1 = k1 * r2 + k2 * r4;
            float d2 = 1.0f / (4.0f * k1 * r2 + 6.0f * k2 * r4 + 8.0f * p1 * t + 8.0f * p2 * s + 1.0f);
            p = (s - d2 * (d1 * s + 2.0f * p1 * s * t + p2 * (r2 + 2.0f * s * s))) * helper.focalLength.x() + helper.center.x();
            q = (t - d2 * (d1 * t + p1 * (r2 + 2.0f * t * t) + 2.0f * p2 * s * t)) * helper.focalLength.y() + helper.center.y();
        } else if (type == RadialOnly) {
            float r2 = s * s + t * t;
            float r4 = r2 * r2;
            float r6 = r4 * r2;
            float r8 = r4 * r4;
            float d = 1.0f + b1 * r2 + b2 * r4 + b3 * r6 + b4 * r8;
            p = s * d * helper.focalLength.x() + helper.center.x();
            q = t * d * helper.focalLength.y() + helper.center.y();
        } else {
            const float s0 = s, t0 = t;
            float squaredError = std::numeric_limits<float>::max();
            for (int i = 0; i < maxIterations && squaredError >= epsilon * epsilon; i++) {
                float r2 = s * s + t * t;
                float r4 = r2 * r2;
                float r6 = r4 * r2;
                float rd = 1.0f + k1 * r2 + k2 * r4 + k3 * r6;
                float invrd = 1.0f / rd;
                float ds = 2.0f * p1 * s * t + p2 * (r2 + 2.0f * s * s);
                float dt = 2.0f * p2 * s * t + p1 * (r2 + 2.0f * t * t);
                s = (s0 - ds) * invrd;
                t = (t0 - dt) * invrd;
                float ep = s * helper.focalLength.x() + helper.center.x();
                float eq = t * helper.focalLength.y() + helper.center.y();
                distort(ep, eq, helper);
                vec2 offsetInPixels = (vec2(ep, eq) - vec2(p, q)) * vec2(helper.width, helper.height);
                squaredError = dot(offsetInPixels, offsetInPixels);
            }
            p = s * helper.focalLength.x() + helper.center.x();
            q = t * helper.focalLength.y() + helper.center.y();
        }
    }
};

class LensDepthOfField
{
public:
    float lensRadius;
    float focusDist;

    LensDepthOfField(float aperture = 0.0f, float focusDist_ = 1.0f) : lensRadius(aperture * 0.5f), focusDist(focusDist_) {}

    vec3 offset(vec3& P, Prng& prng) const
    {
        vec3 O(0.0f);
        if (lensRadius > 0.0f) {
            P *= focusDist;
            O = vec3(lensRadius * Sampler::inUnitDisk(prng.in01x2()), 0.0f);
        }
        return O;
    }
};

class Optics
{
public:
    Projection projection;
    LensDistortion distortion;
    LensDepthOfField depthOfField;

    Optics(const Projection& P = Projection(), const LensDistortion& LD = LensDistortion(),
            const LensDepthOfField& LDOF = LensDepthOfField()) :
        projection(P), distortion(LD), depthOfField(LDOF)
    {
    }
};

}
