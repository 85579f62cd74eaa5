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
        return Helper { projection.center(), projection.focalLength(), projection.inverseFocalLength(), width, height };
    }

private:
    /* normalized image coordinates [0,1]^2 <-> lens coordinates (centred, in focal lengths) */
    static void toLens(float p, float q, const Helper& h, float& s, float& t)
    {
        s = (p - h.center.x()) * h.inverseFocalLength.x();
        t = (q - h.center.y()) * h.inverseFocalLength.y();
    }
    static void toImage(float s, float t, const Helper& h, float& p, float& q)
    {
        p = s * h.focalLength.x() + h.center.x();
        q = t * h.focalLength.y() + h.center.y();
    }

    /* the forward model shared by all three types (unused coefficients are zero): radial polynomial in r^2
     * plus the two tangential terms (reference optics.hpp:226-246) */
    void applyModel(float s, float t, float& sOut, float& tOut) const
    {
        const float r2 = s * s + t * t;
        const float r4 = r2 * r2;
        const float r6 = r4 * r2;
        const float radial = 1.0f + k1 * r2 + k2 * r4 + k3 * r6;
        const float mixed = 2.0f * s * t;
        const float alongS = r2 + 2.0f * s * s;
        const float alongT = r2 + 2.0f * t * t;
        sOut = s * radial + p1 * mixed + p2 * alongS;
        tOut = t * radial + p1 * alongT + p2 * mixed;
    }

public:
    /* p and q are normalized image coordinates in [0,1]^2 */
    void distort(float& p, float& q, const Helper& helper) const
    {
        if (type == None)
            return;
        float s, t, sd, td;
        toLens(p, q, helper, s, t);
        applyModel(s, t, sd, td);
        toImage(sd, td, helper, p, q);
    }

    /* the inverse: closed-form approximations for the two analytic models (optics.hpp:251-278), fixed-point
     * iteration with a tolerance of `epsilon` pixels for the OpenCV model (optics.hpp:279-308) */
    void undistort(float& p, float& q, const Helper& helper) const
    {
        if (type == None)
            return;
        float s, t;
        toLens(p, q, helper, s, t);
        const float r2 = s * s + t * t;
        const float r4 = r2 * r2;
        switch (type) {
        case RadialAndPlanar: {
            const float d1 = k1 * r2 + k2 * r4;
            const float d2 = 1.0f / (4.0f * k1 * r2 + 6.0f * k2 * r4 + 8.0f * p1 * t + 8.0f * p2 * s + 1.0f);
            p = (s - d2 * (d1 * s + 2.0f * p1 * s * t + p2 * (r2 + 2.0f * s * s))) * helper.focalLength.x() + helper.center.x();
            q = (t - d2 * (d1 * t + p1 * (r2 + 2.0f * t * t) + 2.0f * p2 * s * t)) * helper.focalLength.y() + helper.center.y();
            break;
        }
        case RadialOnly: {
            const float r6 = r4 * r2;
            const float r8 = r4 * r4;
            const float d = 1.0f + b1 * r2 + b2 * r4 + b3 * r6 + b4 * r8;
            p = s * d * helper.focalLength.x() + helper.center.x();
            q = t * d * helper.focalLength.y() + helper.center.y();
            break;
        }
        default: {
            const float sTarget = s, tTarget = t;
            float squaredError = std::numeric_limits<float>::max();
            for (int i = 0; i < maxIterations && squaredError >= epsilon * epsilon; i++) {
                const float rr2 = s * s + t * t;
                const float rr4 = rr2 * rr2;
                const float rr6 = rr4 * rr2;
                const float radial = 1.0f + k1 * rr2 + k2 * rr4 + k3 * rr6;
                const float inverseRadial = 1.0f / radial;
                const float shiftS = 2.0f * p1 * s * t + p2 * (rr2 + 2.0f * s * s);
                const float shiftT = 2.0f * p2 * s * t + p1 * (rr2 + 2.0f * t * t);
                s = (sTarget - shiftS) * inverseRadial;
                t = (tTarget - shiftT) * inverseRadial;
                float checkP, checkQ;
                toImage(s, t, helper, checkP, checkQ);
                distort(checkP, checkQ, helper);
                const vec2 offsetInPixels = (vec2(checkP, checkQ) - vec2(p, q)) * vec2(helper.width, helper.height);
                squaredError = dot(offsetInPixels, offsetInPixels);
            }
            toImage(s, t, helper, p, q);
            break;
        }
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
