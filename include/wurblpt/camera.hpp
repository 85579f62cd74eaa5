/* camera.hpp -- Camera (mirror of reference libwurblpt/camera.hpp:42-218: same constructors and
 * public members; getRay() is kept for host-side use, rendering generates rays on the device from
 * the flattened description). */
#pragma once

#include <memory>
#include <vector>

#include "transformation.hpp"
#include "optics.hpp"
#include "animation.hpp"
#include "animation_keyframes.hpp"
#include "ray.hpp"
#include "prng.hpp"
#include "constants.hpp"
#include "flatten.hpp"

namespace WurblPT {

class Camera
{
public:
    enum SurroundMode { Surround_Off, Surround_180, Surround_360 };

    SurroundMode surroundMode;
    float stereoscopicDistance;
    Optics optics;
    Transformation transformation;
    std::unique_ptr<const Animation> animation;

    class RayHelper
    {
    public:
        Transformation cameraAtT0;
        LensDistortion::Helper lensDistortionHelper;
    };

    /* Either a transformation or an animation (which the camera then owns). Surround cameras ignore
     * the optics; stereoscopic cameras store the left view in the upper half of the image. */
    Camera(SurroundMode surroundMode_, float stereoscopicDistance_, const Optics& optics_, const Transformation& transformation_ = Transformation()) :
        surroundMode(surroundMode_), stereoscopicDistance(stereoscopicDistance_), optics(optics_), transformation(transformation_), animation(nullptr)
    {
    }
    Camera(SurroundMode surroundMode_, float stereoscopicDistance_, const Optics& optics_, const Animation* animation_) :
        surroundMode(surroundMode_), stereoscopicDistance(stereoscopicDistance_), optics(optics_), transformation(), animation(animation_)
    {
    }
    Camera(const Optics& optics_, const Transformation& transformation_ = Transformation()) :
        surroundMode(Surround_Off), stereoscopicDistance(0.0f), optics(optics_), transformation(transformation_), animation(nullptr)
    {
    }
    Camera(const Optics& optics_, const Animation* animation_) :
        surroundMode(Surround_Off), stereoscopicDistance(0.0f), optics(optics_), transformation(), animation(animation_)
    {
    }

    Transformation at(float t = 0.0f) const { return (animation.get() ? animation->at(t) : transformation); }

    RayHelper getRayHelper(float t0, unsigned int width, unsigned int height) const
    {
        RayHelper rh;
        rh.cameraAtT0 = at(t0);
        rh.lensDistortionHelper = optics.distortion.getHelper(optics.projection, width, height);
        return rh;
    }

    /* p, q in [0,1]^2 with q = 0 at the bottom */
    Ray getRay(float p, float q, float t0, float t1, const RayHelper& rayHelper, Prng& prng, bool withRandomness = true) const
    {
        float stereoscopicShift = 0.0f;
        if (stereoscopicDistance > 0.0f) {
            q *= 2.0f;
            if (q < 1.0f) {
                stereoscopicShift = -0.5f * stereoscopicDistance;
            } else {
                q -= 1.0f;
                stereoscopicShift = +0.5f * stereoscopicDistance;
            }
        }
        vec3 O, D;
        if (surroundMode == Surround_Off) {
            optics.distortion.undistort(p, q, rayHelper.lensDistortionHelper);
            vec3 P = vec3(mix(optics.projection.l, optics.projection.r, p), mix(optics.projection.b, optics.projection.t, q), -1.0f);
            O = vec3(0.0f);
            if (withRandomness)
                O = optics.depthOfField.offset(P, prng);
            D = P - O;
            O += vec3(stereoscopicShift, 0.0f, 0.0f);
        } else {
            float lon = (2.0f * p - 1.0f) * pi;
            if (surroundMode == Surround_180)
                lon *= 0.5f;
            float lat = (q - 0.5f) * pi;
            D = vec3(cos(lat) * sin(lon), sin(lat), -cos(lat) * cos(lon));
            O = vec3(-cos(lon), 0.0f, -sin(lon)) * stereoscopicShift;
        }
        float t = t0;
        Transformation T = rayHelper.cameraAtT0;
        if (t0 != t1 && withRandomness) {
            t += prng.in01() * (t1 - t0);
            T = at(t);
        }
        vec3 origin = T * O;
        vec3 direction = T.rotation * D;
        return Ray(origin, normalize(direction), t, refractiveIndexOfVacuum);
    }

    /* ---- B200 build only: the C-ABI description; `keyframeStorage` must outlive its use ---- */
    /* `bakeTimes`: where to sample a camera animation that is not an AnimationKeyframes (detail::bakeTimesFor) */
    bool flatten(wpt_camera_desc& c, std::vector<wpt_keyframe>& keyframeStorage,
            const std::vector<float>& bakeTimes = std::vector<float>()) const
    {
        std::memset(&c, 0, sizeof(c));
        c.surround_mode = (surroundMode == Surround_Off ? WPT_SURROUND_OFF : surroundMode == Surround_180 ? WPT_SURROUND_180 : WPT_SURROUND_360);
        c.stereoscopic_distance = stereoscopicDistance;
        c.l = optics.projection.l; c.r = optics.projection.r; c.b = optics.projection.b; c.t = optics.projection.t;
        c.distortion_type = int(optics.distortion.type);
        c.k1 = optics.distortion.k1; c.k2 = optics.distortion.k2; c.k3 = optics.distortion.k3;
        c.p1 = optics.distortion.p1; c.p2 = optics.distortion.p2;
        c.b1 = optics.distortion.b1; c.b2 = optics.distortion.b2; c.b3 = optics.distortion.b3; c.b4 = optics.distortion.b4;
        c.lens_radius = optics.depthOfField.lensRadius;
        c.focus_dist = optics.depthOfField.focusDist;
        detail::store(c.pose_translation, transformation.translation, 3);
        c.pose_rotation[0] = transformation.rotation.x; c.pose_rotation[1] = transformation.rotation.y;
        c.pose_rotation[2] = transformation.rotation.z; c.pose_rotation[3] = transformation.rotation.w;
        detail::store(c.pose_scaling, transformation.scaling, 3);
        keyframeStorage.clear();
        if (animation.get()) {
            const AnimationKeyframes* kf = dynamic_cast<const AnimationKeyframes*>(animation.get());
            if (!kf && bakeTimes.empty())
                return false;
            std::vector<AnimationKeyframes::Keyframe> baked;
            if (!kf) {
                for (float t : bakeTimes)
                    baked.push_back(AnimationKeyframes::Keyframe(t, animation->at(t)));
            }
            for (const auto& k : (kf ? kf->keyframes() : baked)) {
                wpt_keyframe wk;
                wk.t = k.t;
                detail::store(wk.translation, k.transformation.translation, 3);
                wk.rotation[0] = k.transformation.rotation.x; wk.rotation[1] = k.transformation.rotation.y;
                wk.rotation[2] = k.transformation.rotation.z; wk.rotation[3] = k.transformation.rotation.w;
                detail::store(wk.scaling, k.transformation.scaling, 3);
                keyframeStorage.push_back(wk);
            }
            if (keyframeStorage.empty()) {
                /* an empty key-frame list means the default pose (animation_keyframes.hpp:194-195) */
                Transformation T;
                detail::store(c.pose_translation, T.translation, 3);
                c.pose_rotation[0] = c.pose_rotation[1] = c.pose_rotation[2] = 0.0f; c.pose_rotation[3] = 1.0f;
                detail::store(c.pose_scaling, T.scaling, 3);
            }
        }
        c.n_keyframes = keyframeStorage.size();
        c.keyframes = keyframeStorage.data();
        return true;
    }
};

}
