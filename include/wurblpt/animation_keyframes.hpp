/* animation_keyframes.hpp -- key-frame animation (mirror of reference
 * libwurblpt/animation_keyframes.hpp:57-215). This is the animation type the device evaluates. */
#pragma once

#include <vector>

#include "animation.hpp"

namespace WurblPT {

class AnimationKeyframes : public Animation
{
public:
    class Keyframe
    {
    public:
        float t;
        Transformation transformation;
        Keyframe() : t(0), transformation() {}
        Keyframe(float time, const Transformation& transf) : t(time), transformation(transf) {}
    };

private:
    std::vector<Keyframe> _keyframes;   // ascending time

    /* the two key frames around t; equal indices on an exact match */
    void bracket(float t, int& lower, int& higher) const
    {
        int a = 0;
        int b = int(_keyframes.size()) - 1;
        while (b >= a) {
            int c = (a + b) / 2;
            if (_keyframes[c].t < t) {
                a = c + 1;
            } else if (_keyframes[c].t > t) {
                b = c - 1;
            } else {
                lower = higher = c;
                return;
            }
        }
        lower = b;
        higher = a;
    }

public:
    AnimationKeyframes() {}
    virtual ~AnimationKeyframes() {}

    explicit AnimationKeyframes(float t0, const Transformation& T0, float t1, const Transformation& T1)
    {
        addKeyframe(t0, T0);
        addKeyframe(t1, T1);
    }

    explicit AnimationKeyframes(const std::vector<Keyframe>& keyframes) : _keyframes(keyframes) {}

    const std::vector<Keyframe>& keyframes() const { return _keyframes; }

    void addKeyframe(const Keyframe& keyframe)
    {
        if (_keyframes.size() == 0 || keyframe.t > endTime()) {
            _keyframes.push_back(keyframe);
        } else if (keyframe.t < startTime()) {
            _keyframes.insert(_keyframes.begin(), keyframe);
        } else {
            int lower, higher;
            bracket(keyframe.t, lower, higher);
            if (lower == higher)
                _keyframes[lower] = keyframe;
            else
                _keyframes.insert(_keyframes.begin() + higher, keyframe);
        }
    }

    void addKeyframe(float time, const Transformation& transf) { addKeyframe(Keyframe(time, transf)); }

    float startTime() const { return _keyframes.size() == 0 ? 0.0f : _keyframes.front().t; }
    float endTime() const { return _keyframes.size() == 0 ? 0.0f : _keyframes.back().t; }

    virtual Transformation at(float t) const override
    {
        if (_keyframes.size() == 0)
            return Transformation();
        if (t <= startTime())
            return _keyframes.front().transformation;
        if (t >= endTime())
            return _keyframes.back().transformation;
        int lower, higher;
        bracket(t, lower, higher);
        if (lower == higher)
            return _keyframes[lower].transformation;
        float alpha = 1.0f - (_keyframes[higher].t - t) / (_keyframes[higher].t - _keyframes[lower].t);
        return mix(_keyframes[lower].transformation, _keyframes[higher].transformation, alpha);
    }
};

}
