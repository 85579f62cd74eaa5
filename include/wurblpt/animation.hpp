/* animation.hpp -- Animation interface and AnimationCache (mirror of reference
 * libwurblpt/animation.hpp:37-125). On the device, animations are evaluated per sample from
 * key frames; the cache class is kept for API compatibility and host-side queries. */
#pragma once

#include <limits>
#include <vector>

#include "transformation.hpp"

namespace WurblPT {

class Animation
{
public:
    Animation() {}
    virtual ~Animation() {}
    virtual Transformation at(float t) const = 0;
};

namespace detail {

/* Times at which an Animation that is not an AnimationKeyframes is sampled ("baked") into key frames for the
 * device: exact for t0 == t1 (one key frame); 17 samples across an exposure interval, between which the device
 * interpolates like AnimationKeyframes does (translation and scaling linearly, rotation by slerp). */
inline std::vector<float> bakeTimesFor(float t0, float t1)
{
    std::vector<float> times;
    if (!(t1 > t0)) {
        times.push_back(t0);
    } else {
        const int n = 16;
        for (int i = 0; i <= n; i++)
            times.push_back(i == n ? t1 : t0 + (t1 - t0) * (float(i) / float(n)));
    }
    return times;
}

}

class AnimationCache
{
private:
    const std::vector<const Animation*>* _animations;
    float _t;
    std::vector<Transformation> _transformations;
    std::vector<mat4> _matrices;
    std::vector<mat3> _normalMatrices;
    std::vector<bool> _valid;

    void ensure(int i)
    {
        if (!_valid[i]) {
            _transformations[i] = (*_animations)[i]->at(_t);
            _matrices[i] = _transformations[i].toMat4();
            _normalMatrices[i] = _transformations[i].toNormalMatrix();
            _valid[i] = true;
        }
    }

public:
    AnimationCache() : _animations(nullptr), _t(std::numeric_limits<float>::max()) {}

    AnimationCache(const std::vector<const Animation*>& animations) :
        _animations(&animations), _t(std::numeric_limits<float>::max()),
        _transformations(animations.size()), _matrices(animations.size()),
        _normalMatrices(animations.size()), _valid(animations.size(), false)
    {
    }

    AnimationCache(const std::vector<const Animation*>& animations, float t) : AnimationCache(animations)
    {
        _t = t;
        for (size_t i = 0; i < _animations->size(); i++)
            ensure(i);
    }

    void init(float t)
    {
        _t = t;
        for (size_t i = 0; i < _valid.size(); i++)
            _valid[i] = false;
    }

    const Transformation& get(int i) { ensure(i); return _transformations[i]; }
    const mat4& getM(int i) { ensure(i); return _matrices[i]; }
    const mat3& getN(int i) { ensure(i); return _normalMatrices[i]; }
    float t() const { return _t; }
};

}
