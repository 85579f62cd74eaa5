/* hitable.hpp -- Hitable handles (the API-visible part of reference libwurblpt/hitable.hpp:39-165).
 *
 * In the reference a Hitable is a virtual object that intersects rays on the CPU. Here intersection
 * happens on the device; what the API still needs is the *identity* of every hitable, because
 * applications pass the vectors returned by Scene::take() on (e.g. as Medium boundaries) and the
 * renderer reports hitables by id. A Hitable is therefore a light-weight handle.
 */
#pragma once

#include "aabb.hpp"
#include "prng.hpp"
#include "animation.hpp"

namespace WurblPT {

class Material;
class SceneComponent;
class Scene;

class HitRecord
{
public:
    bool haveHit;
    float a;
    vec3 position, normal, tangent;
    vec2 texcoords;
    bool backside;
    const class Hitable* hitable;

    HitRecord() : haveHit(false) {}
    HitRecord(float a_, const vec3& p, const vec3& n, const vec3& t, const vec2& tc, bool bs, const Hitable* h) :
        haveHit(true), a(a_), position(p), normal(n), tangent(t), texcoords(tc), backside(bs), hitable(h)
    {
    }
};

class Hitable
{
public:
    enum Kind { KindTriangle, KindSphere, KindMedium };

private:
    Kind _kind;
    const SceneComponent* _component;
    unsigned int _index;            // triangle number within its mesh instance
    const Material* _material;
    int _animationIndex;
    mutable int _globalId;          // assigned by the Scene that takes the hitable
    friend class Scene;

public:
    Hitable(Kind kind, const SceneComponent* component, unsigned int index, const Material* material, int animationIndex) :
        _kind(kind), _component(component), _index(index), _material(material), _animationIndex(animationIndex), _globalId(-1)
    {
    }
    virtual ~Hitable() {}

    Kind kind() const { return _kind; }
    const SceneComponent* component() const { return _component; }
    unsigned int index() const { return _index; }
    int globalId() const { return _globalId; }
    virtual const Material* material() const { return _material; }
    virtual int animationIndex() const { return _animationIndex; }
};

}
