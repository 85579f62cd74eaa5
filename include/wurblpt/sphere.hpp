/* sphere.hpp -- analytic sphere component (mirror of reference libwurblpt/sphere.hpp:36-86 and of the
 * data held by HitableSphere, hitable_sphere.hpp:33-86). */
#pragma once

#include "scene_component.hpp"
#include "hitable.hpp"
#include "material.hpp"
#include "generator.hpp"

namespace WurblPT {

class Sphere : public SceneComponent
{
private:
    const Material* _material;
    const Transformation _transformation;
    const int _animationIndex;

public:
    Sphere(const Material* material, const Transformation& T) : _material(material), _transformation(T), _animationIndex(-1) {}
    Sphere(const Material* material, int animationIndex = -1) : _material(material), _transformation(), _animationIndex(animationIndex) {}
    Sphere(const vec3& center, float radius, const Material* material) :
        Sphere(material, Transformation(center, quat::null(), vec3(radius)))
    {
    }

    const Material* material() const { return _material; }
    const Transformation& transformation() const { return _transformation; }
    int animationIndex() const { return _animationIndex; }

    virtual std::vector<Hitable*> createHitables() const override
    {
        return std::vector<Hitable*>(1, new Hitable(Hitable::KindSphere, this, 0, _material, _animationIndex));
    }

    /* HitableSphere(T, material, ai): center = T.translation, radius = max(T.scaling), rotation = T.rotation */
    void flatten(Flattener& F, wpt_sphere& out) const
    {
        std::memset(&out, 0, sizeof(out));
        detail::store(out.center, _transformation.translation, 3);
        out.radius = max(_transformation.scaling);
        out.rotation[0] = _transformation.rotation.x;
        out.rotation[1] = _transformation.rotation.y;
        out.rotation[2] = _transformation.rotation.z;
        out.rotation[3] = _transformation.rotation.w;
        out.material = F.material(_material);
        out.animation = _animationIndex;
    }
};

}
