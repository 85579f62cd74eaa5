/* medium.hpp -- homogeneous participating medium bounded by hitables (mirror of reference
 * libwurblpt/medium.hpp:37-57 / hitable_medium.hpp:36-57). */
#pragma once

#include <vector>

#include "scene_component.hpp"
#include "hitable.hpp"
#include "material.hpp"

namespace WurblPT {

class Medium final : public SceneComponent
{
private:
    const std::vector<const Hitable*> _boundaryHitables;
    const float _density;
    const Material* _phaseFunction;

public:
    Medium(const std::vector<const Hitable*>& boundary, float density, const Material* phaseFunction) :
        _boundaryHitables(boundary), _density(density), _phaseFunction(phaseFunction)
    {
    }

    const std::vector<const Hitable*>& boundary() const { return _boundaryHitables; }
    float density() const { return _density; }
    const Material* phaseFunction() const { return _phaseFunction; }

    virtual std::vector<Hitable*> createHitables() const override
    {
        return std::vector<Hitable*>(1, new Hitable(Hitable::KindMedium, this, 0, _phaseFunction, -1));
    }
};

}
