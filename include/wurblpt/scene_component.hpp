/* scene_component.hpp -- base class of everything a Scene can take (mirror of the interface part of
 * reference libwurblpt/scene_component.hpp:40-101; OBJ export is out of scope). */
#pragma once

#include <vector>

namespace WurblPT {

class Hitable;

class SceneComponent
{
public:
    virtual ~SceneComponent() {}

    /* Called by Scene::take to create the hitables of this component. */
    virtual std::vector<Hitable*> createHitables() const { return std::vector<Hitable*>(); }
};

}
