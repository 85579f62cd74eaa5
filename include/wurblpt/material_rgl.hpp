/* material_rgl.hpp -- measured RGL/EPFL BRDFs are OUT OF SCOPE of this build (they need .bsdf data
 * files and the third-party powitacq evaluator). The class exists so that
 * applications that merely mention it (reference wurblpt-cornellbox.cpp:61) still compile; using one
 * in a rendered scene is reported as an error when the scene is prepared. */
#pragma once

#include <string>

#include "material.hpp"

namespace WurblPT {

class MaterialRGL final : public Material
{
public:
    const std::string fileName;

    MaterialRGL(const std::string& fileName_) : fileName(fileName_) {}

    virtual bool flatten(Flattener& F, wpt_material&) const override
    {
        F.fail("MaterialRGL (" + fileName + ") is not supported by the B200 build");
        return false;
    }
};

}
