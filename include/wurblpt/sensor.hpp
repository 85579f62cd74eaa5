/* sensor.hpp -- Sensor interface (mirror of reference libwurblpt/sensor.hpp:36-82). Radiance is
 * accumulated on the device; a sensor here owns the frame and describes itself to the C ABI. A
 * user-defined Sensor subclass cannot run on the device: describe() returns false and mcpt() aborts. */
#pragma once

#include <tgd/array.hpp>

#include "hitable.hpp"
#include "ray.hpp"
#include "flatten.hpp"

namespace WurblPT {

class Sensor
{
public:
    constexpr static int maxPixelComponents = 3;

    virtual ~Sensor() {}
    virtual unsigned int width() const { return 0; }
    virtual unsigned int height() const { return 0; }
    virtual TGD::ArrayContainer* pixelArray() { return nullptr; }
    virtual float aspectRatio() const { return float(width()) / height(); }

    /* fill the C-ABI sensor description */
    virtual bool describe(wpt_sensor_desc& /* out */) const { return false; }
};

}
