/* sensor_rgb.hpp -- RGB sensor with optional gating by distance-to-light and optical path length
 * (mirror of reference libwurblpt/sensor_rgb.hpp:34-97). */
#pragma once

#include <limits>

#include "sensor.hpp"

namespace WurblPT {

class SensorRGB final : public Sensor
{
private:
    TGD::Array<float> _frame;
    const float _minDistToLight, _maxDistToLight;
    const float _minPathLen, _maxPathLen;

public:
    SensorRGB(unsigned int width, unsigned int height,
            float minDistToLight = 0.0f, float maxDistToLight = std::numeric_limits<float>::max(),
            float minPathLen = 0.0f, float maxPathLen = std::numeric_limits<float>::max()) :
        _frame({ width, height }, 3),
        _minDistToLight(minDistToLight), _maxDistToLight(maxDistToLight), _minPathLen(minPathLen), _maxPathLen(maxPathLen)
    {
        _frame.componentTagList(0).set("INTERPRETATION", "RED");
        _frame.componentTagList(1).set("INTERPRETATION", "GREEN");
        _frame.componentTagList(2).set("INTERPRETATION", "BLUE");
    }

    virtual unsigned int width() const override { return _frame.dimension(0); }
    virtual unsigned int height() const override { return _frame.dimension(1); }
    virtual TGD::ArrayContainer* pixelArray() override { return &_frame; }
    const TGD::Array<float>& result() const { return _frame; }

    virtual bool describe(wpt_sensor_desc& out) const override
    {
        std::memset(&out, 0, sizeof(out));
        out.type = WPT_SENSOR_RGB;
        out.width = width();
        out.height = height();
        out.min_dist_to_light = _minDistToLight;
        out.max_dist_to_light = _maxDistToLight;
        out.min_path_len = _minPathLen;
        out.max_path_len = _maxPathLen;
        return true;
    }
};

}
