/* sensor_tof_amcw.hpp -- amplitude-modulated continuous-wave time-of-flight sensor (mirror of
 * reference libwurblpt/sensor_tof_amcw.hpp:54-267; model: Lambers, Hoberg, Kolb, IEEE Sensors J.
 * 15(7), 2015). The per-sample energy split (A, B, total) runs on the device; phase() and result(),
 * which turn four phase images into a distance image, stay on the host as in the reference. */
#pragma once

#include <random>
#include <string>

#include "sensor.hpp"
#include "material.hpp"
#include "constants.hpp"

namespace WurblPT {

class SensorTofAmcw final : public Sensor
{
public:
    constexpr static float dutyCycle = 0.5f;

    unsigned int phaseImageCount;   // multiple of 4
    float wavelength;               // nm
    double modulationFrequency;     // Hz
    float exposureTime;             // per phase image, microseconds
    float readoutTime;              // per phase image, microseconds
    float pauseTime;                // after all phase images, microseconds
    float pixelArea;                // square micrometers
    float contrast;                 // [0,1]
    float quantumEfficiency;
    int maxElectrons;

private:
    int _phaseImageIndex;
    TGD::Array<float> _frame;

public:
    SensorTofAmcw(unsigned int width, unsigned int height) :
        phaseImageCount(4), wavelength(880.0f), modulationFrequency(10e6), exposureTime(1000.0f),
        readoutTime(1000.0f), pauseTime(42000.0f), pixelArea(12.0f * 12.0f), contrast(0.75f),
        quantumEfficiency(0.8f), maxElectrons(100000), _phaseImageIndex(0), _frame({ width, height }, 3)
    {
        _frame.componentTagList(0).set("INTERPRETATION", "a");
        _frame.componentTagList(1).set("INTERPRETATION", "b");
        _frame.componentTagList(2).set("INTERPRETATION", "total");
    }

    void setPhaseIndex(int i) { _phaseImageIndex = i; }
    float tau(unsigned int phaseImageIndex) const { return phaseImageIndex * (2.0f * pi) / phaseImageCount; }
    float fracModfreqC() const { return modulationFrequency / speedOfLight; }
    float fracCModfreq() const { return speedOfLight / modulationFrequency; }

    void setPauseTimeForFPS(float fps)
    {
        float totalFrameDuration = 1e6f / fps;
        float totalPhaseImageDurations = phaseImageCount * (exposureTime + readoutTime);
        pauseTime = totalFrameDuration - totalPhaseImageDurations;
    }
    float frameDuration() const { return (phaseImageCount * (exposureTime + readoutTime) + pauseTime) / 1e6f; }
    float fps() const { return 1.0f / frameDuration(); }
    float phaseImageDuration() const { return (exposureTime + readoutTime) / 1e6f; }

    const TGD::Array<float>& energy() const { return _frame; }

    /* energies -> electrons (+ optional shot noise) -> normalised digital numbers (a-b, a+b, a, b) */
    TGD::Array<float> phase(float shotNoiseFactor, unsigned long long prngSeed = 42) const
    {
        std::mt19937_64 generator(prngSeed);
        std::normal_distribution<float> gaussian(0.0f, 1.0f);
        TGD::Array<float> phaseImage({ _frame.dimension(0), _frame.dimension(1) }, 4);
        phaseImage.globalTagList() = _frame.globalTagList();
        phaseImage.componentTagList(0).set("INTERPRETATION", "a-b");
        phaseImage.componentTagList(1).set("INTERPRETATION", "a+b");
        phaseImage.componentTagList(2).set("INTERPRETATION", "a");
        phaseImage.componentTagList(3).set("INTERPRETATION", "b");
        vec2 maxElectronsVec(maxElectrons, maxElectrons);
        for (size_t i = 0; i < phaseImage.elementCount(); i++) {
            vec2 energy = vec2(_frame.get<float>(i));
            vec2 electrons = quantumEfficiency * wavelength * energy / hc / 10000.0f;
            vec2 shotNoise = sqrt(electrons) * vec2(gaussian(generator), gaussian(generator));
            electrons += shotNoiseFactor * shotNoise;
            vec2 digNums = clamp(electrons, vec2(0.0f), maxElectronsVec) / maxElectronsVec;
            phaseImage.set<float>(i, { digNums.x() - digNums.y(), digNums.x() + digNums.y(), digNums.x(), digNums.y() });
        }
        return phaseImage;
    }

    /* four phase images -> (distance, amplitude, intensity, phase shift) */
    TGD::Array<float> result(const TGD::Array<float>* phases) const
    {
        const int i0 = 0, i90 = 1 * phaseImageCount / 4, i180 = 2 * phaseImageCount / 4, i270 = 3 * phaseImageCount / 4;
        TGD::Array<float> r(phases[0].dimensions(), 4);
        r.globalTagList() = phases[0].globalTagList();
        std::string seconds;
        seconds = phases[i0].globalTagList().value("WURBLPT/CPU_SECONDS", "unknown") + '+';
        seconds += phases[i90].globalTagList().value("WURBLPT/CPU_SECONDS", "unknown") + '+';
        seconds += phases[i180].globalTagList().value("WURBLPT/CPU_SECONDS", "unknown") + '+';
        seconds += phases[i270].globalTagList().value("WURBLPT/CPU_SECONDS", "unknown");
        r.globalTagList().set("WURBLPT/CPU_SECONDS", seconds);
        r.componentTagList(0).set("INTERPRETATION", "distance");
        r.componentTagList(1).set("INTERPRETATION", "amplitude");
        r.componentTagList(2).set("INTERPRETATION", "intensity");
        r.componentTagList(3).set("INTERPRETATION", "phase_shift");
        for (size_t i = 0; i < r.elementCount(); i++) {
            float D[4] = { phases[i0].get<float>(i, 0), phases[i90].get<float>(i, 0),
                phases[i180].get<float>(i, 0), phases[i270].get<float>(i, 0) };
            float phaseShift, distance;
            if (abs(D[0] - D[2]) <= 0.0f && abs(D[1] - D[3]) <= 0.0f) {
                phaseShift = 0.0f;
                distance = 0.0f;
            } else {
                phaseShift = atan(D[3] - D[1], D[0] - D[2]);
                if (phaseShift < 0.0f)
                    phaseShift += 2.0 * pi;
                distance = fracCModfreq() * phaseShift * 0.25f * inv_pi;
            }
            float amplitude = sqrt((D[0] - D[2]) * (D[0] - D[2]) + (D[1] - D[3]) * (D[1] - D[3])) * pi_2;
            float intensity = 0.5f * (D[0] + D[1] + D[2] + D[3]);
            r.set<float>(i, { distance, amplitude, intensity, phaseShift });
        }
        return r;
    }

    virtual unsigned int width() const override { return _frame.dimension(0); }
    virtual unsigned int height() const override { return _frame.dimension(1); }
    virtual TGD::ArrayContainer* pixelArray() override { return &_frame; }

    virtual bool describe(wpt_sensor_desc& out) const override
    {
        std::memset(&out, 0, sizeof(out));
        out.type = WPT_SENSOR_TOF_AMCW;
        out.width = width();
        out.height = height();
        out.pixel_area = pixelArea;
        out.exposure_time = exposureTime;
        out.contrast = contrast;
        out.phase_index = _phaseImageIndex;
        out.phase_count = phaseImageCount;
        out.modulation_frequency = modulationFrequency;
        return true;
    }
};

}
