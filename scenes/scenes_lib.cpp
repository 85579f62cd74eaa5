/*
 * scenes/scenes_lib.cpp -- the benchmark scene set (scenes/configs.hpp) built with this
 * repository's WurblPT API mirror (include/wurblpt/) and exposed to Python as flat descriptions.
 * This is the PRODUCT path for bench.py: it touches nothing under oracle/.
 */
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <tgd/array.hpp>
#include <tgd/io.hpp>
#include <wurblpt/wurblpt.hpp>

#include "configs.hpp"

using namespace WurblPT;

namespace {

struct Entry {
    Scene scene;
    wpt_scene_desc desc;
};

std::vector<std::unique_ptr<Entry>> g_entries;
std::string g_error;

}

extern "C" {

const char* wpts_last_error(void)
{
    return g_error.c_str();
}

/* Build a named scene with the API mirror and flatten it. Returns a handle or -1. */
int wpts_scene_create(const char* name, float scale)
{
    std::unique_ptr<Entry> e(new Entry);
    if (!wptcfg::buildScene(name, e->scene, scale)) {
        g_error = std::string("unknown scene '") + name + "'";
        return -1;
    }
    e->desc = e->scene.describe();
    g_entries.push_back(std::move(e));
    return int(g_entries.size()) - 1;
}

const wpt_scene_desc* wpts_scene_desc(int h)
{
    if (h < 0 || size_t(h) >= g_entries.size() || !g_entries[h])
        return nullptr;
    return &g_entries[h]->desc;
}

void wpts_scene_destroy(int h)
{
    if (h >= 0 && size_t(h) < g_entries.size())
        g_entries[h].reset();
}

/* Default camera / sensor / Parameters / sampling of a named config (scenes/configs.hpp). */
int wpts_default_view(const char* name, wpt_camera_desc* cam, wpt_sensor_desc* sensor, wpt_params* params,
        uint32_t* samplesSqrt, float* t0, float* t1)
{
    static std::vector<wpt_keyframe> keyframes;
    wptcfg::View v = wptcfg::defaultView(name);
    Camera camera = wptcfg::makeCamera(v, float(v.width) / v.height);
    if (!camera.flatten(*cam, keyframes))
        return -1;
    if (v.tof) {
        SensorTofAmcw s(v.width, v.height);
        s.describe(*sensor);
    } else {
        SensorRGB s(v.width, v.height);
        s.describe(*sensor);
    }
    params->max_path_components = v.params.maxPathComponents;
    params->rr_threshold = v.params.rrThreshold;
    params->randomize_ray_over_pixel = v.params.randomizeRayOverPixel;
    params->min_hit_distance = v.params.minHitDistance;
    *samplesSqrt = v.samplesSqrt;
    *t0 = v.t0;
    *t1 = v.t1;
    return 0;
}

}
