/* wurblpt.hpp -- the render entry points: Parameters and mcpt() (mirror of reference
 * libwurblpt/wurblpt.hpp:79-96,279-449, same signatures), plus the convenience includes.
 *
 * mcpt() hands the prepared scene, the camera, the sensor and the parameters to wpt_render()
 * (include/wurblpt_b200.h); the pixel / sample / bounce loops of the reference run inside the
 * sm_100a render kernel. There is no host fallback. */
#pragma once

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <tgd/array.hpp>

#include "scene.hpp"
#include "camera.hpp"
#include "sampler.hpp"
#include "mpi.hpp"

// for convenience, like the reference:
#include "animation_keyframes.hpp"
#include "texture_image.hpp"
#include "envmap.hpp"
#include "fresnel.hpp"
#include "material_lambertian.hpp"
#include "material_mirror.hpp"
#include "material_glass.hpp"
#include "material_modphong.hpp"
#include "material_ggx.hpp"
#include "material_phase_function_isotropic.hpp"
#include "material_rgl.hpp"
#include "light_diffuse.hpp"
#include "light_spot.hpp"
#include "light_tof.hpp"
#include "generator.hpp"
#include "sphere.hpp"
#include "mesh.hpp"
#include "medium.hpp"
#include "postproc.hpp"
#include "sensor.hpp"
#include "sensor_rgb.hpp"
#include "sensor_tof_amcw.hpp"

namespace WurblPT {

class Parameters
{
public:
    unsigned int maxPathComponents; // 1 = sensor to light, 2 = direct illumination, 3 = one indirect bounce, ...
    float rrThreshold;              // Russian roulette starts when the path attenuation falls below this
    bool randomizeRayOverPixel;     // jitter rays over the pixel (antialiasing, ToF flying pixels)
    float minHitDistance;           // minimum distance between consecutive hits

    Parameters() : maxPathComponents(128), rrThreshold(1.0f), randomizeRayOverPixel(true), minHitDistance(0.00001f) {}
};

inline float powerHeuristicWeight(float f, float g)
{
    f *= f;
    g *= g;
    return (f + g > 0.0f ? f / (f + g) : 0.0f);
}

/* statistics of the most recent mcpt() call of this thread (B200 build only) */
inline wpt_stats& lastRenderStats()
{
    static thread_local wpt_stats stats;
    return stats;
}

inline void mcpt(
        MPICoordinator& mpiCoordinator,
        Sensor& sensor,
        const Camera& camera,
        const Scene& scene,
        unsigned int samplesSqrt,
        float t0 = 0.0f, float t1 = 0.0f,
        const Parameters& params = Parameters())
{
    /* the reference asserts !scene.bvhNeedsUpdate(t0, t1) (wurblpt.hpp:289) */
    if (scene.bvhNeedsUpdate(t0, t1) || !scene.deviceScene())
        detail::die("mcpt", "Scene::updateBVH(t0, t1) must be called before rendering");

    wpt_sensor_desc sensorDesc;
    if (!sensor.describe(sensorDesc))
        detail::die("mcpt", "a user-defined Sensor cannot accumulate radiance on the device");
    wpt_camera_desc cameraDesc;
    std::vector<wpt_keyframe> cameraKeyframes;
    if (!camera.flatten(cameraDesc, cameraKeyframes))
        detail::die("mcpt", "only AnimationKeyframes camera animations can be evaluated on the device");
    wpt_params p;
    p.max_path_components = params.maxPathComponents;
    p.rr_threshold = params.rrThreshold;
    p.randomize_ray_over_pixel = params.randomizeRayOverPixel;
    p.min_hit_distance = params.minHitDistance;

    TGD::ArrayContainer* pixelArray = sensor.pixelArray();
    const unsigned int width = sensor.width();
    const unsigned int height = sensor.height();
    const unsigned int samples = samplesSqrt * samplesSqrt;
    fprintf(stderr, "Number of hitables that are hot spots: %zu\n", scene.hotSpots().size());
    fprintf(stderr, "Rendering %ux%u pixels with %u samples for %.3fs-%.3fs.\n", width, height, samples, t0, t1);
    fprintf(stderr, "%s: rendering block of size %u starting at %u\n", mpiCoordinator.processId(), width * height, 0u);

    auto start = std::chrono::steady_clock::now();
    wpt_stats& stats = lastRenderStats();
    if (wpt_render(scene.deviceScene(), &cameraDesc, &sensorDesc, &p, samplesSqrt, t0, t1, nullptr,
                static_cast<float*>(pixelArray->data()), &stats) != WPT_OK)
        detail::die("mcpt", wpt_last_error());
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();

    fprintf(stderr, "%s: progress: 100.0%%\n", mpiCoordinator.processId());
    fprintf(stderr, "%s is done\n", mpiCoordinator.processId());

    if (mpiCoordinator.mainProcess()) {
        pixelArray->globalTagList().set("WURBLPT/COMPILER", "nvcc sm_100a (" + std::string(wpt_version()) + ")");
        pixelArray->globalTagList().set("WURBLPT/SAMPLES_PER_PIXEL", std::to_string(samples));
        pixelArray->globalTagList().set("WURBLPT/MAX_PATH_COMPONENTS", std::to_string(params.maxPathComponents));
        pixelArray->globalTagList().set("WURBLPT/RUSSIAN_ROULETTE_THRESHOLD", std::to_string(params.rrThreshold));
        /* kept under the reference's tag name so that tools reading it keep working; it is the wall
         * time of mcpt() in seconds (the reference stores process CPU time, wurblpt.hpp:393-400) */
        pixelArray->globalTagList().set("WURBLPT/CPU_SECONDS", std::to_string(seconds));
        pixelArray->globalTagList().set("WURBLPT/GPU_KERNEL_SECONDS", std::to_string(stats.kernel_ms * 1e-3));
        pixelArray->globalTagList().set("WURBLPT/RAYS", std::to_string(stats.rays));
    }
}

/* convenience wrapper for programs that do not use MPI */
inline void mcpt(
        Sensor& sensor,
        const Camera& camera,
        const Scene& scene,
        unsigned int samplesSqrt,
        float t0 = 0.0f, float t1 = 0.0f,
        const Parameters& params = Parameters())
{
    MPICoordinator mpiCoordinator;
    mcpt(mpiCoordinator, sensor, camera, scene, samplesSqrt, t0, t1, params);
}

}
