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
#include "texture_noise.hpp"
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
#include "import.hpp"

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

    scene.prepareRender(t0, t1);
    wpt_sensor_desc sensorDesc;
    if (!sensor.describe(sensorDesc))
        detail::die("mcpt", "a user-defined Sensor cannot accumulate radiance on the device");
    wpt_camera_desc cameraDesc;
    std::vector<wpt_keyframe> cameraKeyframes;
    if (!camera.flatten(cameraDesc, cameraKeyframes, detail::bakeTimesFor(t0, t1)))
        detail::die("mcpt", "the camera animation cannot be evaluated on the device");
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

/* Ground truth images (reference wurblpt.hpp:453-769): per pixel, the hit record of the pixel-centre
 * ray. Same class layout (public members and bit constants) as the reference; filled by one device
 * kernel through wpt_get_ground_truth(). */
class GroundTruth
{
private:
    struct ImageSpec {
        TGD::Array<float> GroundTruth::* member;
        unsigned int components;
        const char* tags[3];
    };
    static const ImageSpec* specs()
    {
        /* in bit order; the tags are the reference's INTERPRETATION strings */
        static const ImageSpec table[WPT_GT_FLOAT_IMAGES] = {
            { &GroundTruth::worldSpacePositions, 3, { "x", "y", "z" } },
            { &GroundTruth::worldSpaceGeometryNormals, 3, { "nx", "ny", "nz" } },
            { &GroundTruth::worldSpaceGeometryTangents, 3, { "tx", "ty", "tz" } },
            { &GroundTruth::worldSpaceMaterialNormals, 3, { "nx", "ny", "nz" } },
            { &GroundTruth::worldSpaceMaterialTangents, 3, { "tx", "ty", "tz" } },
            { &GroundTruth::cameraSpacePositions, 3, { "x", "y", "z" } },
            { &GroundTruth::cameraSpaceGeometryNormals, 3, { "nx", "ny", "nz" } },
            { &GroundTruth::cameraSpaceGeometryTangents, 3, { "tx", "ty", "tz" } },
            { &GroundTruth::cameraSpaceMaterialNormals, 3, { "nx", "ny", "nz" } },
            { &GroundTruth::cameraSpaceMaterialTangents, 3, { "tx", "ty", "tz" } },
            { &GroundTruth::cameraSpaceDepths, 1, { "depth", nullptr, nullptr } },
            { &GroundTruth::cameraSpaceDistances, 1, { "distance", nullptr, nullptr } },
            { &GroundTruth::texCoords, 2, { "u", "v", nullptr } },
            { &GroundTruth::worldSpaceOffsetToPrev, 3, { "dx", "dy", "dz" } },
            { &GroundTruth::worldSpaceOffsetToNext, 3, { "dx", "dy", "dz" } },
            { &GroundTruth::cameraSpaceOffsetToPrev, 3, { "dx", "dy", "dz" } },
            { &GroundTruth::cameraSpaceOffsetToNext, 3, { "dx", "dy", "dz" } },
            { &GroundTruth::pixelSpaceOffsetToPrev, 2, { "dx", "dy", nullptr } },
            { &GroundTruth::pixelSpaceOffsetToNext, 2, { "dx", "dy", nullptr } },
        };
        return table;
    }

public:
    static constexpr unsigned int WorldSpacePositions         = WPT_GT_WORLD_POSITIONS;
    static constexpr unsigned int WorldSpaceGeometryNormals   = WPT_GT_WORLD_GEOMETRY_NORMALS;
    static constexpr unsigned int WorldSpaceGeometryTangents  = WPT_GT_WORLD_GEOMETRY_TANGENTS;
    static constexpr unsigned int WorldSpaceMaterialNormals   = WPT_GT_WORLD_MATERIAL_NORMALS;
    static constexpr unsigned int WorldSpaceMaterialTangents  = WPT_GT_WORLD_MATERIAL_TANGENTS;
    static constexpr unsigned int CameraSpacePositions        = WPT_GT_CAMERA_POSITIONS;
    static constexpr unsigned int CameraSpaceGeometryNormals  = WPT_GT_CAMERA_GEOMETRY_NORMALS;
    static constexpr unsigned int CameraSpaceGeometryTangents = WPT_GT_CAMERA_GEOMETRY_TANGENTS;
    static constexpr unsigned int CameraSpaceMaterialNormals  = WPT_GT_CAMERA_MATERIAL_NORMALS;
    static constexpr unsigned int CameraSpaceMaterialTangents = WPT_GT_CAMERA_MATERIAL_TANGENTS;
    static constexpr unsigned int CameraSpaceDepths           = WPT_GT_CAMERA_DEPTHS;
    static constexpr unsigned int CameraSpaceDistances        = WPT_GT_CAMERA_DISTANCES;
    static constexpr unsigned int TexCoords                   = WPT_GT_TEXCOORDS;
    static constexpr unsigned int WorldSpaceOffsetToPrev      = WPT_GT_WORLD_OFFSET_TO_PREV;
    static constexpr unsigned int WorldSpaceOffsetToNext      = WPT_GT_WORLD_OFFSET_TO_NEXT;
    static constexpr unsigned int CameraSpaceOffsetToPrev     = WPT_GT_CAMERA_OFFSET_TO_PREV;
    static constexpr unsigned int CameraSpaceOffsetToNext     = WPT_GT_CAMERA_OFFSET_TO_NEXT;
    static constexpr unsigned int PixelSpaceOffsetToPrev      = WPT_GT_PIXEL_OFFSET_TO_PREV;
    static constexpr unsigned int PixelSpaceOffsetToNext      = WPT_GT_PIXEL_OFFSET_TO_NEXT;
    static constexpr unsigned int Materials                   = WPT_GT_MATERIALS;
    static constexpr unsigned int All                         = WPT_GT_ALL;

    unsigned int bits;
    TGD::Array<float> worldSpacePositions, worldSpaceGeometryNormals, worldSpaceGeometryTangents;
    TGD::Array<float> worldSpaceMaterialNormals, worldSpaceMaterialTangents;
    TGD::Array<float> cameraSpacePositions, cameraSpaceGeometryNormals, cameraSpaceGeometryTangents;
    TGD::Array<float> cameraSpaceMaterialNormals, cameraSpaceMaterialTangents;
    TGD::Array<float> cameraSpaceDepths;        /* == -cameraSpacePosition.z */
    TGD::Array<float> cameraSpaceDistances;     /* == length(cameraSpacePosition) */
    TGD::Array<float> texCoords;
    TGD::Array<float> worldSpaceOffsetToPrev, worldSpaceOffsetToNext;
    TGD::Array<float> cameraSpaceOffsetToPrev, cameraSpaceOffsetToNext;
    TGD::Array<float> pixelSpaceOffsetToPrev, pixelSpaceOffsetToNext;
    TGD::Array<int> materials;

    GroundTruth() : bits(0)
    {
    }

    GroundTruth(unsigned int width, unsigned int height, unsigned int bits = All) : bits(bits)
    {
        const ImageSpec* table = specs();
        for (int i = 0; i < WPT_GT_FLOAT_IMAGES; i++) {
            if (!(bits & (1u << i)))
                continue;
            TGD::Array<float>& image = this->*(table[i].member);
            image = TGD::Array<float>({ width, height }, table[i].components);
            for (unsigned int c = 0; c < table[i].components; c++)
                image.componentTagList(c).set("INTERPRETATION", table[i].tags[c]);
        }
        if (bits & Materials) {
            materials = TGD::Array<int>({ width, height }, 1);
            materials.componentTagList(0).set("INTERPRETATION", "material");
        }
    }

    /* the C-ABI view of the selected images */
    wpt_ground_truth describe()
    {
        wpt_ground_truth out;
        const ImageSpec* table = specs();
        for (int i = 0; i < WPT_GT_FLOAT_IMAGES; i++)
            out.aov[i] = (bits & (1u << i)) ? static_cast<float*>((this->*(table[i].member)).data()) : nullptr;
        out.materials = (bits & Materials) ? static_cast<int32_t*>(materials.data()) : nullptr;
        return out;
    }
};

inline GroundTruth getGroundTruth(const Sensor& sensor, const Camera& camera, const Scene& scene,
        float t0, float tPrev, float tNext,
        unsigned int groundTruthBits = GroundTruth::All,
        const Parameters& params = Parameters())
{
    /* the reference asserts tPrev <= t0 <= tNext and !scene.bvhNeedsUpdate(t0, t0) (wurblpt.hpp:648-651) */
    if (scene.bvhNeedsUpdate(t0, t0) || !scene.deviceScene())
        detail::die("getGroundTruth", "Scene::updateBVH(t0, t0) must be called first");
    scene.prepareGroundTruth(t0, tPrev, tNext);
    wpt_camera_desc cameraDesc;
    std::vector<wpt_keyframe> cameraKeyframes;
    std::vector<float> cameraTimes(1, tPrev);
    if (t0 > tPrev)
        cameraTimes.push_back(t0);
    if (tNext > t0)
        cameraTimes.push_back(tNext);
    if (!camera.flatten(cameraDesc, cameraKeyframes, cameraTimes))
        detail::die("getGroundTruth", "the camera animation cannot be evaluated on the device");
    const unsigned int width = sensor.width();
    const unsigned int height = sensor.height();
    GroundTruth gt(width, height, groundTruthBits);
    fprintf(stderr, "Getting ground truth for %ux%u pixels at %.3fs... ", width, height, t0);
    const std::vector<int32_t> materialIds = scene.materialIdsOfDescription();
    const wpt_ground_truth out = gt.describe();
    if (wpt_get_ground_truth(scene.deviceScene(), &cameraDesc, width, height, t0, tPrev, tNext, params.minHitDistance,
                groundTruthBits, materialIds.data(), materialIds.size(), nullptr, &out) != WPT_OK)
        detail::die("getGroundTruth", wpt_last_error());
    fprintf(stderr, "done\n");
    return gt;
}

inline GroundTruth getGroundTruth(const Sensor& sensor, const Camera& camera, const Scene& scene,
        float t0 = 0.0f,
        unsigned int groundTruthBits = GroundTruth::All,
        const Parameters& params = Parameters())
{
    return getGroundTruth(sensor, camera, scene, t0, t0, t0, groundTruthBits, params);
}

}
