/*
 * capi.cu -- the C ABI (include/wurblpt_b200.h): scene upload, launch set-up, result copies.
 * Host code only; the kernels live in kernels.cu. There is no CPU rendering path in this
 * library: every compute entry point needs a CUDA device and fails loudly without one.
 */
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "wurblpt_b200.h"
#include "scene_host.hpp"
#include "device_scene.cuh"
#include "kernels.hpp"
#include "bvh_device.hpp"

namespace {

thread_local std::string g_lastError;

int fail(int code, const std::string& message)
{
    g_lastError = message;
    return code;
}

int failCuda(cudaError_t e, const char* what)
{
    return fail(WPT_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define WPT_CUDA(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return failCuda(e_, what); } while (0)

/* Device allocation that grows on demand and is freed with its owner. */
struct DeviceBuf {
    void* p = nullptr;
    size_t capacity = 0;

    ~DeviceBuf() { release(); }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        capacity = 0;
    }
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= capacity)
            return cudaSuccess;
        release();
        size_t want = bytes < 256 ? 256 : bytes;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess)
            capacity = want;
        return e;
    }
    cudaError_t upload(const void* src, size_t bytes, cudaStream_t stream = 0)
    {
        cudaError_t e = reserve(bytes);
        if (e != cudaSuccess || bytes == 0)
            return e;
        return cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, stream);
    }
    template<typename T> const T* as() const { return static_cast<const T*>(p); }
};

/* color.hpp:270-278,287-295 */
float srgbToLinear(float x)
{
    return (x <= 0.04045f ? (x * (1.0f / 12.92f)) : std::pow((x + 0.055f) * (1.0f / 1.055f), 2.4f));
}

/* TextureImage::value(x, y) (texture_image.hpp:85-180): decode one texel to RGBA float */
void decodeTexel(const wpt_texture& t, const uint8_t* base, size_t index, float* out)
{
    float c[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    for (int k = 0; k < t.components; k++) {
        if (t.texel_type == WPT_TEXEL_U8)
            c[k] = base[index * t.components + k] / 255.0f;
        else if (t.texel_type == WPT_TEXEL_U16)
            c[k] = reinterpret_cast<const uint16_t*>(base)[index * t.components + k] / 65535.0f;
        else
            c[k] = reinterpret_cast<const float*>(base)[index * t.components + k];
    }
    const bool srgb = t.linearize_srgb != 0;
    if (t.components == 3) {
        out[0] = srgb ? srgbToLinear(c[0]) : c[0];
        out[1] = srgb ? srgbToLinear(c[1]) : c[1];
        out[2] = srgb ? srgbToLinear(c[2]) : c[2];
        out[3] = 1.0f;
    } else if (t.components == 4) {
        out[0] = srgb ? srgbToLinear(c[0]) : c[0];
        out[1] = srgb ? srgbToLinear(c[1]) : c[1];
        out[2] = srgb ? srgbToLinear(c[2]) : c[2];
        out[3] = c[3];
    } else if (t.components == 1) {
        float v = srgb ? srgbToLinear(c[0]) : c[0];
        out[0] = out[1] = out[2] = v;
        out[3] = 1.0f;
    } else {
        float v = srgb ? srgbToLinear(c[0]) : c[0];
        out[0] = out[1] = out[2] = v;
        out[3] = c[1];
    }
}

}

struct wpt_scene {
    /* Several GPUs render one frame (wpt_scene_create with WPT_DEVICE_ALL): the scene the caller holds is the
     * "master" on the first device; it owns the host-side scene and one replica (device arrays, scratch, streams)
     * per further device. A replica's `host` refers to the master's. */
    std::unique_ptr<wpt::HostScene> ownedHost;
    wpt::HostScene& host;
    wpt_scene* master = nullptr;
    std::vector<std::unique_ptr<wpt_scene>> replicas;
    wpt_scene() : ownedHost(new wpt::HostScene), host(*ownedHost) {}
    explicit wpt_scene(wpt_scene& m) : host(m.host), master(&m) {}
    int device = 0;
    int numSMs = 0;
    bool sceneHasMedia = false, sceneHasAnimation = false;
    /* time-independent arrays */
    DeviceBuf materials, instances, meshes, vertices, indices, textures, texels;
    DeviceBuf animations, keyframes, envM, envMs, envMcs, envGuide;
    int envGuideBits = 0;
    /* per updateBVH */
    DeviceBuf node4, node4c, bnode, tri, triShade, sph, med, triGlobal, sphGlobal, medGlobal, hotRefs, triClass, sphClass, medClass;
    /* per call scratch */
    DeviceBuf counters, camKeys, frame, partials, wave[17], qOrigins, qDirections, qIds, qDist, qStats, gtImages, gtMaterialIds, animTable, poolSpill, sharedCounter, blockDone, peerStage;
    /* device builder: inputs in hitable order, scratch */
    DeviceBuf bBox, bTri, bTriShade, bTriClass, bSph, bSphClass, bHot, bScratch;
    DeviceBuf waveBin[4];       /* ray cells, binned trace queue, bin counters, bin offsets + cursors */
    int builder = WPT_BUILDER_HOST_SAH;
    cudaEvent_t evStart = nullptr, evStop = nullptr, evWave[4] = { nullptr, nullptr, nullptr, nullptr };
    cudaStream_t stream2 = nullptr;
    unsigned long long* hostDone = nullptr;     /* pinned, 2 counters */
    wptd::DScene d;
    double uploadSeconds = 0.0;

    ~wpt_scene()
    {
        if (evStart) cudaEventDestroy(evStart);
        if (evStop) cudaEventDestroy(evStop);
        for (int i = 0; i < 4; i++)
            if (evWave[i]) cudaEventDestroy(evWave[i]);
        if (stream2) cudaStreamDestroy(stream2);
        if (hostDone) cudaFreeHost(hostDone);
    }

    uint64_t deviceBytes() const
    {
        const DeviceBuf* all[] = { &materials, &instances, &meshes, &vertices, &indices, &textures, &texels,
            &animations, &keyframes, &envM, &envMs, &envMcs, &envGuide, &node4, &node4c, &bnode, &tri, &triShade, &sph, &med,
            &triGlobal, &sphGlobal, &medGlobal, &hotRefs, &triClass, &sphClass, &medClass };
        uint64_t n = 0;
        for (const DeviceBuf* b : all)
            n += b->capacity;
        return n;
    }
};

namespace {

int selectDevice(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(WPT_E_NO_DEVICE, std::string("no CUDA device available (")
                + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0")
                + "); this library has no CPU fallback");
    if (device < 0 || device >= n)
        return fail(WPT_E_INVALID, "device index out of range");
    WPT_CUDA(cudaSetDevice(device), "cudaSetDevice");
    return WPT_OK;
}

int uploadStatic(wpt_scene* s)
{
    const wpt::HostScene& h = s->host;
    WPT_CUDA(s->materials.upload(h.materials.data(), h.materials.size() * sizeof(wpt_material)), "upload materials");
    WPT_CUDA(s->instances.upload(h.instances.data(), h.instances.size() * sizeof(wpt_instance)), "upload instances");
    WPT_CUDA(s->meshes.upload(h.meshes.data(), h.meshes.size() * sizeof(wpt_mesh)), "upload meshes");
    WPT_CUDA(s->vertices.upload(h.vertices.data(), h.vertices.size() * sizeof(float)), "upload vertices");
    WPT_CUDA(s->indices.upload(h.indices.data(), h.indices.size() * sizeof(uint32_t)), "upload indices");
    WPT_CUDA(s->animations.upload(h.animations.data(), h.animations.size() * sizeof(wpt_animation)), "upload animations");
    WPT_CUDA(s->keyframes.upload(h.keyframes.data(), h.keyframes.size() * sizeof(wpt_keyframe)), "upload keyframes");
    WPT_CUDA(s->envM.upload(h.envM.data(), h.envM.size() * sizeof(float)), "upload envmap M");
    WPT_CUDA(s->envMs.upload(h.envMs.data(), h.envMs.size() * sizeof(int32_t)), "upload envmap Ms");
    WPT_CUDA(s->envMcs.upload(h.envMcs.data(), h.envMcs.size() * sizeof(float)), "upload envmap Mcs");
    /* Guide table for EnvironmentMap::d's search in Mcs (envmap.hpp:190-203): entry j is the first bin whose
     * cumulative sum is >= j / 2^bits, so a draw r only has to search between entries floor(r 2^bits) and
     * the next one. Same answer as the full bisection as long as Mcs never decreases (checked here). */
    s->envGuideBits = 0;
    if (!h.envMcs.empty()) {
        bool monotone = true;
        for (size_t i = 1; i < h.envMcs.size() && monotone; i++)
            monotone = (h.envMcs[i] >= h.envMcs[i - 1]);
        if (monotone) {
            const int bits = 16;
            std::vector<uint32_t> guide((size_t(1) << bits) + 1);
            size_t i = 0;
            for (size_t j = 0; j < guide.size(); j++) {
                const float r = float(j) / float(1u << bits);   /* exact */
                while (i + 1 < h.envMcs.size() && h.envMcs[i] < r)
                    i++;
                guide[j] = uint32_t(i);
            }
            WPT_CUDA(s->envGuide.upload(guide.data(), guide.size() * sizeof(uint32_t)), "upload envmap guide table");
            s->envGuideBits = bits;
        }
    }

    /* textures: descriptors + image texels decoded once to RGBA float (16-byte texel fetches) */
    std::vector<wptd::DTexture> dt(h.textures.size());
    size_t texelCount = 0;
    for (const wpt_texture& t : h.textures)
        if (t.type == WPT_TEX_IMAGE)
            texelCount += size_t(t.width) * t.height;
    if (texelCount >= 0xFFFFFFFFull)
        return fail(WPT_E_UNSUPPORTED, "too many texels");
    std::vector<float> decoded(texelCount * 4);
    size_t cursor = 0;
    for (size_t i = 0; i < h.textures.size(); i++) {
        const wpt_texture& t = h.textures[i];
        wptd::DTexture& o = dt[i];
        std::memset(&o, 0, sizeof(o));
        o.type = t.type;
        o.src = t.src_texture;
        o.width = t.width;
        o.height = t.height;
        o.horiz = t.horiz;
        o.vert = t.vert;
        o.color0 = make_float4(t.color0[0], t.color0[1], t.color0[2], t.color0[3]);
        o.color1 = make_float4(t.color1[0], t.color1[1], t.color1[2], t.color1[3]);
        o.coordFactor = make_float2(t.coord_factor[0], t.coord_factor[1]);
        o.coordOffset = make_float2(t.coord_offset[0], t.coord_offset[1]);
        o.valFactor = make_float4(t.val_factor[0], t.val_factor[1], t.val_factor[2], t.val_factor[3]);
        o.valOffset = make_float4(t.val_offset[0], t.val_offset[1], t.val_offset[2], t.val_offset[3]);
        if (t.type == WPT_TEX_IMAGE) {
            o.texelBase = uint32_t(cursor);
            const uint8_t* base = h.texels.data() + t.texel_offset;
            const size_t n = size_t(t.width) * t.height;
            float* dst = decoded.data() + cursor * 4;
            #pragma omp parallel for schedule(static)
            for (size_t k = 0; k < n; k++)
                decodeTexel(t, base, k, dst + k * 4);
            cursor += n;
        }
    }
    WPT_CUDA(s->textures.upload(dt.data(), dt.size() * sizeof(wptd::DTexture)), "upload textures");
    WPT_CUDA(s->texels.upload(decoded.data(), decoded.size() * sizeof(float)), "upload texels");
    WPT_CUDA(cudaDeviceSynchronize(), "upload scene");
    return WPT_OK;
}

void refreshDeviceScene(wpt_scene* s)
{
    wptd::DScene& d = s->d;
    const wpt::HostScene& h = s->host;
    d.node4 = s->node4.as<float4>();
    d.node4c = s->node4c.as<float4>();
    d.bnode = s->bnode.as<float4>();
    d.tri = s->tri.as<float4>();
    d.triShade = s->triShade.as<uint4>();
    d.sph = s->sph.as<float4>();
    d.med = s->med.as<float4>();
    d.triGlobal = s->triGlobal.as<uint32_t>();
    d.sphGlobal = s->sphGlobal.as<uint32_t>();
    d.medGlobal = s->medGlobal.as<uint32_t>();
    d.hotRefs = s->hotRefs.as<uint32_t>();
    d.triClass = s->triClass.as<uint8_t>();
    d.sphClass = s->sphClass.as<uint8_t>();
    d.medClass = s->medClass.as<uint8_t>();
    d.classMask = h.classMask;
    {
        const float lo[3] = { h.sceneLo[0], h.sceneLo[1], h.sceneLo[2] };
        float scale[3];
        for (int k = 0; k < 3; k++) {
            const float extent = h.sceneHi[k] - h.sceneLo[k];
            scale[k] = (extent > 0.0f && std::isfinite(extent)) ? float(wptd::RAY_CELLS) / extent : 0.0f;
        }
        d.cellLo = make_float3(lo[0], lo[1], lo[2]);
        d.cellScale = make_float3(scale[0], scale[1], scale[2]);
    }
    d.simpleWalk = (s->sceneHasMedia || s->sceneHasAnimation || h.root4 == 0x7FFFFFFF) ? 0 : 1;   /* (an empty scene has no root) */
    d.materials = s->materials.as<wpt_material>();
    d.instances = s->instances.as<wpt_instance>();
    d.meshes = s->meshes.as<wpt_mesh>();
    d.vertices = s->vertices.as<float>();
    d.indices = s->indices.as<uint32_t>();
    d.textures = s->textures.as<wptd::DTexture>();
    d.texels = s->texels.as<float4>();
    d.animations = s->animations.as<wpt_animation>();
    d.keyframes = s->keyframes.as<wpt_keyframe>();
    d.envM = s->envM.as<float>();
    d.envMs = s->envMs.as<int32_t>();
    d.envMcs = s->envMcs.as<float>();
    d.envGuide = s->envGuide.as<uint32_t>();
    d.envGuideBits = s->envGuideBits;
    d.root4 = h.root4;
    d.prefetchLeaves = (h.nTriSlots * 12 * sizeof(float) > (128u << 10)) ? 1 : 0;
    d.nHot = uint32_t(h.hotSpotIds.size());
    d.nAnimations = uint32_t(h.animations.size());
    d.envType = h.env.type;
    d.envCompat = h.env.compat;
    d.envN = h.env.is_n;
    for (int i = 0; i < 6; i++)
        d.envTex[i] = h.env.tex[i];
}

wpt::Trs cameraPose(const wpt_camera_desc* c, float t)
{
    if (c->n_keyframes > 0 && c->keyframes)
        return wpt::keyframesAt(c->keyframes, c->n_keyframes, t);
    return wpt::Trs{ wpt::v3(c->pose_translation[0], c->pose_translation[1], c->pose_translation[2]),
        wpt::Quat{ c->pose_rotation[0], c->pose_rotation[1], c->pose_rotation[2], c->pose_rotation[3] },
        wpt::v3(c->pose_scaling[0], c->pose_scaling[1], c->pose_scaling[2]) };
}

/* camera part of the launch parameters: Camera::getRayHelper (camera.hpp:114-120) */
int fillCamera(wpt_scene* s, const wpt_camera_desc* c, float t0, uint32_t width, uint32_t height,
        wptd::RenderParams& P, cudaStream_t stream)
{
    if (c->surround_mode < WPT_SURROUND_OFF || c->surround_mode > WPT_SURROUND_360
            || c->distortion_type < WPT_DISTORTION_NONE || c->distortion_type > WPT_DISTORTION_OPENCV)
        return fail(WPT_E_INVALID, "camera description is malformed");
    P.surround = c->surround_mode;
    P.stereo = c->stereoscopic_distance;
    P.l = c->l; P.r = c->r; P.b = c->b; P.t = c->t;
    P.distortionType = c->distortion_type;
    P.k1 = c->k1; P.k2 = c->k2; P.k3 = c->k3; P.p1 = c->p1; P.p2 = c->p2;
    P.b1 = c->b1; P.b2 = c->b2; P.b3 = c->b3; P.b4 = c->b4;
    P.lensRadius = c->lens_radius;
    P.focusDist = c->focus_dist;
    const wpt::Trs T = cameraPose(c, t0);
    P.camT = make_float3(T.t.x, T.t.y, T.t.z);
    P.camR = make_float4(T.r.x, T.r.y, T.r.z, T.r.w);
    P.camS = make_float3(T.s.x, T.s.y, T.s.z);
    P.nCamKeys = (c->keyframes ? c->n_keyframes : 0);
    P.camKeys = nullptr;
    if (P.nCamKeys > 0) {
        WPT_CUDA(s->camKeys.upload(c->keyframes, size_t(P.nCamKeys) * sizeof(wpt_keyframe), stream), "upload camera key frames");
        P.camKeys = s->camKeys.as<wpt_keyframe>();
    }
    /* LensDistortion::getHelper (optics.hpp:215-224) with Projection::center/focalLength (optics.hpp:83-97) */
    P.distCenter = make_float2(c->l / (c->l - c->r), c->b / (c->b - c->t));
    P.distFocal = make_float2(1.0f / (c->r - c->l), 1.0f / (c->t - c->b));
    P.distInvFocal = make_float2(c->r - c->l, c->t - c->b);
    P.width = width;
    P.height = height;
    return WPT_OK;
}

/* WPT_PIPELINE / WPT_WAVE_SLOTS: experiment knobs (environment), see DESIGN.md */
int defaultPipeline()
{
    const char* e = std::getenv("WPT_PIPELINE");
    if (e && std::string(e) == "megakernel")
        return WPT_PIPELINE_MEGAKERNEL;
    return WPT_PIPELINE_WAVEFRONT;
}

#ifndef WPT_ITEMS_PER_SLOT
#define WPT_ITEMS_PER_SLOT 4u
#endif

bool waveBinning()
{
    /* Opt-in (WPT_WAVE_BIN=1): the bounce / shadow queries are handed to the trace kernel grid cell by grid cell. The walk
     * itself gains (+22 ... +33 % rays/s in the query kernel, profiles/r2l_coherence_probe.txt), but the queries of a
     * warp then belong to slots all over the 1.9 GB of path state (ray loads and answer stores lose their locality) and
     * the counting sort costs two passes over the queue: C3 660 vs 696 Msamples/s, C1 448 vs 582 (profiles/r2m_bin_ab.txt). */
    const char* e = std::getenv("WPT_WAVE_BIN");
    return e && e[0] == '1';
}

uint32_t waveSlotLimit()
{
    const char* e = std::getenv("WPT_WAVE_SLOTS");
    long v = e ? std::atol(e) : 0;
    return v > 0 ? uint32_t(v) : 12582912u;    /* 12 M slots: 2.3 GB of path state */
}

int sceneFeatures(const wpt_scene* s, bool needOpl)
{
    if (!(s->sceneHasMedia || s->sceneHasAnimation || needOpl))
        return 0;
    return s->sceneHasAnimation ? 7 : 5;    /* F_MEDIA | F_OPL, plus F_ANIM (kernels.cu picks the compiled variant) */
}

/* The device scene for one launch. When all rays of the launch have the same time (t0 == t1: no motion blur, or a
 * query at one time) the animations are evaluated once, here, with the host arithmetic the bounds were computed
 * with, and the kernels read them from a table; otherwise every thread evaluates the key frames at its ray's time
 * into a per-thread cache, which holds MAX_CACHED_ANIMATIONS animations. */
int launchScene(wpt_scene* s, float t0, float t1, cudaStream_t stream, wptd::DScene& d)
{
    d = s->d;
    d.animTable = nullptr;
    const size_t n = s->host.animations.size();
    if (!s->sceneHasAnimation || n == 0)
        return WPT_OK;
    if (t0 == t1) {
        std::vector<float> table(n * 32, 0.0f);
        for (size_t i = 0; i < n; i++) {
            const wpt::Trs T = s->host.animationAt(int(i), t0);
            float* e = table.data() + i * 32;
            e[0] = T.t.x; e[1] = T.t.y; e[2] = T.t.z;
            e[3] = T.r.x; e[4] = T.r.y; e[5] = T.r.z; e[6] = T.r.w;
            e[7] = T.s.x; e[8] = T.s.y; e[9] = T.s.z;
            const wpt::Mat34 M = wpt::toMat34(T);
            std::memcpy(e + 10, M.m, sizeof(M.m));
            wpt::toMat3(T.r, e + 22);
        }
        WPT_CUDA(s->animTable.upload(table.data(), table.size() * sizeof(float), stream), "upload animation table");
        d.animTable = s->animTable.as<float>();
        return WPT_OK;
    }
    /* (more than MAX_CACHED_ANIMATIONS - 1 animations: the others share the cache's last entry, device_trace.cuh) */
    return WPT_OK;
}

/* Spill area of the pooled trace kernel: stack entries beyond those held in shared memory, for every pool entry
 * of every resident block of two concurrent launches. The walk's stack never holds more than 3 * depth + 1 entries. */
int reservePoolSpill(wpt_scene* s, wptd::RenderParams& P)
{
    const uint32_t need = 3 * s->host.wideDepth + 1;
    const uint32_t held = wptd::poolStackEntries();
    P.poolSpillPerRay = (need > held ? need - held : 0) + 1;
    const size_t entries = size_t(2) * wptd::poolGridBlocks(s->numSMs) * wptd::poolRaysPerBlock() * P.poolSpillPerRay;
    WPT_CUDA(s->poolSpill.reserve(entries * sizeof(uint2)), "allocate trace spill area");
    P.poolSpill = static_cast<uint2*>(s->poolSpill.p);
    return WPT_OK;
}

int checkReady(wpt_scene* s)
{
    if (!s)
        return fail(WPT_E_INVALID, "null scene");
    if (s->device == WPT_DEVICE_NONE)
        return fail(WPT_E_NO_DEVICE, "this scene was created without a device (WPT_DEVICE_NONE); "
                "rendering needs a CUDA device, there is no CPU fallback");
    if (!s->host.haveBvh)
        return fail(WPT_E_STATE, "wpt_scene_update_bvh() must be called before rendering "
                "(the reference asserts !scene.bvhNeedsUpdate(), wurblpt.hpp:289)");
    WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    return WPT_OK;
}

}

extern "C" {

const char* wpt_version(void)
{
    return "wurblpt_b200 0.1 (abi 1, sm_100a)";
}

const char* wpt_last_error(void)
{
    return g_lastError.c_str();
}

int wpt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

size_t wpt_abi_sizeof(int which)
{
    switch (which) {
    case 0: return sizeof(wpt_keyframe);
    case 1: return sizeof(wpt_animation);
    case 2: return sizeof(wpt_texture);
    case 3: return sizeof(wpt_material);
    case 4: return sizeof(wpt_mesh);
    case 5: return sizeof(wpt_instance);
    case 6: return sizeof(wpt_sphere);
    case 7: return sizeof(wpt_medium);
    case 8: return sizeof(wpt_component);
    case 9: return sizeof(wpt_envmap);
    case 10: return sizeof(wpt_scene_desc);
    case 11: return sizeof(wpt_camera_desc);
    case 12: return sizeof(wpt_sensor_desc);
    case 13: return sizeof(wpt_params);
    case 14: return sizeof(wpt_render_opts);
    case 15: return sizeof(wpt_stats);
    case 16: return sizeof(wpt_scene_info);
    }
    return 0;
}

/* The devices a scene created with WPT_DEVICE_ALL uses: WURBLPT_DEVICES = "all" (default) or a comma separated
 * list of CUDA device indices. */
static std::vector<int> plannedDevices()
{
    std::vector<int> devices;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    const char* e = std::getenv("WURBLPT_DEVICES");
    if (e && *e && std::string(e) != "all") {
        const std::string list(e);
        size_t pos = 0;
        while (pos < list.size()) {
            size_t comma = list.find(',', pos);
            if (comma == std::string::npos)
                comma = list.size();
            const std::string item = list.substr(pos, comma - pos);
            char* endp = nullptr;
            const long v = std::strtol(item.c_str(), &endp, 10);
            bool known = false;
            for (int d : devices)
                known = known || d == int(v);
            if (endp != item.c_str() && v >= 0 && v < n && !known)
                devices.push_back(int(v));
            pos = comma + 1;
        }
    } else {
        for (int d = 0; d < n; d++)
            devices.push_back(d);
    }
    return devices;
}

int wpt_planned_device_count(void)
{
    return int(plannedDevices().size());
}

/* Everything a scene needs on its device that does not depend on the time range. */
static int initDevice(wpt_scene* s, int device)
{
    int rc = selectDevice(device);
    if (rc != WPT_OK)
        return rc;
    s->device = device;
    cudaDeviceProp prop;
    WPT_CUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(WPT_E_NO_DEVICE, std::string("device '") + prop.name + "' is not a Blackwell (sm_100) GPU; "
                "this library ships sm_100a code only");
    s->numSMs = prop.multiProcessorCount;
    s->sceneHasMedia = !s->host.media.empty();
    for (const wpt_instance& in : s->host.instances)
        s->sceneHasAnimation = s->sceneHasAnimation || in.animation >= 0;
    for (const wpt_sphere& sp : s->host.spheres)
        s->sceneHasAnimation = s->sceneHasAnimation || sp.animation >= 0;
    auto start = std::chrono::steady_clock::now();
    rc = uploadStatic(s);
    if (rc != WPT_OK)
        return rc;
    s->uploadSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    WPT_CUDA(cudaEventCreate(&s->evStart), "cudaEventCreate");
    WPT_CUDA(cudaEventCreate(&s->evStop), "cudaEventCreate");
    for (int i = 0; i < 4; i++)
        WPT_CUDA(cudaEventCreateWithFlags(&s->evWave[i], cudaEventDisableTiming), "cudaEventCreate");
    WPT_CUDA(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking), "cudaStreamCreate");
    WPT_CUDA(cudaMallocHost(&s->hostDone, 2 * sizeof(unsigned long long)), "allocate pinned counters");
    WPT_CUDA(s->counters.reserve(wptd::CTR_WORDS * sizeof(unsigned long long)), "allocate counters");
    return WPT_OK;
}

/* Lets `from` read and write `to`'s memory (the shared work item counter lives on the first device; frames are
 * copied device to device). */
static int enablePeer(int from, int to)
{
    int can = 0;
    WPT_CUDA(cudaDeviceCanAccessPeer(&can, from, to), "cudaDeviceCanAccessPeer");
    if (!can)
        return fail(WPT_E_UNSUPPORTED, "devices " + std::to_string(from) + " and " + std::to_string(to)
                + " cannot access each other's memory (no NVLink / P2P path): cannot render one frame on both");
    WPT_CUDA(cudaSetDevice(from), "cudaSetDevice");
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled)
        cudaGetLastError();
    else if (e != cudaSuccess)
        return failCuda(e, "cudaDeviceEnablePeerAccess");
    return WPT_OK;
}

int wpt_scene_create(const wpt_scene_desc* desc, int device, wpt_scene** out)
{
    if (!out)
        return fail(WPT_E_INVALID, "null output pointer");
    *out = nullptr;
    std::unique_ptr<wpt_scene> s(new wpt_scene);
    int rc = s->host.ingest(desc);
    if (rc != WPT_OK)
        return fail(rc, s->host.error);
    if (const char* e = std::getenv("WPT_LEAF_TRIS")) {     /* experiment knob: triangles per wide-tree leaf, 1..4 */
        const int v = std::atoi(e);
        if (v >= 1 && v <= 4)
            s->host.wideLeafTriangles = uint32_t(v);
    }
    if (const char* e = std::getenv("WURBLPT_BVH_BUILDER"))     /* "device": build the tree on the GPU (wpt_scene_set_builder) */
        s->builder = (std::string(e) == "device" ? WPT_BUILDER_DEVICE_LBVH : WPT_BUILDER_HOST_SAH);
    if (device == WPT_DEVICE_NONE) {
        /* host-only scene: bounds and tree can be built and inspected, nothing can be rendered */
        s->device = WPT_DEVICE_NONE;
        *out = s.release();
        return WPT_OK;
    }
    std::vector<int> devices(1, device);
    if (device == WPT_DEVICE_ALL) {
        devices = plannedDevices();
        if (devices.empty())
            return fail(WPT_E_NO_DEVICE, "no CUDA device available (WURBLPT_DEVICES names none); this library has no CPU fallback");
    }
    rc = initDevice(s.get(), devices[0]);
    if (rc != WPT_OK)
        return rc;
    for (size_t i = 1; i < devices.size(); i++) {
        std::unique_ptr<wpt_scene> r(new wpt_scene(*s));
        rc = initDevice(r.get(), devices[i]);
        if (rc == WPT_OK)
            rc = enablePeer(devices[i], devices[0]);
        if (rc == WPT_OK)
            rc = enablePeer(devices[0], devices[i]);
        if (rc != WPT_OK)
            return rc;
        s->replicas.push_back(std::move(r));
    }
    WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    *out = s.release();
    return WPT_OK;
}

int wpt_ipc_counter_create(int device, void** counters, unsigned char handle[64])
{
    if (!counters || !handle)
        return fail(WPT_E_INVALID, "null argument");
    int rc = selectDevice(device);
    if (rc != WPT_OK)
        return rc;
    void* p = nullptr;
    WPT_CUDA(cudaMalloc(&p, 256), "allocate shared counters");
    WPT_CUDA(cudaMemset(p, 0, 256), "clear shared counters");
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return failCuda(e, "cudaIpcGetMemHandle");
    }
    std::memcpy(handle, &h, 64);
    *counters = p;
    return WPT_OK;
}

int wpt_ipc_counter_open(int device, const unsigned char handle[64], void** counters)
{
    if (!counters || !handle)
        return fail(WPT_E_INVALID, "null argument");
    int rc = selectDevice(device);
    if (rc != WPT_OK)
        return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    WPT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    *counters = p;
    return WPT_OK;
}

int wpt_ipc_counter_close(int device, void* counters, int opened)
{
    if (!counters)
        return WPT_OK;
    int rc = selectDevice(device);
    if (rc != WPT_OK)
        return rc;
    if (opened)
        WPT_CUDA(cudaIpcCloseMemHandle(counters), "cudaIpcCloseMemHandle");
    else
        WPT_CUDA(cudaFree(counters), "free shared counters");
    return WPT_OK;
}

int wpt_ipc_counter_clear(void* counter, void* stream)
{
    if (!counter)
        return fail(WPT_E_INVALID, "null counter");
    WPT_CUDA(cudaMemsetAsync(counter, 0, 8, static_cast<cudaStream_t>(stream)), "clear shared counter");
    return WPT_OK;
}

int wpt_scene_device_count(const wpt_scene* s)
{
    if (!s || s->device == WPT_DEVICE_NONE)
        return 0;
    return 1 + int(s->replicas.size());
}

/* The arrays that depend on the time range the acceleration structure was built for. */
static int uploadBvh(wpt_scene* s)
{
    WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    auto start = std::chrono::steady_clock::now();
    const wpt::HostScene& h = s->host;
    WPT_CUDA(s->node4.upload(h.nodes4.data(), h.nodes4.size() * sizeof(wpt::Node4)), "upload wide nodes");
    WPT_CUDA(s->node4c.reserve(h.nodes4.size() * sizeof(wpt::Node4)), "allocate wide nodes (child-major)");
    WPT_CUDA(wptd::launchNodeTranspose(s->node4.as<float4>(), static_cast<float4*>(s->node4c.p), uint32_t(h.nodes4.size()), 0), "wide nodes child by child");
    WPT_CUDA(s->bnode.upload(h.bnodes.data(), h.bnodes.size() * sizeof(wpt::BNode)), "upload binary nodes");
    WPT_CUDA(s->tri.upload(h.triData.data(), h.triData.size() * sizeof(float)), "upload triangles");
    WPT_CUDA(s->triShade.upload(h.triShade.data(), h.triShade.size() * sizeof(uint32_t)), "upload triangle shading records");
    WPT_CUDA(s->sph.upload(h.sphData.data(), h.sphData.size() * sizeof(float)), "upload spheres");
    WPT_CUDA(s->med.upload(h.medData.data(), h.medData.size() * sizeof(float)), "upload media");
    WPT_CUDA(s->triGlobal.upload(h.triGlobal.data(), h.triGlobal.size() * sizeof(uint32_t)), "upload triangle ids");
    WPT_CUDA(s->sphGlobal.upload(h.sphGlobal.data(), h.sphGlobal.size() * sizeof(uint32_t)), "upload sphere ids");
    WPT_CUDA(s->medGlobal.upload(h.medGlobal.data(), h.medGlobal.size() * sizeof(uint32_t)), "upload medium ids");
    WPT_CUDA(s->hotRefs.upload(h.hotSpotRefs.data(), h.hotSpotRefs.size() * sizeof(uint32_t)), "upload hot spots");
    WPT_CUDA(s->triClass.upload(h.triClass.data(), h.triClass.size()), "upload triangle material classes");
    WPT_CUDA(s->sphClass.upload(h.sphClass.data(), h.sphClass.size()), "upload sphere material classes");
    WPT_CUDA(s->medClass.upload(h.medClass.data(), h.medClass.size()), "upload medium material classes");
    WPT_CUDA(cudaDeviceSynchronize(), "upload acceleration structure");
    refreshDeviceScene(s);
    s->uploadSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    return WPT_OK;
}

/* Scene::updateBVH() with the tree built on the device (bvh_device.cu): uploads the bounds and the primitive data
 * in hitable order, builds, and leaves the arrays the kernels read in slot order. */
static int buildOnDevice(wpt_scene* s, uint32_t* wideNodes, uint32_t* depth, uint32_t* triangles)
{
    WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    auto start = std::chrono::steady_clock::now();
    const wpt::HostScene& h = s->host;
    const uint32_t n = uint32_t(h.prims.size());
    const size_t nTri = h.nTriSlots, nSph = h.nSphSlots;
    WPT_CUDA(s->bBox.upload(h.devBox.data(), h.devBox.size() * sizeof(float)), "upload bounds");
    WPT_CUDA(s->bTri.upload(h.devTriPrim.data(), h.devTriPrim.size() * sizeof(float)), "upload triangles");
    WPT_CUDA(s->bTriShade.upload(h.devTriShadePrim.data(), h.devTriShadePrim.size() * sizeof(uint32_t)), "upload triangle shading records");
    WPT_CUDA(s->bTriClass.upload(h.devTriClassPrim.data(), h.devTriClassPrim.size()), "upload triangle material classes");
    WPT_CUDA(s->bSph.upload(h.devSphPrim.data(), h.devSphPrim.size() * sizeof(float)), "upload spheres");
    WPT_CUDA(s->bSphClass.upload(h.devSphClassPrim.data(), h.devSphClassPrim.size()), "upload sphere material classes");
    WPT_CUDA(s->bHot.upload(h.hotSpotIds.data(), h.hotSpotIds.size() * sizeof(uint32_t)), "upload hot spot ids");
    WPT_CUDA(s->node4.reserve(size_t(n) * sizeof(wpt::Node4)), "allocate wide nodes");
    WPT_CUDA(s->tri.reserve(nTri * 12 * sizeof(float)), "allocate triangles");
    WPT_CUDA(s->triShade.reserve(nTri * 4 * sizeof(uint32_t)), "allocate triangle shading records");
    WPT_CUDA(s->triClass.reserve(nTri), "allocate triangle material classes");
    WPT_CUDA(s->triGlobal.reserve(nTri * sizeof(uint32_t)), "allocate triangle ids");
    WPT_CUDA(s->sph.reserve(nSph * 12 * sizeof(float)), "allocate spheres");
    WPT_CUDA(s->sphClass.reserve(nSph), "allocate sphere material classes");
    WPT_CUDA(s->sphGlobal.reserve(nSph * sizeof(uint32_t)), "allocate sphere ids");
    WPT_CUDA(s->hotRefs.reserve(h.hotSpotIds.size() * sizeof(uint32_t)), "allocate hot spots");
    WPT_CUDA(s->bnode.reserve(64), "allocate binary nodes");
    WPT_CUDA(s->med.reserve(64), "allocate media");
    WPT_CUDA(s->medGlobal.reserve(64), "allocate medium ids");
    WPT_CUDA(s->medClass.reserve(64), "allocate medium classes");
    WPT_CUDA(s->bScratch.reserve(wptd::deviceBuildScratchBytes(n)), "allocate build scratch");
    wptd::BuildIO io;
    io.n = n;
    io.box = s->bBox.as<float4>();
    io.triPrim = s->bTri.as<float4>();
    io.triShadePrim = s->bTriShade.as<uint4>();
    io.triClassPrim = s->bTriClass.as<uint8_t>();
    io.sphPrim = s->bSph.as<float4>();
    io.sphClassPrim = s->bSphClass.as<uint8_t>();
    io.hotPrims = s->bHot.as<uint32_t>();
    io.nHot = uint32_t(h.hotSpotIds.size());
    io.node4 = static_cast<float4*>(s->node4.p);
    io.tri = static_cast<float4*>(s->tri.p);
    io.triShade = static_cast<uint4*>(s->triShade.p);
    io.triClass = static_cast<uint8_t*>(s->triClass.p);
    io.triGlobal = static_cast<uint32_t*>(s->triGlobal.p);
    io.sph = static_cast<float4*>(s->sph.p);
    io.sphClass = static_cast<uint8_t*>(s->sphClass.p);
    io.sphGlobal = static_cast<uint32_t*>(s->sphGlobal.p);
    io.hotRefs = static_cast<uint32_t*>(s->hotRefs.p);
    WPT_CUDA(wptd::launchDeviceBuild(io, s->bScratch.p, h.wideLeafTriangles, 0, wideNodes, depth, triangles), "device tree build");
    WPT_CUDA(cudaDeviceSynchronize(), "device tree build");
    WPT_CUDA(s->node4c.reserve(size_t(*wideNodes) * sizeof(wpt::Node4)), "allocate wide nodes (child-major)");
    WPT_CUDA(wptd::launchNodeTranspose(s->node4.as<float4>(), static_cast<float4*>(s->node4c.p), *wideNodes, 0), "wide nodes child by child");
    WPT_CUDA(cudaDeviceSynchronize(), "wide nodes child by child");
    s->uploadSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    return WPT_OK;
}

int wpt_scene_set_builder(wpt_scene* s, int builder)
{
    if (!s)
        return fail(WPT_E_INVALID, "null scene");
    if (builder != WPT_BUILDER_HOST_SAH && builder != WPT_BUILDER_DEVICE_LBVH)
        return fail(WPT_E_INVALID, "unknown builder");
    s->builder = builder;
    return WPT_OK;
}

int wpt_scene_update_bvh(wpt_scene* s, float t0, float t1)
{
    if (!s)
        return fail(WPT_E_INVALID, "null scene");
    if (s->device != WPT_DEVICE_NONE)
        WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    int rc = WPT_OK;
    if (s->device != WPT_DEVICE_NONE && s->builder == WPT_BUILDER_DEVICE_LBVH) {
        /* bounds on the host, tree and slot order on every device */
        auto start = std::chrono::steady_clock::now();
        rc = s->host.prepareDeviceBuild(t0, t1);
        if (rc == WPT_OK) {
            uint32_t wideNodes = 0, depth = 0, triangles = 0;
            rc = buildOnDevice(s, &wideNodes, &depth, &triangles);
            for (size_t i = 0; i < s->replicas.size() && rc == WPT_OK; i++)
                rc = buildOnDevice(s->replicas[i].get(), &wideNodes, &depth, &triangles);
            if (rc != WPT_OK)
                return rc;
            s->host.finishDeviceBuild(t0, t1, wideNodes, depth, triangles,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count());
            refreshDeviceScene(s);
            for (auto& r : s->replicas)
                refreshDeviceScene(r.get());
            WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
            return WPT_OK;
        }
        if (rc != WPT_E_UNSUPPORTED)
            return fail(rc, s->host.error);
        /* media (nested boundary trees) or fewer than two hitables: the host builder */
    }
    rc = s->host.updateBvh(t0, t1);     /* built once on the host, uploaded to every device */
    if (rc != WPT_OK)
        return fail(rc, s->host.error);
    if (s->device == WPT_DEVICE_NONE)
        return WPT_OK;
    rc = uploadBvh(s);
    for (size_t i = 0; i < s->replicas.size() && rc == WPT_OK; i++)
        rc = uploadBvh(s->replicas[i].get());
    WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    return rc;
}

int wpt_scene_get_info(const wpt_scene* s, wpt_scene_info* info)
{
    if (!s || !info)
        return fail(WPT_E_INVALID, "null argument");
    std::memset(info, 0, sizeof(*info));
    const wpt::HostScene& h = s->host;
    info->hitables = h.topLevel.size();
    info->triangles = h.nTriSlots;
    info->spheres = h.nSphSlots;
    info->media = h.medGlobal.size();
    info->binary_nodes = h.topLevelNodes;
    info->wide_nodes = h.nWideNodes;
    info->binary_depth = h.binaryDepth;
    info->wide_depth = h.wideDepth;
    info->hot_spots = uint32_t(h.hotSpotIds.size());
    for (size_t m = 0; m < h.medGlobal.size(); m++) {
        uint32_t run;
        std::memcpy(&run, &h.medData[m * 4 + 3], sizeof(run));
        info->warp_tested_media += (run != 0u);
    }
    info->device_bytes = s->deviceBytes();
    info->build_seconds = h.buildSeconds;
    info->upload_seconds = s->uploadSeconds;
    return WPT_OK;
}

void wpt_scene_destroy(wpt_scene* s)
{
    if (s) {
        for (auto& r : s->replicas) {
            cudaSetDevice(r->device);
            r.reset();
        }
        s->replicas.clear();
        if (s->device != WPT_DEVICE_NONE)
            cudaSetDevice(s->device);
        delete s;
    }
}

int wpt_scene_get_binary_nodes(const wpt_scene* s, uint64_t capacity, float* boxes, uint32_t* link, int32_t* leaf)
{
    if (!s || !boxes || !link || !leaf)
        return fail(WPT_E_INVALID, "null argument");
    const wpt::HostScene& h = s->host;
    if (!h.haveBvh)
        return fail(WPT_E_STATE, "wpt_scene_update_bvh() has not been called");
    if (h.deviceBuilt)
        return fail(WPT_E_UNSUPPORTED, "the tree was built on the device: there is no binary tree in the reference's layout");
    if (capacity < h.topLevelNodes)
        return fail(WPT_E_INVALID, "capacity too small");
    for (uint64_t i = 0; i < h.topLevelNodes; i++) {
        const wpt::BNode& n = h.bnodes[i];
        for (int k = 0; k < 3; k++) {
            boxes[i * 6 + k] = n.lo[k];
            boxes[i * 6 + 3 + k] = n.hi[k];
        }
        if (n.leaf == wpt::INNER) {
            link[i] = n.link;
            leaf[i] = -1;
        } else {
            const uint32_t type = n.leaf >> wpt::REF_TYPE_SHIFT, slot = n.leaf & wpt::REF_SLOT_MASK;
            link[i] = 0;
            leaf[i] = int32_t(type == wpt::PRIM_TRIANGLE ? h.triGlobal[slot]
                    : type == wpt::PRIM_SPHERE ? h.sphGlobal[slot] : h.medGlobal[slot]);
        }
    }
    return WPT_OK;
}

/* One frame on several GPUs: what the scenes of the devices share. */
struct MultiFrame {
    unsigned long long* itemCounter;    /* in the first device's memory; every device draws its blocks of work items here */
};

static int renderSingle(wpt_scene* s, const wpt_camera_desc* camera, const wpt_sensor_desc* sensor,
        const wpt_params* params, uint32_t samplesSqrt, float t0, float t1,
        const wpt_render_opts* opts, float* frame, wpt_stats* stats, const MultiFrame* multi)
{
    auto wallStart = std::chrono::steady_clock::now();
    int rc = checkReady(s);
    if (rc != WPT_OK)
        return rc;
    if (!camera || !sensor || !params || !frame)
        return fail(WPT_E_INVALID, "null argument");
    if (sensor->width == 0 || sensor->height == 0 || samplesSqrt == 0
            || uint64_t(sensor->width) * sensor->height >= (1ull << 31))
        return fail(WPT_E_INVALID, "invalid image size or sample count");
    if (sensor->type != WPT_SENSOR_RGB && sensor->type != WPT_SENSOR_TOF_AMCW)
        return fail(WPT_E_UNSUPPORTED, "unknown sensor type");
    if (s->host.animations.size() > 0 && (s->host.bvhT0 > t0 || s->host.bvhT1 < t1))
        return fail(WPT_E_STATE, "the acceleration structure does not cover [t0,t1]; call wpt_scene_update_bvh(t0,t1) "
                "(Scene::bvhNeedsUpdate, scene.hpp:145-149)");
    wpt_render_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts)
        o = *opts;
    cudaStream_t stream = static_cast<cudaStream_t>(o.stream);

    wptd::RenderParams P;
    std::memset(&P, 0, sizeof(P));
    rc = fillCamera(s, camera, t0, sensor->width, sensor->height, P, stream);
    if (rc != WPT_OK)
        return rc;
    P.sensorType = sensor->type;
    P.minDistToLight = sensor->min_dist_to_light;
    P.maxDistToLight = sensor->max_dist_to_light;
    P.minPathLen = sensor->min_path_len;
    P.maxPathLen = sensor->max_path_len;
    bool needOpl = false;
    if (sensor->type == WPT_SENSOR_TOF_AMCW) {
        needOpl = true;
        if (sensor->phase_count == 0)
            return fail(WPT_E_INVALID, "ToF sensor needs phase_count > 0");
        P.tofPixelArea = sensor->pixel_area;
        P.tofExposureTime = sensor->exposure_time;
        P.tofContrast = sensor->contrast;
        /* tau(): phaseImageIndex * (2.0f * pi) / phaseImageCount (sensor_tof_amcw.hpp:102-105) */
        P.tofTau = float(sensor->phase_index) * (2.0f * 3.14159274f) / float(sensor->phase_count);
        /* fracModfreqC(): double division, then cast (sensor_tof_amcw.hpp:107-110) */
        P.tofFracModfreqC = float(sensor->modulation_frequency / 299792458.0);
    } else {
        const float fmax = std::numeric_limits<float>::max();
        needOpl = !(sensor->min_dist_to_light <= 0.0f && sensor->max_dist_to_light >= fmax
                && sensor->min_path_len <= 0.0f && sensor->max_path_len >= fmax);
    }
    P.maxPathComponents = params->max_path_components;
    P.rrThreshold = params->rr_threshold;
    P.randomize = params->randomize_ray_over_pixel;
    P.minHit = params->min_hit_distance;
    P.samplesSqrt = samplesSqrt;
    P.samples = samplesSqrt * samplesSqrt;
    P.invSamples = 1.0f / float(P.samples);
    P.invSamplesSqrt = 1.0f / float(samplesSqrt);
    P.invWidth = 1.0f / float(sensor->width);
    P.invHeight = 1.0f / float(sensor->height);
    P.t0 = t0;
    P.t1 = t1;
    if (o.traversal < WPT_TRAVERSAL_WIDE || o.traversal > WPT_TRAVERSAL_QUAD)
        return fail(WPT_E_INVALID, "unknown traversal");
    P.traversal = (o.traversal >= WPT_TRAVERSAL_POOL ? WPT_TRAVERSAL_WIDE : o.traversal);
    if (P.traversal == WPT_TRAVERSAL_WIDE && !(P.minHit > 0.0f))
        P.traversal = WPT_TRAVERSAL_REFERENCE;  /* the ordered walk needs a positive lower bound */
    if (P.traversal == WPT_TRAVERSAL_WIDE && 3 * s->host.wideDepth + 1 > wptd::wideStackEntries())
        P.traversal = WPT_TRAVERSAL_REFERENCE;  /* deeper than the wide walk's stack */
    if (P.traversal == WPT_TRAVERSAL_REFERENCE && s->host.deviceBuilt)
        return fail(WPT_E_UNSUPPORTED, "this tree was built on the device (WPT_BUILDER_DEVICE_LBVH): the reference-order walk needs the "
                "host builder's binary tree; it is also what a non-positive minimum hit distance or a very deep tree falls back to");

    /* sub-streams per pixel: enough lanes to fill the GPU a few times over */
    uint32_t sub = 1;
    if (o.substreams > 0) {
        if (o.substreams > 128 || (o.substreams & (o.substreams - 1)) != 0)
            return fail(WPT_E_INVALID, "substreams must be a power of two <= 128");
        sub = uint32_t(o.substreams);
    } else {
        /* Many short work items: when the items run out, a slot is at most one item away from retiring, so
         * the last rounds of a frame (few live slots) stay short. 32 items per slot, at least one sample
         * per item, and at most 16 GB of partial sums (12 bytes per pixel and sub-stream). */
        const uint64_t pixels = uint64_t(sensor->width) * sensor->height;
        const uint64_t wantItems = 32ull * waveSlotLimit();
        while (sub < 128 && pixels * sub < wantItems && sub * 2 <= P.samples && pixels * (sub * 2) * 12ull <= (16ull << 30))
            sub *= 2;
    }
    while (sub > P.samples && sub > 1)
        sub /= 2;
    P.substreams = sub;
    P.subShift = 0;
    while ((1u << P.subShift) < sub)
        P.subShift++;
    /* a tile is 32 work items, or one pixel's sub-streams if those are more */
    P.tileShift = (P.subShift > 5u ? P.subShift : 5u);
    const uint32_t pixelsPerTile = (1u << P.tileShift) >> P.subShift;
    static const uint32_t tileWs[8] = { 8, 4, 4, 2, 2, 1, 1, 1 };   /* 32, 16, 8, 4, 2, 1, 1, 1 pixels per tile */
    P.tileW = tileWs[P.subShift];
    P.tileH = pixelsPerTile / P.tileW;
    P.tilesX = (sensor->width + P.tileW - 1) / P.tileW;
    const uint32_t tilesY = (sensor->height + P.tileH - 1) / P.tileH;
    P.tilesTotal = P.tilesX * tilesY;
    P.tileStride = (o.tile_stride == 0 ? 1 : o.tile_stride);
    P.tileFirst = o.tile_first;
    if (P.tileFirst >= P.tileStride)
        return fail(WPT_E_INVALID, "tile_first must be smaller than tile_stride");
    P.nLocalTiles = (P.tileFirst < P.tilesTotal ? (P.tilesTotal - P.tileFirst + P.tileStride - 1) / P.tileStride : 0);
    if ((uint64_t(P.nLocalTiles) << P.tileShift) >= 0xFFFFFFFFull)
        return fail(WPT_E_UNSUPPORTED, "image too large for 32-bit work item indices");
    P.nItems = P.nLocalTiles << P.tileShift;

    const size_t frameBytes = size_t(sensor->width) * sensor->height * 3 * sizeof(float);
    if (o.frame_on_device) {
        P.frame = frame;
    } else {
        WPT_CUDA(s->frame.reserve(frameBytes), "allocate frame");
        WPT_CUDA(cudaMemsetAsync(s->frame.p, 0, frameBytes, stream), "clear frame");
        P.frame = static_cast<float*>(s->frame.p);
    }
    P.partials = nullptr;
    if (P.substreams > 1) {
        WPT_CUDA(s->partials.reserve(size_t(sensor->width) * sensor->height * P.substreams * 3 * sizeof(float)), "allocate partial sums");
        P.partials = static_cast<float*>(s->partials.p);
    }
    P.counters = static_cast<unsigned long long*>(s->counters.p);
    WPT_CUDA(cudaMemsetAsync(s->counters.p, 0, wptd::CTR_WORDS * sizeof(unsigned long long), stream), "clear counters");
    P.itemCounter = P.counters + wptd::CTR_ITEM;
    P.blockDone = nullptr;
    MultiFrame fromOpts;
    if (!multi && o.shared_counter) {
        if (o.tile_stride > 1)
            return fail(WPT_E_INVALID, "shared_counter hands out all work items: tile_stride must be 0 or 1");
        fromOpts.itemCounter = static_cast<unsigned long long*>(o.shared_counter);
        multi = &fromOpts;
    }
    if (multi) {
        /* all devices draw blocks of WAVE_WINDOW work items from one counter; a block holds whole pixels (their
         * sub-streams are consecutive items, at most 128), so this device resolves exactly the pixels of its blocks */
        P.itemCounter = multi->itemCounter;
        if (P.substreams > 1) {
            const size_t blocks = size_t(P.nItems) / wptd::WAVE_WINDOW + 1;
            WPT_CUDA(s->blockDone.reserve(blocks), "allocate block flags");
            WPT_CUDA(cudaMemsetAsync(s->blockDone.p, 0, blocks, stream), "clear block flags");
            P.blockDone = static_cast<uint8_t*>(s->blockDone.p);
        }
    }

    /* pipeline */
    int pipeline = o.pipeline;
    if (pipeline < WPT_PIPELINE_AUTO || pipeline > WPT_PIPELINE_WAVEFRONT)
        return fail(WPT_E_INVALID, "unknown pipeline");
    if (pipeline == WPT_PIPELINE_AUTO)
        pipeline = defaultPipeline();
    if (multi)
        pipeline = WPT_PIPELINE_WAVEFRONT;      /* the shared counter hands out whole blocks of work items */
    const int features = sceneFeatures(s, needOpl);
    if (pipeline == WPT_PIPELINE_WAVEFRONT && P.nItems > 0) {
        /* slots: as many as the limit allows for big frames (long launches amortise their ramp and tail), but
         * every slot should see several work items, or the whole frame is one ramp and one tail */
        uint32_t slots = waveSlotLimit();
        const uint32_t share = (std::getenv("WPT_WAVE_SLOTS") ? P.nItems : std::max(P.nItems / WPT_ITEMS_PER_SLOT, 65536u));
        if (slots > share)
            slots = share;
        if (slots > P.nItems)
            slots = P.nItems;
        P.W.nSlots = slots;
        const size_t n = slots;
        /* wave[1], [2], [3] hold interleaved pairs (see WaveState) */
        enum { W_ITEM, W_PRNG_ACC, W_RAY, W_SHADOW, W_ATT, W_HIT, W_CTL, W_CODE, W_OPL, W_RAYRI, W_SHADOWRI, W_SHPRNG,
               W_QUEUE, W_TRACEQ, W_WINDOW };
        const size_t sizes[12] = { sizeof(int4), 2 * sizeof(uint4), 2 * sizeof(float4), 2 * sizeof(float4), sizeof(float4),
            sizeof(uint4), sizeof(uint32_t), sizeof(uint8_t), sizeof(float4), sizeof(float4), sizeof(float4), sizeof(uint4) };
        const int used = (features != 0 ? 12 : 8);
        for (int k = 0; k < used; k++)
            WPT_CUDA(s->wave[k].reserve(n * sizes[k]), "allocate wavefront state");
        P.W.item = static_cast<int4*>(s->wave[W_ITEM].p);
        P.W.prng = static_cast<uint4*>(s->wave[W_PRNG_ACC].p);
        P.W.acc = static_cast<float4*>(s->wave[W_PRNG_ACC].p) + 1;
        P.W.rayO = static_cast<float4*>(s->wave[W_RAY].p);
        P.W.rayD = static_cast<float4*>(s->wave[W_RAY].p) + 1;
        P.W.shD = static_cast<float4*>(s->wave[W_SHADOW].p);
        P.W.pendRad = static_cast<float4*>(s->wave[W_SHADOW].p) + 1;
        P.W.att = static_cast<float4*>(s->wave[W_ATT].p);
        P.W.hit = static_cast<uint4*>(s->wave[W_HIT].p);
        P.W.ctl = static_cast<uint32_t*>(s->wave[W_CTL].p);
        P.W.code = static_cast<uint8_t*>(s->wave[W_CODE].p);
        P.W.opl = P.W.rayRi = P.W.shadowRi = nullptr;
        P.W.shPrng = nullptr;
        if (features != 0) {
            P.W.opl = static_cast<float4*>(s->wave[W_OPL].p);
            P.W.rayRi = static_cast<float4*>(s->wave[W_RAYRI].p);
            P.W.shadowRi = static_cast<float4*>(s->wave[W_SHADOWRI].p);
            P.W.shPrng = static_cast<uint4*>(s->wave[W_SHPRNG].p);
        }
        WPT_CUDA(s->wave[W_QUEUE].reserve(n * wptd::MAT_CLASSES * sizeof(uint32_t)), "allocate shade queues");
        P.W.queue = static_cast<uint32_t*>(s->wave[W_QUEUE].p);
        P.W.queueStride = slots;
        WPT_CUDA(s->wave[W_TRACEQ].reserve(2 * n * sizeof(uint32_t)), "allocate trace queue");
        P.W.traceQueue = static_cast<uint32_t*>(s->wave[W_TRACEQ].p);
        const size_t windows = (n + wptd::WAVE_WINDOW - 1) / wptd::WAVE_WINDOW + 2;
        WPT_CUDA(s->wave[W_WINDOW].reserve(windows * sizeof(uint2)), "allocate window state");
        WPT_CUDA(cudaMemsetAsync(s->wave[W_WINDOW].p, 0, windows * sizeof(uint2), stream), "clear window state");
        P.W.window = static_cast<uint2*>(s->wave[W_WINDOW].p);
        P.W.group = 0;
        P.W.cell = nullptr;
        P.W.sortedQueue = nullptr;
        P.W.binCount = P.W.binBase = nullptr;
        if (waveBinning()) {
            WPT_CUDA(s->waveBin[0].reserve(n * sizeof(uint16_t)), "allocate ray cells");
            WPT_CUDA(s->waveBin[1].reserve(2 * n * sizeof(uint32_t)), "allocate binned trace queue");
            WPT_CUDA(s->waveBin[2].reserve(size_t(2) * wptd::RAY_BINS * sizeof(uint32_t)), "allocate bin counters");
            WPT_CUDA(s->waveBin[3].reserve(size_t(4) * wptd::RAY_BINS * sizeof(uint32_t)), "allocate bin offsets");
            WPT_CUDA(cudaMemsetAsync(s->waveBin[2].p, 0, size_t(2) * wptd::RAY_BINS * sizeof(uint32_t), stream), "clear bin counters");
            P.W.cell = static_cast<uint16_t*>(s->waveBin[0].p);
            P.W.sortedQueue = static_cast<uint32_t*>(s->waveBin[1].p);
            P.W.binCount = static_cast<uint32_t*>(s->waveBin[2].p);
            P.W.binBase = static_cast<uint32_t*>(s->waveBin[3].p);
        }
        if (s->d.simpleWalk) {
            rc = reservePoolSpill(s, P);
            if (rc != WPT_OK)
                return rc;
        }
    }

    const bool timed = (stats != nullptr) || !o.frame_on_device;
    int launches = 0;
    if (timed)
        WPT_CUDA(cudaEventRecord(s->evStart, stream), "cudaEventRecord");
    wptd::DScene dscene;
    rc = launchScene(s, t0, t1, stream, dscene);
    if (rc != WPT_OK)
        return rc;
    if (P.nLocalTiles > 0) {
        if (pipeline == WPT_PIPELINE_WAVEFRONT)
            WPT_CUDA(wptd::launchWavefront(features, dscene, P, s->numSMs, stream, s->stream2, s->hostDone, s->evWave, &launches), "wavefront launch");
        else
            WPT_CUDA(wptd::launchRender(features, dscene, P, s->numSMs, stream, &launches), "render kernel launch");
    }
    if (timed)
        WPT_CUDA(cudaEventRecord(s->evStop, stream), "cudaEventRecord");
    if (!o.frame_on_device)
        WPT_CUDA(cudaMemcpyAsync(frame, s->frame.p, frameBytes, cudaMemcpyDeviceToHost, stream), "copy frame to host");
    if (timed) {
        unsigned long long c[wptd::CTR_NESTED + 1] = { 0 };
        WPT_CUDA(cudaMemcpyAsync(c, s->counters.p, sizeof(c), cudaMemcpyDeviceToHost, stream), "copy counters");
        WPT_CUDA(cudaStreamSynchronize(stream), "render kernel");
        if (stats) {
            std::memset(stats, 0, sizeof(*stats));
            stats->samples = c[wptd::CTR_SAMPLES];
            stats->rays = c[wptd::CTR_RAYS];
            stats->nested_queries = c[wptd::CTR_NESTED];
            stats->tiles = P.nLocalTiles;
            cudaEventElapsedTime(&stats->kernel_ms, s->evStart, s->evStop);
            stats->launches = uint32_t(launches);
            stats->total_ms = float(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wallStart).count());
        }
    }
    return WPT_OK;
}

/* One frame on all devices of the scene (wpt_scene_create with WPT_DEVICE_ALL): one host thread per device renders
 * into that device's own frame, the blocks of work items come from ONE atomic counter in the first device's memory
 * (reached over NVLink by the others), the frames are added up on the first device. */
static int renderMulti(wpt_scene* s, const wpt_camera_desc* camera, const wpt_sensor_desc* sensor,
        const wpt_params* params, uint32_t samplesSqrt, float t0, float t1,
        const wpt_render_opts* opts, float* frame, wpt_stats* stats)
{
    auto wallStart = std::chrono::steady_clock::now();
    int rc = checkReady(s);
    if (rc != WPT_OK)
        return rc;
    if (!camera || !sensor || !params || !frame)
        return fail(WPT_E_INVALID, "null argument");
    wpt_render_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts)
        o = *opts;
    if (o.tile_stride > 1)
        return fail(WPT_E_INVALID, "a scene on several devices renders whole frames (tile_stride must be 0 or 1)");
    cudaStream_t stream = static_cast<cudaStream_t>(o.stream);
    const size_t n = 1 + s->replicas.size();
    const size_t frameFloats = size_t(sensor->width) * sensor->height * 3;
    const size_t frameBytes = frameFloats * sizeof(float);

    WPT_CUDA(s->sharedCounter.reserve(128), "allocate shared work counter");
    WPT_CUDA(cudaMemsetAsync(s->sharedCounter.p, 0, 128, stream), "clear shared work counter");
    float* masterFrame = frame;
    if (!o.frame_on_device) {
        WPT_CUDA(s->frame.reserve(frameBytes), "allocate frame");
        masterFrame = static_cast<float*>(s->frame.p);
    }
    WPT_CUDA(cudaMemsetAsync(masterFrame, 0, frameBytes, stream), "clear frame");
    WPT_CUDA(cudaStreamSynchronize(stream), "prepare frame");      /* the counter is zero before any device draws from it */
    MultiFrame multi;
    multi.itemCounter = static_cast<unsigned long long*>(s->sharedCounter.p);

    std::vector<int> rcs(n, WPT_OK);
    std::vector<std::string> errors(n);
    std::vector<wpt_stats> st(n);
    auto work = [&](size_t i) {
        wpt_scene* d = (i == 0 ? s : s->replicas[i - 1].get());
        wpt_render_opts oi = o;
        oi.frame_on_device = 1;
        oi.tile_first = 0;
        oi.tile_stride = 1;
        float* target = masterFrame;
        if (i > 0) {
            oi.stream = nullptr;
            if (cudaSetDevice(d->device) != cudaSuccess || d->frame.reserve(frameBytes) != cudaSuccess
                    || cudaMemsetAsync(d->frame.p, 0, frameBytes, 0) != cudaSuccess) {
                rcs[i] = WPT_E_CUDA;
                errors[i] = "allocate frame on device " + std::to_string(d->device);
                cudaGetLastError();
                return;
            }
            target = static_cast<float*>(d->frame.p);
        }
        rcs[i] = renderSingle(d, camera, sensor, params, samplesSqrt, t0, t1, &oi, target, &st[i], &multi);
        if (rcs[i] != WPT_OK)
            errors[i] = g_lastError;
    };
    std::vector<std::thread> threads;
    for (size_t i = 1; i < n; i++)
        threads.emplace_back(work, i);
    work(0);
    for (std::thread& t : threads)
        t.join();
    WPT_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    for (size_t i = 0; i < n; i++)
        if (rcs[i] != WPT_OK)
            return fail(rcs[i], "device " + std::to_string(i == 0 ? s->device : s->replicas[i - 1]->device) + ": " + errors[i]);

    /* every pixel was rendered by exactly one device and is zero in the other frames: the sum is an exact gather */
    WPT_CUDA(s->peerStage.reserve(frameBytes), "allocate frame staging");
    for (size_t i = 1; i < n; i++) {
        wpt_scene* d = s->replicas[i - 1].get();
        WPT_CUDA(cudaMemcpyPeerAsync(s->peerStage.p, s->device, d->frame.p, d->device, frameBytes, stream), "copy frame between devices");
        WPT_CUDA(wptd::launchAddFrame(masterFrame, static_cast<const float*>(s->peerStage.p), frameFloats, stream), "add frames");
    }
    if (!o.frame_on_device)
        WPT_CUDA(cudaMemcpyAsync(frame, masterFrame, frameBytes, cudaMemcpyDeviceToHost, stream), "copy frame to host");
    WPT_CUDA(cudaStreamSynchronize(stream), "combine frames");
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (size_t i = 0; i < n; i++) {
            stats->samples += st[i].samples;
            stats->rays += st[i].rays;
            stats->nested_queries += st[i].nested_queries;
            stats->launches += st[i].launches;
            stats->kernel_ms = std::max(stats->kernel_ms, st[i].kernel_ms);
        }
        stats->tiles = st[0].tiles;
        stats->total_ms = float(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wallStart).count());
    }
    return WPT_OK;
}

int wpt_render(wpt_scene* s, const wpt_camera_desc* camera, const wpt_sensor_desc* sensor,
        const wpt_params* params, uint32_t samplesSqrt, float t0, float t1,
        const wpt_render_opts* opts, float* frame, wpt_stats* stats)
{
    if (s && !s->replicas.empty())
        return renderMulti(s, camera, sensor, params, samplesSqrt, t0, t1, opts, frame, stats);
    return renderSingle(s, camera, sensor, params, samplesSqrt, t0, t1, opts, frame, stats, nullptr);
}

static int runQuery(wpt_scene* s, const wpt_camera_desc* camera, uint32_t width, uint32_t height,
        const float* origins, const float* directions, uint64_t nRays, float time, float minHit,
        const wpt_render_opts* opts, int32_t* ids, float* a, uint64_t* counters)
{
    int rc = checkReady(s);
    if (rc != WPT_OK)
        return rc;
    if (!ids || !a)
        return fail(WPT_E_INVALID, "null output pointer");
    wpt_render_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts)
        o = *opts;
    cudaStream_t stream = static_cast<cudaStream_t>(o.stream);
    wptd::RenderParams P;
    std::memset(&P, 0, sizeof(P));
    if (camera) {
        rc = fillCamera(s, camera, time, width, height, P, stream);
        if (rc != WPT_OK)
            return rc;
        nRays = uint64_t(width) * height;
    }
    P.t0 = P.t1 = time;
    P.minHit = minHit;
    if (o.traversal < WPT_TRAVERSAL_WIDE || o.traversal > WPT_TRAVERSAL_QUAD)
        return fail(WPT_E_INVALID, "unknown traversal");
    P.traversal = o.traversal;
    if (o.traversal != WPT_TRAVERSAL_REFERENCE && (!(minHit > 0.0f) || 3 * s->host.wideDepth + 1 > wptd::wideStackEntries()))
        P.traversal = WPT_TRAVERSAL_REFERENCE;
    if (P.traversal == WPT_TRAVERSAL_REFERENCE && s->host.deviceBuilt)
        return fail(WPT_E_UNSUPPORTED, "this tree was built on the device (WPT_BUILDER_DEVICE_LBVH): the reference-order walk needs the "
                "host builder's binary tree; it is also what a non-positive minimum hit distance or a very deep tree falls back to");
    if (P.traversal >= WPT_TRAVERSAL_POOL && (!s->d.simpleWalk || counters))
        return fail(WPT_E_UNSUPPORTED, "WPT_TRAVERSAL_POOL / _PERSISTENT: queries through the pipeline's trace kernels are for scenes "
                "without media and animated hitables, and do not count node visits");
    if (nRays == 0)
        return WPT_OK;
    if (nRays >= (1ull << 31))
        return fail(WPT_E_INVALID, "too many rays for one query call");
    const float* dO = nullptr;
    const float* dD = nullptr;
    if (!camera) {
        if (!origins || !directions)
            return fail(WPT_E_INVALID, "null ray arrays");
        WPT_CUDA(s->qOrigins.upload(origins, nRays * 3 * sizeof(float), stream), "upload ray origins");
        WPT_CUDA(s->qDirections.upload(directions, nRays * 3 * sizeof(float), stream), "upload ray directions");
        dO = s->qOrigins.as<float>();
        dD = s->qDirections.as<float>();
    }
    WPT_CUDA(s->qIds.reserve(nRays * sizeof(int32_t)), "allocate ids");
    WPT_CUDA(s->qDist.reserve(nRays * sizeof(float)), "allocate distances");
    wptd::TraceStats* dStats = nullptr;
    if (counters) {
        WPT_CUDA(s->qStats.reserve(sizeof(wptd::TraceStats)), "allocate stats");
        WPT_CUDA(cudaMemsetAsync(s->qStats.p, 0, sizeof(wptd::TraceStats), stream), "clear stats");
        dStats = static_cast<wptd::TraceStats*>(s->qStats.p);
    }
    wptd::DScene dscene;
    rc = launchScene(s, time, time, stream, dscene);
    if (rc != WPT_OK)
        return rc;
    const bool queryTiming = std::getenv("WPT_QUERY_TIMING") != nullptr;   /* experiment knob: kernel time to stderr */
    if (queryTiming)
        WPT_CUDA(cudaEventRecord(s->evStart, stream), "cudaEventRecord");
    if (P.traversal >= WPT_TRAVERSAL_POOL) {
        const bool pooled = (P.traversal == WPT_TRAVERSAL_POOL);
        const int mode = pooled ? 1 : (P.traversal == WPT_TRAVERSAL_QUAD ? 2 : 0);
        P.traversal = WPT_TRAVERSAL_WIDE;
        /* the rays as wavefront slots: rays (wave[2]), answers (wave[5]), trace queue (wave[13]) */
        WPT_CUDA(s->wave[2].reserve(nRays * 2 * sizeof(float4)), "allocate query slots");
        WPT_CUDA(s->wave[5].reserve(nRays * sizeof(uint4)), "allocate query answers");
        WPT_CUDA(s->wave[13].reserve(nRays * sizeof(uint32_t)), "allocate query queue");
        P.W.rayO = static_cast<float4*>(s->wave[2].p);
        P.W.rayD = static_cast<float4*>(s->wave[2].p) + 1;
        P.W.shD = P.W.rayD;
        P.W.hit = static_cast<uint4*>(s->wave[5].p);
        P.W.traceQueue = static_cast<uint32_t*>(s->wave[13].p);
        P.W.nSlots = uint32_t(nRays);
        P.W.group = 0;
        P.counters = static_cast<unsigned long long*>(s->counters.p);
        rc = reservePoolSpill(s, P);
        if (rc != WPT_OK)
            return rc;
        WPT_CUDA(wptd::launchPoolQuery(dscene, P, s->numSMs, mode, dO, dD, nRays, time, static_cast<int32_t*>(s->qIds.p),
                    static_cast<float*>(s->qDist.p), stream), "pooled query launch");
    } else
    WPT_CUDA(wptd::launchQuery(sceneFeatures(s, false), counters != nullptr, dscene, P, dO, dD, nRays, time,
                static_cast<int32_t*>(s->qIds.p), static_cast<float*>(s->qDist.p), dStats, stream), "query kernel launch");
    if (queryTiming) {
        WPT_CUDA(cudaEventRecord(s->evStop, stream), "cudaEventRecord");
        WPT_CUDA(cudaEventSynchronize(s->evStop), "query kernel");
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, s->evStart, s->evStop);
        std::fprintf(stderr, "wpt query: %llu rays in %.3f ms (%.1f Mrays/s)\n", (unsigned long long)nRays, ms, double(nRays) / ms / 1e3);
    }
    WPT_CUDA(cudaMemcpyAsync(ids, s->qIds.p, nRays * sizeof(int32_t), cudaMemcpyDeviceToHost, stream), "copy ids");
    WPT_CUDA(cudaMemcpyAsync(a, s->qDist.p, nRays * sizeof(float), cudaMemcpyDeviceToHost, stream), "copy distances");
    wptd::TraceStats hs = { 0, 0, 0 };
    if (counters)
        WPT_CUDA(cudaMemcpyAsync(&hs, s->qStats.p, sizeof(hs), cudaMemcpyDeviceToHost, stream), "copy stats");
    WPT_CUDA(cudaStreamSynchronize(stream), "query kernel");
    if (counters) {
        counters[0] = hs.nodes;
        counters[1] = hs.leaves;
        counters[2] = hs.nested;
    }
    return WPT_OK;
}

int wpt_primary_hits(wpt_scene* s, const wpt_camera_desc* camera, uint32_t width, uint32_t height,
        float t, float minHitDistance, const wpt_render_opts* opts, int32_t* ids, float* a, uint64_t* counters)
{
    if (!camera || width == 0 || height == 0)
        return fail(WPT_E_INVALID, "invalid camera or image size");
    return runQuery(s, camera, width, height, nullptr, nullptr, 0, t, minHitDistance, opts, ids, a, counters);
}

int wpt_get_ground_truth(wpt_scene* s, const wpt_camera_desc* camera, uint32_t width, uint32_t height,
        float t0, float tPrev, float tNext, float minHitDistance, uint32_t bits,
        const int32_t* materialIds, uint32_t nMaterialIds, const wpt_render_opts* opts, const wpt_ground_truth* out)
{
    int rc = checkReady(s);
    if (rc != WPT_OK)
        return rc;
    if (!camera || width == 0 || height == 0 || !out)
        return fail(WPT_E_INVALID, "invalid camera, image size or output");
    if (!(tPrev <= t0) || !(t0 <= tNext))
        return fail(WPT_E_INVALID, "getGroundTruth needs t_prev <= t0 <= t_next (wurblpt.hpp:648-649)");
    bits &= uint32_t(WPT_GT_ALL);
    if ((bits & (WPT_GT_PIXEL_OFFSET_TO_PREV | WPT_GT_PIXEL_OFFSET_TO_NEXT))
            && (camera->surround_mode != WPT_SURROUND_OFF || camera->stereoscopic_distance > 0.0f))
        return fail(WPT_E_UNSUPPORTED, "pixel space offsets exist for perspective, non-stereoscopic cameras only (camera.hpp:207-208)");
    for (int i = 0; i < WPT_GT_FLOAT_IMAGES; i++)
        if ((bits & (1u << i)) && !out->aov[i])
            return fail(WPT_E_INVALID, "a selected ground truth image has a null pointer");
    if ((bits & WPT_GT_MATERIALS) && !out->materials)
        return fail(WPT_E_INVALID, "the material image has a null pointer");
    wpt_render_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts)
        o = *opts;
    cudaStream_t stream = static_cast<cudaStream_t>(o.stream);
    wptd::RenderParams P;
    std::memset(&P, 0, sizeof(P));
    rc = fillCamera(s, camera, t0, width, height, P, stream);
    if (rc != WPT_OK)
        return rc;
    P.t0 = P.t1 = t0;
    P.minHit = minHitDistance;
    P.traversal = (o.traversal == WPT_TRAVERSAL_REFERENCE ? WPT_TRAVERSAL_REFERENCE : WPT_TRAVERSAL_WIDE);
    if (P.traversal == WPT_TRAVERSAL_WIDE && (!(minHitDistance > 0.0f) || 3 * s->host.wideDepth + 1 > wptd::wideStackEntries()))
        P.traversal = WPT_TRAVERSAL_REFERENCE;
    if (P.traversal == WPT_TRAVERSAL_REFERENCE && s->host.deviceBuilt)
        return fail(WPT_E_UNSUPPORTED, "this tree was built on the device (WPT_BUILDER_DEVICE_LBVH): the reference-order walk needs the "
                "host builder's binary tree; it is also what a non-positive minimum hit distance or a very deep tree falls back to");

    wptd::GroundTruthParams G;
    std::memset(&G, 0, sizeof(G));
    const float times[3] = { t0, tPrev, tNext };
    for (int i = 0; i < 3; i++) {
        /* inverse(camera.at(t)) (transformation.hpp:157-163); only rotation and translation are used */
        const wpt::Trs T = cameraPose(camera, times[i]);
        const wpt::Quat invR = wpt::Quat{ -T.r.x, -T.r.y, -T.r.z, T.r.w };
        const wpt::V3 invS = wpt::v3(1.0f / T.s.x, 1.0f / T.s.y, 1.0f / T.s.z);
        const wpt::V3 invT = wpt::rotate(invR, wpt::v3(-T.t.x * invS.x, -T.t.y * invS.y, -T.t.z * invS.z));
        G.invCamR[i] = make_float4(invR.x, invR.y, invR.z, invR.w);
        G.invCamT[i] = make_float3(invT.x, invT.y, invT.z);
    }
    G.tPrev = tPrev;
    G.tNext = tNext;
    G.P00 = 2.0f / (camera->r - camera->l);
    G.P11 = 2.0f / (camera->t - camera->b);
    G.P03 = (camera->r + camera->l) / (camera->r - camera->l);
    G.P13 = (camera->t + camera->b) / (camera->t - camera->b);
    G.bits = bits;
    const size_t pixels = size_t(width) * height;
    size_t words = 0;
    for (int i = 0; i < WPT_GT_FLOAT_IMAGES; i++)
        if (bits & (1u << i))
            words += pixels * WPT_GT_COMPONENTS[i];
    if (bits & WPT_GT_MATERIALS)
        words += pixels;
    if (words == 0)
        return WPT_OK;
    WPT_CUDA(s->gtImages.reserve(words * sizeof(float)), "allocate ground truth images");
    float* cursor = static_cast<float*>(s->gtImages.p);
    for (int i = 0; i < WPT_GT_FLOAT_IMAGES; i++) {
        if (bits & (1u << i)) {
            G.aov[i] = cursor;
            cursor += pixels * WPT_GT_COMPONENTS[i];
        }
    }
    if (bits & WPT_GT_MATERIALS)
        G.materials = reinterpret_cast<int32_t*>(cursor);
    if (materialIds && nMaterialIds > 0) {
        WPT_CUDA(s->gtMaterialIds.upload(materialIds, size_t(nMaterialIds) * sizeof(int32_t), stream), "upload material ids");
        G.materialIds = s->gtMaterialIds.as<int32_t>();
        G.nMaterialIds = nMaterialIds;
    }
    wptd::DScene dscene;
    rc = launchScene(s, t0, t0, stream, dscene);
    if (rc != WPT_OK)
        return rc;
    WPT_CUDA(wptd::launchGroundTruth(sceneFeatures(s, false), dscene, P, G, stream), "ground truth kernel launch");
    for (int i = 0; i < WPT_GT_FLOAT_IMAGES; i++)
        if (bits & (1u << i))
            WPT_CUDA(cudaMemcpyAsync(out->aov[i], G.aov[i], pixels * WPT_GT_COMPONENTS[i] * sizeof(float),
                        cudaMemcpyDeviceToHost, stream), "copy ground truth image");
    if (bits & WPT_GT_MATERIALS)
        WPT_CUDA(cudaMemcpyAsync(out->materials, G.materials, pixels * sizeof(int32_t), cudaMemcpyDeviceToHost, stream), "copy material image");
    WPT_CUDA(cudaStreamSynchronize(stream), "ground truth kernel");
    return WPT_OK;
}

int wpt_trace_rays(wpt_scene* s, const float* origins, const float* directions, uint64_t nRays,
        float time, float minHitDistance, const wpt_render_opts* opts, int32_t* ids, float* a, uint64_t* counters)
{
    /* Ray's contract (ray.hpp:47-49, constants.hpp:42): a finite direction of unit length. The walks rely on it:
     * |1 / d| >= 1 keeps the boxes of unused child slots (all planes at FLT_MAX) out of every ray's reach. */
    if (origins && directions) {
        for (uint64_t i = 0; i < nRays; i++) {
            const float* d = directions + 3 * i;
            const float* o = origins + 3 * i;
            const float len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if (!(std::fabs(len2 - 1.0f) < 0.00001f) || !std::isfinite(o[0]) || !std::isfinite(o[1]) || !std::isfinite(o[2]))
                return fail(WPT_E_INVALID, "ray " + std::to_string(i) + ": origin not finite or direction not of unit length (ray.hpp:47-49)");
        }
    }
    return runQuery(s, nullptr, 0, 0, origins, directions, nRays, time, minHitDistance, opts, ids, a, counters);
}

int wpt_prng_kat(const uint32_t* seeds, uint32_t nSeeds, uint32_t nDraws, float* out)
{
    int rc = selectDevice(0);
    if (rc != WPT_OK)
        return rc;
    if (!seeds || !out)
        return fail(WPT_E_INVALID, "null argument");
    DeviceBuf dSeeds, dOut;
    WPT_CUDA(dSeeds.upload(seeds, size_t(nSeeds) * sizeof(uint32_t)), "upload seeds");
    WPT_CUDA(dOut.reserve(size_t(nSeeds) * nDraws * sizeof(float)), "allocate output");
    WPT_CUDA(wptd::launchPrng(dSeeds.as<uint32_t>(), nSeeds, nDraws, static_cast<float*>(dOut.p), 0), "prng kernel launch");
    WPT_CUDA(cudaMemcpy(out, dOut.p, size_t(nSeeds) * nDraws * sizeof(float), cudaMemcpyDeviceToHost), "copy prng output");
    return WPT_OK;
}

}
