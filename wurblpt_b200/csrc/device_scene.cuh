/*
 * device_scene.cuh -- layout of the scene in HBM and of the per-launch parameter blocks.
 *
 * Everything a kernel needs is reachable from two small structs passed by value as
 * __grid_constant__ kernel parameters (constant bank, no global load to find a pointer):
 *   DScene        pointers to the device-resident scene arrays
 *   RenderParams  camera, sensor, Parameters, tiling and output of one mcpt() call
 *
 * Arrays (all cudaMalloc'd, 256-byte aligned; 16-byte elements are read with 128-bit loads):
 *   node4     Node4[ ]         128 B per node: six float4 of child slabs, int4 children, pad
 *   bnode     float4[2*n]      the reference-layout binary tree (32 B per node), top level at 0,
 *                              nested medium-boundary trees behind it
 *   tri       float4[3*slots]  world-space corners; .w of the three = instance, triangle number,
 *                              animation index + 1 (0 = static)
 *   sph       float4[3*slots]  (centre, radius) (rotation quaternion) (material, animation+1, -, -)
 *   med       float4[slots]    (density, material, nested root, -)
 *   *Global   uint32[slots]    slot -> global hitable id (for reporting)
 *   materials wpt_material[ ]  128 B each, the C-ABI struct as is
 *   instances wpt_instance[ ]  128 B each
 *   meshes    wpt_mesh[ ]
 *   vertices  float[ ], indices uint32[ ]   interleaved vertex data for shading normals / uv / tangents
 *   textures  DTexture[ ], texels float4[ ] image textures decoded to RGBA float once at upload
 *   envM / envMs / envMcs      importance tables of the environment map
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "wurblpt_b200.h"

namespace wptd {

struct DTexture {
    int32_t type;
    int32_t src;
    int32_t width, height;
    int32_t horiz, vert;
    uint32_t texelBase;         /* index into DScene::texels */
    int32_t pad;
    float4 color0, color1;
    float2 coordFactor, coordOffset;
    float4 valFactor, valOffset;
};

struct DScene {
    const float4* node4;
    const float4* node4c;       /* the same wide nodes child by child (32 B per child: lo.xyz hi.x | hi.yz child pad): what the
                                   quad trace kernel reads, one child per lane */
    const float4* bnode;
    const float4* tri;
    const uint4* triShade;      /* per triangle slot: float offsets of the 3 vertex attribute blocks in `vertices`, mesh flags */
    const float4* sph;
    const float4* med;
    const uint32_t* triGlobal;
    const uint32_t* sphGlobal;
    const uint32_t* medGlobal;
    const uint32_t* hotRefs;
    const uint8_t* triClass;    /* material class (MC_*) of the hitable in each triangle / sphere / medium slot */
    const uint8_t* sphClass;
    const uint8_t* medClass;
    const wpt_material* materials;
    const wpt_instance* instances;
    const wpt_mesh* meshes;
    const float* vertices;
    const uint32_t* indices;
    const DTexture* textures;
    const float4* texels;
    const wpt_animation* animations;
    const wpt_keyframe* keyframes;
    const float* animTable;     /* all rays of the launch share one time: per animation 32 floats, evaluated on the host
                                   for that time (translation 3, rotation 4, scaling 3, mat4 as 3x4 12, normal matrix 9,
                                   pad); nullptr = every thread evaluates the key frames at its ray's time */
    const float* envM;
    const int32_t* envMs;
    const float* envMcs;
    const uint32_t* envGuide;   /* 2^envGuideBits + 1 entries: first bin with Mcs >= j / 2^envGuideBits */
    int32_t envGuideBits;       /* 0: no guide table */
    int32_t root4;              /* encoded child of the 4-wide tree */
    int32_t prefetchLeaves;     /* triangle data is larger than an SM's L1: ask for a leaf's triangles on arrival */
    uint32_t classMask;         /* bit c: some hitable of the scene has material class c */
    int32_t simpleWalk;         /* no media, no animated hitables: a walk needs no Prng and no animation cache */
    float3 cellLo, cellScale;   /* grid of RAY_CELLS^3 cells over the scene's box: cell = (p - cellLo) * cellScale */
    uint32_t nHot;
    uint32_t nAnimations;
    int32_t envType, envCompat, envN;
    int32_t envTex[6];
};

constexpr int MAX_CACHED_ANIMATIONS = 16;   /* per-thread animation cache entries (AnimationCache), filled lazily; the last
                                               one is shared by all animations beyond the first 15 */

/* Wavefront slot arrays (structure of arrays). The kernels reach the slots through queues, i.e. in
 * scattered order, so fields that are always used together are interleaved in pairs that fill one
 * 32-byte sector: (prng, acc), (rayO, rayD), (shD, pendRad) are indexed [2 * slot]; the second
 * pointer of a pair is the first plus one element. The other arrays are indexed [slot].
 *
 * A slot has up to TWO closest-hit queries pending per round, both leaving the same point rayO:
 *   B  the camera ray / the next path segment (direction rayD), state ST_MAIN
 *   S  the shadow ray of the next-event estimation set up at the same path vertex (direction shD)
 * so a path vertex costs one round (trace both, shade) instead of two. */
struct WaveState {
    int4* item;         /* pixel, sampleIndex, sub-stream, - : only the generate kernel touches it */
    uint32_t* ctl;      /* state (bits 0-2) | kind of the pending shadow query (bits 3-4) | path component << 8 */
    uint8_t* code;      /* what the window-scanning kernels need of a slot in one byte (slotCode()) */
    uint4* prng;        /* xoshiro128+ state of the item's stream */
    float4* acc;        /* sample accumulator of the item */
    float4* rayO;       /* origin xyz of the pending queries, time */
    float4* rayD;       /* query B: direction xyz; .w = geometric path length so far */
    float4* shD;        /* query S: direction xyz; .w = the chosen hot spot's leaf reference */
    float4* pendRad;    /* query S: light sample to add if visible */
    float4* att;        /* path attenuation of the segment B */
    uint4* hit;         /* answers of the trace kernel: (distance bits, leaf reference) of B, then of S */
    float4* opl; float4* rayRi; float4* shadowRi;  /* F_OPL only: optical path length, refractive index of B and of S */
    uint4* shPrng;      /* F_MEDIA only: the stream the walk of query S draws its free-flight samples from */
    uint32_t* queue;    /* filled by the sort kernel: per material class, the slots whose surface hit is to be shaded;
                           class c starts at c * queueStride */
    uint32_t* traceQueue;   /* 2 * nSlots entries (slot << 1 | query) filled by the generate kernel: camera rays
                               from the front, bounce and shadow rays from the back */
    uint2* window;      /* per window of WAVE_WINDOW slots: the block of work items it hands out (next, end) */
    /* Binning of the bounce and shadow queries by the grid cell of their origin (rays that start close to each other
     * walk the same part of the tree): nullptr = the trace kernel takes them in the order the windows listed them. */
    uint16_t* cell;         /* per slot: Morton number of the cell of rayO (written with the queries) */
    uint32_t* sortedQueue;  /* the bounce / shadow part of the trace queue, bin by bin */
    uint32_t* binCount;     /* RAY_BINS counters (generate kernel), zeroed by the scan kernel */
    uint32_t* binBase;      /* RAY_BINS start offsets + RAY_BINS cursors (scan / scatter kernels) */
    uint32_t nSlots;
    uint32_t group;     /* slot group (0 or 1): selects the group's trace ticket / queue length counters */
    uint32_t queueStride;
};
constexpr uint32_t WAVE_WINDOW = 256;       /* slots per window = work items per block */
constexpr uint32_t RAY_CELLS = 32;          /* grid cells per axis for the binning of queries */
constexpr uint32_t RAY_BINS = RAY_CELLS * RAY_CELLS * RAY_CELLS * 2;   /* cell x (path segment | shadow ray) */

/* Material classes of the shading queues: a class has a kernel of its own in which the material type is a
 * compile-time constant. Everything else (lights, mirrors, phase functions, two-sided materials whose sides differ)
 * is shaded by the generic kernel. */
enum : int { MC_GENERIC = 0, MC_LAMBERTIAN = 1, MC_GGX = 2, MC_GLASS = 3, MC_MODPHONG = 4, MAT_CLASSES = 5 };

/* Counter words (64 bit). Every hot counter has a 128-byte line of its own so that atomics on different
 * counters do not meet in one L2 sector. */
enum : int {
    CTR_ITEM = 0,           /* work item ticket */
    CTR_SAMPLES = 16,       /* samples started */
    CTR_RAYS = 32,          /* closest-hit queries */
    CTR_NESTED = 48,        /* nested medium boundary queries */
    CTR_DONE = 64,          /* wavefront slots that found no item left */
    CTR_TICKET = 80,        /* + 16 * group: queue ticket of the persistent trace kernel */
    CTR_TRACE_CAMERA = 112, /* + 16 * group: trace queue, camera rays */
    CTR_TRACE_OTHER = 144,  /* + 16 * group: trace queue, bounce and shadow rays */
    CTR_MATERIAL = 176,     /* + 16 * (group * MAT_CLASSES + class): material queue lengths */
    CTR_WORDS = 176 + 16 * 2 * MAT_CLASSES
};

struct RenderParams {
    /* camera (camera.hpp:42-107, optics.hpp) */
    int32_t surround;
    float stereo;
    float l, r, b, t;
    int32_t distortionType;
    float k1, k2, k3, p1, p2, b1, b2, b3, b4;
    float lensRadius, focusDist;
    float3 camT; float4 camR; float3 camS;      /* camera pose at t0 (RayHelper::cameraAtT0) */
    const wpt_keyframe* camKeys;
    uint32_t nCamKeys;
    float2 distCenter, distFocal, distInvFocal; /* LensDistortion::Helper */
    /* sensor */
    int32_t sensorType;
    uint32_t width, height;
    float minDistToLight, maxDistToLight, minPathLen, maxPathLen;
    float tofPixelArea, tofExposureTime, tofContrast, tofTau, tofFracModfreqC;
    /* Parameters */
    uint32_t maxPathComponents;
    float rrThreshold;
    int32_t randomize;
    float minHit;
    /* sampling */
    uint32_t samplesSqrt, samples;
    float invSamples, invSamplesSqrt, invWidth, invHeight;
    float t0, t1;
    /* tiling */
    uint32_t substreams, subShift, tileShift;   /* work items per tile = 1 << tileShift */
    uint32_t tileW, tileH, tilesX, tilesTotal;
    uint32_t tileFirst, tileStride, nLocalTiles;
    uint32_t nItems;                /* nLocalTiles << tileShift work items (pixel, sub-stream) */
    /* output */
    float* frame;
    float* partials;                /* [pixel][sub-stream][3] partial sums when substreams > 1 */
    unsigned long long* counters;   /* CTR_WORDS words, see CTR_* */
    unsigned long long* itemCounter;    /* the work item ticket: counters + CTR_ITEM, or - when several GPUs render one
                                           frame - ONE counter in the first GPU's memory that all of them draw their
                                           blocks of work items from over NVLink (the MPICoordinator block hand-out,
                                           mpi.hpp:74-107, without the messages) */
    uint8_t* blockDone;             /* several GPUs: [item / WAVE_WINDOW] = this GPU rendered that block of items */
    WaveState W;
    uint2* poolSpill;               /* pooled trace kernel: stack entries beyond POOL_STACK, per pool entry (2 groups) */
    uint32_t poolSpillPerRay;
    int32_t traversal;
};

/* getGroundTruth() launch parameters (wurblpt.hpp:643-663) */
struct GroundTruthParams {
    float4 invCamR[3];          /* inverse(camera.at(t)).rotation for t0, tPrev, tNext */
    float3 invCamT[3];          /* ... .translation */
    float tPrev, tNext;
    float P00, P11, P03, P13;   /* Camera::ImageSpaceHelper (camera.hpp:194-203) */
    uint32_t bits;
    float* aov[WPT_GT_FLOAT_IMAGES];
    int32_t* materials;
    const int32_t* materialIds; /* description material index -> Scene::materialIndex(), or nullptr */
    uint32_t nMaterialIds;
};

/* counters for the ray-query kernels */
struct TraceStats {
    unsigned long long nodes, leaves, nested;
};

}
