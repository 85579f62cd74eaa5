/* bvh_device.hpp -- the device-side Scene::updateBVH() (bvh_device.cu) */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wptd {

/* Device pointers. Inputs are in hitable (global id) order, outputs in slot order. */
struct BuildIO {
    uint32_t n;                     /* hitables: triangles and spheres of the top level (>= 2) */
    const float4* box;              /* 2 per hitable: lo xyz | type (PRIM_*), hi xyz | index among the hitables of its type */
    const float4* triPrim;          /* 3 per triangle: world-space corners | instance, triangle number, animation + 1 */
    const uint4* triShadePrim;      /* per triangle: vertex attribute offsets, mesh flags */
    const uint8_t* triClassPrim;    /* per triangle: material class */
    const float4* sphPrim;          /* 3 per sphere */
    const uint8_t* sphClassPrim;
    const uint32_t* hotPrims;       /* global ids of the hot spots */
    uint32_t nHot;
    float4* node4;                  /* out: 8 float4 per wide node, root at 0 (capacity: n nodes) */
    float4* tri; uint4* triShade; uint8_t* triClass; uint32_t* triGlobal;
    float4* sph; uint8_t* sphClass; uint32_t* sphGlobal;
    uint32_t* hotRefs;              /* out: leaf references of the hot spots */
};

size_t deviceBuildScratchBytes(uint32_t n);

/* n >= 2. Synchronises `stream` (the host reads the size of every tree level). */
cudaError_t launchDeviceBuild(const BuildIO& io, void* scratch, uint32_t maxRun, cudaStream_t stream,
        uint32_t* wideNodes, uint32_t* wideDepth, uint32_t* triangles);

}
