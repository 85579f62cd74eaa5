/* kernels.hpp -- host-callable launchers of kernels.cu */
#pragma once

#include <cuda_runtime.h>

#include "device_scene.cuh"

namespace wptd {

/* features: 0 = static scene, RGB sensor with open gates; anything else selects the full variant
 * (media, animations, optical path length tracking). */
cudaError_t launchRender(int features, const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream, int* launches);

/* Wavefront pipeline; P.W must point at allocated slot arrays. stream2: an internal second stream;
 * hostDone: 2 pinned counters; events: 4 events. */
cudaError_t launchWavefront(int features, const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream,
        cudaStream_t stream2, unsigned long long* hostDone, cudaEvent_t* events, int* launches);

/* origins == nullptr: pixel-centre camera rays of P (width*height rays). */
cudaError_t launchQuery(int features, bool count, const DScene& S, const RenderParams& P,
        const float* origins, const float* directions, unsigned long long nRays, float time,
        int32_t* ids, float* dist, TraceStats* stats, cudaStream_t stream);

/* The same queries through a trace kernel of the wavefront pipeline - pooled, or the persistent per-lane one (scenes without media /
 * animated hitables):
 * P.W.rayO / rayD / traceQueue / hit must hold nRays slots, P.poolSpill the spill area, P.counters the counters. */
cudaError_t launchPoolQuery(const DScene& S, const RenderParams& P, int numSMs, int mode /* 0 persistent per-lane, 1 pooled, 2 quad */, const float* origins, const float* directions,
        unsigned long long nRays, float time, int32_t* ids, float* dist, cudaStream_t stream);
unsigned int poolGridBlocks(int numSMs);
/* node4c from node4 (see DScene): nNodes wide nodes, child by child */
cudaError_t launchNodeTranspose(const float4* node4, float4* node4c, uint32_t nNodes, cudaStream_t stream);
unsigned int poolRaysPerBlock();
unsigned int poolStackEntries();
unsigned int wideStackEntries();     /* capacity of the per-lane walk's stack (shared memory + local spill) */

/* getGroundTruth(): the images selected by G.bits for P.width x P.height pixels (device pointers in G). */
cudaError_t launchGroundTruth(int features, const DScene& S, const RenderParams& P, const GroundTruthParams& G, cudaStream_t stream);

/* frame[i] += other[i] (both on the current device; combining the frames of several GPUs) */
cudaError_t launchAddFrame(float* frame, const float* other, size_t nFloats, cudaStream_t stream);

cudaError_t launchPrng(const uint32_t* seeds, uint32_t nSeeds, uint32_t nDraws, float* out, cudaStream_t stream);

}
