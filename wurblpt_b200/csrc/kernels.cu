/*
 * kernels.cu -- the sm_100a kernels of the path-tracing core and their launchers.
 *
 *   renderKernel<F>     mcpt() + tracePath() (reference wurblpt.hpp:108-436) as a persistent
 *                       megakernel: one CTA set per SM, warps pull pixel tiles from an atomic
 *                       ticket counter, every lane owns one (pixel, sub-stream) and runs a small
 *                       state machine with ONE tree-walk site, so lanes tracing camera rays,
 *                       bounce rays and shadow rays walk the tree together.
 *   queryKernel<F>      closest hit for explicit rays or pixel-centre camera rays (the loop of
 *                       getGroundTruth(), wurblpt.hpp:659-668); used for bit-exact parity checks.
 *   prngKernel          Prng known answers.
 *
 * Compiled with --fmad=false (no contraction) so FP32 results round like the strict oracle.
 */
#include <cstdio>

#include "kernels.hpp"
#include "device_shade.cuh"

namespace wptd {

constexpr int BLOCK = 128;

enum { ST_NEW = 0, ST_MAIN = 1, ST_SHADOW_HOT = 2, ST_SHADOW_ENV = 3 };

template<int F>
WPT_DEV Hit closestHit(const DScene& S, const RenderParams& P, const Ray& ray, const RayHelper& h,
        Ctx<F>& ctx, uint2* smStack)
{
    if (P.traversal == WPT_TRAVERSAL_REFERENCE)
        return traverseReference<F, true>(S, 0, ray, h, P.minHit, MAXVAL, P.minHit, ctx);
    return traverseWide<F>(S, ray, h, P.minHit, MAXVAL, P.minHit, ctx, smStack, BLOCK);
}

WPT_DEV bool isTofLight(const wpt_material* m)
{
    return m != nullptr && m->type == WPT_MAT_LIGHT_TOF;
}

template<int F>
__global__ void __launch_bounds__(BLOCK) renderKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    extern __shared__ uint2 smStackAll[];
    uint2* smStack = smStackAll + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const float4 one4 = f4(1.0f);

    unsigned long long nSamples = 0, nRays = 0, nNested = 0;

    for (;;) {
        /* ---- fetch a tile ticket (MPICoordinator::getBlock, mpi.hpp:241-254, as a device atomic) ---- */
        unsigned int ticket = 0;
        if (lane == 0)
            ticket = (unsigned int)atomicAdd(P.counters + 0, 1ull);
        ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
        if (ticket >= P.nLocalTiles)
            break;
        const unsigned int tile = P.tileFirst + ticket * P.tileStride;
        const unsigned int tileX = tile % P.tilesX, tileY = tile / P.tilesX;
        const unsigned int sub = lane & (P.substreams - 1u);
        const unsigned int pit = lane >> P.subShift;
        const unsigned int x = tileX * P.tileW + pit % P.tileW;
        const unsigned int y = tileY * P.tileH + pit / P.tileW;
        const bool valid = (x < P.width && y < P.height);
        const unsigned int pixel = y * P.width + x;

        Ctx<F> ctx;
        ctx.prng = prngInit(pixel);                     /* Prng prng(pixel) (wurblpt.hpp:342) */
        for (unsigned int k = 0; k < sub; k++)
            prngJump(ctx.prng);
        ctx.nested = 0;
        ctx.nodes = ctx.leaves = 0;

        float acc[3] = { 0.0f, 0.0f, 0.0f };
        unsigned int sampleIndex = sub;
        int state = valid ? ST_NEW : -1;

        /* path state */
        Ray ray;
        float4 att = one4, opl = f4(0.0f), rayRi = one4;
        float pathLength = 0.0f;
        unsigned int pathComponent = 0;
        /* pending state across a shadow-ray walk */
        float4 nextAtt = one4, nextRi = one4, pendRad = f4(0.0f), shadowRi = one4;
        float3 nextDir = f3(0.0f);
        uint32_t pendHot = NO_HIT;
        ray.o = f3(0.0f); ray.d = f3(0.0f, 0.0f, 1.0f); ray.time = 0.0f;
        Ray walkRay = ray;

        while (state >= 0) {
            if (state == ST_NEW) {
                if (sampleIndex >= P.samples)
                    break;
                /* stratified jitter (wurblpt.hpp:349-359); first draw goes to the row offset, see in01x2 */
                float u = float(x), v = float(y);
                if (P.randomize) {
                    const unsigned int j = sampleIndex / P.samplesSqrt;
                    const unsigned int i = sampleIndex % P.samplesSqrt;
                    const float rj = in01(ctx.prng);
                    const float ri = in01(ctx.prng);
                    u += (float(i) + ri) * P.invSamplesSqrt;
                    v += (float(j) + rj) * P.invSamplesSqrt;
                } else {
                    u += 0.5f;
                    v += 0.5f;
                }
                u *= P.invWidth;
                v *= P.invHeight;
                ray = cameraRay(P, u, v, ctx.prng, true);
                animationCacheInit<F>(S, ctx, ray.time);
                att = one4;
                pathLength = 0.0f;
                opl = f4(0.0f);
                rayRi = one4;       /* refractiveIndexOfVacuum (camera.hpp:184) */
                pathComponent = 0;
                sampleIndex += P.substreams;
                nSamples++;
                walkRay = ray;
                state = ST_MAIN;
            }

            /* ---- the one tree-walk site ---- */
            const RayHelper h = makeRayHelper(walkRay);
            const Hit hit = closestHit<F>(S, P, walkRay, h, ctx, smStack);
            nRays++;

            bool advance = false;
            if (state == ST_MAIN) {
                if (hit.ref == NO_HIT) {
                    /* no hit: environment map, end of path (wurblpt.hpp:136-146) */
                    if (S.envType != WPT_ENV_NONE)
                        accumulate<F>(P, acc, f4(MAXVAL), MAXVAL, att * envRadiance(S, ray.d), false);
                    state = ST_NEW;
                    continue;
                }
                pathLength += hit.a;
                if (F & F_OPL)
                    opl = opl + hit.a * rayRi;
                if (!(pathComponent + 1 < P.maxPathComponents)) {
                    state = ST_NEW;
                    continue;
                }
                HitRec hr = buildHitRecord<F>(S, hit, ray, h, ctx);
                const wpt_material* mat = resolveMaterial(S, hr.material, hr);
                const ScatterRec sr = materialScatter(S, mat, ray, rayRi, hr, ctx.prng);
                accumulate<F>(P, acc, opl, (pathComponent == 0 ? 0.0f : hit.a),
                        att * materialEmitted(S, mat, ray, hr), isTofLight(mat));
                if (sr.type == SCATTER_NONE) {
                    state = ST_NEW;
                    continue;
                }
                nextAtt = att * sr.att;
                if (sr.type == SCATTER_RANDOM) {
                    if (sr.pdf > 0.0f)
                        nextAtt = nextAtt / sr.pdf;
                    else
                        nextAtt = f4(0.0f);
                }
                nextDir = sr.dir;
                if (F & F_OPL)
                    nextRi = sr.ri;
                advance = true;
                if (sr.type == SCATTER_RANDOM && S.nHot > 0) {
                    /* next event estimation towards a hot spot with MIS (wurblpt.hpp:179-220) */
                    const float hotPdf = hotSpotsPdf<F>(S, hr.p, sr.dir, ray.time, ctx);
                    nextAtt = nextAtt * powerHeuristic(sr.pdf, hotPdf);
                    unsigned int hotIndex = (unsigned int)(in01(ctx.prng) * float(S.nHot));
                    hotIndex = min(hotIndex, S.nHot - 1u);
                    const uint32_t hotRef = __ldg(S.hotRefs + hotIndex);
                    const float3 directDir = hotSpotDirection<F>(S, hotRef, hr.p, ctx);
                    const float directPdf = hotSpotsPdf<F>(S, hr.p, directDir, ray.time, ctx);
                    if (directPdf > 0.0f) {
                        const ScatterRec dsr = materialScatterToDirection(S, mat, ray, rayRi, hr, directDir);
                        if (dsr.pdf > 0.0f) {
                            const float weight = powerHeuristic(directPdf, dsr.pdf);
                            pendRad = ((att * dsr.att) / directPdf) * weight;
                            pendHot = hotRef;
                            if (F & F_OPL)
                                shadowRi = dsr.ri;
                            walkRay.o = hr.p;
                            walkRay.d = directDir;
                            walkRay.time = ray.time;
                            state = ST_SHADOW_HOT;
                            advance = false;
                        }
                    }
                } else if (sr.type == SCATTER_RANDOM && S.envType != WPT_ENV_NONE && S.envN > 0) {
                    /* next event estimation towards the environment map (wurblpt.hpp:221-252) */
                    const float lightsP = envPdf(S, sr.dir);
                    nextAtt = nextAtt * powerHeuristic(sr.pdf, lightsP);
                    const float3 lightDir = envSample(S, ctx.prng);
                    const float lightDirP = envPdf(S, lightDir);
                    const ScatterRec lsr = materialScatterToDirection(S, mat, ray, rayRi, hr, lightDir);
                    if (lsr.pdf > 0.0f) {
                        const float weight = powerHeuristic(lightDirP, lsr.pdf);
                        pendRad = (((att * lsr.att) / lightDirP) * weight) * envRadiance(S, lightDir);
                        walkRay.o = hr.p;
                        walkRay.d = lightDir;
                        walkRay.time = ray.time;
                        state = ST_SHADOW_ENV;
                        advance = false;
                    }
                }
                /* the next path segment starts at the hit point (wurblpt.hpp:256) */
                ray.o = hr.p;
            } else if (state == ST_SHADOW_HOT) {
                /* the chosen hot spot must be the closest hit (wurblpt.hpp:208) */
                if (hit.ref == pendHot) {
                    HitRec dhr = buildHitRecord<F>(S, hit, walkRay, h, ctx);
                    const wpt_material* lightMat = resolveMaterial(S, dhr.material, dhr);
                    const float4 rad = pendRad * materialEmitted(S, lightMat, walkRay, dhr);
                    float4 oplAcc = opl;
                    if (F & F_OPL)
                        oplAcc = opl + hit.a * shadowRi;
                    accumulate<F>(P, acc, oplAcc, hit.a, rad, isTofLight(lightMat));
                }
                advance = true;
            } else {
                /* the environment must be visible (wurblpt.hpp:240) */
                if (hit.ref == NO_HIT)
                    accumulate<F>(P, acc, f4(MAXVAL), MAXVAL, pendRad, false);
                advance = true;
            }

            if (advance) {
                /* update attenuation and ray, then Russian roulette (wurblpt.hpp:255-273) */
                att = nextAtt;
                ray.d = nextDir;
                if (F & F_OPL)
                    rayRi = nextRi;
                walkRay = ray;
                state = ST_MAIN;
                const float maxAtt = vmaxcomp(att);
                if (maxAtt < P.rrThreshold && pathComponent >= 5) {
                    const float q = rclamp(1.0f - maxAtt, 0.0f, 0.95f);
                    if (in01(ctx.prng) < q) {
                        state = ST_NEW;
                    } else {
                        const float rrWeight = 1.0f / (1.0f - q);
                        att = att * rrWeight;
                    }
                }
                pathComponent++;
            }
        }
        nNested += ctx.nested;

        /* ---- combine the sub-streams of a pixel (fixed order => deterministic) and store ---- */
        for (unsigned int off = P.substreams >> 1; off > 0; off >>= 1) {
            acc[0] += __shfl_xor_sync(0xFFFFFFFFu, acc[0], off);
            acc[1] += __shfl_xor_sync(0xFFFFFFFFu, acc[1], off);
            acc[2] += __shfl_xor_sync(0xFFFFFFFFu, acc[2], off);
        }
        if (valid && sub == 0) {
            /* Sensor::finishPixel (sensor_rgb.hpp:82-87) */
            float* out = P.frame + size_t(pixel) * 3;
            out[0] = P.invSamples * acc[0];
            out[1] = P.invSamples * acc[1];
            out[2] = P.invSamples * acc[2];
        }
    }

    /* statistics: one atomic per warp and counter */
    for (int off = 16; off > 0; off >>= 1) {
        nSamples += __shfl_xor_sync(0xFFFFFFFFu, nSamples, off);
        nRays += __shfl_xor_sync(0xFFFFFFFFu, nRays, off);
        nNested += __shfl_xor_sync(0xFFFFFFFFu, nNested, off);
    }
    if (lane == 0) {
        atomicAdd(P.counters + 1, nSamples);
        atomicAdd(P.counters + 2, nRays);
        atomicAdd(P.counters + 3, nNested);
    }
}

/* ------------------------------------------------------------ query kernels */

template<int F>
__global__ void __launch_bounds__(BLOCK) queryKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P,
        const float* origins, const float* directions, unsigned long long nRays, float time,
        int32_t* ids, float* dist, TraceStats* stats)
{
    extern __shared__ uint2 smStackAll[];
    uint2* smStack = smStackAll + threadIdx.x;
    const unsigned long long i = blockIdx.x * (unsigned long long)BLOCK + threadIdx.x;
    Ctx<F> ctx;
    ctx.nested = 0;
    ctx.nodes = ctx.leaves = 0;
    if (i < nRays) {
        ctx.prng = prngInit((unsigned int)i);
        Ray ray;
        if (origins) {
            ray.o = f3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
            ray.d = f3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
            ray.time = time;
        } else {
            /* pixel-centre camera ray, randomness off (wurblpt.hpp:663-665) */
            const unsigned int x = (unsigned int)(i % P.width), y = (unsigned int)(i / P.width);
            ray = cameraRay(P, (float(x) + 0.5f) / float(P.width), (float(y) + 0.5f) / float(P.height), ctx.prng, false);
        }
        animationCacheInit<F>(S, ctx, ray.time);
        const RayHelper h = makeRayHelper(ray);
        const Hit hit = closestHit<F>(S, P, ray, h, ctx, smStack);
        int32_t id = -1;
        if (hit.ref != NO_HIT) {
            const uint32_t type = hit.ref >> REF_TYPE_SHIFT, slot = hit.ref & REF_SLOT_MASK;
            id = int32_t(type == PRIM_TRIANGLE ? S.triGlobal[slot] : type == PRIM_SPHERE ? S.sphGlobal[slot] : S.medGlobal[slot]);
        }
        ids[i] = id;
        dist[i] = (hit.ref != NO_HIT ? hit.a : 0.0f);
    }
    if (stats) {
        unsigned long long n = ctx.nodes, l = ctx.leaves, q = ctx.nested;
        for (int off = 16; off > 0; off >>= 1) {
            n += __shfl_xor_sync(0xFFFFFFFFu, n, off);
            l += __shfl_xor_sync(0xFFFFFFFFu, l, off);
            q += __shfl_xor_sync(0xFFFFFFFFu, q, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&stats->nodes, n);
            atomicAdd(&stats->leaves, l);
            atomicAdd(&stats->nested, q);
        }
    }
}

__global__ void prngKernel(const uint32_t* seeds, uint32_t nSeeds, uint32_t nDraws, float* out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSeeds)
        return;
    Prng p = prngInit(seeds[i]);
    for (uint32_t k = 0; k < nDraws; k++)
        out[size_t(i) * nDraws + k] = in01(p);
}

/* ---------------------------------------------------------------- launchers */

static size_t stackBytes()
{
    return size_t(SM_STACK) * BLOCK * sizeof(uint2);
}

template<int F>
static cudaError_t launchRenderT(const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream, int* gridOut)
{
    int perSM = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, renderKernel<F>, BLOCK, stackBytes());
    if (e != cudaSuccess)
        return e;
    if (perSM < 1)
        perSM = 1;
    /* persistent: exactly as many CTAs as fit, but not more warps than there are tiles */
    long long grid = (long long)numSMs * perSM;
    long long needed = ((long long)P.nLocalTiles + (BLOCK / 32) - 1) / (BLOCK / 32);
    if (needed < grid)
        grid = needed < 1 ? 1 : needed;
    *gridOut = int(grid);
    renderKernel<F><<<int(grid), BLOCK, stackBytes(), stream>>>(S, P);
    return cudaGetLastError();
}

cudaError_t launchRender(int features, const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream, int* grid)
{
    if (features == 0)
        return launchRenderT<0>(S, P, numSMs, stream, grid);
    return launchRenderT<F_MEDIA | F_ANIM | F_OPL>(S, P, numSMs, stream, grid);
}

cudaError_t launchQuery(int features, bool count, const DScene& S, const RenderParams& P,
        const float* origins, const float* directions, unsigned long long nRays, float time,
        int32_t* ids, float* dist, TraceStats* stats, cudaStream_t stream)
{
    const unsigned int grid = (unsigned int)((nRays + BLOCK - 1) / BLOCK);
    if (grid == 0)
        return cudaSuccess;
    const bool full = (features != 0);
    if (count) {
        if (full)
            queryKernel<F_MEDIA | F_ANIM | F_COUNT><<<grid, BLOCK, stackBytes(), stream>>>(S, P, origins, directions, nRays, time, ids, dist, stats);
        else
            queryKernel<F_COUNT><<<grid, BLOCK, stackBytes(), stream>>>(S, P, origins, directions, nRays, time, ids, dist, stats);
    } else {
        if (full)
            queryKernel<F_MEDIA | F_ANIM><<<grid, BLOCK, stackBytes(), stream>>>(S, P, origins, directions, nRays, time, ids, dist, nullptr);
        else
            queryKernel<0><<<grid, BLOCK, stackBytes(), stream>>>(S, P, origins, directions, nRays, time, ids, dist, nullptr);
    }
    return cudaGetLastError();
}

cudaError_t launchPrng(const uint32_t* seeds, uint32_t nSeeds, uint32_t nDraws, float* out, cudaStream_t stream)
{
    if (nSeeds == 0)
        return cudaSuccess;
    prngKernel<<<(nSeeds + 127) / 128, 128, 0, stream>>>(seeds, nSeeds, nDraws, out);
    return cudaGetLastError();
}

}
