/*
 * kernels.cu -- the sm_100a kernels of the path-tracing core and their launchers.
 *
 * Two pipelines run the same per-path state machine (device_path.cuh):
 *
 *   wavefront    (default) path state lives in HBM, one slot per in-flight work item (WaveState). A round is
 *                  waveTracePersistentKernel   closest hit for every slot listed in the trace queue
 *                  waveSortKernel              surface hits that are to be shaded -> material queue (slot order)
 *                  waveMaterialKernel          hit record, emission, scatter, next-event estimation set-up
 *                  waveGenerateKernel          misses, shadow-ray results, Russian roulette, path end; finishPixel,
 *                                              next work item, next camera ray; builds the trace queue
 *                The slots are split into two groups whose rounds run on two streams. The host enqueues rounds in
 *                batches and reads a "slots retired" counter one batch late, so the GPU never waits for the host.
 *   megakernel   renderKernel<F>: every lane keeps one work item's state in registers and loops
 *                startSample -> closest hit -> processHit with ONE tree-walk site; kept for A/B comparisons.
 *
 *   groundTruthKernel<F>  getGroundTruth() (wurblpt.hpp:643-759): one pixel-centre ray per pixel, 20 images.
 *   queryKernel<F>        closest hit for explicit rays or pixel-centre camera rays; bit-exact parity checks.
 *   resolveKernel         fixed-order sum of sub-stream partial sums.
 *   prngKernel            Prng known answers.
 *
 * Compiled with --fmad=false (no contraction) so FP32 results round like the strict oracle.
 */
#include <cstdio>
#include <cstdlib>

#include "kernels.hpp"
#include "device_path.cuh"

namespace wptd {

#ifndef WPT_BLOCK
#define WPT_BLOCK 128
#endif
#ifndef WPT_MIN_BLOCKS
#define WPT_MIN_BLOCKS 8
#endif
#ifndef WPT_TRACE_MIN_BLOCKS
#define WPT_TRACE_MIN_BLOCKS 8
#endif
#ifndef WPT_TRACE_MIN_BLOCKS_STATIC
#define WPT_TRACE_MIN_BLOCKS_STATIC 9   /* persistent trace kernel of static scenes without media (F == 0): 56 registers, +1.8 % on
                                           C3 over 8 blocks (63 registers), profiles/r2x_quad_kernel_ab.txt */
#endif
#ifndef WPT_SHADE_MIN_BLOCKS
#define WPT_SHADE_MIN_BLOCKS 4
#endif
#ifndef WPT_LOGIC_MIN_BLOCKS
#define WPT_LOGIC_MIN_BLOCKS 6
#endif
#ifndef WPT_MATERIAL_MIN_BLOCKS
#define WPT_MATERIAL_MIN_BLOCKS 6   /* 85 registers: +1.5 % on C3, +5 % on C1 over 4 blocks (128 registers), profiles/r2f_ab.txt */
#endif
#ifndef WPT_PREFETCH_LEAF
#define WPT_PREFETCH_LEAF 1
#endif
#ifndef WPT_PREFETCH_NODE
#define WPT_PREFETCH_NODE 0     /* 1: the persistent trace kernel asks for the next inner node (one 128-byte line) the moment it is
                                   chosen or popped, a loop-top's worth of instructions before its rows are loaded */
#endif
#ifndef WPT_CAMERA_REFILL_MIN
#define WPT_CAMERA_REFILL_MIN 32
#endif
#ifndef WPT_REFILL_MIN
#define WPT_REFILL_MIN 12   /* idle lanes that must have gathered before the persistent trace kernel fetches new rays */
#endif
#ifndef WPT_LEAF_UNIT
#define WPT_LEAF_UNIT 0     /* 1: a lane tests one primitive per leaf round instead of all primitives of its leaf (up to four); leaf lanes 7 -> 12 of 32 but node lanes 21 -> 18: -3 % on C3 */
#endif
#ifndef WPT_LEAF_COOP
#define WPT_LEAF_COOP 0     /* 1: leaf rounds of triangle leaves pool the triangles of all waiting lanes and test one per lane
                               (static scenes; see waveTracePersistentKernel). Bit-identical answers; the shuffles, the owner
                               table and the in-order combine cost what the fuller lanes save: C3 +0.7 %, C1 -3 %, C2 -3 %
                               (profiles/r2q_coop_ab.txt) */
#endif
#ifndef WPT_ANY_HIT
#define WPT_ANY_HIT 1       /* shadow queries towards the environment stop at the first hit (persistent trace kernel) */
#endif
#ifndef WPT_NODE_REPEAT
#define WPT_NODE_REPEAT 1       /* node steps per node round of the persistent trace kernel (1 = one step per loop iteration;
                                   2: +0.6 % on C3, -0.7 % on C2: noise level, profiles/r2y_trace_micro_ab.txt) */
#endif
#ifndef WPT_NODE_REPEAT_MIN
#define WPT_NODE_REPEAT_MIN 16  /* lanes that must still be at an inner node for another step of the same round */
#endif
#ifndef WPT_LEAF_MIN
#define WPT_LEAF_MIN 12     /* lanes that must be at a leaf before the persistent trace kernel runs a leaf round */
#endif
constexpr int BLOCK = WPT_BLOCK;

template<int F>
WPT_DEV Hit closestHit(const DScene& S, const RenderParams& P, const Ray& ray, const RayHelper& h,
        Ctx<F>& ctx, uint2* smStack)
{
    if (P.traversal == WPT_TRAVERSAL_REFERENCE)
        return traverseReference<F, true>(S, 0, ray, h, P.minHit, MAXVAL, P.minHit, ctx);
    return traverseWide<F>(S, ray, h, P.minHit, MAXVAL, P.minHit, ctx, smStack, BLOCK);
}

/* ------------------------------------------------------------------ megakernel */

template<int F>
__global__ void __launch_bounds__(BLOCK, WPT_MIN_BLOCKS) renderKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    extern __shared__ uint2 smStackAll[];
    uint2* smStack = smStackAll + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneBelow = (1u << lane) - 1u;
    constexpr unsigned FULL = 0xFFFFFFFFu;

    unsigned long long nSamples = 0, nRays = 0;
    Ctx<F> ctx;
    ctx.prng = prngInit(0);
    ctx.nested = 0;
    ctx.nodes = ctx.leaves = 0;
    PathState ps;
    ps.state = ST_IDLE;
    ps.pixel = ps.sub = ps.sampleIndex = 0;
    ps.acc[0] = ps.acc[1] = ps.acc[2] = 0.0f;
    ps.walkRay.o = f3(0.0f); ps.walkRay.d = f3(0.0f, 0.0f, 1.0f); ps.walkRay.time = 0.0f;
    ps.att = ps.opl = ps.rayRi = ps.nextAtt = ps.pendRad = ps.nextRi = ps.shadowRi = f4(0.0f);
    ps.nextDir = f3(0.0f);
    ps.pathLength = 0.0f;
    ps.pathComponent = 0;
    ps.pendHot = NO_HIT;
    bool exhausted = false;

    for (;;) {
        /* ---- hand new items to idle lanes (MPICoordinator::getBlock, mpi.hpp:241-254, as a device
         * atomic; the whole warp passes here once per walk) ---- */
        const unsigned wantMask = __ballot_sync(FULL, ps.state == ST_IDLE && !exhausted);
        if (wantMask != 0u) {
            const int leader = __ffs(wantMask) - 1;
            unsigned int base = 0;
            if (int(lane) == leader)
                base = (unsigned int)atomicAdd(P.itemCounter, (unsigned long long)__popc(wantMask));
            base = __shfl_sync(FULL, base, leader);
            if (ps.state == ST_IDLE && !exhausted) {
                const unsigned int item = base + __popc(wantMask & laneBelow);
                if (item >= P.nItems) {
                    exhausted = true;
                } else {
                    const WorkItem w = decodeItem(P, item);
                    if (w.valid)
                        beginItem<F>(ps, ctx, w);
                }
            }
        }
        if (__all_sync(FULL, exhausted))
            break;
        if (ps.state == ST_IDLE)
            continue;
        if (ps.state == ST_NEW) {
            if (!startSample<F>(S, P, ps, ctx)) {
                finishItem(P, ps);
                ps.state = ST_IDLE;
                continue;
            }
            animationCacheInit<F>(S, ctx, ps.walkRay.time);
            nSamples++;
        }
        /* ---- the one tree-walk site ---- */
        const RayHelper h = makeRayHelper(ps.walkRay);
        const Hit hit = closestHit<F>(S, P, ps.walkRay, h, ctx, smStack);
        nRays++;
        processHit<F>(S, P, ps, ctx, hit, h);
    }

    /* statistics: one atomic per warp and counter */
    unsigned long long nNested = ctx.nested;
    for (int off = 16; off > 0; off >>= 1) {
        nSamples += __shfl_xor_sync(FULL, nSamples, off);
        nRays += __shfl_xor_sync(FULL, nRays, off);
        nNested += __shfl_xor_sync(FULL, nNested, off);
    }
    if (lane == 0) {
        atomicAdd(P.counters + CTR_SAMPLES, nSamples);
        atomicAdd(P.counters + CTR_RAYS, nRays);
        atomicAdd(P.counters + CTR_NESTED, nNested);
    }
}

/* Adds the sub-stream partial sums of every pixel of this rank's tiles in a fixed order
 * (deterministic result) and applies Sensor::finishPixel's 1/samples. */
__global__ void resolveKernel(const __grid_constant__ RenderParams P)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;       /* (local tile, pixel in tile) */
    const unsigned int pixelsPerTile = (1u << P.tileShift) >> P.subShift;
    if (i >= P.nLocalTiles * pixelsPerTile)
        return;
    const unsigned int localTile = i / pixelsPerTile, pit = i % pixelsPerTile;
    const unsigned int item = (localTile << P.tileShift) | (pit << P.subShift);
    if (P.blockDone && !P.blockDone[item / WAVE_WINDOW])
        return;         /* another GPU rendered this pixel */
    const WorkItem w = decodeItem(P, item);
    if (!w.valid)
        return;
    const float* part = P.partials + size_t(w.pixel) * P.substreams * 3;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    for (unsigned int s = 0; s < P.substreams; s++) {
        r += part[s * 3 + 0];
        g += part[s * 3 + 1];
        b += part[s * 3 + 2];
    }
    float* out = P.frame + size_t(w.pixel) * 3;
    out[0] = P.invSamples * r;
    out[1] = P.invSamples * g;
    out[2] = P.invSamples * b;
}

/* Several GPUs: adds another GPU's frame (zero wherever that GPU rendered nothing) to this one's. Every pixel is
 * rendered by exactly one GPU, so the sum is an exact gather (x + 0 = x). */
__global__ void addFrameKernel(float4* frame, const float4* other, size_t n4, float* tail, const float* otherTail, unsigned int nTail)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n4) {
        const float4 a = frame[i], b = other[i];
        frame[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    if (i < nTail)
        tail[i] += otherTail[i];
}

cudaError_t launchAddFrame(float* frame, const float* other, size_t nFloats, cudaStream_t stream)
{
    const size_t n4 = nFloats / 4;
    const unsigned int nTail = (unsigned int)(nFloats % 4);
    if (nFloats == 0)
        return cudaSuccess;
    const size_t threads = n4 > nTail ? n4 : nTail;
    addFrameKernel<<<(unsigned int)((threads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<float4*>(frame),
            reinterpret_cast<const float4*>(other), n4, frame + n4 * 4, other + n4 * 4, nTail);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------- wavefront */

/* Slot control word: state in bits 0-2 (7 = idle), kind of the pending shadow query in bits 3-4, path component from bit 8. */
WPT_DEV unsigned int makeCtl(int state, int shadow, unsigned int pathComponent)
{
    return (unsigned int)(state & 7) | ((unsigned int)shadow << 3) | (pathComponent << 8);
}
WPT_DEV int ctlState(unsigned int ctl) { return (ctl & 7u) == 7u ? ST_IDLE : int(ctl & 7u); }
WPT_DEV int ctlShadow(unsigned int ctl) { return int((ctl >> 3) & 3u); }
WPT_DEV unsigned int ctlPathComponent(unsigned int ctl) { return ctl >> 8; }

/* One byte per slot for the kernels that scan whole windows (sort, generate): state in bits 0-2 (7 = idle),
 * bit 3 = the pending query B is on the last allowed path component, bit 4 = a shadow query S is pending,
 * bit 5 = "fresh": the queries were issued by the material kernel of this round and are not answered yet,
 * bits 6-7 = min(path component, 3). */
constexpr unsigned int CODE_IDLE = 0xFFu, CODE_LAST = 8u, CODE_SHADOW = 16u, CODE_FRESH = 32u;
WPT_DEV unsigned int slotCode(const RenderParams& P, int state, int shadow, unsigned int pathComponent, bool fresh)
{
    if (state == ST_IDLE)
        return CODE_IDLE;
    const unsigned int last = (pathComponent + 1u < P.maxPathComponents) ? 0u : CODE_LAST;
    return (unsigned int)state | last | (shadow != SHADOW_NONE ? CODE_SHADOW : 0u) | (fresh ? CODE_FRESH : 0u)
        | (min(pathComponent, 3u) << 6);
}
WPT_DEV int codeState(unsigned int code) { return code == CODE_IDLE ? ST_IDLE : int(code & 7u); }

/* ---- slot state in HBM <-> PathState in registers ---- */

/* The work item of a slot whose path has ended (or that is idle): what the generate kernel needs to start the next
 * sample or to finish the item. */
template<int F>
WPT_DEV void loadItem(const RenderParams& P, unsigned int slot, bool idle, PathState& ps, Ctx<F>& ctx)
{
    ps.state = idle ? ST_IDLE : ST_NEW;
    ps.shadow = SHADOW_NONE;
    ps.pathComponent = 0;
    if (idle) {
        ps.pixel = ps.sampleIndex = ps.sub = 0;
        return;
    }
    const int4 item = P.W.item[slot];
    ps.pixel = (unsigned int)item.x;
    ps.sampleIndex = (unsigned int)item.y;
    ps.sub = (unsigned int)item.z;
    const uint4 r = P.W.prng[2 * slot];
    ctx.prng.s0 = r.x; ctx.prng.s1 = r.y; ctx.prng.s2 = r.z; ctx.prng.s3 = r.w;
    const float4 acc = P.W.acc[2 * slot];
    ps.acc[0] = acc.x; ps.acc[1] = acc.y; ps.acc[2] = acc.z;
}

/* Everything a shade kernel needs of a slot whose queries have been answered, as independent loads (none of them
 * waits for the slot's control word). ITEM: the work item too (the path may end here). */
template<int F, bool ITEM>
WPT_DEV void loadAnswered(const RenderParams& P, unsigned int slot, PathState& ps, Ctx<F>& ctx, Hit& hitB, Hit& hitS)
{
    const unsigned int ctl = P.W.ctl[slot];
    const uint4 r = P.W.prng[2 * slot];
    const float4 acc = P.W.acc[2 * slot];
    const float4 o = P.W.rayO[2 * slot], d = P.W.rayD[2 * slot];
    const float4 sd = P.W.shD[2 * slot], pr = P.W.pendRad[2 * slot];
    const float4 att = P.W.att[slot];
    const uint4 hw = P.W.hit[slot];
    if (F & F_OPL) {
        ps.opl = P.W.opl[slot];
        ps.rayRi = P.W.rayRi[slot];
        ps.shadowRi = P.W.shadowRi[slot];
    }
    if (ITEM) {
        const int4 item = P.W.item[slot];
        ps.pixel = (unsigned int)item.x;
        ps.sampleIndex = (unsigned int)item.y;
        ps.sub = (unsigned int)item.z;
    }
    ps.state = ctlState(ctl);
    ps.shadow = ctlShadow(ctl);
    ps.pathComponent = ctlPathComponent(ctl);
    ctx.prng.s0 = r.x; ctx.prng.s1 = r.y; ctx.prng.s2 = r.z; ctx.prng.s3 = r.w;
    ps.acc[0] = acc.x; ps.acc[1] = acc.y; ps.acc[2] = acc.z;
    ps.walkRay.o = xyz(o); ps.walkRay.time = o.w;
    ps.walkRay.d = xyz(d); ps.pathLength = d.w;
    ps.shDir = xyz(sd); ps.pendHot = __float_as_uint(sd.w);
    ps.pendRad = pr;
    ps.att = att;
    hitB.a = __uint_as_float(hw.x); hitB.ref = hw.y;
    hitS.a = __uint_as_float(hw.z); hitS.ref = hw.w;
}

/* A PathState's own stream for the walk of its shadow query (media draw free-flight samples inside the walk,
 * hitable_medium.hpp:89, and the two walks of a slot run concurrently): two draws of the item's stream seed it. */
WPT_DEV uint4 forkPrng(Prng& prng)
{
    const uint64_t seed = (uint64_t(prngNext(prng)) << 32) | uint64_t(prngNext(prng));
    const uint64_t s01 = splitmix64(seed);
    const uint64_t s23 = splitmix64(s01);
    return make_uint4(uint32_t(s01 >> 32), uint32_t(s01), uint32_t(s23 >> 32), uint32_t(s23));
}

/* Morton number (15 bits) of the grid cell of a point: the binning key of the queries that start there. */
WPT_DEV unsigned int rayCell(const DScene& S, float3 p)
{
    const float top = float(RAY_CELLS) - 1.0f;
    const unsigned int x = (unsigned int)fminf(fmaxf((p.x - S.cellLo.x) * S.cellScale.x, 0.0f), top);
    const unsigned int y = (unsigned int)fminf(fmaxf((p.y - S.cellLo.y) * S.cellScale.y, 0.0f), top);
    const unsigned int z = (unsigned int)fminf(fmaxf((p.z - S.cellLo.z) * S.cellScale.z, 0.0f), top);
    unsigned int key = 0;
#pragma unroll
    for (int b = 0; b < 5; b++)
        key |= (((x >> b) & 1u) << (3 * b + 2)) | (((y >> b) & 1u) << (3 * b + 1)) | (((z >> b) & 1u) << (3 * b));
    return key;
}

/* ITEM: the work item's fields may have changed (generate kernel). `fresh`: see slotCode(). */
template<int F, bool ITEM>
WPT_DEV void storeSlot(const DScene& S, const RenderParams& P, unsigned int slot, PathState& ps, Ctx<F>& ctx, bool fresh)
{
    if (ITEM)
        P.W.item[slot] = make_int4(int(ps.pixel), int(ps.sampleIndex), int(ps.sub), 0);
    P.W.ctl[slot] = makeCtl(ps.state, ps.shadow, ps.pathComponent);
    P.W.code[slot] = uint8_t(slotCode(P, ps.state, ps.shadow, ps.pathComponent, fresh));
    if (ps.state == ST_EXHAUSTED)
        return;
    if ((F & F_MEDIA) && ps.shadow != SHADOW_NONE)
        P.W.shPrng[slot] = forkPrng(ctx.prng);
    P.W.prng[2 * slot] = make_uint4(ctx.prng.s0, ctx.prng.s1, ctx.prng.s2, ctx.prng.s3);
    P.W.acc[2 * slot] = f4(ps.acc[0], ps.acc[1], ps.acc[2], 0.0f);
    if (ps.state == ST_NEW)
        return;     /* path ended: the generate kernel starts the next sample */
    P.W.rayO[2 * slot] = f4(ps.walkRay.o, ps.walkRay.time);
    if (!ITEM && P.W.cell)
        P.W.cell[slot] = uint16_t(rayCell(S, ps.walkRay.o));
    if (ps.state == ST_MAIN) {
        P.W.rayD[2 * slot] = f4(ps.walkRay.d, ps.pathLength);
        P.W.att[slot] = ps.att;
        if (F & F_OPL)
            P.W.rayRi[slot] = ps.rayRi;
    }
    if (F & F_OPL)
        P.W.opl[slot] = ps.opl;
    if (ps.shadow != SHADOW_NONE) {
        P.W.shD[2 * slot] = f4(ps.shDir, __uint_as_float(ps.pendHot));
        P.W.pendRad[2 * slot] = ps.pendRad;
        if (F & F_OPL)
            P.W.shadowRi[slot] = ps.shadowRi;
    }
}

/* ---- per-group counters ---- */
WPT_DEV unsigned long long* ctrTicket(const RenderParams& P) { return P.counters + CTR_TICKET + 16 * P.W.group; }
WPT_DEV unsigned long long* ctrMaterial(const RenderParams& P, int cls) { return P.counters + CTR_MATERIAL + 16 * (P.W.group * MAT_CLASSES + cls); }
WPT_DEV unsigned long long* ctrTraceCamera(const RenderParams& P) { return P.counters + CTR_TRACE_CAMERA + 16 * P.W.group; }
WPT_DEV unsigned long long* ctrTraceOther(const RenderParams& P) { return P.counters + CTR_TRACE_OTHER + 16 * P.W.group; }

/* Entry p of the trace queue (camera rays from the front, the others from the back): slot << 1 | query (0 = B, 1 = S). */
WPT_DEV unsigned int traceQueueEntry(const RenderParams& P, unsigned int p, unsigned int nCamera)
{
    if (p < nCamera)
        return P.W.traceQueue[p];
    return P.W.sortedQueue ? P.W.sortedQueue[p - nCamera] : P.W.traceQueue[2u * P.W.nSlots - 1u - (p - nCamera)];
}

/* Adds one to the counter of every lane's bin (lanes with `valid`), one atomic per distinct bin of the warp; returns
 * the lane's rank among the warp's lanes of the same bin plus the counter's value before the warp's add. */
WPT_DEV unsigned int binAdd(uint32_t* counters, unsigned int bin, bool valid)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned active = __ballot_sync(0xFFFFFFFFu, valid);
    unsigned int result = 0;
    if (valid) {
        const unsigned peers = __match_any_sync(active, bin);
        const int leader = __ffs(peers) - 1;
        unsigned int base = 0;
        if (int(lane) == leader)
            base = atomicAdd(counters + bin, (unsigned int)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        result = base + __popc(peers & ((1u << lane) - 1u));
    }
    return result;
}

WPT_DEV void countSamples(const RenderParams& P, unsigned int newSamples)
{
    const unsigned int n = __reduce_add_sync(0xFFFFFFFFu, newSamples);
    if ((threadIdx.x & 31u) == 0u && n != 0u)
        atomicAdd(P.counters + CTR_SAMPLES, (unsigned long long)n);
}

/* Shading queue of a hit: the material class of the hitable (resolved on the host, scene_host.cpp). */
WPT_DEV unsigned int hitMaterialClass(const DScene& S, uint32_t ref)
{
    const uint32_t type = ref >> REF_TYPE_SHIFT, slot = ref & REF_SLOT_MASK;
    return type == PRIM_TRIANGLE ? __ldg(S.triClass + slot) : type == PRIM_SPHERE ? __ldg(S.sphClass + slot) : __ldg(S.medClass + slot);
}

/* Sort: after the trace kernel, every warp looks at a window of consecutive slots and lists the ones whose camera
 * ray / path segment hit a surface that is to be shaded, one list per material class (the material queues,
 * W.queue). Entries stay in slot order inside a window, so a warp of a material kernel reads near neighbours in the
 * slot arrays (whole sectors), every one of its lanes has work, and all of them run the same material code.
 * One atomic per window and class. Everything else that was answered (rays that left the scene or reached the
 * last path component, paths that only wait for their last shadow query) is left to waveGenerateKernel, which
 * looks at every slot anyway. */
constexpr unsigned int SORT_WINDOW = 256;       /* slots per warp */
constexpr unsigned int TRACE_CLASSES = 8;       /* see waveGenerateKernel */
constexpr unsigned int SORT_BLOCK = 128;
static_assert(SORT_BLOCK == BLOCK && SORT_WINDOW == WAVE_WINDOW, "the sort and generate kernels share one launch geometry");
static_assert(BLOCK == WPT_STACK_STRIDE, "the shared-memory stack columns are BLOCK entries apart");

__global__ void __launch_bounds__(SORT_BLOCK) waveSortKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    __shared__ uint8_t classOf[SORT_BLOCK / 32][SORT_WINDOW];
    __shared__ uint32_t lists[SORT_BLOCK / 32][SORT_WINDOW];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *ctrTicket(P) = 0ull;       /* ticket of the next (persistent) trace launch */
        *ctrTraceCamera(P) = 0ull;  /* trace queue lengths: the trace kernel is done with them, the generate kernel */
        *ctrTraceOther(P) = 0ull;   /* of this round counts them up again */
    }
    const unsigned int windowBase = (blockIdx.x * (SORT_BLOCK / 32u) + warp) * SORT_WINDOW;
    if (windowBase >= P.W.nSlots)
        return;
    unsigned int present = 0;
#pragma unroll
    for (unsigned int k = 0; k < SORT_WINDOW / 32u; k++) {
        const unsigned int slot = windowBase + k * 32u + lane;
        unsigned int cls = MAT_CLASSES;     /* not to be shaded */
        if (slot < P.W.nSlots) {
            const unsigned int code = P.W.code[slot];
            if ((code & 0x0Fu) == (unsigned int)ST_MAIN) {  /* a camera ray / path segment, not on the last path component */
                const uint32_t ref = P.W.hit[slot].y;
                if (ref != NO_HIT)
                    cls = hitMaterialClass(S, ref);
            }
        }
        classOf[warp][k * 32u + lane] = uint8_t(cls);
        present |= (cls < MAT_CLASSES ? 1u << cls : 0u);
    }
    present = __reduce_or_sync(FULL, present);
    __syncwarp();
    for (unsigned int c = 0; c < MAT_CLASSES; c++) {
        if (!((present >> c) & 1u))
            continue;
        unsigned int n = 0;
#pragma unroll
        for (unsigned int k = 0; k < SORT_WINDOW / 32u; k++) {
            const bool mine = (classOf[warp][k * 32u + lane] == c);
            const unsigned mask = __ballot_sync(FULL, mine);
            if (mine)
                lists[warp][n + __popc(mask & below)] = windowBase + k * 32u + lane;
            n += __popc(mask);
        }
        __syncwarp();
        unsigned int base = 0;
        if (lane == 0)
            base = (unsigned int)atomicAdd(ctrMaterial(P, c), (unsigned long long)n);
        base = __shfl_sync(FULL, base, 0);
        uint32_t* q = P.W.queue + size_t(c) * P.W.queueStride;
        for (unsigned int e = lane; e < n; e += 32u)
            q[base + e] = lists[warp][e];
        __syncwarp();
    }
}

/* Generate: finishes the paths the material kernels did not continue (rays that left the scene or reached the last
 * path component, the last shadow query of an ended path), starts the next sample (camera ray) of every slot whose
 * path has ended, hands out work items, retires slots when none are left, and lists the queries to trace. One warp
 * per window of WAVE_WINDOW consecutive slots; a window takes its work items from a block of WAVE_WINDOW
 * consecutive items at a time (MPICoordinator::getBlock, mpi.hpp:241-254, as a device atomic), so the slots of a
 * window always work on neighbouring pixels and their camera rays - listed first in the trace queue, window by
 * window - walk the tree together. */
template<int F>
__global__ void __launch_bounds__(BLOCK, WPT_LOGIC_MIN_BLOCKS) waveGenerateKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    /* per warp: the trace class of query B and of query S of each slot of the window, then the queries listed class by class */
    __shared__ uint8_t classOf[BLOCK / 32][2 * WAVE_WINDOW];
    __shared__ uint32_t lists[BLOCK / 32][2 * WAVE_WINDOW];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    const unsigned int w = blockIdx.x * (BLOCK / 32u) + warp;
    const unsigned int windowBase = w * WAVE_WINDOW;
    if (windowBase >= P.W.nSlots)
        return;
    uint2 block = P.W.window[w];            /* x = next item, y = end of the block; x = ~0: no items left */
    unsigned int newSamples = 0, retired = 0;
    for (unsigned int k = 0; k < WAVE_WINDOW / 32u; k++) {
        const unsigned int slot = windowBase + k * 32u + lane;
        const bool valid = slot < P.W.nSlots;
        const unsigned int code = valid ? (unsigned int)P.W.code[slot] : (unsigned int)ST_EXHAUSTED;
        int state = codeState(code);
        bool shadow = (state == ST_MAIN || state == ST_TAIL) && (code & CODE_SHADOW);
        unsigned int pathComponent = code >> 6;     /* min(path component, 3): all the class needs */
        const bool fresh = (state == ST_MAIN || state == ST_TAIL) && (code & CODE_FRESH);
        const bool generate = valid && (state == ST_NEW || state == ST_IDLE);
        /* answered queries the material kernels did not take: the path ends here */
        const bool logic = (state == ST_MAIN || state == ST_TAIL) && !fresh;
        if (fresh)
            P.W.code[slot] = uint8_t(code & ~CODE_FRESH);   /* answered by the next trace launch */
        if (__any_sync(FULL, generate || logic)) {
            Ctx<F> ctx;
            ctx.nested = 0;
            ctx.nodes = ctx.leaves = 0;
            PathState ps;
            ps.state = ST_EXHAUSTED;
            ps.shadow = SHADOW_NONE;
            if (logic) {
                Hit hitB, hitS;
                loadAnswered<F, true>(P, slot, ps, ctx, hitB, hitS);
                animationCacheInit<F>(S, ctx, ps.walkRay.time);
                if (ps.shadow != SHADOW_NONE)
                    consumeShadow<F>(S, P, ps, ctx, hitS);
                if (ps.state == ST_MAIN)
                    waveEndPath<F>(S, P, ps, hitB);
                ps.state = ST_NEW;
            } else if (generate) {
                loadItem<F>(P, slot, state == ST_IDLE, ps, ctx);
            }
            bool working = generate || logic;
            while (__any_sync(FULL, working)) {
                if (working && ps.state == ST_NEW) {
                    if (startSample<F>(S, P, ps, ctx)) {
                        newSamples++;
                        working = false;
                    } else {
                        finishItem(P, ps);
                        ps.state = ST_IDLE;
                    }
                }
                const bool wantItem = working && ps.state == ST_IDLE;
                const unsigned wantMask = __ballot_sync(FULL, wantItem);
                if (wantMask != 0u) {
                    if (block.x == block.y && block.x != 0xFFFFFFFFu) {     /* the window's block is used up */
                        unsigned int base = 0;
                        if (lane == 0) {
                            base = (unsigned int)atomicAdd_system(P.itemCounter, (unsigned long long)WAVE_WINDOW);
                            if (P.blockDone && base < P.nItems)
                                P.blockDone[base / WAVE_WINDOW] = 1;    /* this GPU renders (and resolves) these items' pixels */
                        }
                        base = __shfl_sync(FULL, base, 0);
                        if (base >= P.nItems)
                            block = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
                        else
                            block = make_uint2(base, min(base + WAVE_WINDOW, P.nItems));
                    }
                    const unsigned int avail = block.y - block.x;
                    const unsigned int rank = __popc(wantMask & below);
                    if (wantItem) {
                        if (block.x == 0xFFFFFFFFu) {
                            ps.state = ST_EXHAUSTED;
                            working = false;
                            retired++;
                        } else if (rank < avail) {
                            const WorkItem wi = decodeItem(P, block.x + rank);
                            if (wi.valid)
                                beginItem<F>(ps, ctx, wi);      /* -> ST_NEW */
                        }
                    }
                    if (block.x != 0xFFFFFFFFu)
                        block.x += min((unsigned int)__popc(wantMask), avail);
                }
            }
            if (generate || logic) {
                ps.shadow = SHADOW_NONE;
                storeSlot<F, true>(S, P, slot, ps, ctx, false);
                state = ps.state;
                shadow = false;
                pathComponent = 0;
            }
        }
        /* Trace classes: rays of one class from one window (= a few neighbouring pixels) are similar rays.
         * 0 = camera rays; 1..3 = shadow rays set up at path component 0, 1, 2+ (same region of the scene, and
         * most of them towards the brightest part of the environment); 5..7 = path segments 1, 2, 3+. */
        unsigned int clsB = TRACE_CLASSES, clsS = TRACE_CLASSES;       /* nothing to trace */
        if (state == ST_MAIN)
            clsB = (pathComponent == 0u ? 0u : 4u + pathComponent);
        if (shadow)
            clsS = max(pathComponent, 1u);     /* the slot's path component already counts the next segment */
        classOf[warp][2u * (k * 32u + lane)] = uint8_t(clsB);
        classOf[warp][2u * (k * 32u + lane) + 1u] = uint8_t(min(clsS, TRACE_CLASSES));
    }
    __syncwarp();
    /* counting sort of the window's queries by class; slot order is kept inside a class */
    unsigned int nCam = 0, nListed = 0;
    for (unsigned int c = 0; c < TRACE_CLASSES; c++) {
#pragma unroll
        for (unsigned int k = 0; k < 2u * WAVE_WINDOW / 32u; k++) {
            const bool mine = (classOf[warp][k * 32u + lane] == c);
            const unsigned mask = __ballot_sync(FULL, mine);
            if (mine)
                lists[warp][nListed + __popc(mask & below)] = 2u * windowBase + k * 32u + lane;     /* slot << 1 | query */
            nListed += __popc(mask);
        }
        if (c == 0)
            nCam = nListed;
    }
    __syncwarp();
    const unsigned int nOther = nListed - nCam;
    unsigned int camBase = 0, otherBase = 0;
    if (lane == 0) {
        P.W.window[w] = block;
        if (nCam)
            camBase = (unsigned int)atomicAdd(ctrTraceCamera(P), (unsigned long long)nCam);
        if (nOther)
            otherBase = (unsigned int)atomicAdd(ctrTraceOther(P), (unsigned long long)nOther);
    }
    camBase = __shfl_sync(FULL, camBase, 0);
    otherBase = __shfl_sync(FULL, otherBase, 0);
    for (unsigned int e = lane; e < nCam; e += 32u)
        P.W.traceQueue[camBase + e] = lists[warp][e];
    for (unsigned int e0 = 0; e0 < nOther; e0 += 32u) {
        const unsigned int e = e0 + lane;
        const bool valid = e < nOther;
        const unsigned int entry = valid ? lists[warp][nCam + e] : 0u;
        if (valid)
            P.W.traceQueue[2u * P.W.nSlots - 1u - (otherBase + e)] = entry;
        if (P.W.binCount)      /* how many queries start in each grid cell (bounce and shadow rays apart) */
            binAdd(P.W.binCount, valid ? ((unsigned int)P.W.cell[entry >> 1] << 1) | (entry & 1u) : 0u, valid);
    }
    countSamples(P, newSamples);
    const unsigned int r = __reduce_add_sync(FULL, retired);
    if (lane == 0 && r != 0u)
        atomicAdd(P.counters + CTR_DONE, (unsigned long long)r);
}

/* Binning of the bounce / shadow queries by the grid cell of their origin, after the generate kernel has counted
 * them per bin: start offset of every bin (one block), then every query goes to the next free place of its bin. The
 * order inside a bin does not matter (every query is answered on its own). Measured with the query kernel on C3's
 * bounce rays: +22 ... +33 % rays/s when the rays are taken cell by cell (profiles/r2l_coherence_probe.txt). */
__global__ void __launch_bounds__(1024) waveBinScanKernel(const __grid_constant__ RenderParams P)
{
    __shared__ unsigned int warpSums[32];
    constexpr unsigned int PER = RAY_BINS / 1024;
    const unsigned int first = threadIdx.x * PER;
    unsigned int sum = 0;
    for (unsigned int k = 0; k < PER; k++)
        sum += P.W.binCount[first + k];
    unsigned int x = sum;
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, x, off);
        if ((threadIdx.x & 31u) >= (unsigned)off)
            x += y;
    }
    if ((threadIdx.x & 31u) == 31u)
        warpSums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned int w = warpSums[threadIdx.x];
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, w, off);
            if (threadIdx.x >= (unsigned)off)
                w += y;
        }
        warpSums[threadIdx.x] = w;
    }
    __syncthreads();
    unsigned int run = ((threadIdx.x >> 5) ? warpSums[(threadIdx.x >> 5) - 1] : 0u) + x - sum;
    for (unsigned int k = 0; k < PER; k++) {
        const unsigned int c = P.W.binCount[first + k];
        P.W.binBase[first + k] = run;
        P.W.binBase[RAY_BINS + first + k] = 0u;     /* cursor */
        P.W.binCount[first + k] = 0u;               /* for the next round's generate kernel */
        run += c;
    }
}

__global__ void __launch_bounds__(256) waveBinScatterKernel(const __grid_constant__ RenderParams P)
{
    const unsigned int nOther = (unsigned int)*ctrTraceOther(P);
    const unsigned int rounded = (nOther + 31u) & ~31u;
    for (unsigned int p = blockIdx.x * 256u + threadIdx.x; p < rounded; p += gridDim.x * 256u) {
        const bool valid = p < nOther;
        const unsigned int entry = valid ? P.W.traceQueue[2u * P.W.nSlots - 1u - p] : 0u;
        const unsigned int bin = valid ? ((unsigned int)P.W.cell[entry >> 1] << 1) | (entry & 1u) : 0u;
        const unsigned int pos = binAdd(P.W.binBase + RAY_BINS, bin, valid);
        if (valid)
            P.W.sortedQueue[P.W.binBase[bin] + pos] = entry;
    }
}

/* Material: shade the surface hits of camera rays and path segments, one kernel per material class (MC_*; the
 * material type is a compile-time constant in all but the generic one). A fixed grid strides over the class's queue. */
template<int MC> struct ClassMaterialType { static constexpr int value = -1; };
template<> struct ClassMaterialType<MC_LAMBERTIAN> { static constexpr int value = WPT_MAT_LAMBERTIAN; };
template<> struct ClassMaterialType<MC_GGX> { static constexpr int value = WPT_MAT_GGX; };
template<> struct ClassMaterialType<MC_GLASS> { static constexpr int value = WPT_MAT_GLASS; };
template<> struct ClassMaterialType<MC_MODPHONG> { static constexpr int value = WPT_MAT_MODPHONG; };

template<int F, int MC>
__global__ void __launch_bounds__(BLOCK, WPT_MATERIAL_MIN_BLOCKS) waveMaterialKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    const unsigned int count = (unsigned int)*ctrMaterial(P, MC);
    const uint32_t* queue = P.W.queue + size_t(MC) * P.W.queueStride;
    for (unsigned int i = blockIdx.x * BLOCK + threadIdx.x; i < count; i += gridDim.x * BLOCK) {
        const unsigned int slot = queue[i];
        Ctx<F> ctx;
        ctx.nested = 0;
        ctx.nodes = ctx.leaves = 0;
        PathState ps;
        Hit hitB, hitS;
        loadAnswered<F, false>(P, slot, ps, ctx, hitB, hitS);
        animationCacheInit<F>(S, ctx, ps.walkRay.time);
        if (ps.shadow != SHADOW_NONE)
            consumeShadow<F>(S, P, ps, ctx, hitS);
        const RayHelper h = makeRayHelper(ps.walkRay);
        waveShade<F, ClassMaterialType<MC>::value>(S, P, ps, ctx, hitB, h);
        storeSlot<F, false>(S, P, slot, ps, ctx, true);
    }
}

/* Trace: the closest-hit query of every entry of the trace queue, one per thread (reference-order walk, and the
 * wide walk without lane refill as an experiment). */
template<int F>
__global__ void __launch_bounds__(BLOCK, WPT_TRACE_MIN_BLOCKS) waveTraceKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    extern __shared__ uint2 smStackAll[];
    uint2* smStack = smStackAll + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  /* queue lengths of this round (the sort kernel that follows counts them up) */
        for (int c = 0; c < MAT_CLASSES; c++)
            *ctrMaterial(P, c) = 0ull;
    }
    const unsigned int nCamera = (unsigned int)*ctrTraceCamera(P);
    const unsigned int nQueued = nCamera + (unsigned int)*ctrTraceOther(P);
    unsigned int traced = 0;
    Ctx<F> ctx;
    ctx.nested = 0;
    ctx.nodes = ctx.leaves = 0;
    for (unsigned int p = blockIdx.x * BLOCK + threadIdx.x; p < nQueued; p += gridDim.x * BLOCK) {
        const unsigned int e = traceQueueEntry(P, p, nCamera);
        const unsigned int slot = e >> 1, which = e & 1u;
        const float4 o = P.W.rayO[2 * slot], d = (which ? P.W.shD[2 * slot] : P.W.rayD[2 * slot]);
        Ray ray;
        ray.o = xyz(o); ray.time = o.w; ray.d = xyz(d);
        if (F & F_MEDIA) {
            const uint4 r = (which ? P.W.shPrng[slot] : P.W.prng[2 * slot]);
            ctx.prng.s0 = r.x; ctx.prng.s1 = r.y; ctx.prng.s2 = r.z; ctx.prng.s3 = r.w;
        }
        animationCacheInit<F>(S, ctx, ray.time);
        const RayHelper h = makeRayHelper(ray);
        const Hit hit = closestHit<F>(S, P, ray, h, ctx, smStack);
        reinterpret_cast<uint2*>(P.W.hit)[e] = make_uint2(__float_as_uint(hit.a), hit.ref);
        if ((F & F_MEDIA) && !which)
            P.W.prng[2 * slot] = make_uint4(ctx.prng.s0, ctx.prng.s1, ctx.prng.s2, ctx.prng.s3);
        traced++;
    }
    /* statistics: one atomic per warp */
    const unsigned int n = __reduce_add_sync(0xFFFFFFFFu, traced);
    const unsigned int q = __reduce_add_sync(0xFFFFFFFFu, ctx.nested);
    if ((threadIdx.x & 31) == 0) {
        if (n)
            atomicAdd(P.counters + CTR_RAYS, (unsigned long long)n);
        if (q)
            atomicAdd(P.counters + CTR_NESTED, (unsigned long long)q);
    }
}

/* Trace, persistent form (4-wide walk only): every lane walks the tree for one query and, the moment its walk
 * ends, stores the answer and takes the next entry of the trace queue from a ticket counter -- inside the walk
 * loop, so a lane never waits for the longest ray of its warp. One loop iteration handles one inner node or one
 * primitive of a leaf. */
/* A leaf round can be run cooperatively when every waiting lane's leaf holds triangles only (warp-uniform result). */
WPT_DEV bool coopLeafRound(unsigned leafMask, bool atLeaf, int32_t cur)
{
    const uint32_t type = uint32_t(~cur) >> REF_TYPE_SHIFT;
    const bool other = atLeaf && !(type == PRIM_TRIANGLE || type == PRIM_TRIANGLE_RUN);
    return leafMask != 0u && __ballot_sync(0xFFFFFFFFu, other) == 0u;
}

template<int F>
__global__ void __launch_bounds__(BLOCK, (F == 0 ? WPT_TRACE_MIN_BLOCKS_STATIC : WPT_TRACE_MIN_BLOCKS)) waveTracePersistentKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    extern __shared__ uint2 smStackAll[];
    __shared__ float4 coopResult[(WPT_LEAF_COOP && !(F & (F_ANIM | F_MEDIA))) ? (BLOCK / 32) * 128 : 1];
    __shared__ uint8_t coopOwner[(WPT_LEAF_COOP && !(F & (F_ANIM | F_MEDIA))) ? (BLOCK / 32) * 128 : 1];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneBelow = (1u << lane) - 1u;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    Ctx<F> ctx;
    ctx.nested = 0;
    ctx.nodes = ctx.leaves = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  /* queue lengths of this round (the sort kernel that follows counts them up) */
        for (int c = 0; c < MAT_CLASSES; c++)
            *ctrMaterial(P, c) = 0ull;
    }

    const unsigned int nCamera = (unsigned int)*ctrTraceCamera(P);
    const unsigned int nQueued = nCamera + (unsigned int)*ctrTraceOther(P);
    if (blockIdx.x == 0 && threadIdx.x == 0 && nQueued != 0u)       /* statistics: every queued query is answered by this launch */
        atomicAdd(P.counters + CTR_RAYS, (unsigned long long)nQueued);
    /* Without hot spots every shadow query S asks whether the environment is visible (wurblpt.hpp:236-251): only
     * "nothing hit" is read of its answer (consumeShadow), so its walk may stop at the first hit instead of looking for
     * the closest one. Not with media: their walks draw from the Prng, the draws must not depend on the visiting order. */
    const bool anyHitScene = WPT_ANY_HIT && !(F & F_MEDIA) && S.nHot == 0u;
    /* media bounded by a small triangle mesh are tested by the whole warp in leaf rounds (mediumRoundCoop) */
    constexpr bool MEDIUM_COOP = WPT_MEDIUM_COOP && (F & F_MEDIA) && !(F & F_ANIM);
    unsigned int entry = 0;             /* slot << 1 | query of the walk in progress */
    bool active = false;
    Ray ray;
    ray.o = f3(0.0f); ray.d = f3(0.0f, 0.0f, 1.0f); ray.time = 0.0f;
    RayHelper h = makeRayHelper(ray);
#if WPT_NODE_FAST
    RowOffsets rows = rowOffsets(ray.d);
#else
    uint32_t rows = 0;                  /* octantRows(ray.d); only used with WPT_OCTANT_ROWS */
#endif
    WideBest best;
    best.hit.a = MAXVAL;
    best.hit.ref = NO_HIT;
    best.Ts = best.det = 0.0f;
    int32_t cur = EMPTY_CHILD;
    /* the walk's stack: a column of the block's shared-memory array, a local-memory spill, the pointer in a register */
    const uint32_t smBase = uint32_t(__cvta_generic_to_shared(smStackAll + threadIdx.x));
    uint2 spill[LOCAL_STACK];
    int sp = 0;
#if WPT_STACK_TOS
    StackTop top;
    top.code = uint32_t(EMPTY_CHILD);
    top.tmin = 0u;
#define WPT_PUSH_FAR(nk) pushFarChildrenTop(smBase, spill, sp, top, nk)
#define WPT_POP_NEXT() stackPopNextTop(smBase, spill, sp, top, best.hit.a)
#else
#define WPT_PUSH_FAR(nk) pushFarChildren(smBase, spill, sp, nk)
#define WPT_POP_NEXT() stackPopNext(smBase, spill, sp, best.hit.a)
#endif
    const float amin = P.minHit;

    /* Tickets are drawn from the global counter a chunk at a time and handed out inside the warp
     * without further atomics; idle lanes are only refilled once enough of them have gathered (or
     * none is busy), because a refill stalls the whole warp on the loads of the new rays. */
    /* 256 entries per ticket when the queue is long; down to 32 when it is short (the last rounds of a frame, small
     * frames), so that the few rays are spread over all resident warps instead of queueing up in a few of them */
    const unsigned int perWarp = nQueued / (gridDim.x * (BLOCK / 32u));
    const unsigned int CHUNK = perWarp >= 2048u ? 256u : perWarp >= 512u ? 128u : perWarp >= 128u ? 64u : 32u;
    constexpr int REFILL_MIN = WPT_REFILL_MIN;
    unsigned int warpNext = 0, warpEnd = 0;
    bool drained = false;           /* the global ticket counter has run past the last entry */
    unsigned exhaustedMask = 0u;    /* lanes that found the queue empty (warp-uniform; only a refill changes it) */
    for (;;) {
        const unsigned busyMask = __ballot_sync(FULL, active);
        const unsigned wantMask = ~busyMask & ~exhaustedMask;
        unsigned activeMask = busyMask;
        /* camera rays of one window are coherent: they are started 32 at a time so that they stay together */
        const int refillMin = (warpNext < warpEnd && warpNext < nCamera) ? WPT_CAMERA_REFILL_MIN : REFILL_MIN;
        if (wantMask != 0u && (__popc(wantMask) >= refillMin || busyMask == 0u)) {
            if (warpNext >= warpEnd && !drained) {
                unsigned int base = 0;
                if (lane == 0)
                    base = (unsigned int)atomicAdd(ctrTicket(P), (unsigned long long)CHUNK);
                base = __shfl_sync(FULL, base, 0);
                warpNext = base;
                warpEnd = min(base + CHUNK, nQueued);
                if (base >= nQueued) {
                    drained = true;
                    warpNext = warpEnd = 0;
                }
            }
            const unsigned int avail = warpEnd - warpNext;
            const unsigned int rank = __popc(wantMask & laneBelow);
            const bool got = ((wantMask >> lane) & 1u) && rank < avail;
            if (got) {
                const unsigned int e = traceQueueEntry(P, warpNext + rank, nCamera);
                const unsigned int s = e >> 1, which = e & 1u;
                const float4 o = P.W.rayO[2 * s], d = (which ? P.W.shD[2 * s] : P.W.rayD[2 * s]);
                ray.o = xyz(o); ray.time = o.w; ray.d = xyz(d);
                if (F & F_MEDIA) {
                    const uint4 r = (which ? P.W.shPrng[s] : P.W.prng[2 * s]);
                    ctx.prng.s0 = r.x; ctx.prng.s1 = r.y; ctx.prng.s2 = r.z; ctx.prng.s3 = r.w;
                }
                animationCacheInit<F>(S, ctx, ray.time);
                h = makeRayHelper(ray);
#if WPT_NODE_FAST
                rows = rowOffsets(ray.d);
#else
                rows = octantRows(ray.d);
#endif
                entry = e;
                active = true;
                best.hit.a = MAXVAL;
                best.hit.ref = NO_HIT;
                sp = 0;
#if WPT_STACK_TOS
                top.code = uint32_t(EMPTY_CHILD);
#endif
                cur = S.root4;
            }
            const unsigned gotMask = __ballot_sync(FULL, got);
            activeMask |= gotMask;
            if (drained)        /* the lanes that got nothing will never get anything */
                exhaustedMask |= wantMask & ~gotMask;
            warpNext += min((unsigned int)__popc(wantMask), avail);
        }
        if (exhaustedMask == FULL)
            break;
        /* One iteration runs EITHER the inner-node code OR the leaf code, never both: lanes that have
         * reached a leaf wait until LEAF_MIN lanes have one (or no lane has an inner node left), then
         * those lanes test their primitives together. Without this, the few lanes that are at a leaf in
         * any given iteration make the whole warp pay for the leaf code every time. */
        const bool atLeaf = active && cur < 0;
        const unsigned leafMask = __ballot_sync(FULL, atLeaf);
        const unsigned innerMask = activeMask & ~leafMask;
        const bool leafRound = (__popc(leafMask) >= WPT_LEAF_MIN) || innerMask == 0u;
#if WPT_NODE_REPEAT > 1
        if (!active && leafRound && !(WPT_LEAF_COOP && !(F & (F_ANIM | F_MEDIA))) && !MEDIUM_COOP)
            continue;       /* (idle lanes take part in the ballots of a node round, and test triangles for the others
                               in a cooperative leaf round) */
#else
        const bool atInner = active && cur >= 0 && cur != EMPTY_CHILD;   /* (EMPTY_CHILD: the root of an empty scene) */
        if (!active && !(((WPT_LEAF_COOP && !(F & (F_ANIM | F_MEDIA))) || MEDIUM_COOP) && leafRound))
            continue;       /* (idle lanes test triangles for the others in a cooperative leaf round) */
#endif

        bool mediumDone = false;
        if (MEDIUM_COOP && leafRound) {
            uint32_t run = 0u;
            if (atLeaf && (uint32_t(~cur) >> REF_TYPE_SHIFT) == PRIM_MEDIUM)
                run = __float_as_uint(__ldg(S.med + (uint32_t(~cur) & REF_SLOT_MASK)).w);
            if (__any_sync(FULL, run != 0u)) {
                mediumRoundCoop<F>(S, run, uint32_t(~cur), ray, h, amin, P.minHit, ctx, best);
                if (run != 0u) {
                    cur = WPT_POP_NEXT();
                    mediumDone = true;
                }
            }
        }
#if WPT_NODE_REPEAT <= 1
        if (!leafRound) {
            if (atInner) {
#if WPT_NODE_FAST
                const NodeKeys nk = nodeKeysFast(S, cur, ray.o, h.inv, amin, best.hit.a, rows);
#else
                const NodeKeys nk = nodeKeys(S, cur, ray.o, h.inv, amin, best.hit.a, rows);
#endif
                if (!keyMissed(nk.k0)) {
                    WPT_PUSH_FAR(nk);
                    cur = sel4i(nk.ch, nk.k0 & 3u);
                } else {
                    cur = WPT_POP_NEXT();
                }
            }
        } else
#else
        if (!leafRound) {
            /* A node round runs up to WPT_NODE_REPEAT node steps while at least WPT_NODE_REPEAT_MIN lanes are still at
             * an inner node: the bookkeeping at the top and the bottom of the loop (~55 instructions: refill and
             * leaf-round decisions, three ballots) is then paid once for several steps. Lanes that arrive at a leaf or
             * finish in the first step wait for the next iteration. A lane's own walk is the same sequence either way. */
#pragma unroll 1
            for (int rep = 0;; rep++) {
                if (active && cur >= 0 && cur != EMPTY_CHILD) {
#if WPT_NODE_FAST
                    const NodeKeys nk = nodeKeysFast(S, cur, ray.o, h.inv, amin, best.hit.a, rows);
#else
                    const NodeKeys nk = nodeKeys(S, cur, ray.o, h.inv, amin, best.hit.a, rows);
#endif
                    if (!keyMissed(nk.k0)) {
                        WPT_PUSH_FAR(nk);
                        cur = sel4i(nk.ch, nk.k0 & 3u);
                    } else {
                        cur = WPT_POP_NEXT();
                    }
#if WPT_PREFETCH_LEAF
                    if (S.prefetchLeaves && cur < 0)
                        prefetchLeaf(S, uint32_t(~cur));
#endif
                }
                if (rep + 1 >= WPT_NODE_REPEAT)
                    break;
                if (__popc(__ballot_sync(FULL, active && cur >= 0 && cur != EMPTY_CHILD)) < WPT_NODE_REPEAT_MIN)
                    break;
            }
        } else
#endif
        if (WPT_LEAF_COOP && !(F & (F_ANIM | F_MEDIA)) && coopLeafRound(leafMask, atLeaf, cur)) {
            /* Cooperative leaf round: a leaf holds 1-4 triangles and only ~10 lanes wait at one, so looping over "my"
             * triangles runs the (long) triangle test up to four times at a third of the lanes. Instead the triangles of
             * all waiting lanes are numbered (prefix sum of the counts) and every lane of the warp tests ONE of them for
             * its owner, whose ray it fetches by shuffle; the owners then take their results in slot order exactly as the
             * loop would: a later triangle of the run only counts under the distance an earlier one has set. */
            uint32_t first = 0, cnt = 0;
            if (atLeaf) {
                const uint32_t ref = uint32_t(~cur);
                const bool run = (ref >> REF_TYPE_SHIFT) == PRIM_TRIANGLE_RUN;
                first = run ? (ref & RUN_SLOT_MASK) : (ref & REF_SLOT_MASK);
                cnt = run ? ((ref >> RUN_COUNT_SHIFT) & 3u) + 1u : 1u;
            }
            uint32_t inc = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, inc, off);
                if (lane >= (unsigned)off)
                    inc += y;
            }
            const uint32_t total = __shfl_sync(FULL, inc, 31);
            const uint32_t excl = inc - cnt;
            uint8_t* ownerOf = coopOwner + (threadIdx.x >> 5) * 128;
            float4* result = coopResult + (threadIdx.x >> 5) * 128;
            for (uint32_t i = 0; i < cnt; i++)
                ownerOf[excl + i] = uint8_t(lane);
            __syncwarp();
            const uint32_t kpack = uint32_t(h.kx) | (uint32_t(h.ky) << 2) | (uint32_t(h.kz) << 4);
            for (uint32_t base = 0; base < total; base += 32u) {
                const uint32_t j = base + lane;
                const bool valid = j < total;
                const unsigned own = valid ? ownerOf[j] : 0u;
                Ray oray;
                oray.o = f3(__shfl_sync(FULL, ray.o.x, own), __shfl_sync(FULL, ray.o.y, own), __shfl_sync(FULL, ray.o.z, own));
                oray.d = f3(0.0f);
                oray.time = 0.0f;
                RayHelper oh;
                oh.inv = f3(0.0f);
                oh.Sx = __shfl_sync(FULL, h.Sx, own);
                oh.Sy = __shfl_sync(FULL, h.Sy, own);
                oh.Sz = __shfl_sync(FULL, h.Sz, own);
                const uint32_t ok = __shfl_sync(FULL, kpack, own);
                oh.kx = int(ok & 3u); oh.ky = int((ok >> 2) & 3u); oh.kz = int(ok >> 4);
                const float oamax = __shfl_sync(FULL, best.hit.a, own);
                const uint32_t ofirst = __shfl_sync(FULL, first, own);
                const uint32_t oexcl = __shfl_sync(FULL, excl, own);
                if (valid) {
                    const float4* tp = S.tri + size_t(ofirst + (j - oexcl)) * 3;
                    const float4 t0 = __ldg(tp + 0), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                    float ta = 0.0f, U, V, W, det, Ts = 0.0f, da = 0.0f;
                    const bool found = triangleTest(xyz(t0), xyz(t1), xyz(t2), oray, oh, amin, oamax, true, ta, U, V, W, det, Ts, da);
                    result[j] = f4(ta, Ts, da, found ? 1.0f : 0.0f);
                }
            }
            __syncwarp();
            if (atLeaf) {
                for (uint32_t i = 0; i < cnt; i++) {
                    const float4 r = result[excl + i];
                    /* (the test ran under the distance the round started with: redo its upper bound, hitable_triangle.hpp:263-264) */
                    if (r.w != 0.0f && !(r.y > best.hit.a * r.z))
                        considerHit(best, (PRIM_TRIANGLE << REF_TYPE_SHIFT) | (first + i), r.x, r.y, r.z);
                }
                cur = WPT_POP_NEXT();
            }
            __syncwarp();
        } else if (atLeaf && !mediumDone) {
#if WPT_LEAF_UNIT
            /* one primitive per lane and leaf round; what is left of a run stays this lane's leaf */
            const uint32_t rest = wideLeafStep<F>(S, uint32_t(~cur), ray, h, amin, P.minHit, ctx, best);
            cur = (rest != NO_HIT ? int32_t(~rest) : WPT_POP_NEXT());
#else
            wideLeaf<F>(S, uint32_t(~cur), ray, h, amin, P.minHit, ctx, best);
            /* occlusion query: the first hit answers it */
            cur = (anyHitScene && (entry & 1u) && best.hit.ref != NO_HIT) ? EMPTY_CHILD : WPT_POP_NEXT();
#endif
        }
#if WPT_PREFETCH_LEAF
#if WPT_NODE_REPEAT > 1
        if (S.prefetchLeaves && cur < 0 && leafRound && atLeaf)
#else
        if (S.prefetchLeaves && cur < 0 && ((!leafRound && atInner) || (leafRound && atLeaf)))
#endif
            prefetchLeaf(S, uint32_t(~cur));
#endif
#if WPT_PREFETCH_NODE && WPT_NODE_REPEAT <= 1
        if (S.prefetchLeaves && cur >= 0 && cur != EMPTY_CHILD && ((!leafRound && atInner) || (leafRound && atLeaf)))
            asm volatile("prefetch.global.L1 [%0];" :: "l"(S.node4 + size_t(cur) * 8));
#endif
        if (active && cur == EMPTY_CHILD) {
            reinterpret_cast<uint2*>(P.W.hit)[entry] = make_uint2(__float_as_uint(best.hit.a), best.hit.ref);
            if ((F & F_MEDIA) && !(entry & 1u))
                P.W.prng[2 * (entry >> 1)] = make_uint4(ctx.prng.s0, ctx.prng.s1, ctx.prng.s2, ctx.prng.s3);
            active = false;
        }
    }
    const unsigned int q = __reduce_add_sync(FULL, ctx.nested);
    if (lane == 0 && q)
        atomicAdd(P.counters + CTR_NESTED, (unsigned long long)q);
}
#undef WPT_PUSH_FAR
#undef WPT_POP_NEXT

/* ------------------------------------------------------------ quad trace kernel */

#ifndef WPT_TRACE_QUAD_DEFAULT
#define WPT_TRACE_QUAD_DEFAULT 0    /* 1: static scenes without media are traced by waveTraceQuadKernel (WPT_TRACE_QUAD=0/1 overrides) */
#endif
#ifndef WPT_QUAD_STACK
#define WPT_QUAD_STACK 40       /* shared-memory stack entries per quad (deeper entries: local memory) */
#endif
#ifndef WPT_QUAD_REFILL_MIN
#define WPT_QUAD_REFILL_MIN 12  /* idle LANES (3 quads) that must have gathered before new queries are fetched */
#endif
#ifndef WPT_QUAD_LEAF_MIN
#define WPT_QUAD_LEAF_MIN 12    /* LANES (3 quads) that must wait at a leaf before a leaf round runs */
#endif
#ifndef WPT_QUAD_MIN_BLOCKS
#define WPT_QUAD_MIN_BLOCKS 8
#endif
constexpr int QUAD_STACK = WPT_QUAD_STACK;
constexpr int QUAD_SPILL = 96 - QUAD_STACK;         /* same capacity as the per-lane walk: wideStackEntries() */
constexpr int QUADS_PER_BLOCK = BLOCK / 4;
constexpr int QUAD_STRIDE = QUADS_PER_BLOCK + 1;    /* entries of one stack level are QUAD_STRIDE apart: the +1 spreads the
                                                       levels that the lanes of one quad push together over the banks */
static_assert(SM_STACK + LOCAL_STACK == QUAD_STACK + QUAD_SPILL, "both walks have the stack capacity capi.cu checks against");

static size_t quadStackBytes()
{
    return size_t(QUAD_STACK) * QUAD_STRIDE * sizeof(uint2);
}

/* The wide nodes child by child: node4 holds one 128-byte node as six rows of four child planes plus the child row (what
 * a lane that tests all four children wants); node4c holds the same 128 bytes as four 32-byte child records
 * (lo.x lo.y lo.z hi.x | hi.y hi.z child 0), what a lane that tests ONE child wants: two 128-bit loads. */
__global__ void __launch_bounds__(256) nodeTransposeKernel(const float4* __restrict__ node4, float4* __restrict__ node4c, uint32_t nNodes)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;     /* (node, child) */
    if (i >= nNodes * 4u)
        return;
    const uint32_t node = i >> 2, k = i & 3u;
    const float* src = reinterpret_cast<const float*>(node4 + size_t(node) * 8);
    const float lox = src[0 + k], hix = src[4 + k], loy = src[8 + k], hiy = src[12 + k], loz = src[16 + k], hiz = src[20 + k];
    const float child = src[24 + k];
    node4c[size_t(node) * 8 + 2 * k] = f4(lox, loy, loz, hix);
    node4c[size_t(node) * 8 + 2 * k + 1] = f4(hiy, hiz, child, 0.0f);
}

cudaError_t launchNodeTranspose(const float4* node4, float4* node4c, uint32_t nNodes, cudaStream_t stream)
{
    if (nNodes == 0)
        return cudaSuccess;
    nodeTransposeKernel<<<(nNodes * 4u + 255u) / 256u, 256, 0, stream>>>(node4, node4c, nNodes);
    return cudaGetLastError();
}

WPT_DEV void quadStackStore(uint32_t smBase, int pos, uint32_t code, uint32_t tminBits)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(smBase + uint32_t(pos) * uint32_t(QUAD_STRIDE * sizeof(uint2))),
            "r"(code), "r"(tminBits) : "memory");
}

/* Next stack entry whose entry distance is still under `limit` (all four lanes of a quad run this with the same
 * arguments and read the same entries: broadcasts). */
WPT_DEV int32_t quadStackPopNext(uint32_t smBase, const uint2* spill, int& sp, float limit)
{
    while (sp > 0) {
        --sp;
        uint2 e;
        if (sp < QUAD_STACK)
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y)
                    : "r"(smBase + uint32_t(sp) * uint32_t(QUAD_STRIDE * sizeof(uint2))) : "memory");
        else
            e = spill[sp - QUAD_STACK];
        if (__uint_as_float(e.y) <= limit)
            return int32_t(e.x);
    }
    return EMPTY_CHILD;
}

/* Trace, quad form (static scenes without media: F == 0). Four neighbouring lanes walk the tree for ONE query:
 *
 *   inner node   lane j of the quad fetches child j of the node (two 128-bit loads from node4c: the four lanes read
 *                the node's 128-byte line once per load, where the per-lane walk reads it seven times, one row per
 *                load) and runs the one slab test; the four keys are exchanged by three shuffles, every lane ranks
 *                its child among the hit ones, pushes it at its place of the quad's stack (farthest first) or, if
 *                it is the nearest, hands it to the quad as the next node;
 *   leaf         lane j tests triangle j of the leaf's run (a leaf holds 1-4 triangles; the per-lane walk loops over
 *                them at a third of the lanes); the results are taken in slot order exactly as that loop would, a later
 *                triangle only under the distance an earlier one has set. Spheres are tested by all four lanes alike.
 *
 * The state of a walk (ray, best hit, current node, stack pointer) is replicated in the quad's four lanes and every
 * branch on it is uniform across the quad. The arithmetic of the slab, triangle and sphere tests and the order in which
 * hits are considered are those of the per-lane walk: the answers are bit-identical (tests/test_gpu_parity.py). Queue
 * tickets, refill of idle quads inside the walk loop, node rounds vs. leaf rounds: as in waveTracePersistentKernel. */
template<int F>
__global__ void __launch_bounds__(BLOCK, WPT_QUAD_MIN_BLOCKS) waveTraceQuadKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    static_assert(!(F & (F_MEDIA | F_ANIM)), "the quad walk is for static scenes without media");
    extern __shared__ uint2 smStackAll[];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneBelow = (1u << lane) - 1u;
    const unsigned j = lane & 3u;                   /* my child / my triangle */
    const unsigned qshift = lane & ~3u;             /* first lane of my quad */
    constexpr unsigned FULL = 0xFFFFFFFFu;
    Ctx<F> ctx;
    ctx.nested = 0;
    ctx.nodes = ctx.leaves = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  /* queue lengths of this round (the sort kernel that follows counts them up) */
        for (int c = 0; c < MAT_CLASSES; c++)
            *ctrMaterial(P, c) = 0ull;
    }
    const unsigned int nCamera = (unsigned int)*ctrTraceCamera(P);
    const unsigned int nQueued = nCamera + (unsigned int)*ctrTraceOther(P);
    if (blockIdx.x == 0 && threadIdx.x == 0 && nQueued != 0u)
        atomicAdd(P.counters + CTR_RAYS, (unsigned long long)nQueued);
    const bool anyHitScene = WPT_ANY_HIT && S.nHot == 0u;      /* see waveTracePersistentKernel */
    unsigned int entry = 0;
    bool active = false;
    Ray ray;
    ray.o = f3(0.0f); ray.d = f3(0.0f, 0.0f, 1.0f); ray.time = 0.0f;
    RayHelper h = makeRayHelper(ray);
    uint32_t rows = 0;
    WideBest best;
    best.hit.a = MAXVAL;
    best.hit.ref = NO_HIT;
    best.Ts = best.det = 0.0f;
    int32_t cur = EMPTY_CHILD;
    const uint32_t smBase = uint32_t(__cvta_generic_to_shared(smStackAll + (threadIdx.x >> 2)));
    uint2 spill[QUAD_SPILL];
    int sp = 0;
    const float amin = P.minHit;

    const unsigned int perWarp = nQueued / (gridDim.x * (BLOCK / 32u));
    const unsigned int CHUNK = perWarp >= 512u ? 64u : perWarp >= 128u ? 32u : perWarp >= 32u ? 16u : 8u;    /* queries per ticket */
    unsigned int warpNext = 0, warpEnd = 0;
    bool drained = false;
    unsigned exhaustedMask = 0u;
    for (;;) {
        const unsigned busyMask = __ballot_sync(FULL, active);
        const unsigned wantMask = ~busyMask & ~exhaustedMask;
        unsigned activeMask = busyMask;
        const int refillMin = (warpNext < warpEnd && warpNext < nCamera) ? 32 : WPT_QUAD_REFILL_MIN;
        if (wantMask != 0u && (__popc(wantMask) >= refillMin || busyMask == 0u)) {
            if (warpNext >= warpEnd && !drained) {
                unsigned int base = 0;
                if (lane == 0)
                    base = (unsigned int)atomicAdd(ctrTicket(P), (unsigned long long)CHUNK);
                base = __shfl_sync(FULL, base, 0);
                warpNext = base;
                warpEnd = min(base + CHUNK, nQueued);
                if (base >= nQueued) {
                    drained = true;
                    warpNext = warpEnd = 0;
                }
            }
            const unsigned int avail = warpEnd - warpNext;
            const unsigned int rank = __popc(wantMask & laneBelow) >> 2;       /* idle quads before mine */
            const bool got = ((wantMask >> lane) & 1u) && rank < avail;
            if (got) {
                const unsigned int e = traceQueueEntry(P, warpNext + rank, nCamera);
                const unsigned int s = e >> 1, which = e & 1u;
                const float4 o = P.W.rayO[2 * s], d = (which ? P.W.shD[2 * s] : P.W.rayD[2 * s]);
                ray.o = xyz(o); ray.time = o.w; ray.d = xyz(d);
                h = makeRayHelper(ray);
                rows = octantRows(ray.d);
                entry = e;
                active = true;
                best.hit.a = MAXVAL;
                best.hit.ref = NO_HIT;
                sp = 0;
                cur = S.root4;
            }
            const unsigned gotMask = __ballot_sync(FULL, got);
            activeMask |= gotMask;
            if (drained)
                exhaustedMask |= wantMask & ~gotMask;
            warpNext += min((unsigned int)__popc(wantMask) >> 2, avail);
        }
        if (exhaustedMask == FULL)
            break;
        const bool atLeaf = active && cur < 0;
        const bool atInner = active && cur >= 0 && cur != EMPTY_CHILD;
        const unsigned leafMask = __ballot_sync(FULL, atLeaf);
        const unsigned innerMask = activeMask & ~leafMask;
        const bool leafRound = (__popc(leafMask) >= WPT_QUAD_LEAF_MIN) || innerMask == 0u;

        /* Both kinds of round are entered by the whole warp (leafRound is warp-uniform) and every shuffle / ballot below
         * runs with all 32 lanes converged: quads that have nothing to do in the round take part with a key of "missed" /
         * "not found". (Sub-warp masks would do, but nvcc guards every *_sync with a run-time mask by MATCH / REDUX / a
         * divergence check.) */
        if (!leafRound) {
            /* my child's record */
            uint32_t key = 0xFFFFFFFFu;
            int32_t child = EMPTY_CHILD;
            if (atInner) {
                const float4* n = S.node4c + size_t(cur) * 8 + 2u * j;
                const float4 ra = __ldg(n), rb = __ldg(n + 1);
                child = __float_as_int(rb.z);
                const bool sx = rows & 1u, sy = rows & 2u, sz = rows & 4u;
                key = childKeyNearFar(sx ? ra.w : ra.x, sx ? ra.x : ra.w, sy ? rb.x : ra.y, sy ? ra.y : rb.x,
                        sz ? rb.y : ra.z, sz ? ra.z : rb.y, ray.o, h.inv, amin, best.hit.a, j);
                if (child == EMPTY_CHILD)
                    key = 0xFFFFFFFFu;
            }
            const uint32_t k1 = __shfl_xor_sync(FULL, key, 1), k2 = __shfl_xor_sync(FULL, key, 2), k3 = __shfl_xor_sync(FULL, key, 3);
            const bool hit = key != 0xFFFFFFFFu;
            const int nHit = int(hit) + int(k1 != 0xFFFFFFFFu) + int(k2 != 0xFFFFFFFFu) + int(k3 != 0xFFFFFFFFu);
            /* keys of hit children differ in their low bits: `rank` numbers them by entry distance, 0 = nearest */
            const int rank = int(k1 < key) + int(k2 < key) + int(k3 < key);
            const uint32_t nearest = min(min(key, k1), min(k2, k3));
            const int32_t nearestChild = __shfl_sync(FULL, child, int(qshift + (nearest & 3u)));
            const int newSp = sp + (nHit > 0 ? nHit - 1 : 0);
            const bool deep = newSp > QUAD_STACK;       /* (rare) the push reaches into the local-memory part of the stack */
            if (__any_sync(FULL, deep)) {
                /* every lane of a quad keeps a copy of the local-memory part: the entries are handed round */
                for (int r = 3; r >= 1; r--) {
                    const unsigned m = (__ballot_sync(FULL, hit && rank == r) >> qshift) & 0xFu;
                    const int src = int(qshift) + (m ? __ffs(int(m)) - 1 : 0);
                    const uint32_t c = __shfl_sync(FULL, uint32_t(child), src), t = __shfl_sync(FULL, key & ~3u, src);
                    if (deep && r <= nHit - 1) {
                        const int pos = sp + (nHit - 1 - r);
                        if (pos < QUAD_STACK) {
                            if (j == 0u)
                                quadStackStore(smBase, pos, c, t);
                        } else {
                            spill[pos - QUAD_STACK] = make_uint2(c, t);
                        }
                    }
                }
            }
            if (atInner) {
                if (nHit == 0) {
                    cur = quadStackPopNext(smBase, spill, sp, best.hit.a);
                } else {
                    if (!deep && hit && rank >= 1)      /* farthest first: the nearest of the pushed ones ends up on top */
                        quadStackStore(smBase, sp + (nHit - 1 - rank), uint32_t(child), key & ~3u);
                    sp = newSp;
                    cur = nearestChild;
                }
            }
        } else {
            const uint32_t ref = uint32_t(~cur);
            const uint32_t type = ref >> REF_TYPE_SHIFT;
            const bool triLeaf = atLeaf && (type == PRIM_TRIANGLE_RUN || type == PRIM_TRIANGLE);
            const bool run = (type == PRIM_TRIANGLE_RUN);
            const uint32_t first = run ? (ref & RUN_SLOT_MASK) : (ref & REF_SLOT_MASK);
            const uint32_t cnt = run ? ((ref >> RUN_COUNT_SHIFT) & 3u) + 1u : 1u;
            float ta = 0.0f, Ts = 0.0f, da = 0.0f;
            bool found = false;
            if (triLeaf && j < cnt) {
                const float4* tp = S.tri + size_t(first + j) * 3;
                const float4 t0 = __ldg(tp + 0), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                float U, V, W, det;
                found = triangleTest(xyz(t0), xyz(t1), xyz(t2), ray, h, amin, best.hit.a, true, ta, U, V, W, det, Ts, da);
            }
            const unsigned foundAll = __ballot_sync(FULL, found);
            const unsigned foundBits = (foundAll >> qshift) & 0xFu;
#pragma unroll
            for (int i = 0; i < 4; i++) {       /* in slot order */
                if (foundAll & (0x11111111u << i)) {    /* (warp-uniform) some quad's triangle i was hit */
                    const int src = int(qshift) + i;
                    const float ia = __shfl_sync(FULL, ta, src), iTs = __shfl_sync(FULL, Ts, src), ida = __shfl_sync(FULL, da, src);
                    /* (the test ran under the distance the round started with: redo its upper bound, hitable_triangle.hpp:263-264) */
                    if (((foundBits >> i) & 1u) && !(iTs > best.hit.a * ida))
                        considerHit(best, (PRIM_TRIANGLE << REF_TYPE_SHIFT) | (first + uint32_t(i)), ia, iTs, ida);
                }
            }
            if (atLeaf) {
                if (!triLeaf)       /* a sphere: all four lanes alike */
                    wideLeaf<F>(S, ref, ray, h, amin, P.minHit, ctx, best);
                cur = (anyHitScene && (entry & 1u) && best.hit.ref != NO_HIT) ? EMPTY_CHILD : quadStackPopNext(smBase, spill, sp, best.hit.a);
            }
        }
#if WPT_PREFETCH_LEAF
        if (S.prefetchLeaves && cur < 0 && ((!leafRound && atInner) || (leafRound && atLeaf)))
            prefetchLeaf(S, uint32_t(~cur));
#endif
        if (active && cur == EMPTY_CHILD) {
            if (j == 0u)
                reinterpret_cast<uint2*>(P.W.hit)[entry] = make_uint2(__float_as_uint(best.hit.a), best.hit.ref);
            active = false;
        }
    }
}

/* Drain: the last rounds of a frame hold few live slots but cost a full round each (five launches that scan or
 * stride over all slots, ~70 us). Once the work items have run out and few slots are alive, this kernel finishes
 * every one of them in a thread of its own: the slot's pending queries, the rest of its path, the remaining samples
 * of its item and whatever items are left in its window's block - the same functions in the same order with the
 * same Prng draws as the rounds would run, so the frame does not depend on when the host switches over. */
template<int F>
__global__ void __launch_bounds__(BLOCK, 4) waveDrainKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P)
{
    extern __shared__ uint2 smStackAll[];
    uint2* smStack = smStackAll + threadIdx.x;
    const unsigned int slot = blockIdx.x * BLOCK + threadIdx.x;
    unsigned int nSamples = 0, nRays = 0, retired = 0;
    Ctx<F> ctx;
    ctx.nested = 0;
    ctx.nodes = ctx.leaves = 0;
    int state = ST_EXHAUSTED;
    if (slot < P.W.nSlots)
        state = codeState((unsigned int)P.W.code[slot]);
    if (state == ST_MAIN || state == ST_TAIL || state == ST_NEW) {
        PathState ps;
        Hit hitB, hitS;
        uint4 forked = make_uint4(0u, 0u, 0u, 0u);
        if (state == ST_NEW) {
            loadItem<F>(P, slot, false, ps, ctx);
        } else {
            loadAnswered<F, true>(P, slot, ps, ctx, hitB, hitS);    /* (the answers are not in yet: traced below) */
            if ((F & F_MEDIA) && ps.shadow != SHADOW_NONE)
                forked = P.W.shPrng[slot];
        }
        uint2* window = P.W.window + slot / WAVE_WINDOW;
        for (;;) {
            if (ps.state == ST_MAIN || ps.state == ST_TAIL) {
                animationCacheInit<F>(S, ctx, ps.walkRay.time);
                if (ps.shadow != SHADOW_NONE) {
                    Ray ray = ps.walkRay;
                    ray.d = ps.shDir;
                    const Prng mine = ctx.prng;
                    if (F & F_MEDIA) {
                        ctx.prng.s0 = forked.x; ctx.prng.s1 = forked.y; ctx.prng.s2 = forked.z; ctx.prng.s3 = forked.w;
                    }
                    hitS = closestHit<F>(S, P, ray, makeRayHelper(ray), ctx, smStack);
                    ctx.prng = mine;
                    nRays++;
                }
                if (ps.state == ST_MAIN) {
                    const RayHelper h = makeRayHelper(ps.walkRay);
                    hitB = closestHit<F>(S, P, ps.walkRay, h, ctx, smStack);
                    nRays++;
                    if (ps.shadow != SHADOW_NONE)
                        consumeShadow<F>(S, P, ps, ctx, hitS);
                    if (hitB.ref == NO_HIT || !(ps.pathComponent + 1 < P.maxPathComponents)) {
                        waveEndPath<F>(S, P, ps, hitB);
                    } else {
                        waveShade<F, -1>(S, P, ps, ctx, hitB, h);
                        if ((F & F_MEDIA) && ps.shadow != SHADOW_NONE)
                            forked = forkPrng(ctx.prng);
                    }
                } else {
                    consumeShadow<F>(S, P, ps, ctx, hitS);
                    ps.state = ST_NEW;
                }
            }
            if (ps.state == ST_NEW) {
                if (startSample<F>(S, P, ps, ctx)) {
                    nSamples++;
                    continue;
                }
                finishItem(P, ps);
                /* another item of this window's block, if any is left (the global counter has run out) */
                const uint2 block = *window;
                unsigned int item = 0xFFFFFFFFu;
                if (block.x != 0xFFFFFFFFu && block.x < block.y) {
                    item = atomicAdd(&window->x, 1u);
                    if (item >= block.y)
                        item = 0xFFFFFFFFu;
                }
                if (item == 0xFFFFFFFFu)
                    break;
                const WorkItem wi = decodeItem(P, item);
                if (!wi.valid)
                    continue;       /* ps.state stays ST_NEW with no samples left... */
                beginItem<F>(ps, ctx, wi);
            }
        }
        P.W.code[slot] = uint8_t(ST_EXHAUSTED);
        P.W.ctl[slot] = makeCtl(ST_EXHAUSTED, SHADOW_NONE, 0u);
        retired = 1;
    }
    const unsigned int ns = __reduce_add_sync(0xFFFFFFFFu, nSamples);
    const unsigned int nr = __reduce_add_sync(0xFFFFFFFFu, nRays);
    const unsigned int nq = __reduce_add_sync(0xFFFFFFFFu, ctx.nested);
    const unsigned int nd = __reduce_add_sync(0xFFFFFFFFu, retired);
    if ((threadIdx.x & 31u) == 0u) {
        if (ns) atomicAdd(P.counters + CTR_SAMPLES, (unsigned long long)ns);
        if (nr) atomicAdd(P.counters + CTR_RAYS, (unsigned long long)nr);
        if (nq) atomicAdd(P.counters + CTR_NESTED, (unsigned long long)nq);
        if (nd) atomicAdd(P.counters + CTR_DONE, (unsigned long long)nd);
    }
}

/* ---- Trace, pooled form --------------------------------------------------------------------------------
 *
 * The persistent kernel above ties a query to a lane for the whole walk; in every iteration a lane either wants
 * an inner-node step or a leaf step, and the warp can only run one of the two: measured on the 1M-triangle scene,
 * node code runs with 18-21 of 32 lanes, leaf code with 7-13 (profiles/r2a_*).
 *
 * Here a WARP owns a pool of POOL_RAYS = 64 queries whose walk state lives in shared memory, and three FIFO queues
 * of pool indices: queries whose next step is an inner node, queries whose next step is a leaf primitive, and free
 * entries. Every iteration takes up to 32 entries from ONE of the two ready queues, so all lanes run the same
 * code with (nearly) all lanes busy; a step ends by putting the query into the queue of its next step. The query
 * that was just queued is not needed again before the others in front of it are done, so its next node / its
 * leaf's triangles are prefetched while it waits.
 *
 * Per query in shared memory (64 B + stack):
 *   A  origin xyz, distance of the best hit          B  1/direction xyz, next node or leaf (encoded child)
 *   C  shear constants Sx Sy Sz, kx|ky<<2|kz<<4 | octant rows << 8 | stack size << 16
 *   D  best hit's leaf reference, its T and |det| (near-tie rule), trace queue entry
 *   stack  POOL_STACK entries (encoded child, entry distance); deeper entries go to a spill area in HBM
 * Scenes with media or animated hitables (Prng draws and animation caches inside the walk) use the persistent
 * kernel. Same tests, same answers as traverseWide (see considerHit); WPT_TRAVERSAL_POOL queries run through it
 * (tests/test_gpu_parity.py::test_pooled_trace_kernel_answers_like_the_wide_walk). Renders use it only with
 * WPT_WAVE_POOL=1: see wavePool() for the measurement that keeps the per-lane kernel the default. */
#ifndef WPT_POOL_STACK
#define WPT_POOL_STACK 12
#endif
constexpr int POOL_RAYS = 64;
constexpr int POOL_STACK = WPT_POOL_STACK;
constexpr int POOL_WARPS = BLOCK / 32;
#ifndef WPT_POOL_MIN_BLOCKS
#define WPT_POOL_MIN_BLOCKS 5
#endif
#ifndef WPT_POOL_PREFETCH
#define WPT_POOL_PREFETCH 1
#endif

struct PoolShared {
    float4 A[POOL_WARPS][POOL_RAYS];
    float4 B[POOL_WARPS][POOL_RAYS];
    float4 C[POOL_WARPS][POOL_RAYS];
    float4 D[POOL_WARPS][POOL_RAYS];
    uint2 stack[POOL_WARPS][POOL_STACK][POOL_RAYS];
    uint8_t qNode[POOL_WARPS][POOL_RAYS], qLeaf[POOL_WARPS][POOL_RAYS], qFree[POOL_WARPS][POOL_RAYS];
};

struct PoolWalk {       /* what a lane holds of its query during one step */
    uint2* stack;       /* this query's column of the warp's stack array, stride POOL_RAYS */
    uint2* spill;       /* this query's spill entries in HBM */
    int sp;
};

WPT_DEV void poolPush(PoolWalk& w, uint32_t code, float tmin)
{
    const uint2 e = make_uint2(code, __float_as_uint(tmin));
    if (w.sp < POOL_STACK)
        w.stack[w.sp * POOL_RAYS] = e;
    else
        w.spill[w.sp - POOL_STACK] = e;
    w.sp++;
}

WPT_DEV int32_t poolPopNext(PoolWalk& w, float limit)
{
    while (w.sp > 0) {
        --w.sp;
        const uint2 e = (w.sp < POOL_STACK ? w.stack[w.sp * POOL_RAYS] : w.spill[w.sp - POOL_STACK]);
        if (__uint_as_float(e.y) <= limit)
            return int32_t(e.x);
    }
    return EMPTY_CHILD;
}

enum { POOL_TO_NONE = 0, POOL_TO_NODE = 1, POOL_TO_LEAF = 2, POOL_TO_FREE = 3 };

__global__ void __launch_bounds__(BLOCK, WPT_POOL_MIN_BLOCKS) waveTracePoolKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P,
        uint2* spillArea, unsigned int spillPerRay)
{
    __shared__ PoolShared sm;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr unsigned QMASK = POOL_RAYS - 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  /* queue lengths of this round (the sort kernel that follows counts them up) */
        for (int c = 0; c < MAT_CLASSES; c++)
            *ctrMaterial(P, c) = 0ull;
    }
    float4* const sA = sm.A[warp];
    float4* const sB = sm.B[warp];
    float4* const sC = sm.C[warp];
    float4* const sD = sm.D[warp];
    uint8_t* const qNode = sm.qNode[warp];
    uint8_t* const qLeaf = sm.qLeaf[warp];
    uint8_t* const qFree = sm.qFree[warp];
    uint2* const spillWarp = spillArea + size_t(blockIdx.x * POOL_WARPS + warp) * POOL_RAYS * spillPerRay;

    const unsigned int nCamera = (unsigned int)*ctrTraceCamera(P);
    const unsigned int nQueued = nCamera + (unsigned int)*ctrTraceOther(P);
    const float amin = P.minHit;
    const bool rootIsNode = (S.root4 >= 0);     /* a scene of one hitable has a leaf for a root */
    /* queue state (warp-uniform): FIFO rings of POOL_RAYS entries */
    unsigned int nodeHead = 0, nodeCnt = 0, leafHead = 0, leafCnt = 0, freeHead = 0, freeCnt = POOL_RAYS;
    qFree[lane] = uint8_t(lane);
    qFree[lane + 32] = uint8_t(lane + 32);
    __syncwarp();

    const unsigned int perWarp = nQueued / (gridDim.x * POOL_WARPS);
    const unsigned int CHUNK = perWarp >= 2048u ? 256u : perWarp >= 512u ? 128u : perWarp >= 128u ? 64u : 32u;
    unsigned int warpNext = 0, warpEnd = 0, traced = 0;
    bool drained = false;
    /* (the bound only guards against a walk that never ends: every pass ends a step of at least one query) */
    for (unsigned long long pass = 0; pass < (1ull << 36); pass++) {
        /* ---- refill: new queries into free pool entries, 32 at a time (or whatever is needed to have a full pass) ---- */
        if (!drained && (freeCnt >= 32u || nodeCnt + leafCnt < 32u) && freeCnt > 0u) {
            if (warpNext >= warpEnd) {
                unsigned int base = 0;
                if (lane == 0)
                    base = (unsigned int)atomicAdd(ctrTicket(P), (unsigned long long)CHUNK);
                base = __shfl_sync(FULL, base, 0);
                warpNext = base;
                warpEnd = min(base + CHUNK, nQueued);
                if (base >= nQueued) {
                    drained = true;
                    warpNext = warpEnd = 0;
                }
            }
            const unsigned int take = min(min(32u, freeCnt), warpEnd - warpNext);
            if (lane < take) {
                const unsigned int r = qFree[(freeHead + lane) & QMASK];
                const unsigned int e = traceQueueEntry(P, warpNext + lane, nCamera);
                const unsigned int s = e >> 1, which = e & 1u;
                const float4 o = P.W.rayO[2 * s], d = (which ? P.W.shD[2 * s] : P.W.rayD[2 * s]);
                Ray ray;
                ray.o = xyz(o); ray.time = o.w; ray.d = xyz(d);
                const RayHelper h = makeRayHelper(ray);
                sA[r] = f4(ray.o, MAXVAL);
                sB[r] = f4(h.inv, __int_as_float(S.root4));
                sC[r] = f4(h.Sx, h.Sy, h.Sz, __uint_as_float(uint32_t(h.kx) | (uint32_t(h.ky) << 2) | (uint32_t(h.kz) << 4) | (octantRows(ray.d) << 8)));
                sD[r] = f4(__uint_as_float(NO_HIT), 0.0f, 0.0f, __uint_as_float(e));
                if (rootIsNode)
                    qNode[(nodeHead + nodeCnt + lane) & QMASK] = uint8_t(r);
                else
                    qLeaf[(leafHead + leafCnt + lane) & QMASK] = uint8_t(r);
            }
            freeHead += take; freeCnt -= take; warpNext += take;
            if (rootIsNode)
                nodeCnt += take;
            else
                leafCnt += take;
            __syncwarp();
        }
        if (nodeCnt + leafCnt == 0u) {
            if (drained)
                break;
            continue;
        }
        /* ---- one pass: inner nodes or leaf primitives, whichever fills the warp (else the longer queue) ---- */
        const bool nodePass = (nodeCnt >= 32u) || (leafCnt < 32u && nodeCnt >= leafCnt);
        int dest = POOL_TO_NONE;
        unsigned int r = 0;
        if (nodePass) {
            const unsigned int n = min(nodeCnt, 32u);
            if (lane < n) {
                r = qNode[(nodeHead + lane) & QMASK];
                const float4 a = sA[r], b = sB[r];
                const uint32_t cw = __float_as_uint(sC[r].w);
                PoolWalk w;
                w.stack = &sm.stack[warp][0][r];
                w.spill = spillWarp + size_t(r) * spillPerRay;
                w.sp = int(cw >> 16);
                int32_t cur = __float_as_int(b.w);
                const NodeKeys nk = nodeKeys(S, cur, xyz(a), xyz(b), amin, a.w, (cw >> 8) & 7u);
                if (nk.k0 != 0xFFFFFFFFu) {
                    if (nk.k3 != 0xFFFFFFFFu)
                        poolPush(w, uint32_t(sel4i(nk.ch, nk.k3 & 3u)), __uint_as_float(nk.k3 & ~3u));
                    if (nk.k2 != 0xFFFFFFFFu)
                        poolPush(w, uint32_t(sel4i(nk.ch, nk.k2 & 3u)), __uint_as_float(nk.k2 & ~3u));
                    if (nk.k1 != 0xFFFFFFFFu)
                        poolPush(w, uint32_t(sel4i(nk.ch, nk.k1 & 3u)), __uint_as_float(nk.k1 & ~3u));
                    cur = sel4i(nk.ch, nk.k0 & 3u);
                } else {
                    cur = poolPopNext(w, a.w);
                }
                sB[r].w = __int_as_float(cur);
                sC[r].w = __uint_as_float((cw & 0xFFFFu) | (uint32_t(w.sp) << 16));
                dest = (cur == EMPTY_CHILD ? POOL_TO_FREE : cur < 0 ? POOL_TO_LEAF : POOL_TO_NODE);
#if WPT_POOL_PREFETCH
                if (dest == POOL_TO_NODE)
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(S.node4 + size_t(cur) * 8));
                else if (dest == POOL_TO_LEAF)
                    prefetchLeaf(S, uint32_t(~cur));
#endif
            }
            nodeHead += n; nodeCnt -= n;
        } else {
            const unsigned int n = min(leafCnt, 32u);
            if (lane < n) {
                r = qLeaf[(leafHead + lane) & QMASK];
                const float4 a = sA[r], b = sB[r], c = sC[r], dd = sD[r];
                const uint32_t cw = __float_as_uint(c.w);
                PoolWalk w;
                w.stack = &sm.stack[warp][0][r];
                w.spill = spillWarp + size_t(r) * spillPerRay;
                w.sp = int(cw >> 16);
                Ray ray;
                ray.o = xyz(a); ray.d = f3(0.0f); ray.time = 0.0f;     /* the triangle test needs the origin and the shear constants only */
                RayHelper h;
                h.inv = xyz(b);
                h.kx = int(cw & 3u); h.ky = int((cw >> 2) & 3u); h.kz = int((cw >> 4) & 3u);
                h.Sx = c.x; h.Sy = c.y; h.Sz = c.z;
                WideBest best;
                best.hit.a = a.w;
                best.hit.ref = __float_as_uint(dd.x);
                best.Ts = dd.y;
                best.det = dd.z;
                /* one primitive of the leaf; what is left of a run of triangles stays this query's leaf */
                const uint32_t ref = uint32_t(~__float_as_int(b.w));
                uint32_t tref = ref, rest = NO_HIT;
                if ((ref >> REF_TYPE_SHIFT) == PRIM_TRIANGLE_RUN) {
                    const uint32_t first = ref & RUN_SLOT_MASK;
                    const uint32_t more = (ref >> RUN_COUNT_SHIFT) & 3u;
                    tref = (PRIM_TRIANGLE << REF_TYPE_SHIFT) | first;
                    if (more != 0u)
                        rest = (PRIM_TRIANGLE_RUN << REF_TYPE_SHIFT) | ((more - 1u) << RUN_COUNT_SHIFT) | (first + 1u);
                }
                bool found = false;
                float ta = 0.0f, Ts = 0.0f, da = 0.0f;
                if ((tref >> REF_TYPE_SHIFT) == PRIM_TRIANGLE) {
                    const float4* tp = S.tri + size_t(tref & REF_SLOT_MASK) * 3;
                    const float4 t0 = __ldg(tp + 0), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                    float U, V, W, det;
                    found = triangleTest(xyz(t0), xyz(t1), xyz(t2), ray, h, amin, best.hit.a, true, ta, U, V, W, det, Ts, da);
                } else if ((tref >> REF_TYPE_SHIFT) == PRIM_SPHERE) {
                    /* the sphere test needs the direction: it is still where the query came from */
                    const uint32_t e = __float_as_uint(dd.w);
                    ray.d = xyz((e & 1u) ? P.W.shD[2 * (e >> 1)] : P.W.rayD[2 * (e >> 1)]);
                    const float4* p = S.sph + size_t(tref & REF_SLOT_MASK) * 3;
                    const float4 s0 = __ldg(p);
                    found = sphereTest(xyz(s0), s0.w, ray, amin, best.hit.a, ta);
                }
                if (found) {
                    considerHit(best, tref, ta, Ts, da);
                    sA[r].w = best.hit.a;
                    sD[r] = f4(__uint_as_float(best.hit.ref), best.Ts, best.det, dd.w);
                }
                int32_t cur = (rest != NO_HIT ? int32_t(~rest) : poolPopNext(w, best.hit.a));
                sB[r].w = __int_as_float(cur);
                sC[r].w = __uint_as_float((cw & 0xFFFFu) | (uint32_t(w.sp) << 16));
                dest = (cur == EMPTY_CHILD ? POOL_TO_FREE : cur < 0 ? POOL_TO_LEAF : POOL_TO_NODE);
#if WPT_POOL_PREFETCH
                if (dest == POOL_TO_NODE)
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(S.node4 + size_t(cur) * 8));
                else if (dest == POOL_TO_LEAF && rest == NO_HIT)
                    prefetchLeaf(S, uint32_t(~cur));
#endif
            }
            leafHead += n; leafCnt -= n;
        }
        /* ---- the query's next step decides its queue; finished walks are answered and their entry freed ---- */
        if (dest == POOL_TO_FREE) {
            const float4 a = sA[r], dd = sD[r];
            reinterpret_cast<uint2*>(P.W.hit)[__float_as_uint(dd.w)] = make_uint2(__float_as_uint(a.w), __float_as_uint(dd.x));
            traced++;
        }
        const unsigned mN = __ballot_sync(FULL, dest == POOL_TO_NODE);
        const unsigned mL = __ballot_sync(FULL, dest == POOL_TO_LEAF);
        const unsigned mF = __ballot_sync(FULL, dest == POOL_TO_FREE);
        if (dest == POOL_TO_NODE)
            qNode[(nodeHead + nodeCnt + __popc(mN & below)) & QMASK] = uint8_t(r);
        else if (dest == POOL_TO_LEAF)
            qLeaf[(leafHead + leafCnt + __popc(mL & below)) & QMASK] = uint8_t(r);
        else if (dest == POOL_TO_FREE)
            qFree[(freeHead + freeCnt + __popc(mF & below)) & QMASK] = uint8_t(r);
        nodeCnt += __popc(mN); leafCnt += __popc(mL); freeCnt += __popc(mF);
        __syncwarp();
    }
    const unsigned int n = __reduce_add_sync(FULL, traced);
    if (lane == 0 && n)
        atomicAdd(P.counters + CTR_RAYS, (unsigned long long)n);
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

static size_t stackBytes();

/* Closest-hit queries through the pooled trace kernel (WPT_TRAVERSAL_POOL): the rays are laid out like wavefront
 * slots, traced by waveTracePoolKernel, and the answers converted to ids. This is how the bit-exact parity
 * tests reach the kernel that renders. */
__global__ void __launch_bounds__(BLOCK) poolQuerySetupKernel(const __grid_constant__ RenderParams P,
        const float* origins, const float* directions, unsigned long long nRays, float time)
{
    const unsigned long long i = blockIdx.x * (unsigned long long)BLOCK + threadIdx.x;
    if (i == 0) {
        *ctrTicket(P) = 0ull;
        *ctrTraceCamera(P) = nRays;
        *ctrTraceOther(P) = 0ull;
    }
    if (i >= nRays)
        return;
    Ray ray;
    if (origins) {
        ray.o = f3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        ray.d = f3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
        ray.time = time;
    } else {
        Prng prng = prngInit((unsigned int)i);
        const unsigned int x = (unsigned int)(i % P.width), y = (unsigned int)(i / P.width);
        ray = cameraRay(P, (float(x) + 0.5f) / float(P.width), (float(y) + 0.5f) / float(P.height), prng, false);
    }
    P.W.rayO[2 * i] = f4(ray.o, ray.time);
    P.W.rayD[2 * i] = f4(ray.d, 0.0f);
    P.W.traceQueue[i] = (unsigned int)(i << 1);
}

__global__ void __launch_bounds__(BLOCK) poolQueryFinishKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P,
        unsigned long long nRays, int32_t* ids, float* dist)
{
    const unsigned long long i = blockIdx.x * (unsigned long long)BLOCK + threadIdx.x;
    if (i >= nRays)
        return;
    const uint4 hw = P.W.hit[i];
    int32_t id = -1;
    if (hw.y != NO_HIT) {
        const uint32_t type = hw.y >> REF_TYPE_SHIFT, slot = hw.y & REF_SLOT_MASK;
        id = int32_t(type == PRIM_TRIANGLE ? S.triGlobal[slot] : type == PRIM_SPHERE ? S.sphGlobal[slot] : S.medGlobal[slot]);
    }
    ids[i] = id;
    dist[i] = (hw.y != NO_HIT ? __uint_as_float(hw.x) : 0.0f);
}

cudaError_t launchPoolQuery(const DScene& S, const RenderParams& P, int numSMs, int mode, const float* origins, const float* directions,
        unsigned long long nRays, float time, int32_t* ids, float* dist, cudaStream_t stream)
{
    const unsigned int grid = (unsigned int)((nRays + BLOCK - 1) / BLOCK);
    if (grid == 0)
        return cudaSuccess;
    const bool pooled = (mode == 1);
    poolQuerySetupKernel<<<grid, BLOCK, 0, stream>>>(P, origins, directions, nRays, time);
    if (mode == 2) {
        int perSM = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, waveTraceQuadKernel<0>, BLOCK, quadStackBytes());
        if (e != cudaSuccess)
            return e;
        unsigned int pgrid = (unsigned int)(numSMs * (perSM < 1 ? 1 : perSM));
        const unsigned int needed = (unsigned int)((nRays + BLOCK / 4 - 1) / (BLOCK / 4));     /* a block walks BLOCK / 4 queries at a time */
        if (pgrid > needed)
            pgrid = needed;
        waveTraceQuadKernel<0><<<pgrid, BLOCK, quadStackBytes(), stream>>>(S, P);
    } else if (!pooled) {
        int perSM = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, waveTracePersistentKernel<0>, BLOCK, stackBytes());
        if (e != cudaSuccess)
            return e;
        unsigned int pgrid = (unsigned int)(numSMs * (perSM < 1 ? 1 : perSM));
        if (pgrid > grid)
            pgrid = grid;
        waveTracePersistentKernel<0><<<pgrid, BLOCK, stackBytes(), stream>>>(S, P);
    } else
    waveTracePoolKernel<<<poolGridBlocks(numSMs), BLOCK, 0, stream>>>(S, P, P.poolSpill, P.poolSpillPerRay);
    poolQueryFinishKernel<<<grid, BLOCK, 0, stream>>>(S, P, nRays, ids, dist);
    return cudaGetLastError();
}

/* Animation index of a hitable (Hitable::animationIndex(), hitable.hpp:131): -1 = not animated. */
WPT_DEV int hitAnimationIndex(const DScene& S, uint32_t ref)
{
    const uint32_t type = ref >> REF_TYPE_SHIFT, slot = ref & REF_SLOT_MASK;
    uint32_t anim = 0;
    if (type == PRIM_TRIANGLE)
        anim = __float_as_uint(__ldg(S.tri + size_t(slot) * 3 + 2).w);
    else if (type == PRIM_SPHERE)
        anim = __float_as_uint(__ldg(S.sph + size_t(slot) * 3 + 2).y);
    return int(anim) - 1;
}

/* inverse(Transformation) (transformation.hpp:157-163) */
WPT_DEV Trs trsInverse(const Trs& T)
{
    const float3 invT = -T.t;
    const float4 invR = f4(-T.r.x, -T.r.y, -T.r.z, T.r.w);
    const float3 invS = 1.0f / T.s;
    Trs R;
    R.t = qrotate(invR, invT * invS);
    R.r = invR;
    R.s = invS;
    return R;
}

/* Camera::cameraSpaceToImageSpace (camera.hpp:205-217), times the image size */
WPT_DEV float2 cameraToPixelSpace(const RenderParams& P, const GroundTruthParams& G, float3 p)
{
    const float px = G.P00 * p.x + G.P03 * p.z, py = G.P11 * p.y + G.P13 * p.z, pw = -p.z;
    float ix = 0.5f * (px / pw) + 0.5f;
    float iy = 0.5f * (py / pw) + 0.5f;
    lensDistort(P, ix, iy);
    return make_float2(ix * float(P.width), iy * float(P.height));
}

WPT_DEV void store3(float* image, size_t pixel, float3 v)
{
    image[pixel * 3 + 0] = v.x; image[pixel * 3 + 1] = v.y; image[pixel * 3 + 2] = v.z;
}

/* getGroundTruth() (wurblpt.hpp:656-755): one thread per pixel. */
template<int F>
__global__ void __launch_bounds__(BLOCK) groundTruthKernel(const __grid_constant__ DScene S, const __grid_constant__ RenderParams P,
        const __grid_constant__ GroundTruthParams G)
{
    extern __shared__ uint2 smStackAll[];
    uint2* smStack = smStackAll + threadIdx.x;
    const size_t pixel = blockIdx.x * size_t(BLOCK) + threadIdx.x;
    if (pixel >= size_t(P.width) * P.height)
        return;
    Ctx<F> ctx;
    ctx.nested = 0;
    ctx.nodes = ctx.leaves = 0;
    ctx.prng = prngInit((unsigned int)pixel);
    const unsigned int x = (unsigned int)(pixel % P.width), y = (unsigned int)(pixel / P.width);
    const float invWidth = 1.0f / float(P.width), invHeight = 1.0f / float(P.height);
    const float2 pixelCoord = make_float2((float(x) + 0.5f) * invWidth, (float(y) + 0.5f) * invHeight);
    const Ray ray = cameraRay(P, pixelCoord.x, pixelCoord.y, ctx.prng, false);
    animationCacheInit<F>(S, ctx, ray.time);
    const RayHelper h = makeRayHelper(ray);
    const Hit hit = closestHit<F>(S, P, ray, h, ctx, smStack);

    float3 wsPos = f3(0.0f), wsGNrm = f3(0.0f), wsGTan = f3(0.0f), wsMNrm = f3(0.0f), wsMTan = f3(0.0f);
    float3 csPos = f3(0.0f), csGNrm = f3(0.0f), csGTan = f3(0.0f), csMNrm = f3(0.0f), csMTan = f3(0.0f);
    float csDepth = 0.0f, csDist = 0.0f;
    float2 txCor = make_float2(0.0f, 0.0f);
    float3 wsOP = f3(0.0f), wsON = f3(0.0f), csOP = f3(0.0f), csON = f3(0.0f);
    float2 psOP = make_float2(0.0f, 0.0f), psON = make_float2(0.0f, 0.0f);
    int matInd = -1;
    if (hit.ref != NO_HIT) {
        const HitRec hr = buildHitRecord<F>(S, hit, ray, h, ctx);
        wsPos = hr.p;
        wsGNrm = hr.n;
        wsGTan = hr.t;
        TangentSpace ts;
        if (hr.material >= 0)
            ts = tangentSpaceAt(S, S.materials + hr.material, hr);
        else
            ts = (dot(hr.t, hr.t) > EPSILON ? tangentSpace(hr.n, hr.t) : tangentSpace(hr.n));
        wsMNrm = ts.normal;
        wsMTan = ts.tangent;
        csPos = qrotate(G.invCamR[0], wsPos) + G.invCamT[0];
        csGNrm = qrotate(G.invCamR[0], wsGNrm);
        csGTan = qrotate(G.invCamR[0], wsGTan);
        csMNrm = qrotate(G.invCamR[0], wsMNrm);
        csMTan = qrotate(G.invCamR[0], wsMTan);
        csDepth = -csPos.z;
        csDist = sqrtf(dot(csPos, csPos));
        txCor = hr.tc;
        float3 wsPosPrev = wsPos, wsPosNext = wsPos;
        const int anim = hitAnimationIndex(S, hit.ref);
        if (anim >= 0 && uint32_t(anim) < S.nAnimations) {
            const wpt_animation an = S.animations[anim];
            const wpt_keyframe* kf = S.keyframes + an.first_keyframe;
            const float3 posOrig = trsApply(trsInverse(keyframesAt(kf, an.keyframe_count, ray.time)), wsPos);
            wsPosPrev = trsApply(keyframesAt(kf, an.keyframe_count, G.tPrev), posOrig);
            wsPosNext = trsApply(keyframesAt(kf, an.keyframe_count, G.tNext), posOrig);
        }
        wsOP = wsPosPrev - wsPos;
        wsON = wsPosNext - wsPos;
        const float3 csPosPrev = qrotate(G.invCamR[1], wsPosPrev) + G.invCamT[1];
        const float3 csPosNext = qrotate(G.invCamR[2], wsPosNext) + G.invCamT[2];
        csOP = csPosPrev - csPos;
        csON = csPosNext - csPos;
        if (G.bits & (WPT_GT_PIXEL_OFFSET_TO_PREV | WPT_GT_PIXEL_OFFSET_TO_NEXT)) {
            const float2 psPos = make_float2(pixelCoord.x * float(P.width), pixelCoord.y * float(P.height));
            const float2 psPrev = cameraToPixelSpace(P, G, csPosPrev);
            const float2 psNext = cameraToPixelSpace(P, G, csPosNext);
            psOP = make_float2(psPrev.x - psPos.x, psPrev.y - psPos.y);
            psON = make_float2(psNext.x - psPos.x, psNext.y - psPos.y);
        }
        matInd = hr.material;
        if (G.materialIds && matInd >= 0)
            matInd = (uint32_t(matInd) < G.nMaterialIds ? G.materialIds[matInd] : -1);
    }
    if (G.bits & WPT_GT_WORLD_POSITIONS) store3(G.aov[0], pixel, wsPos);
    if (G.bits & WPT_GT_WORLD_GEOMETRY_NORMALS) store3(G.aov[1], pixel, wsGNrm);
    if (G.bits & WPT_GT_WORLD_GEOMETRY_TANGENTS) store3(G.aov[2], pixel, wsGTan);
    if (G.bits & WPT_GT_WORLD_MATERIAL_NORMALS) store3(G.aov[3], pixel, wsMNrm);
    if (G.bits & WPT_GT_WORLD_MATERIAL_TANGENTS) store3(G.aov[4], pixel, wsMTan);
    if (G.bits & WPT_GT_CAMERA_POSITIONS) store3(G.aov[5], pixel, csPos);
    if (G.bits & WPT_GT_CAMERA_GEOMETRY_NORMALS) store3(G.aov[6], pixel, csGNrm);
    if (G.bits & WPT_GT_CAMERA_GEOMETRY_TANGENTS) store3(G.aov[7], pixel, csGTan);
    if (G.bits & WPT_GT_CAMERA_MATERIAL_NORMALS) store3(G.aov[8], pixel, csMNrm);
    if (G.bits & WPT_GT_CAMERA_MATERIAL_TANGENTS) store3(G.aov[9], pixel, csMTan);
    if (G.bits & WPT_GT_CAMERA_DEPTHS) G.aov[10][pixel] = csDepth;
    if (G.bits & WPT_GT_CAMERA_DISTANCES) G.aov[11][pixel] = csDist;
    if (G.bits & WPT_GT_TEXCOORDS) { G.aov[12][pixel * 2] = txCor.x; G.aov[12][pixel * 2 + 1] = txCor.y; }
    if (G.bits & WPT_GT_WORLD_OFFSET_TO_PREV) store3(G.aov[13], pixel, wsOP);
    if (G.bits & WPT_GT_WORLD_OFFSET_TO_NEXT) store3(G.aov[14], pixel, wsON);
    if (G.bits & WPT_GT_CAMERA_OFFSET_TO_PREV) store3(G.aov[15], pixel, csOP);
    if (G.bits & WPT_GT_CAMERA_OFFSET_TO_NEXT) store3(G.aov[16], pixel, csON);
    if (G.bits & WPT_GT_PIXEL_OFFSET_TO_PREV) { G.aov[17][pixel * 2] = psOP.x; G.aov[17][pixel * 2 + 1] = psOP.y; }
    if (G.bits & WPT_GT_PIXEL_OFFSET_TO_NEXT) { G.aov[18][pixel * 2] = psON.x; G.aov[18][pixel * 2 + 1] = psON.y; }
    if (G.bits & WPT_GT_MATERIALS) G.materials[pixel] = matInd;
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

/* The quad trace kernel (waveTraceQuadKernel) renders static scenes without media unless WPT_TRACE_QUAD=0. */
static bool waveQuadTrace(const DScene& S)
{
    const char* e = std::getenv("WPT_TRACE_QUAD");
    const bool on = e ? (e[0] != '0') : (WPT_TRACE_QUAD_DEFAULT != 0);
    return on && S.simpleWalk && S.node4c != nullptr;
}

static cudaError_t launchResolve(const RenderParams& P, cudaStream_t stream)
{
    if (P.substreams == 1)
        return cudaSuccess;
    const unsigned int pixels = P.nLocalTiles * ((1u << P.tileShift) >> P.subShift);
    resolveKernel<<<(pixels + 255) / 256, 256, 0, stream>>>(P);
    return cudaGetLastError();
}

template<int F>
static cudaError_t launchRenderT(const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream, int* launches)
{
    int perSM = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, renderKernel<F>, BLOCK, stackBytes());
    if (e != cudaSuccess)
        return e;
    if (perSM < 1)
        perSM = 1;
    /* persistent: as many CTAs as are resident at once (a multiple of the SM count), fewer only
     * when there are not enough work items to give every lane one */
    long long grid = (long long)numSMs * perSM;
    long long needed = ((long long)P.nItems + BLOCK - 1) / BLOCK;
    if (needed < grid)
        grid = needed < 1 ? 1 : needed;
    renderKernel<F><<<int(grid), BLOCK, stackBytes(), stream>>>(S, P);
    e = cudaGetLastError();
    *launches = (P.substreams == 1 ? 1 : 2);
    if (e != cudaSuccess)
        return e;
    return launchResolve(P, stream);
}

cudaError_t launchRender(int features, const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream, int* launches)
{
    if (features == 0)
        return launchRenderT<0>(S, P, numSMs, stream, launches);
    return launchRenderT<F_MEDIA | F_ANIM | F_OPL>(S, P, numSMs, stream, launches);
}

/* Wavefront driver. `hostDone` is pinned host memory for two counters, `events` two CUDA events.
 * Batches of (trace, shade) pairs are enqueued; the "slots exhausted" counter is copied after every
 * batch and inspected one batch late, so that a batch is always queued behind the running one. */
static int waveGroups()
{
    const char* e = std::getenv("WPT_WAVE_GROUPS");         /* experiment knob: 1 = a single slot group */
    return (e && e[0] == '1') ? 1 : 2;
}

static bool wavePgridFull()
{
    const char* e = std::getenv("WPT_WAVE_PGRID_FULL");    /* experiment knob: 0 = the groups' trace launches share the resident CTAs */
    return !(e && e[0] == '0');
}

static bool wavePlainTrace()
{
    const char* e = std::getenv("WPT_WAVE_PLAIN_TRACE");   /* experiment knob: one ray per thread, no refill */
    return e && e[0] == '1';
}

/* Offsets every slot array of `P.W` by `first` slots and narrows it to `count` slots. */
static RenderParams slotGroup(const RenderParams& P, uint32_t group, uint32_t first, uint32_t count)
{
    RenderParams G = P;
    G.W.group = group;
    G.W.nSlots = count;
    G.W.item += first; G.W.ctl += first; G.W.code += first;
    G.W.prng += 2 * first; G.W.acc += 2 * first; G.W.rayO += 2 * first; G.W.rayD += 2 * first;
    G.W.shD += 2 * first; G.W.pendRad += 2 * first; G.W.att += first; G.W.hit += first;
    if (G.W.opl) { G.W.opl += first; G.W.rayRi += first; G.W.shadowRi += first; }
    if (G.W.shPrng) G.W.shPrng += first;
    G.W.queue += first;             /* every class's list of the group starts `first` entries into the class's array */
    G.W.traceQueue += 2 * first;
    G.W.window += first / WAVE_WINDOW;
    if (G.W.cell) {
        G.W.cell += first;
        G.W.sortedQueue += 2 * first;
        G.W.binCount += size_t(group) * RAY_BINS;
        G.W.binBase += size_t(group) * 2 * RAY_BINS;
    }
    return G;
}

/* Resident blocks of the pooled trace kernel: its grid, and what the spill area is sized for. */
unsigned int poolGridBlocks(int numSMs)
{
    static int perSM = 0;
    if (perSM == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, waveTracePoolKernel, BLOCK, 0) != cudaSuccess || perSM < 1)
            perSM = 1;
        if (const char* ev = std::getenv("WPT_POOL_CTAS")) {    /* experiment knob: resident pool CTAs per SM */
            const int v = std::atoi(ev);
            if (v >= 1 && v < perSM)
                perSM = v;
        }
    }
    return (unsigned int)(numSMs * perSM);
}
unsigned int poolRaysPerBlock() { return POOL_WARPS * POOL_RAYS; }
unsigned int poolStackEntries() { return POOL_STACK; }
unsigned int wideStackEntries() { return SM_STACK + LOCAL_STACK; }

static bool wavePool(const DScene& S, const RenderParams& P)
{
    /* Opt-in (WPT_WAVE_POOL=1). Measured on the 1M-triangle scene (profiles/r2d_*): 20 of 32 lanes per instruction
     * instead of 15, node passes at 29 lanes, but a pass costs ~350 warp instructions (queue bookkeeping, state
     * round trips through shared memory, stack spills to HBM instead of local memory): 615 vs 679 Msamples/s. */
    const char* e = std::getenv("WPT_WAVE_POOL");
    return (e && e[0] == '1') && S.simpleWalk && P.poolSpill != nullptr;
}

/* One material kernel per class present in the scene; fixed grids (resident blocks) that stride over the queues. */
template<int F, int MC>
static void launchMaterialClass(const DScene& S, const RenderParams& R, int numSMs, unsigned int maxBlocks, cudaStream_t stream)
{
    static int perSM = 0;
    if (perSM == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, waveMaterialKernel<F, MC>, BLOCK, 0) != cudaSuccess || perSM < 1)
            perSM = 1;
    }
    unsigned int grid = (unsigned int)(numSMs * perSM);
    if (grid > maxBlocks)
        grid = maxBlocks;
    waveMaterialKernel<F, MC><<<grid, BLOCK, 0, stream>>>(S, R);
}

template<int F>
static int launchMaterial(const DScene& S, const RenderParams& R, int numSMs, unsigned int maxBlocks, cudaStream_t stream)
{
    int n = 0;
    if (S.classMask & (1u << MC_GENERIC)) { launchMaterialClass<F, MC_GENERIC>(S, R, numSMs, maxBlocks, stream); n++; }
    if (S.classMask & (1u << MC_LAMBERTIAN)) { launchMaterialClass<F, MC_LAMBERTIAN>(S, R, numSMs, maxBlocks, stream); n++; }
    if (S.classMask & (1u << MC_GGX)) { launchMaterialClass<F, MC_GGX>(S, R, numSMs, maxBlocks, stream); n++; }
    if (S.classMask & (1u << MC_GLASS)) { launchMaterialClass<F, MC_GLASS>(S, R, numSMs, maxBlocks, stream); n++; }
    if (S.classMask & (1u << MC_MODPHONG)) { launchMaterialClass<F, MC_MODPHONG>(S, R, numSMs, maxBlocks, stream); n++; }
    return n;
}

/* Wavefront driver. The slots are split into two groups that run the same (trace, logic, material)
 * rounds on two streams: while one group's kernel drains (the tail of a persistent trace launch, or
 * the last blocks of a short shade launch) the other group's kernel fills the SMs.
 * `hostDone` is pinned host memory for two counters, `events` four CUDA events (2 for the batch
 * hand-shake, 2 for fork/join with the second stream). Batches of rounds are enqueued; the "slots
 * exhausted" counter is copied after every batch and inspected one batch late, so that a batch is
 * always queued behind the running one. */
template<int F>
static cudaError_t launchWavefrontT(const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream,
        cudaStream_t stream2, unsigned long long* hostDone, cudaEvent_t* events, int* launches)
{
    const int batch = 8;
    int perSM = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, waveTracePersistentKernel<F>, BLOCK, stackBytes());
    if (e != cudaSuccess)
        return e;
    if (const char* ev = std::getenv("WPT_TRACE_CTAS")) {   /* experiment knob: resident trace CTAs per SM */
        const int v = std::atoi(ev);
        if (v >= 1 && v < perSM)
            perSM = v;
    }
    const bool persistent = (P.traversal == WPT_TRAVERSAL_WIDE) && !wavePlainTrace();
    const bool pool = wavePool(S, P);
    bool quad = false;
    int quadPerSM = 0;
    if constexpr (F == 0) {
        quad = persistent && !pool && waveQuadTrace(S);
        if (quad) {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&quadPerSM, waveTraceQuadKernel<0>, BLOCK, quadStackBytes());
            if (e != cudaSuccess)
                return e;
            if (const char* ev = std::getenv("WPT_TRACE_CTAS")) {
                const int v = std::atoi(ev);
                if (v >= 1 && v < quadPerSM)
                    quadPerSM = v;
            }
            perSM = quadPerSM;
        }
    }
    const unsigned int poolGrid = poolGridBlocks(numSMs);
    unsigned long long drainSlots = 32768;      /* live slots at which waveDrainKernel takes over (0: never); C1 +6 %,
                                                   media +2 %, larger values cost media up to 6 % (profiles/r2h_ab.txt) */
    if (const char* ev = std::getenv("WPT_DRAIN_SLOTS"))
        drainSlots = std::strtoull(ev, nullptr, 10);
    const int nGroups = (waveGroups() >= 2 && P.W.nSlots >= 4096) ? 2 : 1;
    RenderParams G[2];
    unsigned int grid[2], sgrid[2], pgrid[2];
    cudaStream_t st[2] = { stream, stream2 };
    const uint32_t half = (nGroups == 2 ? (P.W.nSlots / 2 + WAVE_WINDOW - 1) / WAVE_WINDOW * WAVE_WINDOW : P.W.nSlots);
    for (int g = 0; g < nGroups; g++) {
        const uint32_t first = (g == 0 ? 0 : half);
        const uint32_t count = (g == 0 ? half : P.W.nSlots - half);
        G[g] = slotGroup(P, uint32_t(g), first, count);
        sgrid[g] = (count + SORT_WINDOW * (SORT_BLOCK / 32u) - 1u) / (SORT_WINDOW * (SORT_BLOCK / 32u));   /* sort kernel */
        grid[g] = (count + BLOCK - 1) / BLOCK;
        /* persistent trace kernel: the resident CTAs are shared between the groups */
        pgrid[g] = (unsigned int)(numSMs * (perSM < 1 ? 1 : perSM)) / (unsigned int)(wavePgridFull() ? 1 : nGroups);
        if (pgrid[g] > grid[g])
            pgrid[g] = grid[g];
        if (pgrid[g] < 1)
            pgrid[g] = 1;
    }
    int n = 0;
    e = cudaMemsetAsync(P.W.code, 0xFF, size_t(P.W.nSlots), stream);   /* all slots idle */
    if (e != cudaSuccess)
        return e;
    if (nGroups == 2) {     /* fork */
        if ((e = cudaEventRecord(events[2], stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream2, events[2], 0)) != cudaSuccess) return e;
    }
    for (int g = 0; g < nGroups; g++)
        waveGenerateKernel<F><<<sgrid[g], BLOCK, 0, st[g]>>>(S, G[g]);      /* every slot gets an item and its first camera ray */
    n += nGroups;
    hostDone[0] = hostDone[1] = 0;
    /* every round retires at least one path component of every live slot: a bound on the rounds of a frame */
    const unsigned long long itemsPerSlot = (P.nItems + P.W.nSlots - 1) / P.W.nSlots;
    const unsigned long long samplesPerItem = (P.samples + P.substreams - 1) / P.substreams;
    const unsigned long long maxRounds = (itemsPerSlot + 1) * samplesPerItem * (P.maxPathComponents + 2ull) + 64ull;
    for (int k = 0;; k++) {
        if ((unsigned long long)k * batch > maxRounds)
            return cudaErrorLaunchFailure;      /* the slots never retired: a bug, not a long frame */
        for (int i = 0; i < batch; i++) {
            for (int g = 0; g < nGroups; g++) {
                const RenderParams& R = G[g];
                if (persistent && pool)
                    waveTracePoolKernel<<<poolGrid, BLOCK, 0, st[g]>>>(S, R,
                            P.poolSpill + size_t(g) * poolGrid * poolRaysPerBlock() * P.poolSpillPerRay, P.poolSpillPerRay);
                else if (quad) {
                    if constexpr (F == 0)
                        waveTraceQuadKernel<0><<<pgrid[g], BLOCK, quadStackBytes(), st[g]>>>(S, R);
                } else if (persistent)
                    waveTracePersistentKernel<F><<<pgrid[g], BLOCK, stackBytes(), st[g]>>>(S, R);
                else
                    waveTraceKernel<F><<<pgrid[g], BLOCK, stackBytes(), st[g]>>>(S, R);
                waveSortKernel<<<sgrid[g], SORT_BLOCK, 0, st[g]>>>(S, R);
                n += 3 + launchMaterial<F>(S, R, numSMs, grid[g], st[g]);
                waveGenerateKernel<F><<<sgrid[g], BLOCK, 0, st[g]>>>(S, R);
                if (R.W.cell) {
                    waveBinScanKernel<<<1, 1024, 0, st[g]>>>(R);
                    waveBinScatterKernel<<<numSMs * 4, 256, 0, st[g]>>>(R);
                    n += 2;
                }
            }
        }
        if (nGroups == 2) {     /* join the second stream before reading the counter */
            if ((e = cudaEventRecord(events[3], stream2)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(stream, events[3], 0)) != cudaSuccess) return e;
        }
        e = cudaMemcpyAsync(hostDone + (k & 1), P.counters + CTR_DONE, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess)
            return e;
        e = cudaEventRecord(events[k & 1], stream);
        if (e != cudaSuccess)
            return e;
        if (nGroups == 2) {     /* fork again for the next batch */
            if ((e = cudaStreamWaitEvent(stream2, events[k & 1], 0)) != cudaSuccess) return e;
        }
        if (k >= 1) {
            e = cudaEventSynchronize(events[(k - 1) & 1]);
            if (e != cudaSuccess)
                return e;
            const unsigned long long done = hostDone[(k - 1) & 1];
            if (done >= P.W.nSlots)
                break;
            if (done > 0 && P.W.nSlots - done <= drainSlots) {
                /* the work items have run out (a slot only retires when it finds none) and few slots are alive:
                 * finish them without further rounds, behind the batch that is already queued */
                for (int g = 0; g < nGroups; g++)
                    waveDrainKernel<F><<<grid[g], BLOCK, stackBytes(), st[g]>>>(S, G[g]);
                n += nGroups;
                break;
            }
        }
    }
    if (nGroups == 2) {     /* final join: everything queued on the second stream precedes what follows on `stream` */
        if ((e = cudaEventRecord(events[3], stream2)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream, events[3], 0)) != cudaSuccess) return e;
    }
    *launches = n + (P.substreams == 1 ? 0 : 1);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    return launchResolve(P, stream);
}

cudaError_t launchWavefront(int features, const DScene& S, const RenderParams& P, int numSMs, cudaStream_t stream,
        cudaStream_t stream2, unsigned long long* hostDone, cudaEvent_t* events, int* launches)
{
    if (features == 0)
        return launchWavefrontT<0>(S, P, numSMs, stream, stream2, hostDone, events, launches);
    if (!(features & F_ANIM))   /* media / ToF / gated sensor without animated hitables: no per-thread animation cache */
        return launchWavefrontT<F_MEDIA | F_OPL>(S, P, numSMs, stream, stream2, hostDone, events, launches);
    return launchWavefrontT<F_MEDIA | F_ANIM | F_OPL>(S, P, numSMs, stream, stream2, hostDone, events, launches);
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

cudaError_t launchGroundTruth(int features, const DScene& S, const RenderParams& P, const GroundTruthParams& G, cudaStream_t stream)
{
    const unsigned long long pixels = (unsigned long long)P.width * P.height;
    const unsigned int grid = (unsigned int)((pixels + BLOCK - 1) / BLOCK);
    if (grid == 0)
        return cudaSuccess;
    if (features != 0)
        groundTruthKernel<F_MEDIA | F_ANIM><<<grid, BLOCK, stackBytes(), stream>>>(S, P, G);
    else
        groundTruthKernel<0><<<grid, BLOCK, stackBytes(), stream>>>(S, P, G);
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
