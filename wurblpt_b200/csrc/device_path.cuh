/*
 * device_path.cuh -- tracePath() (reference wurblpt.hpp:108-275) and the per-pixel sample loop of
 * mcpt() (wurblpt.hpp:336-368) as a small state machine around ONE closest-hit query:
 *
 *      startSample()   ST_NEW  -> camera ray                               -> ST_MAIN
 *      [closest-hit query for state.walkRay]
 *      processHit()    ST_MAIN        -> emission, scatter, set up next-event estimation
 *                                        -> ST_SHADOW_HOT / ST_SHADOW_ENV (shadow ray in walkRay)
 *                                        -> ST_MAIN (next bounce in walkRay) or ST_NEW (path ended)
 *                      ST_SHADOW_*    -> add the light sample if visible, then next bounce / path end
 *
 * Both pipelines are built from these two functions: the megakernel keeps a PathState in registers
 * and loops; the wavefront pipeline keeps it in HBM between its shade and trace kernels.
 */
#pragma once

#include "device_shade.cuh"

namespace wptd {

enum { ST_IDLE = -1, ST_NEW = 0, ST_MAIN = 1, ST_SHADOW_HOT = 2, ST_SHADOW_ENV = 3, ST_EXHAUSTED = 4,
       ST_TAIL = 5 };   /* wavefront only: the path has ended, its last shadow query is still pending */
enum { SHADOW_NONE = 0, SHADOW_HOT = 1, SHADOW_ENV = 2 };

struct PathState {
    int state;
    /* work item: one (pixel, sub-stream) */
    unsigned int pixel, sub, sampleIndex;
    float acc[3];
    /* path */
    Ray walkRay;            /* the ray of the pending / next closest-hit query */
    float4 att;
    float pathLength;
    unsigned int pathComponent;
    float4 opl, rayRi;      /* optical path length and refractive index (F_OPL only) */
    /* carried across a shadow-ray query */
    float4 nextAtt, pendRad;
    float3 nextDir;
    uint32_t pendHot;
    float4 nextRi, shadowRi;    /* F_OPL only */
    /* wavefront pipeline: the shadow query that travels with the next path segment */
    int shadow;                 /* SHADOW_* */
    float3 shDir;
};

WPT_DEV bool isTofLight(const wpt_material* m)
{
    return m != nullptr && m->type == WPT_MAT_LIGHT_TOF;
}

/* Begin the next sample of the item (wurblpt.hpp:348-365). Returns false when the item has no
 * samples left (the caller then stores the pixel and fetches a new item). */
template<int F>
WPT_DEV bool startSample(const DScene& S, const RenderParams& P, PathState& ps, Ctx<F>& ctx)
{
    if (ps.sampleIndex >= P.samples)
        return false;
    const unsigned int px = ps.pixel % P.width, py = ps.pixel / P.width;
    /* stratified jitter (wurblpt.hpp:349-359); the first draw goes to the row offset, see in01x2 */
    float u = float(px), v = float(py);
    if (P.randomize) {
        const unsigned int j = ps.sampleIndex / P.samplesSqrt;
        const unsigned int i = ps.sampleIndex % P.samplesSqrt;
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
    ps.walkRay = cameraRay(P, u, v, ctx.prng, true);
    ps.att = f4(1.0f);
    ps.pathLength = 0.0f;
    ps.opl = f4(0.0f);
    ps.rayRi = f4(1.0f);        /* refractiveIndexOfVacuum (camera.hpp:184) */
    ps.pathComponent = 0;
    ps.sampleIndex += P.substreams;
    ps.state = ST_MAIN;
    return true;
}

/* Update attenuation and ray for the next path segment, then Russian roulette (wurblpt.hpp:255-273). */
WPT_DEV void advancePath(const RenderParams& P, PathState& ps, Prng& prng, bool trackRi)
{
    ps.att = ps.nextAtt;
    ps.walkRay.d = ps.nextDir;
    if (trackRi)
        ps.rayRi = ps.nextRi;
    ps.state = ST_MAIN;
    const float maxAtt = vmaxcomp(ps.att);
    if (maxAtt < P.rrThreshold && ps.pathComponent >= 5) {
        const float q = rclamp(1.0f - maxAtt, 0.0f, 0.95f);
        if (in01(prng) < q) {
            ps.state = ST_NEW;
        } else {
            const float rrWeight = 1.0f / (1.0f - q);
            ps.att = ps.att * rrWeight;
        }
    }
    ps.pathComponent++;
}

/* The cheap part of tracePath() after a closest-hit query: a camera/bounce ray that left the scene
 * or reached the last allowed path component, and the results of shadow rays. Returns true when the
 * query was a camera/bounce ray that hit a surface which still has to be shaded (surfaceShade()). */
template<int F>
WPT_DEV bool processLogic(const DScene& S, const RenderParams& P, PathState& ps, Ctx<F>& ctx, const Hit& hit, const RayHelper& h)
{
    const Ray ray = ps.walkRay;
    if (ps.state == ST_MAIN) {
        if (hit.ref == NO_HIT) {
            /* no hit: environment map, end of path (wurblpt.hpp:136-146) */
            if (S.envType != WPT_ENV_NONE)
                accumulate<F>(P, ps.acc, f4(MAXVAL), MAXVAL, ps.att * envRadiance(S, ray.d), false);
            ps.state = ST_NEW;
            return false;
        }
        ps.pathLength += hit.a;
        if (F & F_OPL)
            ps.opl = ps.opl + hit.a * ps.rayRi;
        if (!(ps.pathComponent + 1 < P.maxPathComponents)) {
            ps.state = ST_NEW;
            return false;
        }
        return true;
    }
    if (ps.state == ST_SHADOW_HOT) {
        /* the chosen hot spot must be the closest hit (wurblpt.hpp:208) */
        if (hit.ref == ps.pendHot) {
            HitRec dhr = buildHitRecord<F>(S, hit, ray, h, ctx);
            const wpt_material* lightMat = resolveMaterial(S, dhr.material, dhr);
            const float4 rad = ps.pendRad * materialEmitted(S, lightMat, ray, dhr);
            float4 oplAcc = ps.opl;
            if (F & F_OPL)
                oplAcc = ps.opl + hit.a * ps.shadowRi;
            accumulate<F>(P, ps.acc, oplAcc, hit.a, rad, isTofLight(lightMat));
        }
    } else {
        /* the environment must be visible (wurblpt.hpp:240) */
        if (hit.ref == NO_HIT)
            accumulate<F>(P, ps.acc, f4(MAXVAL), MAXVAL, ps.pendRad, false);
    }
    advancePath(P, ps, ctx.prng, (F & F_OPL) != 0);
    return false;
}

/* The expensive part: hit record, emission, scatter, next-event estimation set-up (wurblpt.hpp:156-252).
 * Leaves the slot with its next query (bounce or shadow ray) or in ST_NEW (path ended). */
template<int F>
WPT_DEV void surfaceShade(const DScene& S, const RenderParams& P, PathState& ps, Ctx<F>& ctx, const Hit& hit, const RayHelper& h)
{
    const Ray ray = ps.walkRay;
    HitRec hr = buildHitRecord<F>(S, hit, ray, h, ctx);
    const wpt_material* mat = resolveMaterial(S, hr.material, hr);
    const ScatterRec sr = materialScatter(S, mat, ray, ps.rayRi, hr, ctx.prng);
    accumulate<F>(P, ps.acc, ps.opl, (ps.pathComponent == 0 ? 0.0f : hit.a),
            ps.att * materialEmitted(S, mat, ray, hr), isTofLight(mat));
    if (sr.type == SCATTER_NONE) {
        ps.state = ST_NEW;
        return;
    }
    ps.nextAtt = ps.att * sr.att;
    if (sr.type == SCATTER_RANDOM) {
        if (sr.pdf > 0.0f)
            ps.nextAtt = ps.nextAtt / sr.pdf;
        else
            ps.nextAtt = f4(0.0f);
    }
    ps.nextDir = sr.dir;
    if (F & F_OPL)
        ps.nextRi = sr.ri;
    bool advance = true;
    if (sr.type == SCATTER_RANDOM && S.nHot > 0) {
        /* next event estimation towards a hot spot with MIS (wurblpt.hpp:179-220) */
        const float hotPdf = hotSpotsPdf<F>(S, hr.p, sr.dir, ray.time, ctx);
        ps.nextAtt = ps.nextAtt * powerHeuristic(sr.pdf, hotPdf);
        unsigned int hotIndex = (unsigned int)(in01(ctx.prng) * float(S.nHot));
        hotIndex = min(hotIndex, S.nHot - 1u);
        const uint32_t hotRef = __ldg(S.hotRefs + hotIndex);
        const float3 directDir = hotSpotDirection<F>(S, hotRef, hr.p, ctx);
        const float directPdf = hotSpotsPdf<F>(S, hr.p, directDir, ray.time, ctx);
        if (directPdf > 0.0f) {
            const ScatterRec dsr = materialScatterToDirection(S, mat, ray, ps.rayRi, hr, directDir);
            if (dsr.pdf > 0.0f) {
                const float weight = powerHeuristic(directPdf, dsr.pdf);
                ps.pendRad = ((ps.att * dsr.att) / directPdf) * weight;
                ps.pendHot = hotRef;
                if (F & F_OPL)
                    ps.shadowRi = dsr.ri;
                ps.walkRay.d = directDir;
                ps.state = ST_SHADOW_HOT;
                advance = false;
            }
        }
    } else if (sr.type == SCATTER_RANDOM && S.envType != WPT_ENV_NONE && S.envN > 0) {
        /* next event estimation towards the environment map (wurblpt.hpp:221-252) */
        const float lightsP = envPdf(S, sr.dir);
        ps.nextAtt = ps.nextAtt * powerHeuristic(sr.pdf, lightsP);
        const float3 lightDir = envSample(S, ctx.prng);
        const float lightDirP = envPdf(S, lightDir);
        const ScatterRec lsr = materialScatterToDirection(S, mat, ray, ps.rayRi, hr, lightDir);
        if (lsr.pdf > 0.0f) {
            const float weight = powerHeuristic(lightDirP, lsr.pdf);
            ps.pendRad = (((ps.att * lsr.att) / lightDirP) * weight) * envRadiance(S, lightDir);
            ps.walkRay.d = lightDir;
            ps.state = ST_SHADOW_ENV;
            advance = false;
        }
    }
    /* the shadow ray and the next path segment both start at the hit point (wurblpt.hpp:203,256) */
    ps.walkRay.o = hr.p;
    if (advance)
        advancePath(P, ps, ctx.prng, (F & F_OPL) != 0);
}

/* Everything tracePath() does after a closest-hit query returned `hit` for ps.walkRay. */
template<int F>
WPT_DEV void processHit(const DScene& S, const RenderParams& P, PathState& ps, Ctx<F>& ctx, const Hit& hit, const RayHelper& h)
{
    if (processLogic<F>(S, P, ps, ctx, hit, h))
        surfaceShade<F>(S, P, ps, ctx, hit, h);
}

/* ---- wavefront pipeline: one round per path vertex ----
 *
 * tracePath() sets up the light sample of a vertex (shadow ray) and the next path segment at the same hit point;
 * the two rays do not depend on each other, only the ORDER in which their results are consumed matters
 * (wurblpt.hpp:203-220 before :255-273). The wavefront pipeline therefore traces both in the same round:
 *   waveShade()      at a surface hit: emission, scatter, light sample -> shadow query S and next segment B
 *                    (throughput update and Russian roulette of wurblpt.hpp:255-273 included)
 *   consumeShadow()  when both are answered: the light sample is added if the shadow query says "visible";
 *                    then B's hit is shaded (waveShade again) or ends the path.
 * The Prng draws of Russian roulette come before those of the next vertex either way; images are the same
 * Monte-Carlo estimate. */

/* The result of the pending shadow query S (wurblpt.hpp:204-219, :236-251). */
template<int F>
WPT_DEV void consumeShadow(const DScene& S, const RenderParams& P, PathState& ps, Ctx<F>& ctx, const Hit& hit)
{
    if (ps.shadow == SHADOW_HOT) {
        /* the chosen hot spot must be the closest hit (wurblpt.hpp:208) */
        if (hit.ref == ps.pendHot) {
            Ray ray;
            ray.o = ps.walkRay.o;
            ray.d = ps.shDir;
            ray.time = ps.walkRay.time;
            const RayHelper h = makeRayHelper(ray);
            HitRec dhr = buildHitRecord<F>(S, hit, ray, h, ctx);
            const wpt_material* lightMat = resolveMaterial(S, dhr.material, dhr);
            const float4 rad = ps.pendRad * materialEmitted(S, lightMat, ray, dhr);
            float4 oplAcc = ps.opl;
            if (F & F_OPL)
                oplAcc = ps.opl + hit.a * ps.shadowRi;
            accumulate<F>(P, ps.acc, oplAcc, hit.a, rad, isTofLight(lightMat));
        }
    } else if (ps.shadow == SHADOW_ENV) {
        /* the environment must be visible (wurblpt.hpp:240) */
        if (hit.ref == NO_HIT)
            accumulate<F>(P, ps.acc, f4(MAXVAL), MAXVAL, ps.pendRad, false);
    }
    ps.shadow = SHADOW_NONE;
}

/* A camera ray or path segment that left the scene, or reached the last allowed path component
 * (wurblpt.hpp:136-154): the path ends. */
template<int F>
WPT_DEV void waveEndPath(const DScene& S, const RenderParams& P, PathState& ps, const Hit& hit)
{
    if (hit.ref == NO_HIT && S.envType != WPT_ENV_NONE)
        accumulate<F>(P, ps.acc, f4(MAXVAL), MAXVAL, ps.att * envRadiance(S, ps.walkRay.d), false);
    ps.state = ST_NEW;
}

/* Surface hit of the camera ray / path segment B that is to be shaded (not on the last path component):
 * wurblpt.hpp:149-273 for this vertex. MT >= 0: the material type, known from the shading queue's class. */
template<int F, int MT>
WPT_DEV void waveShade(const DScene& S, const RenderParams& P, PathState& ps, Ctx<F>& ctx, const Hit& hit, const RayHelper& h)
{
    const Ray ray = ps.walkRay;
    ps.pathLength += hit.a;
    if (F & F_OPL)
        ps.opl = ps.opl + hit.a * ps.rayRi;
    HitRec hr = buildHitRecord<F>(S, hit, ray, h, ctx);
    const wpt_material* mat = resolveMaterial(S, hr.material, hr);
    ScatterRec sr;
    if (MT >= 0)
        sr = materialScatterT<MT>(S, mat, ray, ps.rayRi, hr, ctx.prng);
    else
        sr = materialScatter(S, mat, ray, ps.rayRi, hr, ctx.prng);
    /* only lights and ModPhong emit */
    if (MT < 0 || MT == WPT_MAT_MODPHONG)
        accumulate<F>(P, ps.acc, ps.opl, (ps.pathComponent == 0 ? 0.0f : hit.a),
                ps.att * materialEmitted(S, mat, ray, hr), isTofLight(mat));
    ps.shadow = SHADOW_NONE;
    if (sr.type == SCATTER_NONE) {
        ps.state = ST_NEW;
        return;
    }
    float4 nextAtt = ps.att * sr.att;
    if (sr.type == SCATTER_RANDOM) {
        if (sr.pdf > 0.0f)
            nextAtt = nextAtt / sr.pdf;
        else
            nextAtt = f4(0.0f);
    }
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
            ScatterRec dsr;
            if (MT >= 0)
                dsr = materialScatterToDirectionT<MT>(S, mat, ray, ps.rayRi, hr, directDir);
            else
                dsr = materialScatterToDirection(S, mat, ray, ps.rayRi, hr, directDir);
            if (dsr.pdf > 0.0f) {
                const float weight = powerHeuristic(directPdf, dsr.pdf);
                ps.pendRad = ((ps.att * dsr.att) / directPdf) * weight;
                ps.pendHot = hotRef;
                if (F & F_OPL)
                    ps.shadowRi = dsr.ri;
                ps.shDir = directDir;
                ps.shadow = SHADOW_HOT;
            }
        }
    } else if (sr.type == SCATTER_RANDOM && S.envType != WPT_ENV_NONE && S.envN > 0) {
        /* next event estimation towards the environment map (wurblpt.hpp:221-252) */
        const float lightsP = envPdf(S, sr.dir);
        nextAtt = nextAtt * powerHeuristic(sr.pdf, lightsP);
        const float3 lightDir = envSample(S, ctx.prng);
        const float lightDirP = envPdf(S, lightDir);
        ScatterRec lsr;
        if (MT >= 0)
            lsr = materialScatterToDirectionT<MT>(S, mat, ray, ps.rayRi, hr, lightDir);
        else
            lsr = materialScatterToDirection(S, mat, ray, ps.rayRi, hr, lightDir);
        if (lsr.pdf > 0.0f) {
            const float weight = powerHeuristic(lightDirP, lsr.pdf);
            ps.pendRad = (((ps.att * lsr.att) / lightDirP) * weight) * envRadiance(S, lightDir);
            ps.pendHot = NO_HIT;
            ps.shDir = lightDir;
            ps.shadow = SHADOW_ENV;
        }
    }
    /* the shadow ray and the next path segment both start at the hit point (wurblpt.hpp:203,256) */
    ps.walkRay.o = hr.p;
    /* next segment: attenuation, ray, Russian roulette (wurblpt.hpp:255-273) */
    ps.att = nextAtt;
    ps.walkRay.d = sr.dir;
    if (F & F_OPL)
        ps.rayRi = sr.ri;
    ps.state = ST_MAIN;
    const float maxAtt = vmaxcomp(ps.att);
    if (maxAtt < P.rrThreshold && ps.pathComponent >= 5) {
        const float q = rclamp(1.0f - maxAtt, 0.0f, 0.95f);
        if (in01(ctx.prng) < q) {
            ps.state = ST_NEW;
        } else {
            const float rrWeight = 1.0f / (1.0f - q);
            ps.att = ps.att * rrWeight;
        }
    }
    ps.pathComponent++;
    /* A segment on the last allowed path component contributes only if it leaves the scene into an environment map
     * (wurblpt.hpp:136-154: its hit is dropped before emission): without one there is nothing to trace. */
    if (ps.state == ST_MAIN && !(ps.pathComponent + 1 < P.maxPathComponents) && S.envType == WPT_ENV_NONE)
        ps.state = ST_NEW;
    if (ps.state == ST_NEW && ps.shadow != SHADOW_NONE)
        ps.state = ST_TAIL;
}

/* A work item is one (pixel, sub-stream): all samples of that sub-stream, run back to back with one
 * Prng (wurblpt.hpp:342-368). Items are numbered tile by tile (1 << tileShift per tile) so that neighbouring
 * items are neighbouring pixels. */
struct WorkItem {
    unsigned int pixel, sub;
    bool valid;
};

WPT_DEV WorkItem decodeItem(const RenderParams& P, unsigned int item)
{
    WorkItem w;
    const unsigned int localTile = item >> P.tileShift, inTile = item & ((1u << P.tileShift) - 1u);
    const unsigned int tile = P.tileFirst + localTile * P.tileStride;
    const unsigned int tileX = tile % P.tilesX, tileY = tile / P.tilesX;
    w.sub = inTile & (P.substreams - 1u);
    const unsigned int pit = inTile >> P.subShift;
    const unsigned int x = tileX * P.tileW + pit % P.tileW;
    const unsigned int y = tileY * P.tileH + pit / P.tileW;
    w.valid = (x < P.width && y < P.height);
    w.pixel = y * P.width + x;
    return w;
}

template<int F>
WPT_DEV void beginItem(PathState& ps, Ctx<F>& ctx, const WorkItem& w)
{
    ps.pixel = w.pixel;
    ps.sub = w.sub;
    ctx.prng = prngInitSub(w.pixel, w.sub); /* sub 0: Prng prng(pixel) (wurblpt.hpp:342) */
    ps.acc[0] = ps.acc[1] = ps.acc[2] = 0.0f;
    ps.sampleIndex = w.sub;
    ps.state = ST_NEW;
}

/* Sensor::finishPixel (sensor_rgb.hpp:82-87), or a partial sum that resolveKernel adds up in a fixed order */
WPT_DEV void finishItem(const RenderParams& P, const PathState& ps)
{
    if (P.substreams == 1) {
        float* out = P.frame + size_t(ps.pixel) * 3;
        out[0] = P.invSamples * ps.acc[0];
        out[1] = P.invSamples * ps.acc[1];
        out[2] = P.invSamples * ps.acc[2];
    } else {
        float* out = P.partials + (size_t(ps.pixel) * P.substreams + ps.sub) * 3;
        out[0] = ps.acc[0];
        out[1] = ps.acc[1];
        out[2] = ps.acc[2];
    }
}

}
