/*
 * bvh_device.cu -- Scene::updateBVH() on the device (reference scene.hpp:145-169, bvh.hpp:93-193):
 * builds the 4-wide tree the production walk reads, and the primitive arrays in slot order, from the
 * bounds of the hitables. SURVEY.md section 8(f) rank 2.
 *
 * The reference's builder is a full-sort SAH sweep, one recursion level after the other; its tree is
 * restated on the host (scene_host.cpp) because the parity walk must visit the reference's nodes. The
 * production walk only needs SOME good tree over the same hitables (closest hits do not depend on the
 * tree; exact distance ties are the counted exception), so this builder is free to be parallel:
 *
 *   bounds of every hitable over [t0,t1]   host, the reference's formulas (hitable_triangle.hpp:336-395)
 *   boundsKernel      box of the centroids                                      atomics on ordered ints
 *   mortonKernel      63-bit Morton code of every centroid                      HBM streaming
 *   bitonic*Kernel    sort (code, hitable) pairs                                HBM / shared memory
 *   hierarchyKernel   binary radix tree over the sorted codes (Karras 2012)     one thread per inner node
 *   fitKernel         boxes bottom-up, second arrival continues                 atomics
 *   slot*Kernel       triangle / sphere slot of every sorted position (scan)
 *   gatherKernel      primitive arrays in slot order (tri, triShade, classes, ids)
 *   collapseKernel    4-wide nodes level by level: opens the child with the largest surface area until
 *                     four are held; subtrees of <= 4 triangles become one leaf (a run of slots)
 *
 * Slots are in Morton order, so neighbouring slots are neighbouring triangles (like the reference's
 * depth-first storage order). All kernels are bound by HBM streaming or launch latency: 1 M triangles
 * build in a few milliseconds.
 */
#include <cstdint>
#include <cstdio>

#include "bvh_device.hpp"

namespace wptd {

namespace {

constexpr uint32_t LEAF_FLAG = 0x80000000u;     /* child of a binary node: leaf at sorted position (child & ~LEAF_FLAG) */
constexpr uint32_t REF_TYPE_SHIFT_ = 29;
constexpr uint32_t RUN_COUNT_SHIFT_ = 27;
constexpr int32_t EMPTY_CHILD_ = 0x7FFFFFFF;
constexpr uint32_t PRIM_TRIANGLE_ = 0, PRIM_TRIANGLE_RUN_ = 3;

__device__ __forceinline__ int orderedInt(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float orderedFloat(int i)
{
    return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF);
}

/* centroid bounds: bounds[0..2] = min, [3..5] = max as ordered ints */
__global__ void boundsKernel(const float4* __restrict__ box, uint32_t n, int* bounds)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    const bool valid = i < n;
    if (valid) {
        const float4 lo = box[2 * i], hi = box[2 * i + 1];
        cx = 0.5f * (lo.x + hi.x); cy = 0.5f * (lo.y + hi.y); cz = 0.5f * (lo.z + hi.z);
    }
    float mnx = valid ? cx : 3.4e38f, mny = valid ? cy : 3.4e38f, mnz = valid ? cz : 3.4e38f;
    float mxx = valid ? cx : -3.4e38f, mxy = valid ? cy : -3.4e38f, mxz = valid ? cz : -3.4e38f;
    for (int off = 16; off > 0; off >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xFFFFFFFFu, mnx, off)); mny = fminf(mny, __shfl_xor_sync(0xFFFFFFFFu, mny, off));
        mnz = fminf(mnz, __shfl_xor_sync(0xFFFFFFFFu, mnz, off));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xFFFFFFFFu, mxx, off)); mxy = fmaxf(mxy, __shfl_xor_sync(0xFFFFFFFFu, mxy, off));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xFFFFFFFFu, mxz, off));
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicMin(bounds + 0, orderedInt(mnx)); atomicMin(bounds + 1, orderedInt(mny)); atomicMin(bounds + 2, orderedInt(mnz));
        atomicMax(bounds + 3, orderedInt(mxx)); atomicMax(bounds + 4, orderedInt(mxy)); atomicMax(bounds + 5, orderedInt(mxz));
    }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v)
{
    uint64_t x = v & 0x1FFFFFu;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

/* keys[i] = Morton code of hitable i's centroid, order[i] = i; the padding up to the power of two sorts last */
__global__ void mortonKernel(const float4* __restrict__ box, uint32_t n, uint32_t nPadded, const int* bounds,
        uint64_t* keys, uint32_t* order)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nPadded)
        return;
    if (i >= n) {
        keys[i] = ~0ull;
        order[i] = 0xFFFFFFFFu;
        return;
    }
    const float3 mn = make_float3(orderedFloat(bounds[0]), orderedFloat(bounds[1]), orderedFloat(bounds[2]));
    const float3 mx = make_float3(orderedFloat(bounds[3]), orderedFloat(bounds[4]), orderedFloat(bounds[5]));
    const float4 lo = box[2 * i], hi = box[2 * i + 1];
    const float sx = mx.x > mn.x ? 2097151.0f / (mx.x - mn.x) : 0.0f;
    const float sy = mx.y > mn.y ? 2097151.0f / (mx.y - mn.y) : 0.0f;
    const float sz = mx.z > mn.z ? 2097151.0f / (mx.z - mn.z) : 0.0f;
    const uint32_t qx = (uint32_t)fminf(fmaxf((0.5f * (lo.x + hi.x) - mn.x) * sx, 0.0f), 2097151.0f);
    const uint32_t qy = (uint32_t)fminf(fmaxf((0.5f * (lo.y + hi.y) - mn.y) * sy, 0.0f), 2097151.0f);
    const uint32_t qz = (uint32_t)fminf(fmaxf((0.5f * (lo.z + hi.z) - mn.z) * sz, 0.0f), 2097151.0f);
    keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    order[i] = i;
}

/* ---- bitonic sort of (key, order) pairs, ascending by key then by order ---- */
__device__ __forceinline__ bool pairGreater(uint64_t ka, uint32_t oa, uint64_t kb, uint32_t ob)
{
    return ka > kb || (ka == kb && oa > ob);
}

constexpr uint32_t SORT_CHUNK = 2048;       /* elements sorted / merged in shared memory by one block of 1024 threads */

/* all stages with k <= SORT_CHUNK */
__global__ void __launch_bounds__(1024) bitonicSortSharedKernel(uint64_t* keys, uint32_t* order)
{
    __shared__ uint64_t sk[SORT_CHUNK];
    __shared__ uint32_t so[SORT_CHUNK];
    const uint32_t base = blockIdx.x * SORT_CHUNK;
    for (uint32_t t = threadIdx.x; t < SORT_CHUNK; t += 1024) {
        sk[t] = keys[base + t];
        so[t] = order[base + t];
    }
    __syncthreads();
    for (uint32_t k = 2; k <= SORT_CHUNK; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const uint32_t t = threadIdx.x;
            const uint32_t a = ((t & ~(j - 1)) << 1) | (t & (j - 1));       /* index with bit j clear */
            const uint32_t b = a | j;
            const bool ascending = (((base + a) & k) == 0);
            const uint64_t ka = sk[a], kb = sk[b];
            const uint32_t oa = so[a], ob = so[b];
            if (pairGreater(ka, oa, kb, ob) == ascending) {
                sk[a] = kb; sk[b] = ka;
                so[a] = ob; so[b] = oa;
            }
            __syncthreads();
        }
    }
    for (uint32_t t = threadIdx.x; t < SORT_CHUNK; t += 1024) {
        keys[base + t] = sk[t];
        order[base + t] = so[t];
    }
}

/* one step (j >= SORT_CHUNK) of stage k in HBM: one thread per pair */
__global__ void bitonicStepKernel(uint64_t* keys, uint32_t* order, uint32_t j, uint32_t k, uint32_t nPairs)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nPairs)
        return;
    const uint32_t a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const uint32_t b = a | j;
    const bool ascending = ((a & k) == 0);
    const uint64_t ka = keys[a], kb = keys[b];
    const uint32_t oa = order[a], ob = order[b];
    if (pairGreater(ka, oa, kb, ob) == ascending) {
        keys[a] = kb; keys[b] = ka;
        order[a] = ob; order[b] = oa;
    }
}

/* the steps j = SORT_CHUNK/2 .. 1 of stage k (k > SORT_CHUNK) in shared memory */
__global__ void __launch_bounds__(1024) bitonicMergeSharedKernel(uint64_t* keys, uint32_t* order, uint32_t k)
{
    __shared__ uint64_t sk[SORT_CHUNK];
    __shared__ uint32_t so[SORT_CHUNK];
    const uint32_t base = blockIdx.x * SORT_CHUNK;
    for (uint32_t t = threadIdx.x; t < SORT_CHUNK; t += 1024) {
        sk[t] = keys[base + t];
        so[t] = order[base + t];
    }
    __syncthreads();
    const bool ascending = ((base & k) == 0);
    for (uint32_t j = SORT_CHUNK >> 1; j > 0; j >>= 1) {
        const uint32_t t = threadIdx.x;
        const uint32_t a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const uint32_t b = a | j;
        const uint64_t ka = sk[a], kb = sk[b];
        const uint32_t oa = so[a], ob = so[b];
        if (pairGreater(ka, oa, kb, ob) == ascending) {
            sk[a] = kb; sk[b] = ka;
            so[a] = ob; so[b] = oa;
        }
        __syncthreads();
    }
    for (uint32_t t = threadIdx.x; t < SORT_CHUNK; t += 1024) {
        keys[base + t] = sk[t];
        order[base + t] = so[t];
    }
}

/* ---- binary radix tree (Karras, "Maximizing parallelism in the construction of BVHs, octrees, and k-d trees") ---- */
__device__ __forceinline__ int commonPrefix(const uint64_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n)
        return -1;
    const uint64_t x = keys[i] ^ keys[j];
    if (x == 0ull)
        return 64 + __clz(uint32_t(i) ^ uint32_t(j));      /* equal codes: the position decides */
    return __clzll((long long)x);
}

/* inner node i in [0, n-2]: children left[i], right[i] (LEAF_FLAG | sorted position for leaves), covered range */
__global__ void hierarchyKernel(const uint64_t* __restrict__ keys, int n, uint32_t* left, uint32_t* right,
        uint32_t* rangeFirst, uint32_t* rangeLast, uint32_t* parentOfInner, uint32_t* parentOfLeaf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1)
        return;
    const int d = (commonPrefix(keys, n, i, i + 1) - commonPrefix(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int minPrefix = commonPrefix(keys, n, i, i - d);
    int lmax = 2;
    while (commonPrefix(keys, n, i, i + lmax * d) > minPrefix)
        lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (commonPrefix(keys, n, i, i + (l + t) * d) > minPrefix)
            l += t;
    const int j = i + l * d;
    const int nodePrefix = commonPrefix(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1) {
        if (commonPrefix(keys, n, i, i + (s + t) * d) > nodePrefix)
            s += t;
        if (t <= 1)
            break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const uint32_t l_ = (first == gamma ? (LEAF_FLAG | uint32_t(gamma)) : uint32_t(gamma));
    const uint32_t r_ = (last == gamma + 1 ? (LEAF_FLAG | uint32_t(gamma + 1)) : uint32_t(gamma + 1));
    left[i] = l_;
    right[i] = r_;
    rangeFirst[i] = uint32_t(first);
    rangeLast[i] = uint32_t(last);
    if (l_ & LEAF_FLAG) parentOfLeaf[gamma] = uint32_t(i); else parentOfInner[gamma] = uint32_t(i);
    if (r_ & LEAF_FLAG) parentOfLeaf[gamma + 1] = uint32_t(i); else parentOfInner[gamma + 1] = uint32_t(i);
    if (i == 0)
        parentOfInner[0] = 0xFFFFFFFFu;
}

/* boxes bottom-up: node boxes are stored as 2 float4 (lo.xyz | "all triangles" flag, hi.xyz | surface area);
 * inner node i at [i], leaf at sorted position p at [n - 1 + p] */
__device__ __forceinline__ float boxArea(float4 lo, float4 hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

__global__ void fitKernel(const float4* __restrict__ box, const uint32_t* __restrict__ order, int n,
        const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
        const uint32_t* __restrict__ parentOfInner, const uint32_t* __restrict__ parentOfLeaf,
        float4* nodeBox, uint32_t* arrived)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const uint32_t prim = order[p];
    float4 lo = box[2 * prim], hi = box[2 * prim + 1];
    const uint32_t type = __float_as_uint(lo.w);
    lo.w = (type == PRIM_TRIANGLE_ ? 1.0f : 0.0f);
    hi.w = boxArea(lo, hi);
    nodeBox[2 * (n - 1 + p)] = lo;
    nodeBox[2 * (n - 1 + p) + 1] = hi;
    if (n == 1)
        return;
    __threadfence();
    uint32_t node = parentOfLeaf[p];
    while (node != 0xFFFFFFFFu) {
        if (atomicAdd(arrived + node, 1u) == 0u)
            return;                     /* the sibling's subtree is not done: its thread continues */
        __threadfence();
        const uint32_t l = left[node], r = right[node];
        const uint32_t li = (l & LEAF_FLAG) ? uint32_t(n - 1) + (l & ~LEAF_FLAG) : l;
        const uint32_t ri = (r & LEAF_FLAG) ? uint32_t(n - 1) + (r & ~LEAF_FLAG) : r;
        const float4 llo = __ldcg(nodeBox + 2 * li), lhi = __ldcg(nodeBox + 2 * li + 1);
        const float4 rlo = __ldcg(nodeBox + 2 * ri), rhi = __ldcg(nodeBox + 2 * ri + 1);
        float4 mlo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), llo.w * rlo.w);
        float4 mhi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
        mhi.w = boxArea(mlo, mhi);
        nodeBox[2 * node] = mlo;
        nodeBox[2 * node + 1] = mhi;
        __threadfence();
        node = parentOfInner[node];
    }
}

/* ---- slots: exclusive scan of "is a triangle" over the sorted positions ---- */
constexpr uint32_t SCAN_BLOCK = 1024;

__global__ void __launch_bounds__(SCAN_BLOCK) scanBlocksKernel(const float4* __restrict__ box, const uint32_t* __restrict__ order, uint32_t n,
        uint32_t* triRank, uint32_t* blockSums)
{
    __shared__ uint32_t warpSums[32];
    const uint32_t p = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    uint32_t v = 0;
    if (p < n)
        v = (__float_as_uint(box[2 * order[p]].w) == PRIM_TRIANGLE_) ? 1u : 0u;
    uint32_t x = v;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
        if ((threadIdx.x & 31u) >= (uint32_t)off)
            x += y;
    }
    if ((threadIdx.x & 31u) == 31u)
        warpSums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = warpSums[threadIdx.x];
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, off);
            if (threadIdx.x >= (uint32_t)off)
                w += y;
        }
        warpSums[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t before = (threadIdx.x >> 5) ? warpSums[(threadIdx.x >> 5) - 1] : 0u;
    if (p < n)
        triRank[p] = before + x - v;        /* exclusive within the block */
    if (threadIdx.x == SCAN_BLOCK - 1)
        blockSums[blockIdx.x] = before + x;
}

/* exclusive scan of the block sums by one block (a few thousand entries) */
__global__ void __launch_bounds__(1024) scanSumsKernel(uint32_t* blockSums, uint32_t nBlocks, uint32_t* total)
{
    __shared__ uint32_t carry;
    __shared__ uint32_t warpSums[32];
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nBlocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nBlocks ? blockSums[i] : 0u;
        uint32_t x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
            if ((threadIdx.x & 31u) >= (uint32_t)off)
                x += y;
        }
        if ((threadIdx.x & 31u) == 31u)
            warpSums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warpSums[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, off);
                if (threadIdx.x >= (uint32_t)off)
                    w += y;
            }
            warpSums[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t before = carry + ((threadIdx.x >> 5) ? warpSums[(threadIdx.x >> 5) - 1] : 0u);
        if (i < nBlocks)
            blockSums[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023)
            carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = carry;
}

/* Primitive arrays in slot order. Triangles get the slots 0.. in sorted order, spheres likewise. */
__global__ void gatherKernel(BuildIO io, const uint32_t* __restrict__ order, const uint32_t* __restrict__ triRank,
        const uint32_t* __restrict__ blockSums, uint32_t* leafRef, uint32_t* refOfPrim)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= io.n)
        return;
    const uint32_t prim = order[p];
    const float4 lo = io.box[2 * prim], hi = io.box[2 * prim + 1];
    const uint32_t type = __float_as_uint(lo.w), typeIndex = __float_as_uint(hi.w);
    const uint32_t trisBefore = triRank[p] + blockSums[p / SCAN_BLOCK];
    uint32_t ref;
    if (type == PRIM_TRIANGLE_) {
        const uint32_t slot = trisBefore;
        io.tri[3 * slot + 0] = io.triPrim[3 * typeIndex + 0];
        io.tri[3 * slot + 1] = io.triPrim[3 * typeIndex + 1];
        io.tri[3 * slot + 2] = io.triPrim[3 * typeIndex + 2];
        io.triShade[slot] = io.triShadePrim[typeIndex];
        io.triClass[slot] = io.triClassPrim[typeIndex];
        io.triGlobal[slot] = prim;
        ref = (PRIM_TRIANGLE_ << REF_TYPE_SHIFT_) | slot;
    } else {
        const uint32_t slot = p - trisBefore;
        io.sph[3 * slot + 0] = io.sphPrim[3 * typeIndex + 0];
        io.sph[3 * slot + 1] = io.sphPrim[3 * typeIndex + 1];
        io.sph[3 * slot + 2] = io.sphPrim[3 * typeIndex + 2];
        io.sphClass[slot] = io.sphClassPrim[typeIndex];
        io.sphGlobal[slot] = prim;
        ref = (type << REF_TYPE_SHIFT_) | slot;
    }
    leafRef[p] = ref;
    refOfPrim[prim] = ref;
}

__global__ void hotRefsKernel(const uint32_t* __restrict__ hotPrims, uint32_t nHot, const uint32_t* __restrict__ refOfPrim, uint32_t* hotRefs)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nHot)
        hotRefs[i] = refOfPrim[hotPrims[i]];
}

/* ---- 4-wide nodes, one tree level per launch ---- */
struct CollapseItem {
    uint32_t binary;    /* inner node of the radix tree */
    uint32_t wide;      /* index of the wide node it becomes */
};

__device__ __forceinline__ bool becomesLeaf(uint32_t child, int n, const float4* __restrict__ nodeBox,
        const uint32_t* __restrict__ rangeFirst, const uint32_t* __restrict__ rangeLast, uint32_t maxRun)
{
    if (child & LEAF_FLAG)
        return true;
    /* a subtree of at most maxRun hitables that are all triangles: consecutive slots, one leaf */
    return (rangeLast[child] - rangeFirst[child] + 1u <= maxRun) && nodeBox[2 * child].w != 0.0f;
}

__global__ void collapseKernel(const CollapseItem* __restrict__ in, uint32_t nIn, CollapseItem* out, uint32_t* counters /* [0] = wide nodes, [1] = out items */,
        int n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
        const uint32_t* __restrict__ rangeFirst, const uint32_t* __restrict__ rangeLast,
        const float4* __restrict__ nodeBox, const uint32_t* __restrict__ leafRef, uint32_t maxRun, float4* node4)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nIn)
        return;
    const CollapseItem it = in[t];
    uint32_t kids[4];
    int count = 2;
    kids[0] = left[it.binary];
    kids[1] = right[it.binary];
    while (count < 4) {
        int pick = -1;
        float best = -1.0f;
        for (int k = 0; k < count; k++) {
            if (becomesLeaf(kids[k], n, nodeBox, rangeFirst, rangeLast, maxRun))
                continue;
            const float area = nodeBox[2 * kids[k] + 1].w;
            if (area > best) {
                best = area;
                pick = k;
            }
        }
        if (pick < 0)
            break;
        const uint32_t open = kids[pick];
        for (int k = count; k > pick + 1; k--)
            kids[k] = kids[k - 1];
        kids[pick] = left[open];
        kids[pick + 1] = right[open];
        count++;
    }
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
    int32_t child[4];
    const float far = 3.40282347e+38f;
    for (int k = 0; k < 4; k++) {
        if (k >= count) {
            lox[k] = hix[k] = loy[k] = hiy[k] = loz[k] = hiz[k] = far;
            child[k] = EMPTY_CHILD_;
            continue;
        }
        const uint32_t c = kids[k];
        const uint32_t bi = (c & LEAF_FLAG) ? uint32_t(n - 1) + (c & ~LEAF_FLAG) : c;
        const float4 lo = nodeBox[2 * bi], hi = nodeBox[2 * bi + 1];
        lox[k] = lo.x; loy[k] = lo.y; loz[k] = lo.z;
        hix[k] = hi.x; hiy[k] = hi.y; hiz[k] = hi.z;
        if (c & LEAF_FLAG) {
            child[k] = int32_t(~leafRef[c & ~LEAF_FLAG]);
        } else if (becomesLeaf(c, n, nodeBox, rangeFirst, rangeLast, maxRun)) {
            const uint32_t first = rangeFirst[c], cnt = rangeLast[c] - first + 1u;
            const uint32_t firstSlot = leafRef[first] & ((1u << RUN_COUNT_SHIFT_) - 1u);    /* all triangles: consecutive slots */
            child[k] = int32_t(~((PRIM_TRIANGLE_RUN_ << REF_TYPE_SHIFT_) | ((cnt - 1u) << RUN_COUNT_SHIFT_) | firstSlot));
        } else {
            const uint32_t w = atomicAdd(counters + 0, 1u);
            const uint32_t o = atomicAdd(counters + 1, 1u);
            out[o].binary = c;
            out[o].wide = w;
            child[k] = int32_t(w);
        }
    }
    float4* dst = node4 + size_t(it.wide) * 8;
    dst[0] = make_float4(lox[0], lox[1], lox[2], lox[3]);
    dst[1] = make_float4(hix[0], hix[1], hix[2], hix[3]);
    dst[2] = make_float4(loy[0], loy[1], loy[2], loy[3]);
    dst[3] = make_float4(hiy[0], hiy[1], hiy[2], hiy[3]);
    dst[4] = make_float4(loz[0], loz[1], loz[2], loz[3]);
    dst[5] = make_float4(hiz[0], hiz[1], hiz[2], hiz[3]);
    dst[6] = make_float4(__int_as_float(child[0]), __int_as_float(child[1]), __int_as_float(child[2]), __int_as_float(child[3]));
    dst[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

__global__ void initKernel(int* bounds, uint32_t* counters, CollapseItem* queue)
{
    bounds[0] = bounds[1] = bounds[2] = 0x7FFFFFFF;
    bounds[3] = bounds[4] = bounds[5] = int(0x80000000);
    counters[0] = 1;        /* wide node 0 = the root */
    counters[1] = 0;
    queue[0].binary = 0;
    queue[0].wide = 0;
}

template<typename T> T* carve(char*& p, size_t count)
{
    T* r = reinterpret_cast<T*>(p);
    p += (count * sizeof(T) + 255) / 256 * 256;
    return r;
}

/* The scratch arrays of one build, carved out of one allocation. */
struct Scratch {
    uint64_t* keys; uint32_t* order;
    uint32_t *left, *right, *rangeFirst, *rangeLast, *parentOfInner, *arrived, *parentOfLeaf;
    float4* nodeBox;
    uint32_t *triRank, *blockSums, *leafRef, *refOfPrim;
    CollapseItem *queueA, *queueB;
    int* bounds; uint32_t* counters;
    size_t bytes;

    Scratch(void* base, uint32_t n, uint32_t nPadded)
    {
        char* p = static_cast<char*>(base);
        keys = carve<uint64_t>(p, nPadded);
        order = carve<uint32_t>(p, nPadded);
        left = carve<uint32_t>(p, n);
        right = carve<uint32_t>(p, n);
        rangeFirst = carve<uint32_t>(p, n);
        rangeLast = carve<uint32_t>(p, n);
        parentOfInner = carve<uint32_t>(p, n);
        arrived = carve<uint32_t>(p, n);
        parentOfLeaf = carve<uint32_t>(p, n);
        nodeBox = carve<float4>(p, size_t(2) * n * 2);
        triRank = carve<uint32_t>(p, n);
        blockSums = carve<uint32_t>(p, size_t(n) / SCAN_BLOCK + 2);
        leafRef = carve<uint32_t>(p, n);
        refOfPrim = carve<uint32_t>(p, n);
        queueA = carve<CollapseItem>(p, n);
        queueB = carve<CollapseItem>(p, n);
        bounds = carve<int>(p, 8);
        counters = carve<uint32_t>(p, 8);
        bytes = size_t(p - static_cast<char*>(base));
    }
};

uint32_t paddedCount(uint32_t n)
{
    uint32_t nPadded = SORT_CHUNK;
    while (nPadded < n)
        nPadded <<= 1;
    return nPadded;
}

}

size_t deviceBuildScratchBytes(uint32_t n)
{
    return Scratch(nullptr, n, paddedCount(n)).bytes + 256;
}

/* Builds the tree for io.n >= 2 hitables. On return (stream synchronised): io.node4 holds *wideNodes nodes with the
 * root at index 0, the primitive arrays are in slot order, *triangles / *wideDepth are set. */
cudaError_t launchDeviceBuild(const BuildIO& io, void* scratch, uint32_t maxRun, cudaStream_t stream,
        uint32_t* wideNodes, uint32_t* wideDepth, uint32_t* triangles)
{
    const uint32_t n = io.n;
    const uint32_t nPadded = paddedCount(n);
    const Scratch sc(scratch, n, nPadded);
    uint64_t* keys = sc.keys;
    uint32_t* order = sc.order;
    uint32_t *left = sc.left, *right = sc.right, *rangeFirst = sc.rangeFirst, *rangeLast = sc.rangeLast;
    uint32_t *parentOfInner = sc.parentOfInner, *arrived = sc.arrived, *parentOfLeaf = sc.parentOfLeaf;
    float4* nodeBox = sc.nodeBox;
    uint32_t *triRank = sc.triRank, *blockSums = sc.blockSums, *leafRef = sc.leafRef, *refOfPrim = sc.refOfPrim;
    CollapseItem *queueA = sc.queueA, *queueB = sc.queueB;
    int* bounds = sc.bounds;
    uint32_t* counters = sc.counters;
    uint32_t* total = counters + 4;

    const uint32_t T = 256;
    initKernel<<<1, 1, 0, stream>>>(bounds, counters, queueA);
    cudaError_t e = cudaMemsetAsync(arrived, 0, size_t(n) * sizeof(uint32_t), stream);
    if (e != cudaSuccess)
        return e;
    boundsKernel<<<(n + T - 1) / T, T, 0, stream>>>(io.box, n, bounds);
    mortonKernel<<<(nPadded + T - 1) / T, T, 0, stream>>>(io.box, n, nPadded, bounds, keys, order);
    bitonicSortSharedKernel<<<nPadded / SORT_CHUNK, 1024, 0, stream>>>(keys, order);
    for (uint32_t k = SORT_CHUNK << 1; k <= nPadded; k <<= 1) {
        for (uint32_t j = k >> 1; j >= SORT_CHUNK; j >>= 1)
            bitonicStepKernel<<<(nPadded / 2 + T - 1) / T, T, 0, stream>>>(keys, order, j, k, nPadded / 2);
        bitonicMergeSharedKernel<<<nPadded / SORT_CHUNK, 1024, 0, stream>>>(keys, order, k);
    }
    hierarchyKernel<<<(n + T - 1) / T, T, 0, stream>>>(keys, int(n), left, right, rangeFirst, rangeLast, parentOfInner, parentOfLeaf);
    fitKernel<<<(n + T - 1) / T, T, 0, stream>>>(io.box, order, int(n), left, right, parentOfInner, parentOfLeaf, nodeBox, arrived);
    const uint32_t scanBlocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    scanBlocksKernel<<<scanBlocks, SCAN_BLOCK, 0, stream>>>(io.box, order, n, triRank, blockSums);
    scanSumsKernel<<<1, 1024, 0, stream>>>(blockSums, scanBlocks, total);
    gatherKernel<<<(n + T - 1) / T, T, 0, stream>>>(io, order, triRank, blockSums, leafRef, refOfPrim);
    if (io.nHot)
        hotRefsKernel<<<(io.nHot + T - 1) / T, T, 0, stream>>>(io.hotPrims, io.nHot, refOfPrim, io.hotRefs);
    /* level by level: the host reads how many nodes the level produced */
    uint32_t host[8] = { 0 };
    uint32_t nIn = 1, depth = 0;
    CollapseItem* in = queueA;
    CollapseItem* out = queueB;
    while (nIn > 0) {
        depth++;
        collapseKernel<<<(nIn + T - 1) / T, T, 0, stream>>>(in, nIn, out, counters, int(n), left, right, rangeFirst, rangeLast,
                nodeBox, leafRef, maxRun, io.node4);
        e = cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess)
            return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess)
            return e;
        nIn = host[1];
        e = cudaMemsetAsync(counters + 1, 0, sizeof(uint32_t), stream);
        if (e != cudaSuccess)
            return e;
        CollapseItem* tmp = in; in = out; out = tmp;
        if (depth > 4096)
            return cudaErrorLaunchFailure;
    }
    *wideNodes = host[0];
    *wideDepth = depth;
    *triangles = host[4];
    return cudaGetLastError();
}

}
