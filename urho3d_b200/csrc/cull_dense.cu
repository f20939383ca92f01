// Octree query + occlusion culling for a whole batch of views as dense, barrier-free launches.  Same results as cull.cu's
// one-CTA-per-view kernel (which stays for single views and as the on-device fallback when this path's scratch overflows):
//   Octree::GetDrawables(OctreeQuery&) -> Octant::GetDrawablesInternal        (Octree.cpp:495-499, 229-255)
//   FrustumOctreeQuery / OccludedFrustumOctreeQuery / ShadowCasterOctreeQuery (OctreeQuery.cpp:98-118, View.cpp:116-158, 59-84)
//   CheckVisibilityWork's occlusion predicate + the in-view stage             (View.cpp:177-211)
//
// The work of ALL views is pooled, so no launch depends on how much one view has to do:
//   k_dense_seed      one thread per view: the root (never tested by OctreeQuery walks, Octree.cpp:231) and its children
//   k_dense_level     per octree level: a global queue of (view, octant, inside) pairs, 32 per warp, TestOctant on dense warps;
//                     survivors record their state and append their children to the next level's part of the queue
//   k_dense_groups    per view and group of 32 octants (DFS order): result words (32 drawables each) the group's surviving octants need
//   k_dense_scan      per view: exclusive scan of those -> every octant's drawables own a run of words in DFS order
//   k_dense_words     one warp per surviving group: describes every word (first DFS slot, view, valid drawables, inside flag)
//   k_dense_frustum   flat over the words: TestDrawables, 32 drawables per warp step (two coalesced float4 loads each)
//   k_dense_occlusion flat over the words: the drawables that still need the occlusion predicate, handed out 32 at a time
//   k_dense_emit      per view: scan of the words' popcounts -> ids in result_ order (= ascending DFS slot)
// No block-wide barrier anywhere except the emit kernel's scan; boxes with large screen rects are walked by the whole warp.
#include "u3d_device.cuh"
#include "u3d_kernels.h"

#include <algorithm>

namespace u3d
{

// Block size of the persistent / grid-stride kernels.  Small blocks: a block's SM resources are released when its LAST warp is done, and
// with several batches in flight on their own streams the next launch's blocks move in as soon as they are (measured: see profiles).
#ifndef U3D_DENSE_THREADS
#define U3D_DENSE_THREADS 128 // A/B at 4096 views, 8 batches in flight: 622 k views/s with 256, 652 k with 128 or 64
#endif
constexpr int kDenseThreads = U3D_DENSE_THREADS;
constexpr int kDenseWarps = kDenseThreads / 32;
#ifndef U3D_TREE_UNITS
#define U3D_TREE_UNITS 4 // 64 registers (A/B, 4096 views: 663 k views/s at 85 registers, 698 k at 64)
#endif
constexpr int kDenseMinBlocks = U3D_TREE_UNITS * 256 / kDenseThreads; // the octree kernels: 1024 threads (<= 64 registers) per SM
constexpr int kEmitThreads = 256, kEmitWarps = kEmitThreads / 32; // k_dense_emit: one block per view
constexpr int kDenseWStack = 128;  // per-warp texel stack of the cooperative range test

__device__ int g_denseHeavyExtent = 64; // rect extent (pixels) from which a box is walked by the whole warp (tuning hook "dense_heavy")

__device__ int g_denseWalkPool = 0;     // 1: the boxes of a warp share one texel stack (range_visible_pool) instead of one walk per lane ("dense_pool")

// Diagnostic ("cull_prof" option): per kernel family k (0 = levels, 1 = occlusion) [4k] earliest warp start, [4k+1] latest warp end (ns,
// globaltimer), [4k+2] sum of the warps' busy times, [4k+3] warps.  sum / (warps * (end - start)) = how evenly the warps finish.
__device__ unsigned long long g_denseProf[8 + 4 * 32]; // [8 + 4 * level ...]: the same four numbers per octree level
__device__ int g_denseProfOn;
__device__ __forceinline__ unsigned long long dense_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void dense_prof_end(int k, unsigned long long t0)
{
    if (g_denseProfOn && (threadIdx.x & 31) == 0)
    {
        const unsigned long long t1 = dense_now();
        atomicMin(&g_denseProf[4 * k], t0);
        atomicMax(&g_denseProf[4 * k + 1], t1);
        atomicAdd(&g_denseProf[4 * k + 2], t1 - t0);
        atomicAdd(&g_denseProf[4 * k + 3], 1ull);
    }
}

enum : int { CTR_LEVEL0 = 0, CTR_GROUPS = 32, CTR_POOL = 33, CTR_FALLBACK = 34, CTR_TICKET_OCC = 35, CTR_TICKET_FRUSTUM = 36, CTR_TICKET_LEVEL0 = 64,
    CTR_WORDS = 128 };

/// Dynamic work distribution for the persistent grids: a warp takes the next chunk of a kernel's work list (one atomic per chunk).
/// Static slices leave the warps that happen to get the expensive boxes running long after the others have finished.
__device__ __forceinline__ uint32_t next_chunk(uint32_t* ticket)
{
    uint32_t c = 0;
    if ((threadIdx.x & 31) == 0)
        c = atomicAdd(ticket, 1u);
    return __shfl_sync(0xffffffffu, c, 0);
}

/// Position of the n-th set bit (n counts from 0; the mask has more than n bits set).
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t m, uint32_t n)
{
    uint32_t pos = 0;
#pragma unroll
    for (int half = 16; half; half >>= 1)
    {
        const uint32_t c = (uint32_t)__popc(m & ((1u << half) - 1u));
        if (n >= c)
        {
            n -= c;
            pos += half;
            m >>= half;
        }
    }
    return pos;
}

/// OcclusionBuffer::IsVisible for one box per lane, each lane with its own view.  Small rects are walked by their lane, large ones
/// by the whole warp one after another.  Every lane of the warp must call.
__device__ __forceinline__ bool dense_is_visible(const BufGeom& g, const ViewConst* __restrict__ views, const int* __restrict__ depthAll,
    const int2* __restrict__ mipAll, bool useMips, bool need, int v, const float4& mn, const float4& mx, uint32_t* wstack)
{
    const int lane = threadIdx.x & 31;
    BoxRect r;
    r.left = r.top = r.right = r.bottom = r.z = 0;
    bool vis = false, heavy = false, pooled = false;
    const bool pool = useMips && g_denseWalkPool && g.nlev <= 8 && g.mw[0] < 4096;
    if (need)
    {
        const int* depth = depthAll + (size_t)v * g.w * g.h;
        const int2* mips = mipAll + (size_t)v * g.mipTexels;
        if (ob_project(g, views[v].vp, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, r))
            vis = true;
        else if (useMips && max(r.right - r.left, r.bottom - r.top) >= g_denseHeavyExtent)
            heavy = true;
        else if (pool)
            pooled = true;
        else
            vis = range_visible_thread(g, depth, mips, useMips, r);
    }
    if (pool)
    {
        const bool pv = range_visible_pool(g, depthAll, mipAll, v, r, pooled, wstack, kDenseWStack);
        if (pooled)
            vis = pv;
    }
    uint32_t todo = __ballot_sync(0xffffffffu, heavy);
    while (todo)
    {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        BoxRect rr;
        rr.left = __shfl_sync(0xffffffffu, r.left, src);
        rr.top = __shfl_sync(0xffffffffu, r.top, src);
        rr.right = __shfl_sync(0xffffffffu, r.right, src);
        rr.bottom = __shfl_sync(0xffffffffu, r.bottom, src);
        rr.z = __shfl_sync(0xffffffffu, r.z, src);
        const int vv = __shfl_sync(0xffffffffu, v, src);
        const bool rv = range_visible_warp(g, depthAll + (size_t)vv * g.w * g.h, mipAll + (size_t)vv * g.mipTexels, rr, wstack, kDenseWStack);
        if (lane == src)
            vis = rv;
    }
    return vis;
}

__device__ __forceinline__ void dense_mark_alive(const SceneDev& sc, const DenseDev& dd, int v, int o, unsigned char st, unsigned char* __restrict__ stateAll)
{
    stateAll[(size_t)v * dd.stateStride + o] = st;
    if (sc.octCount[o] > 0)
        atomicOr(&dd.aliveBits[(size_t)v * dd.aliveWords + (o >> 10)], 1u << ((o >> 5) & 31));
}

__global__ void __launch_bounds__(kDenseThreads) k_dense_seed(SceneDev sc, const ViewConst* __restrict__ views, int numViews, unsigned char* __restrict__ stateAll,
    uint32_t* __restrict__ outCounts, DenseDev dd)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= numViews)
        return;
    outCounts[2 * v] = 0; // k_dense_frustum adds the query result size up with atomics
    outCounts[2 * v + 1] = 0;
    const ViewConst& vc = views[v];
    // the root is never tested by OctreeQuery walks (Octree.cpp:231); the ray walk tests every octant (Octree.cpp:257-261)
    bool rootAlive = true;
    if (vc.queryMode == QUERY_RAY || vc.queryMode == QUERY_RAY_CANDIDATES)
    {
        const float4 mn = sc.octMin[0], mx = sc.octMax[0];
        rootAlive = shape_test<false>(vc.queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z) != OUTSIDE;
    }
    if (!rootAlive)
        return;
    dense_mark_alive(sc, dd, v, 0, 1, stateAll);
    const int2 ch = sc.octChild[0];
    if (ch.y)
    {
        const uint32_t pos = atomicAdd(&dd.ctr[CTR_LEVEL0 + 1], (uint32_t)ch.y);
        if (pos + (uint32_t)ch.y > dd.queueCap)
        {
            dd.ctr[CTR_FALLBACK] = 1;
            return;
        }
        for (int j = 0; j < ch.y; ++j)
            dd.queue[pos + j] = make_uint2((uint32_t)v, (uint32_t)sc.bfs[ch.x + j]);
    }
}

/// One warp per (view, block of 1024 octants), one lane per group of 32 octants: result words the group's surviving octants need.
/// Groups with work are appended to the group list.
__global__ void __launch_bounds__(kDenseThreads) k_dense_groups(SceneDev sc, int numViews, const unsigned char* __restrict__ stateAll, DenseDev dd)
{
    if (dd.ctr[CTR_FALLBACK])
        return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t totalWarps = gridDim.x * kDenseWarps;
    const uint32_t items = (uint32_t)numViews * dd.aliveWords;
    for (uint32_t it = blockIdx.x * kDenseWarps + wid; it < items; it += totalWarps)
    {
        const int v = (int)(it / dd.aliveWords), blk = (int)(it % dd.aliveWords);
        const uint32_t bits = dd.aliveBits[(size_t)v * dd.aliveWords + blk];
        const int grp = blk * 32 + lane;
        uint32_t words = 0;
        if ((bits >> lane) & 1u)
        {
            const unsigned char* st = stateAll + (size_t)v * dd.stateStride + grp * 32;
            const uint4 s0 = *reinterpret_cast<const uint4*>(st), s1 = *reinterpret_cast<const uint4*>(st + 16);
            const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if ((sw[j >> 2] >> ((j & 3) * 8)) & 0xFFu)
                    words += ((uint32_t)sc.octCount[grp * 32 + j] + 31u) >> 5;
        }
        if (grp < dd.groupsPerView)
            dd.groupCount[(size_t)v * dd.groupsPerView + grp] = words;
        const uint32_t bal = __ballot_sync(0xffffffffu, words != 0);
        if (bal)
        {
            uint32_t pos = 0;
            if (lane == 0)
                pos = atomicAdd(&dd.ctr[CTR_GROUPS], (uint32_t)__popc(bal));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (words)
                dd.groupList[pos + __popc(bal & ((1u << lane) - 1u))] = make_uint2((uint32_t)v, (uint32_t)grp);
        }
    }
}

/// One warp per view: exclusive scan of the groups' word counts; the view's words get a region of the pool.
__global__ void __launch_bounds__(kDenseThreads) k_dense_scan(int numViews, DenseDev dd)
{
    if (dd.ctr[CTR_FALLBACK])
        return;
    const int lane = threadIdx.x & 31;
    const int v = (blockIdx.x * kDenseThreads + threadIdx.x) >> 5;
    if (v >= numViews)
        return;
    const uint32_t* cnt = dd.groupCount + (size_t)v * dd.groupsPerView;
    uint32_t* offs = dd.groupOff + (size_t)v * dd.groupsPerView;
    uint32_t carry = 0;
    for (int b = 0; b < dd.groupsPerView; b += 32)
    {
        const uint32_t c = b + lane < dd.groupsPerView ? cnt[b + lane] : 0u;
        uint32_t x = c;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, x, s);
            if (lane >= s)
                x += t;
        }
        if (b + lane < dd.groupsPerView)
            offs[b + lane] = carry + x - c;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0)
    {
        const uint32_t pos = atomicAdd(&dd.ctr[CTR_POOL], carry);
        dd.wordBase[v] = pos;
        dd.viewWords[v] = carry;
        if ((uint64_t)pos + carry > dd.poolCap)
            dd.ctr[CTR_FALLBACK] = 1;
    }
}

/// One warp per group with work: every surviving octant of the group describes its result words (first DFS slot, view, number of valid
/// drawables, inside flag), so that the drawable tests can run flat over the pool.
__global__ void __launch_bounds__(kDenseThreads) k_dense_words(SceneDev sc, const unsigned char* __restrict__ stateAll, DenseDev dd)
{
    if (dd.ctr[CTR_FALLBACK])
        return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t nGroups = dd.ctr[CTR_GROUPS];
    const uint32_t totalWarps = gridDim.x * kDenseWarps;
    for (uint32_t it = blockIdx.x * kDenseWarps + wid; it < nGroups; it += totalWarps)
    {
        const uint2 item = dd.groupList[it];
        const int v = (int)item.x, grp = (int)item.y;
        const uint32_t wordBase = dd.wordBase[v] + dd.groupOff[(size_t)v * dd.groupsPerView + grp];
        const int o = grp * 32 + lane;
        uint32_t cnt = 0, first = 0, ins = 0;
        if (o < sc.numOctants)
        {
            const unsigned char st = stateAll[(size_t)v * dd.stateStride + o];
            if (st)
            {
                cnt = (uint32_t)sc.octCount[o];
                first = (uint32_t)__float_as_int(sc.octMax[o].w);
                ins = st == 2 ? 1u : 0u;
            }
        }
        uint32_t woff = (cnt + 31u) >> 5;
        {
            const uint32_t mine = woff;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, woff, s);
                if (lane >= s)
                    woff += t;
            }
            woff -= mine;
        }
        const uint32_t common = (uint32_t)v | (ins << 21);
        const uint32_t staticBase = cnt ? sc.octWordStart[o] : 0u;
        if (cnt && cnt <= 256)
            for (uint32_t k = 0; k < cnt; k += 32)
            {
                const uint32_t w = wordBase + woff + (k >> 5);
                dd.wordSlot[w] = first + k;
                dd.wordMeta[w] = common | ((min(cnt - k, 32u) - 1u) << 16);
                dd.wordStatic[w] = staticBase + (k >> 5);
            }
        // octants with many drawables: the whole warp writes their descriptors
        uint32_t big = __ballot_sync(0xffffffffu, cnt > 256);
        while (big)
        {
            const int j = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t bCnt = __shfl_sync(0xffffffffu, cnt, j), bFirst = __shfl_sync(0xffffffffu, first, j);
            const uint32_t bWord = wordBase + __shfl_sync(0xffffffffu, woff, j), bCommon = __shfl_sync(0xffffffffu, common, j);
            const uint32_t bStatic = __shfl_sync(0xffffffffu, staticBase, j);
            for (uint32_t k = lane * 32u; k < bCnt; k += 1024u)
            {
                dd.wordSlot[bWord + (k >> 5)] = bFirst + k;
                dd.wordMeta[bWord + (k >> 5)] = bCommon | ((min(bCnt - k, 32u) - 1u) << 16);
                dd.wordStatic[bWord + (k >> 5)] = bStatic + (k >> 5);
            }
        }
    }
}

/// FrustumOctreeQuery::TestDrawables (OctreeQuery.cpp:106-118) / ShadowCasterOctreeQuery (View.cpp:70-82) / ZoneOccluderOctreeQuery
/// (View.cpp:99-112) flat over the pool: one word = 32 consecutive drawables of one surviving octant of one view per warp step, two coalesced
/// float4 loads per drawable.  Writes, per word, the drawables that are in the result without a buffer test and those that still need
/// CheckVisibilityWork's occlusion predicate (k_dense_occlusion); the in-view stage's draw-distance test (View.cpp:179-186) is applied here.
// BULK = true: the 32 boxes of a word (two runs of <= 512 contiguous bytes, one per array) arrive in shared memory by two bulk
// asynchronous copies (TMA, cp.async.bulk + mbarrier complete_tx) issued one word ahead by lane 0, double-buffered per warp; the lanes
// then read their box with two 16-byte shared loads.  BULK = false: two coalesced 16-byte global loads per lane.  A/B: profiles/README.md.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool BULK> __global__ void __launch_bounds__(kDenseThreads) k_dense_frustum(SceneDev sc, const ViewConst* __restrict__ views, int stage,
    uint32_t* __restrict__ outCounts, DenseDev dd)
{
    __shared__ __align__(16) float4 sBox[BULK ? kDenseWarps : 1][2][2][32]; // [warp][buffer][min / max][lane]
    __shared__ __align__(8) unsigned long long sBar[BULK ? kDenseWarps : 1][2];
    if (dd.ctr[CTR_FALLBACK])
        return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t total = dd.ctr[CTR_POOL];
    const uint32_t totalWarps = gridDim.x * kDenseWarps;
    const bool inViewStage = stage == STAGE_INVIEW;
    uint32_t phase[2] = {0u, 0u};
    if (BULK)
    {
        if (lane == 0)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sBar[wid][0])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sBar[wid][1])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    // lane 0: both copies of one word into buffer `buf`
    auto fetch = [&](int buf, uint32_t s0, uint32_t nValid) {
        const uint32_t bytes = nValid * 16u;
        const uint32_t bar = smem_addr(&sBar[wid][buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(&sBox[wid][buf][0][0])),
                     "l"(sc.drwMin + s0), "r"(bytes), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(&sBox[wid][buf][1][0])),
                     "l"(sc.drwMax + s0), "r"(bytes), "r"(bar)
                     : "memory");
    };
    for (uint32_t it = blockIdx.x * kDenseWarps + wid; it * 32u < total; it += totalWarps)
    {
        const uint32_t w0 = it * 32u, myW = w0 + lane;
        uint32_t meta = 0, slot0 = 0;
        if (myW < total)
        {
            meta = dd.wordMeta[myW];
            slot0 = dd.wordSlot[myW];
        }
        const int nW = (int)min(32u, total - w0);
        uint32_t myVis = 0, myTest = 0, myQuery = 0;
        // Words of octants that are INSIDE the query volume: every drawable passes the shape test, so the word's result follows from
        // its static summary (flags / view mask if uniform over the word, occludee and shadow-caster bits) -- one 16-byte load by the
        // word's own lane instead of 2 x 512 bytes of boxes and a warp step.  (Not for the in-view stage, which needs per-drawable
        // distances, nor for the zone / occluder query, whose flag test reads the occluder bit.)
        bool fast = false;
        if (!BULK && myW < total && !inViewStage && ((meta >> 21) & 1u))
        {
            const ViewConst& vcw = views[meta & 0xFFFFu];
            const uint4 info = sc.wordInfo[dd.wordStatic[myW]];
            if (info.x != 0xFFFFFFFFu && vcw.queryMode != QUERY_ZONE_OCCLUDER)
            {
                const uint32_t nv = ((meta >> 16) & 31u) + 1u;
                uint32_t q = ((info.x & vcw.drawableFlags) != 0u && (info.y & vcw.viewMask) != 0u) ? (nv == 32u ? 0xFFFFFFFFu : (1u << nv) - 1u) : 0u;
                if (vcw.queryMode == QUERY_SHADOWCASTER)
                    q &= info.w;
                myTest = (stage >= STAGE_VISIBLE && vcw.useOcclusion != 0) ? (q & info.z) : 0u;
                myVis = q & ~myTest;
                myQuery = (uint32_t)__popc(q);
                fast = true;
            }
        }
        if (BULK)
        {
            __syncwarp(); // every lane is done with both buffers
            const uint32_t m0 = __shfl_sync(0xffffffffu, meta, 0), f0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (lane == 0)
                fetch(0, f0, ((m0 >> 16) & 31u) + 1u);
        }
        const uint32_t slow = __ballot_sync(0xffffffffu, myW < total && !fast);
        for (int j = 0; j < nW; ++j)
        {
            if (!BULK && !((slow >> j) & 1u))
                continue;
            const uint32_t m = __shfl_sync(0xffffffffu, meta, j), s0 = __shfl_sync(0xffffffffu, slot0, j);
            if (BULK)
            {
                const uint32_t mNext = __shfl_sync(0xffffffffu, meta, (j + 1) & 31), sNext = __shfl_sync(0xffffffffu, slot0, (j + 1) & 31);
                __syncwarp(); // buffer (j + 1) & 1 was read two words ago by every lane
                if (lane == 0 && j + 1 < nW)
                    fetch((j + 1) & 1, sNext, ((mNext >> 16) & 31u) + 1u);
                const uint32_t bar = smem_addr(&sBar[wid][j & 1]);
                uint32_t done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase[j & 1]) : "memory");
                phase[j & 1] ^= 1u;
            }
            const ViewConst& vc = views[m & 0xFFFFu];
            const int queryMode = vc.queryMode;
            const bool sInside = (m >> 21) & 1u;
            const uint32_t d = s0 + lane;
            bool pass = false, test = false, vis0 = false;
            if ((uint32_t)lane <= ((m >> 16) & 31u))
            {
                const float4 mn = BULK ? sBox[wid][j & 1][0][lane] : sc.drwMin[d];
                const float4 mx = BULK ? sBox[wid][j & 1][1][lane] : sc.drwMax[d];
                const uint32_t dm = __float_as_uint(mn.w);
                const bool flagsOk = queryMode != QUERY_ZONE_OCCLUDER ? (dm & 0xFFu & vc.drawableFlags) != 0
                                                                      : ((dm & 0xFFu) == 4u || ((dm & 0xFFu) == 1u && (dm & kMetaOccluder)));
                if (flagsOk && (__float_as_uint(mx.w) & vc.viewMask) && (queryMode != QUERY_SHADOWCASTER || (dm & kMetaCastShadows)))
                    pass = sInside || shape_test<true>(queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z) != OUTSIDE;
                if (pass)
                {
                    // !buffer || !drawable->IsOccludee() || buffer->IsVisible(box), View.cpp:177: the last term is k_dense_occlusion's
                    if (stage >= STAGE_VISIBLE && vc.useOcclusion != 0 && (dm & kMetaOccludee))
                        test = true;
                    else
                        vis0 = true;
                    if (inViewStage)
                    {
                        // Drawable::UpdateBatches distance (Drawable.cpp:134-137) and the draw-distance test (View.cpp:179-186)
                        const float maxDistance = sc.drwDrawDist[d];
                        if (maxDistance > 0.0f)
                        {
                            const float cx = (mx.x + mn.x) * 0.5f, cy = (mx.y + mn.y) * 0.5f, cz = (mx.z + mn.z) * 0.5f; // BoundingBox::Center
                            float dist;
                            if (!vc.ortho)
                            {
                                const float dx = cx - vc.camPos[0], dy = cy - vc.camPos[1], dz = cz - vc.camPos[2];
                                dist = sqrtf(dx * dx + dy * dy + dz * dz); // (worldPos - cameraPos).Length(), Camera.cpp:487-493
                            }
                            else
                                // Abs((GetView() * worldPos).z_), Camera.cpp:495, in the operation order of the reference's default (URHO3D_SSE)
                                // build of Matrix3x4 * Vector3 (Matrix3x4.h:256-274): (m20*x + m22*z) + (m21*y + m23)
                                dist = fabsf((vc.viewZ[0] * cx + vc.viewZ[2] * cz) + (vc.viewZ[1] * cy + vc.viewZ[3]));
                            if (dist > maxDistance)
                            {
                                vis0 = false;
                                test = false;
                            }
                        }
                    }
                }
            }
            const uint32_t q = __ballot_sync(0xffffffffu, pass);
            const uint32_t v0 = __ballot_sync(0xffffffffu, vis0);
            const uint32_t tb = __ballot_sync(0xffffffffu, test);
            if (lane == j)
            {
                myVis = v0;
                myTest = tb;
                myQuery = (uint32_t)__popc(q);
            }
        }
        if (myW < total)
        {
            dd.visWord[myW] = myVis;
            dd.testMask[myW] = myTest;
        }
        // query result size per view (counts[v][0]): the words of a step nearly always belong to one view
        const uint32_t myView = meta & 0xFFFFu;
        const uint32_t firstView = __shfl_sync(0xffffffffu, myView, 0);
        if (__all_sync(0xffffffffu, myW >= total || myView == firstView))
        {
            const uint32_t sum = __reduce_add_sync(0xffffffffu, myQuery);
            if (lane == 0 && sum)
                atomicAdd(&outCounts[2 * firstView], sum);
        }
        else if (myQuery)
            atomicAdd(&outCounts[2 * myView], myQuery);
    }
}

/// CheckVisibilityWork's occlusion predicate (View.cpp:177) over the drawables k_dense_frustum left to it.  A warp step takes 32 words,
/// and hands their marked drawables out 32 at a time (candidate c belongs to the word whose running count covers c), so that
/// OcclusionBuffer::IsVisible always runs on full warps.
__global__ void __launch_bounds__(kDenseThreads) k_dense_occlusion(SceneDev sc, BufGeom g, const ViewConst* __restrict__ views, const int* __restrict__ depthAll,
    const int2* __restrict__ mipAll, int mipsValid, DenseDev dd)
{
    __shared__ uint32_t wstack[kDenseWarps][kDenseWStack];
    if (dd.ctr[CTR_FALLBACK])
        return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool useMips = mipsValid != 0;
    const uint32_t total = dd.ctr[CTR_POOL];
    const uint32_t totalWarps = gridDim.x * kDenseWarps;
    for (uint32_t it = blockIdx.x * kDenseWarps + wid; it * 32u < total; it += totalWarps)
    {
        const uint32_t myW = it * 32u + lane;
        uint32_t mask = 0, view = 0, slot0 = 0;
        if (myW < total)
            mask = dd.testMask[myW];
        if (!__any_sync(0xffffffffu, mask != 0))
            continue;
        if (mask)
        {
            view = dd.wordMeta[myW] & 0xFFFFu;
            slot0 = dd.wordSlot[myW];
        }
        const uint32_t cnt = (uint32_t)__popc(mask);
        uint32_t excl = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, excl, s);
            if (lane >= s)
                excl += t;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, excl, 31);
        excl -= cnt;
        for (uint32_t c0 = 0; c0 < tot; c0 += 32)
        {
            const uint32_t c = min(c0 + lane, tot - 1);
            const bool have = c0 + lane < tot;
            // owner = the last lane whose running count is <= c (it has drawables left: the next lane's count is beyond c)
            int owner = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1)
            {
                const uint32_t e = __shfl_sync(0xffffffffu, excl, (owner + step) & 31);
                if (owner + step < 32 && e <= c)
                    owner += step;
            }
            const uint32_t oMask = __shfl_sync(0xffffffffu, mask, owner), oExcl = __shfl_sync(0xffffffffu, excl, owner);
            const uint32_t oSlot = __shfl_sync(0xffffffffu, slot0, owner), oView = __shfl_sync(0xffffffffu, view, owner);
            const uint32_t bit = nth_set_bit(oMask, c - oExcl);
            float4 mn = make_float4(0.f, 0.f, 0.f, 0.f), mx = mn;
            if (have)
            {
                mn = sc.drwMin[oSlot + bit];
                mx = sc.drwMax[oSlot + bit];
            }
            const bool vis = dense_is_visible(g, views, depthAll, mipAll, useMips, have, (int)oView, mn, mx, wstack[wid]);
            if (have && vis)
                atomicOr(&dd.visWord[it * 32u + owner], 1u << bit);
        }
    }
}

// -----------------------------------------------------------------------------------------------------------------------------
// OcclusionBuffer::IsVisible (OB.cpp:336-456) for a stream of boxes, in two tiers.
//   Tier 1, dense: project the box, then classify ALL texels of the <= 3 x 3 grid that covers its rect one level below the level whose
//   texels match the rect's size -- nine independent loads, no stack, no divergence.  On the benchmark scene this decides 80 % of the
//   boxes (a texel is conclusive unless min < z <= max, and a mixed texel inside the rect proves visibility).
//   Tier 2: the rest wait in a per-warp buffer with the texels that still have to be refined, and are walked 32 at a time, one per
//   lane (depth-first over the border-cut mixed texels only; mean 5 steps).
// Rects whose first grid is larger than 3 x 3 (start level clamped to the coarsest one) are walked by the whole warp.
// -----------------------------------------------------------------------------------------------------------------------------
constexpr int kWalkBuf = 64;      // buffered boxes per warp (drained whenever 32 are waiting)
constexpr int kWalkStack = 64;    // texels per lane: <= 9 to start with, + 3 per finer level

struct WalkShared
{
    uint32_t lt[kDenseWarps][kWalkBuf], rb[kDenseWarps][kWalkBuf], z[kDenseWarps][kWalkBuf], meta[kDenseWarps][kWalkBuf], tag[kDenseWarps][kWalkBuf];
    uint32_t wstack[kDenseWarps][kDenseWStack];
};

/// Level of the first grid: one below the level whose texels match the rect's size (so the grid is at most 3 x 3), or the coarsest level
/// for rects larger than its texels (the grid is then at most 8 x 8: the coarsest level is, OB.cpp:97-106).
__device__ __forceinline__ int walk_top_level(const BufGeom& g, const BoxRect& r)
{
    const int extent = max(r.right - r.left, r.bottom - r.top) + 1;
    const int lv = extent <= 2 ? 0 : (32 - __clz(extent - 1)) - 1;
    return min(max(lv - 1, 0), g.nlev - 1);
}

constexpr uint32_t kWalkBigGrid = 1u << 9; // in the mask: the first grid is larger than 3 x 3, tier 2 starts from its border texels

/// Tier 1 for one rect.  Returns 1 = visible, 0 = nothing visible, 2 = undecided with `mask` = the texels (bit j * 3 + i of the grid at
/// walk_top_level(r)) that must be refined, or kWalkBigGrid.
__device__ __forceinline__ int walk_top(const BufGeom& g, const int2* __restrict__ mips, const BoxRect& r, uint32_t& mask)
{
    const int lv = walk_top_level(g, r), s = lv + 1;
    const int rx0 = r.left >> s, ry0 = r.top >> s;
    const int nx = (r.right >> s) - rx0 + 1, ny = (r.bottom >> s) - ry0 + 1;
    mask = 0;
    const int2* row = mips + g.moff[lv] + ry0 * g.mw[lv] + rx0;
    const int stride = g.mw[lv];
    bool vis = false;
    if (nx > 3 || ny > 3)
    {
        // Large rect on the coarsest level: a texel inside the rect shows the box unless even its farthest pixel is nearer (z > max);
        // a texel cut by the rect's border shows it when all its pixels are at least as far (z <= min) and must be refined when mixed.
        bool cut = false;
        for (int j = 0; j < ny; ++j)
        {
            const bool rIn = ((ry0 + j) << s) >= r.top && min(((ry0 + j + 1) << s) - 1, g.h - 1) <= r.bottom;
            for (int i = 0; i < nx; ++i)
            {
                const int2 t = row[j * stride + i];
                const bool in = rIn && ((rx0 + i) << s) >= r.left && min(((rx0 + i + 1) << s) - 1, g.w - 1) <= r.right;
                vis |= r.z <= (in ? t.y : t.x);
                cut |= !in && r.z > t.x && r.z <= t.y;
            }
            if (vis)
                return 1;
        }
        mask = cut ? kWalkBigGrid : 0u;
        return cut ? 2 : 0;
    }
    // Nine loads (indices clamped into the grid, duplicates hit L1) and nine-bit masks instead of branches: bit j * 3 + i = texel (i, j).
    //   valid: the texel belongs to the grid (i < nx, j < ny)
    //   in:    its pixels all lie inside the rect -- its column and its row do (the last column / row is cut by the buffer's edge)
    //   nearer: z <= min (every pixel under it is at least as far as the box), mixed: min < z <= max
    // A mixed texel inside the rect proves visibility (the pixel that holds its max lies in the rect); mixed texels cut by the border are
    // the ones that have to be refined.
    uint32_t colIn = 0, rowIn = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        colIn |= (((rx0 + i) << s) >= r.left && min(((rx0 + i + 1) << s) - 1, g.w - 1) <= r.right) ? (1u << i) : 0u;
        rowIn |= (((ry0 + i) << s) >= r.top && min(((ry0 + i + 1) << s) - 1, g.h - 1) <= r.bottom) ? (1u << i) : 0u;
    }
    auto spread = [](uint32_t rows) { return (rows & 1u) | ((rows & 2u) << 2) | ((rows & 4u) << 4); }; // row bit j -> bit 3 j
    const uint32_t valid = ((1u << nx) - 1u) * spread((1u << ny) - 1u);
    const uint32_t in = colIn * spread(rowIn);
    const int c1 = nx > 1 ? 1 : 0, c2 = nx > 2 ? 2 : nx - 1;
    const int r1 = ny > 1 ? stride : 0, r2 = ny > 2 ? 2 * stride : (ny - 1) * stride;
    const int colOff[3] = {0, c1, c2}, rowOff[3] = {0, r1, r2};
    uint32_t nearer = 0, reach = 0; // reach: z <= max
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            const int2 t = row[rowOff[j] + colOff[i]];
            nearer |= r.z <= t.x ? (1u << (j * 3 + i)) : 0u;
            reach |= r.z <= t.y ? (1u << (j * 3 + i)) : 0u;
        }
    const uint32_t mixed = reach & ~nearer;
    if (valid & (nearer | (mixed & in)))
        return 1;
    mask = valid & mixed & ~in;
    return mask ? 2 : 0;
}

/// Tier 2 for one rect, one lane: depth-first refinement of the texels in `mask`, at most `budget` steps.  Returns 1 = visible,
/// 0 = nothing visible, 2 = budget spent with `sp` texels left on `stack` (the warp finishes those together, walk_drain).
__device__ int g_walkBudgetBusy = 96;  // steps a lane walks alone while the launch has plenty of other work ("walk_budget_busy")
__device__ int g_walkBudgetIdle = 8;   // ... and in launches with little work, where a long walk is what everybody waits for ("walk_budget_idle")
__device__ __forceinline__ int walk_deep(const BufGeom& g, const int* __restrict__ depth, const int2* __restrict__ mips, const BoxRect& r, uint32_t mask,
    uint32_t* stack, int& sp, int budget)
{
    const int lv0 = walk_top_level(g, r), s0 = lv0 + 1;
    const int rx0 = r.left >> s0, ry0 = r.top >> s0;
    sp = 0;
    if (mask & kWalkBigGrid)
    {
        // the border texels of the first grid (<= 28 of 8 x 8); they are classified again when popped
        const int nx = (r.right >> s0) - rx0 + 1, ny = (r.bottom >> s0) - ry0 + 1;
        for (int j = 0; j < ny; ++j)
        {
            const bool rIn = ((ry0 + j) << s0) >= r.top && min(((ry0 + j + 1) << s0) - 1, g.h - 1) <= r.bottom;
            for (int i = 0; i < nx; ++i)
                if (!(rIn && ((rx0 + i) << s0) >= r.left && min(((rx0 + i + 1) << s0) - 1, g.w - 1) <= r.right))
                    stack[sp++] = pack_texel(lv0, rx0 + i, ry0 + j);
        }
        if (sp > budget)
            return 2; // a long border: straight to the warp
    }
    else
    {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if ((mask >> (j * 3 + i)) & 1u)
                    stack[sp++] = pack_texel(lv0, rx0 + i, ry0 + j) | 0x80000000u; // bit 31: known to be mixed and cut by the rect's border
    }
    for (int step = 0; sp; ++step)
    {
        if (step == budget)
            return 2;
        const uint32_t e = stack[--sp];
        const int lv = (int)((e >> 27) & 15u), x = (int)((e >> 14) & 0x1FFF), y = (int)(e & 0x3FFF);
        const int c = (e >> 31) ? 2 : classify_texel(g, mips, r, lv, x, y);
        if (c == 1)
            return 1;
        if (c == 0)
            continue;
        if (lv == 0)
        {
            if (texel_pixels_visible(g, depth, r, x, y))
                return 1;
        }
        else
        {
            // the <= 2 x 2 children that overlap the rect (and exist: the finer level may end one texel earlier), without loops
            const int cl = lv - 1;
            const int xLo = r.left >> lv, xHi = min(r.right >> lv, g.mw[cl] - 1), yLo = r.top >> lv, yHi = min(r.bottom >> lv, g.mh[cl] - 1);
            const bool xa = 2 * x >= xLo && 2 * x <= xHi, xb = 2 * x + 1 >= xLo && 2 * x + 1 <= xHi;
            const bool ya = 2 * y >= yLo && 2 * y <= yHi, yb = 2 * y + 1 >= yLo && 2 * y + 1 <= yHi;
            const uint32_t base = pack_texel(cl, 2 * x, 2 * y);
            if (xa && ya) stack[sp++] = base;
            if (xb && ya) stack[sp++] = base + (1u << 14);
            if (xa && yb) stack[sp++] = base + 1u;
            if (xb && yb) stack[sp++] = base + (1u << 14) + 1u;
        }
    }
    return 0;
}

/// Tier 1 for one box per lane with `need`: returns the lane's state 0 / 1 (decided) or 2 (undecided, `r` and `mask` valid).
__device__ __forceinline__ int walk_project(const BufGeom& g, const ViewConst* __restrict__ views, const int2* __restrict__ mipAll, bool need, uint32_t view,
    const float4& mn, const float4& mx, BoxRect& r, uint32_t& mask)
{
    r.left = r.top = r.right = r.bottom = r.z = 0;
    mask = 0;
    int state = 0;
    if (need)
    {
        if (ob_project(g, views[view].vp, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, r))
            state = 1;
        else
            state = walk_top(g, mipAll + (size_t)view * g.mipTexels, r, mask);
    }
    return state;
}

/// Lanes with `und` append their box to the warp's buffer.  meta = view | mask << 16 (10 bits) | user bits << 26.
__device__ __forceinline__ void walk_push(WalkShared& sh, int wid, int& buffered, bool und, const BoxRect& r, uint32_t meta, uint32_t tag)
{
    const uint32_t bal = __ballot_sync(0xffffffffu, und);
    if (und)
    {
        const int p = buffered + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
        sh.lt[wid][p] = (uint32_t)r.left | ((uint32_t)r.top << 16);
        sh.rb[wid][p] = (uint32_t)r.right | ((uint32_t)r.bottom << 16);
        sh.z[wid][p] = (uint32_t)r.z;
        sh.meta[wid][p] = meta;
        sh.tag[wid][p] = tag;
    }
    buffered += __popc(bal);
    __syncwarp();
}

/// Tier 2 for up to 32 buffered boxes, one per lane.  Returns true on lanes that held a box (`meta`, `tag` = what walk_push stored), and
/// the answer in `vis`.
__device__ __forceinline__ bool walk_drain(const BufGeom& g, const int* __restrict__ depthAll, const int2* __restrict__ mipAll, WalkShared& sh, int wid,
    int& buffered, uint32_t& meta, uint32_t& tag, bool& vis, int budget)
{
    const int lane = threadIdx.x & 31;
    const int take = min(buffered, 32);
    buffered -= take;
    const bool had = lane < take;
    vis = false;
    meta = tag = 0;
    BoxRect r;
    r.left = r.top = r.right = r.bottom = r.z = 0;
    uint32_t stack[kWalkStack];
    int sp = 0, state = 0;
    if (had)
    {
        const int p = buffered + lane;
        const uint32_t a = sh.lt[wid][p], b = sh.rb[wid][p];
        r.left = (int)(a & 0xFFFFu);
        r.top = (int)(a >> 16);
        r.right = (int)(b & 0xFFFFu);
        r.bottom = (int)(b >> 16);
        r.z = (int)sh.z[wid][p];
        meta = sh.meta[wid][p];
        tag = sh.tag[wid][p];
        const uint32_t vv = meta & 0xFFFFu;
        state = walk_deep(g, depthAll + (size_t)vv * g.w * g.h, mipAll + (size_t)vv * g.mipTexels, r, (meta >> 16) & 0x3FFu, stack, sp, budget);
        vis = state == 1;
    }
    // walks that ran out of their budget (long borders): the warp finishes them one after another, 32 texels per step
    uint32_t todo = __ballot_sync(0xffffffffu, state == 2);
    while (todo)
    {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        __syncwarp();
        if (lane == src)
            for (int k = 0; k < sp; ++k)
                sh.wstack[wid][k] = stack[k];
        __syncwarp();
        BoxRect rr;
        rr.left = __shfl_sync(0xffffffffu, r.left, src);
        rr.top = __shfl_sync(0xffffffffu, r.top, src);
        rr.right = __shfl_sync(0xffffffffu, r.right, src);
        rr.bottom = __shfl_sync(0xffffffffu, r.bottom, src);
        rr.z = __shfl_sync(0xffffffffu, r.z, src);
        const uint32_t vv = __shfl_sync(0xffffffffu, meta, src) & 0xFFFFu;
        const int n = __shfl_sync(0xffffffffu, sp, src);
        const bool rv = range_visible_warp_from(g, depthAll + (size_t)vv * g.w * g.h, mipAll + (size_t)vv * g.mipTexels, rr, sh.wstack[wid], n, kDenseWStack);
        if (lane == src)
            vis = rv;
    }
    __syncwarp();
    return had;
}

/// Warp-wide: lanes with `alive` record their octant's state and append its children to the next level's part of the queue (one atomic
/// per warp).
__device__ __forceinline__ void level_finalize(const SceneDev& sc, const DenseDev& dd, unsigned char* __restrict__ stateAll, int level, uint32_t nextBase,
    bool alive, int v, int o, bool inside)
{
    const int lane = threadIdx.x & 31;
    int2 ch = make_int2(0, 0);
    if (alive)
    {
        dense_mark_alive(sc, dd, v, o, inside ? 2 : 1, stateAll);
        ch = sc.octChild[o];
    }
    int off = ch.y;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, off, s);
        if (lane >= s)
            off += t;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    if (!total)
        return;
    off -= ch.y;
    uint32_t pos = 0;
    if (lane == 0)
        pos = atomicAdd(&dd.ctr[CTR_LEVEL0 + level + 1], (uint32_t)total);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if ((uint64_t)nextBase + pos + (uint32_t)total > dd.queueCap)
    {
        if (lane == 0)
            dd.ctr[CTR_FALLBACK] = 1;
        return;
    }
    const uint32_t ins = inside ? 0x80000000u : 0u;
    for (int j = 0; j < ch.y; ++j)
        dd.queue[nextBase + pos + off + j] = make_uint2((uint32_t)v, (uint32_t)sc.bfs[ch.x + j] | ins);
}

/// TestOctant for every queued (view, octant) pair of one level; the pairs of level `level` sit behind those of the levels before it.
__global__ void __launch_bounds__(kDenseThreads, kDenseMinBlocks) k_dense_level(SceneDev sc, BufGeom g, const ViewConst* __restrict__ views, const int* __restrict__ depthAll,
    const int2* __restrict__ mipAll, int mipsValid, unsigned char* __restrict__ stateAll, DenseDev dd, int level)
{
    __shared__ WalkShared sh;
    if (dd.ctr[CTR_FALLBACK])
        return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t base = 0;
    for (int l = 1; l < level; ++l)
        base += dd.ctr[CTR_LEVEL0 + l];
    const uint32_t n = dd.ctr[CTR_LEVEL0 + level];
    const uint32_t nextBase = base + n;
    const bool useMips = mipsValid != 0;
    const uint32_t totalWarps = gridDim.x * kDenseWarps;
    const unsigned long long profT0 = g_denseProfOn ? dense_now() : 0ull;
    // Upper levels hold few pairs, and their boxes have large screen rects that a warp walks one after another: spread them over all
    // warps of the grid instead of filling a few warps with 32 each.
    const uint32_t per = n >= totalWarps * 32u ? 32u : max(1u, (n + totalWarps - 1u) / totalWarps);
    const uint32_t nSteps = (n + per - 1u) / per;
    // dynamic chunks of 4 steps once there are enough of them; small levels go one step per warp
    const uint32_t kChunk = nSteps >= totalWarps * 8u ? 4u : 1u;
    const int budget = nSteps >= totalWarps * 4u ? g_walkBudgetBusy : g_walkBudgetIdle;
    int buffered = 0;
    // octants whose visibility needs the deeper walk: meta = view | mask << 16 | inside << 26, tag = octant
    auto drain = [&]() {
        uint32_t meta, tag;
        bool vis;
        const bool had = walk_drain(g, depthAll, mipAll, sh, wid, buffered, meta, tag, vis, budget);
        level_finalize(sc, dd, stateAll, level, nextBase, had && vis, (int)(meta & 0xFFFFu), (int)tag, (meta >> 26) & 1u);
    };
    uint32_t* ticket = &dd.ctr[CTR_TICKET_LEVEL0 + level];
    for (uint32_t chunk = next_chunk(ticket); chunk * kChunk < nSteps; chunk = next_chunk(ticket))
    for (uint32_t it = chunk * kChunk; it < min((chunk + 1u) * kChunk, nSteps); ++it)
    {
        const uint32_t idx = it * per + lane;
        int v = 0, o = 0, res = OUTSIDE;
        bool occluded = false;
        float4 mn = make_float4(0.f, 0.f, 0.f, 0.f), mx = mn;
        if ((uint32_t)lane < per && idx < n)
        {
            const uint2 e = dd.queue[base + idx];
            v = (int)e.x;
            o = (int)(e.y & 0x7FFFFFFFu);
            const ViewConst& vc = views[v];
            mn = sc.octMin[o];
            mx = sc.octMax[o];
            if (e.y >> 31)
                res = INSIDE; // FrustumOctreeQuery::TestOctant, OctreeQuery.cpp:98-104
            else
                res = shape_test<false>(vc.queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z);
            occluded = vc.useOcclusion != 0 && vc.queryMode == QUERY_OCCLUDED;
        }
        // OccludedFrustumOctreeQuery::TestOctant, View.cpp:128-139
        bool undecided = false;
        if (__any_sync(0xffffffffu, occluded && res != OUTSIDE))
        {
            const bool need = occluded && res != OUTSIDE;
            if (useMips)
            {
                BoxRect r;
                uint32_t topMask;
                const int state = walk_project(g, views, mipAll, need, (uint32_t)v, mn, mx, r, topMask);
                if (need && state == 0)
                    res = OUTSIDE;
                undecided = need && state == 2;
                walk_push(sh, wid, buffered, undecided, r, (uint32_t)v | (topMask << 16) | (res == INSIDE ? 1u << 26 : 0u), (uint32_t)o);
            }
            else
            {
                const bool vis = dense_is_visible(g, views, depthAll, mipAll, false, need, v, mn, mx, nullptr);
                if (need && !vis)
                    res = OUTSIDE;
            }
        }
        level_finalize(sc, dd, stateAll, level, nextBase, res != OUTSIDE && !undecided, v, o, res == INSIDE);
        if (buffered >= 32)
            drain();
    }
    if (buffered)
        drain();
    dense_prof_end(0, profT0);
    dense_prof_end(2 + level, profT0);
}

// -----------------------------------------------------------------------------------------------------------------------------
// All octree levels in ONE persistent launch.  A child octant can be tested as soon as its parent has survived -- nothing needs a
// whole level to finish first -- so the (view, octant) pairs of every level go through one global work queue: warps claim 32 slots
// at a time, test what has been published in them, and append the survivors' children behind the queue's tail.  There is one ramp-up
// and one tail for the whole tree instead of one per level, and no launch in between.
//   entry  = (view | epoch << 16, octant | inside << 31): one aligned 8-byte store, so a consumer that sees the epoch sees the
//            payload; the epoch changes every run, stale entries of earlier runs never match (the queue is wiped when it wraps)
//   head   = next slot to claim, tail = next slot to reserve, work = pairs published and not finished yet (children are added
//            before their parent is subtracted): 0 means the walk is complete
// -----------------------------------------------------------------------------------------------------------------------------
enum : int { CTR_QHEAD = 40, CTR_QTAIL = 41, CTR_QWORK = 42 };

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p)
{
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

/// Warp-wide: lanes with `alive` record their octant's state and publish its children behind the queue's tail.
__device__ __forceinline__ void tree_finalize(const SceneDev& sc, const DenseDev& dd, unsigned char* __restrict__ stateAll, uint32_t epochTag, bool alive,
    int v, int o, bool inside, const int2* early = nullptr /* the octant's child range when the caller has loaded it already */)
{
    const int lane = threadIdx.x & 31;
    int2 ch = make_int2(0, 0);
    if (alive)
    {
        dense_mark_alive(sc, dd, v, o, inside ? 2 : 1, stateAll);
        ch = early ? *early : sc.octChild[o];
    }
    int off = ch.y;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, off, s);
        if (lane >= s)
            off += t;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    if (!total)
        return;
    off -= ch.y;
    uint32_t pos = 0;
    if (lane == 0)
    {
        atomicAdd(&dd.ctr[CTR_QWORK], (uint32_t)total); // before this warp reports the parents as finished
        pos = atomicAdd(&dd.ctr[CTR_QTAIL], (uint32_t)total);
    }
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if ((uint64_t)pos + (uint32_t)total > dd.queueCap)
    {
        if (lane == 0)
            dd.ctr[CTR_FALLBACK] = 1;
        return;
    }
    const uint32_t ins = inside ? 0x80000000u : 0u;
    for (int j = 0; j < ch.y; ++j)
        dd.queue[pos + off + j] = make_uint2((uint32_t)v | epochTag, (uint32_t)sc.bfs[ch.x + j] | ins);
}

__global__ void __launch_bounds__(kDenseThreads) k_dense_tree_seed(SceneDev sc, const ViewConst* __restrict__ views, int numViews,
    unsigned char* __restrict__ stateAll, uint32_t* __restrict__ outCounts, DenseDev dd, uint32_t epochTag)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= numViews)
        return;
    outCounts[2 * v] = 0; // k_dense_frustum adds the query result size up with atomics
    outCounts[2 * v + 1] = 0;
    const ViewConst& vc = views[v];
    // the root is never tested by OctreeQuery walks (Octree.cpp:231); the ray walk tests every octant (Octree.cpp:257-261)
    if (vc.queryMode == QUERY_RAY || vc.queryMode == QUERY_RAY_CANDIDATES)
    {
        const float4 mn = sc.octMin[0], mx = sc.octMax[0];
        if (shape_test<false>(vc.queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z) == OUTSIDE)
            return;
    }
    dense_mark_alive(sc, dd, v, 0, 1, stateAll);
    const int2 ch = sc.octChild[0];
    if (ch.y)
    {
        atomicAdd(&dd.ctr[CTR_QWORK], (uint32_t)ch.y);
        const uint32_t pos = atomicAdd(&dd.ctr[CTR_QTAIL], (uint32_t)ch.y);
        if (pos + (uint32_t)ch.y > dd.queueCap)
        {
            dd.ctr[CTR_FALLBACK] = 1;
            return;
        }
        for (int j = 0; j < ch.y; ++j)
            dd.queue[pos + j] = make_uint2((uint32_t)v | epochTag, (uint32_t)sc.bfs[ch.x + j]);
    }
}

__global__ void __launch_bounds__(kDenseThreads, kDenseMinBlocks) k_dense_tree(SceneDev sc, BufGeom g, const ViewConst* __restrict__ views, const int* __restrict__ depthAll,
    const int2* __restrict__ mipAll, int mipsValid, unsigned char* __restrict__ stateAll, DenseDev dd, uint32_t epochTag)
{
    __shared__ WalkShared sh;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool useMips = mipsValid != 0;
    const uint32_t totalLanes = gridDim.x * kDenseThreads;
    const unsigned long long profT0 = g_denseProfOn ? dense_now() : 0ull;
    const int budgetBusy = g_walkBudgetBusy, budgetIdle = g_walkBudgetIdle;
    int budget = budgetIdle;
    int buffered = 0;
    uint32_t finished = 0;   // pairs this warp has finished and not yet subtracted from the work counter (warp-uniform)
    uint32_t claimBase = 0, claimMask = 0;
    auto report = [&]() {
        if (finished && lane == 0)
            atomicSub(&dd.ctr[CTR_QWORK], finished);
        finished = 0;
    };
    // octants whose visibility needs the deeper walk: meta = view | mask << 16 | inside << 26, tag = octant
    auto drain = [&]() {
        uint32_t meta, tag;
        bool vis;
        const bool had = walk_drain(g, depthAll, mipAll, sh, wid, buffered, meta, tag, vis, budget);
        tree_finalize(sc, dd, stateAll, epochTag, had && vis, (int)(meta & 0xFFFFu), (int)tag, (meta >> 26) & 1u);
        finished += (uint32_t)__popc(__ballot_sync(0xffffffffu, had));
    };
    for (;;)
    {
        if (!claimMask)
        {
            uint32_t tail = 0;
            if (lane == 0)
            {
                claimBase = atomicAdd(&dd.ctr[CTR_QHEAD], 32u);
                tail = ld_volatile_u32(&dd.ctr[CTR_QTAIL]);
            }
            claimBase = __shfl_sync(0xffffffffu, claimBase, 0);
            tail = __shfl_sync(0xffffffffu, tail, 0);
            claimMask = 0xffffffffu;
            // with a backlog, a lane may walk long borders alone; when the queue runs dry, a long walk is what everybody waits for
            budget = (tail > claimBase && tail - claimBase > totalLanes) ? budgetBusy : budgetIdle;
        }
        const uint32_t slot = claimBase + lane;
        uint2 e = make_uint2(0u, 0u);
        bool ready = false;
        if (((claimMask >> lane) & 1u) && slot < dd.queueCap)
        {
            e = ld_volatile_u2(dd.queue + slot);
            ready = (e.x & 0xFFFF0000u) == epochTag;
        }
        const uint32_t readyMask = __ballot_sync(0xffffffffu, ready);
        if (!readyMask)
        {
            if (buffered)
            {
                drain();
                continue;
            }
            report();
            uint32_t stop = 0;
            if (lane == 0)
                stop = (ld_volatile_u32(&dd.ctr[CTR_QWORK]) == 0u || ld_volatile_u32(&dd.ctr[CTR_FALLBACK]) != 0u) ? 1u : 0u;
            if (__shfl_sync(0xffffffffu, stop, 0))
                break;
            __nanosleep(100);
            continue;
        }
        claimMask &= ~readyMask;
        int v = 0, o = 0, res = OUTSIDE;
        bool occluded = false;
        float4 mn = make_float4(0.f, 0.f, 0.f, 0.f), mx = mn;
        int2 ch = make_int2(0, 0);
        if (ready)
        {
            v = (int)(e.x & 0xFFFFu);
            o = (int)(e.y & 0x7FFFFFFFu);
            const ViewConst& vc = views[v];
            mn = sc.octMin[o];
            mx = sc.octMax[o];
            ch = sc.octChild[o]; // needed only if the octant survives, but asked for now: its round trip overlaps the tests
            if (e.y >> 31)
                res = INSIDE; // FrustumOctreeQuery::TestOctant, OctreeQuery.cpp:98-104
            else
                res = shape_test<false>(vc.queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z);
            occluded = vc.useOcclusion != 0 && vc.queryMode == QUERY_OCCLUDED;
        }
        // OccludedFrustumOctreeQuery::TestOctant, View.cpp:128-139
        bool undecided = false;
        if (__any_sync(0xffffffffu, occluded && res != OUTSIDE))
        {
            const bool need = occluded && res != OUTSIDE;
            if (useMips)
            {
                BoxRect r;
                uint32_t topMask;
                const int state = walk_project(g, views, mipAll, need, (uint32_t)v, mn, mx, r, topMask);
                if (need && state == 0)
                    res = OUTSIDE;
                undecided = need && state == 2;
                walk_push(sh, wid, buffered, undecided, r, (uint32_t)v | (topMask << 16) | (res == INSIDE ? 1u << 26 : 0u), (uint32_t)o);
            }
            else
            {
                const bool vis = dense_is_visible(g, views, depthAll, mipAll, false, need, v, mn, mx, nullptr);
                if (need && !vis)
                    res = OUTSIDE;
            }
        }
        tree_finalize(sc, dd, stateAll, epochTag, ready && res != OUTSIDE && !undecided, v, o, res == INSIDE, &ch);
        finished += (uint32_t)__popc(__ballot_sync(0xffffffffu, ready && !undecided));
        if (buffered >= 32)
            drain();
        if (finished >= 1024u)
            report();
    }
    dense_prof_end(0, profT0);
}

/// CheckVisibilityWork's occlusion predicate (View.cpp:177) over the drawables k_dense_frustum left to it.
#ifndef U3D_OCC_MINB
#define U3D_OCC_MINB (1024 / U3D_DENSE_THREADS) // 64 registers: the kernel is bound by the latency of its dependent loads, more warps per SM win
                                                 // over fewer spills (A/B, 4096 views: 649 k views/s at 80 registers, 673 k at 64, 667 k at 48)
#endif
__global__ void __launch_bounds__(kDenseThreads, U3D_OCC_MINB) k_dense_occlusion_walk(SceneDev sc, BufGeom g, const ViewConst* __restrict__ views,
    const int* __restrict__ depthAll, const int2* __restrict__ mipAll, DenseDev dd)
{
    __shared__ WalkShared sh;
    if (dd.ctr[CTR_FALLBACK])
        return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t total = dd.ctr[CTR_POOL];
    const uint32_t totalWarps = gridDim.x * kDenseWarps;
    const unsigned long long profT0 = g_denseProfOn ? dense_now() : 0ull;
    const int budget = g_walkBudgetBusy;
    int buffered = 0;
    auto drain = [&]() {
        uint32_t meta, tag;
        bool vis;
        if (walk_drain(g, depthAll, mipAll, sh, wid, buffered, meta, tag, vis, budget) && vis)
            atomicOr(&dd.visWord[tag], 1u << (meta >> 26));
    };
    // words per ticket: 32 (one per lane) when there is plenty of work; small launches hand out 8 at a time -- a ticket is up to 32
    // candidates per word, and with few tickets per warp the last ticket is what the launch ends on
    const uint32_t wpt = total >= totalWarps * 32u * 8u ? 32u : 8u;
    const uint32_t nBlocks = (total + wpt - 1u) / wpt;
    for (uint32_t it = next_chunk(&dd.ctr[CTR_TICKET_OCC]); it < nBlocks; it = next_chunk(&dd.ctr[CTR_TICKET_OCC]))
    {
        const uint32_t myW = it * wpt + lane;
        uint32_t mask = 0, view = 0, slot0 = 0;
        if ((uint32_t)lane < wpt && myW < total)
            mask = dd.testMask[myW];
        if (!__any_sync(0xffffffffu, mask != 0))
            continue;
        if (mask)
        {
            view = dd.wordMeta[myW] & 0xFFFFu;
            slot0 = dd.wordSlot[myW];
        }
        const uint32_t cnt = (uint32_t)__popc(mask);
        uint32_t excl = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, excl, s);
            if (lane >= s)
                excl += t;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, excl, 31);
        excl -= cnt;
        for (uint32_t c0 = 0; c0 < tot; c0 += 32)
        {
            const uint32_t c = min(c0 + lane, tot - 1);
            const bool cand = c0 + lane < tot;
            // owner = the last lane whose running count is <= c (it has drawables left: the next lane's count is beyond c)
            int owner = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1)
            {
                const uint32_t e = __shfl_sync(0xffffffffu, excl, (owner + step) & 31);
                if (owner + step < 32 && e <= c)
                    owner += step;
            }
            const uint32_t oMask = __shfl_sync(0xffffffffu, mask, owner), oExcl = __shfl_sync(0xffffffffu, excl, owner);
            const uint32_t oSlot = __shfl_sync(0xffffffffu, slot0, owner), oView = __shfl_sync(0xffffffffu, view, owner);
            const uint32_t bit = nth_set_bit(oMask, c - oExcl);
            const uint32_t word = it * wpt + owner;
            float4 mn = make_float4(0.f, 0.f, 0.f, 0.f), mx = mn;
            if (cand)
            {
                mn = sc.drwMin[oSlot + bit];
                mx = sc.drwMax[oSlot + bit];
            }
            BoxRect r;
            uint32_t topMask;
            const int state = walk_project(g, views, mipAll, cand, oView, mn, mx, r, topMask);
            if (state == 1)
                atomicOr(&dd.visWord[word], 1u << bit);
            walk_push(sh, wid, buffered, state == 2, r, oView | (topMask << 16) | (bit << 26), word);
            if (buffered >= 32)
                drain();
        }
    }
    if (buffered)
        drain();
    dense_prof_end(1, profT0);
}

/// Per view: ids of the set bits of the view's words in order (= the reference's result_ order), counts, and for the in-view stage the
/// view-space Z range of the geometries (View.cpp:198-211, merged as View.cpp:926-951).
__global__ void __launch_bounds__(kEmitThreads) k_dense_emit(SceneDev sc, const ViewConst* __restrict__ views, int stage, uint32_t* __restrict__ outIdx,
    uint32_t outCap, uint32_t* __restrict__ outCounts, float* __restrict__ zRangeAll, uint32_t* __restrict__ flags, DenseDev dd, uint32_t chunkWords)
{
    __shared__ uint32_t warpTot[kEmitWarps];
    __shared__ float red[2 * kEmitWarps];
    if (dd.ctr[CTR_FALLBACK])
        return;
    // grid (chunks, views): a view with a very long result (a shadow cascade over most of the scene) is emitted by several blocks, each
    // over `chunkWords` of the view's words; a block finds its first output position by counting the bits of the words before its chunk
    const int v = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const ViewConst& vc = views[v];
    const bool inViewStage = stage == STAGE_INVIEW;
    const uint32_t allWords = dd.viewWords[v];
    // equal shares of the view's words (in units of 256), whatever their number
    const uint32_t share = gridDim.x > 1 ? max(256u, ((allWords + gridDim.x - 1) / gridDim.x + 255u) / 256u * 256u) : chunkWords;
    const uint32_t first = blockIdx.x * share;
    if (first >= allWords && blockIdx.x != 0)
        return;
    const bool lastChunk = first + share >= allWords;
    const uint32_t nWords = min(allWords, first + share);
    const uint32_t* words = dd.visWord + dd.wordBase[v];
    const uint32_t* slots = dd.wordSlot + dd.wordBase[v];
    uint32_t* out = outIdx + (size_t)v * outCap;
    uint32_t cursor = 0;
    if (first)
    {
        uint32_t sum = 0;
        for (uint32_t i = tid; i < first; i += kEmitThreads)
            sum += (uint32_t)__popc(words[i]);
#pragma unroll
        for (int s = 16; s; s >>= 1)
            sum += __shfl_xor_sync(0xffffffffu, sum, s);
        if (lane == 0)
            warpTot[wid] = sum;
        __syncthreads();
        for (int w = 0; w < kEmitWarps; ++w)
            cursor += warpTot[w];
        __syncthreads();
    }
    float zMin = __int_as_float(0x7f800000), zMax = 0.0f; // PerThreadSceneResult::minZ_/maxZ_ start values, View.cpp:893-894
    for (uint32_t base = first; base < nWords; base += kEmitThreads)
    {
        const uint32_t i = base + tid;
        uint32_t word = i < nWords ? words[i] : 0u;
        const uint32_t c = (uint32_t)__popc(word);
        // block-wide exclusive scan
        uint32_t x = c;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, x, s);
            if (lane >= s)
                x += t;
        }
        __syncthreads(); // warpTot of the previous round has been read
        if (lane == 31)
            warpTot[wid] = x;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kEmitWarps; ++w)
        {
            const uint32_t t = warpTot[w];
            if (w < wid)
                pre += t;
            tot += t;
        }
        uint32_t pos = cursor + pre + x - c;
        if (word)
        {
            const uint32_t s0 = slots[i];
            while (word)
            {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                const uint32_t d = s0 + (uint32_t)b;
                if (pos < outCap)
                    out[pos] = sc.drwId[d];
                ++pos;
                if (inViewStage)
                {
                    const float4 mn = sc.drwMin[d];
                    if (__float_as_uint(mn.w) & 1u) // DRAWABLE_GEOMETRY
                    {
                        const float4 mx = sc.drwMax[d];
                        const float cx = (mx.x + mn.x) * 0.5f, cy = (mx.y + mn.y) * 0.5f, cz = (mx.z + mn.z) * 0.5f;
                        const float ex = (mx.x - mn.x) * 0.5f, ey = (mx.y - mn.y) * 0.5f, ez = (mx.z - mn.z) * 0.5f;
                        if (ex * ex + ey * ey + ez * ez < 100000000.0f * 100000000.0f) // M_LARGE_VALUE, MathDefs.h:55
                        {
                            const float viewCenterZ = vc.viewZ[0] * cx + vc.viewZ[1] * cy + vc.viewZ[2] * cz + vc.viewZ[3];
                            const float viewEdgeZ = fabsf(vc.viewZ[0]) * ex + fabsf(vc.viewZ[1]) * ey + fabsf(vc.viewZ[2]) * ez;
                            const float lo = viewCenterZ - viewEdgeZ, hi = viewCenterZ + viewEdgeZ;
                            zMin = lo < zMin ? lo : zMin;
                            zMax = hi > zMax ? hi : zMax;
                        }
                    }
                }
            }
        }
        cursor += tot;
    }
    if (inViewStage && zRangeAll)
    {
        for (int s = 16; s; s >>= 1)
        {
            const float a = __shfl_xor_sync(0xffffffffu, zMin, s), b = __shfl_xor_sync(0xffffffffu, zMax, s);
            zMin = a < zMin ? a : zMin;
            zMax = b > zMax ? b : zMax;
        }
        if (lane == 0)
        {
            red[2 * wid] = zMin;
            red[2 * wid + 1] = zMax;
        }
        __syncthreads();
        if (tid == 0)
        {
            for (int w = 1; w < kEmitWarps; ++w)
            {
                zMin = red[2 * w] < zMin ? red[2 * w] : zMin;
                zMax = red[2 * w + 1] > zMax ? red[2 * w + 1] : zMax;
            }
            if (zMin == __int_as_float(0x7f800000))
                zMin = 0.0f; // View.cpp:950-951
            zRangeAll[2 * v] = zMin;
            zRangeAll[2 * v + 1] = zMax;
        }
    }
    if (tid == 0 && lastChunk)
    {
        outCounts[2 * v + 1] = cursor;
        if (cursor > outCap)
            atomicOr(flags, kFlagOutOverflow);
    }
}

// -----------------------------------------------------------------------------------------------------------------------------
// Two-phase occluder drawing (see u3d_kernels.h): near / far split of a view's draw items and the visibility test of the far ones.
// -----------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t item_depth_bin(const float* vp, const float* b)
{
    // clip-space w of the box centre = its distance along the view direction (perspective) -- any monotone key will do, the split only
    // decides what is drawn first.  Positive floats order like their bit patterns: exponent + 2 mantissa bits = 1024 bins.
    const float cx = (b[0] + b[3]) * 0.5f, cy = (b[1] + b[4]) * 0.5f, cz = (b[2] + b[5]) * 0.5f;
    const float w = vp[12] * cx + vp[13] * cy + vp[14] * cz + vp[15];
    return w > 0.0f ? min(__float_as_uint(w) >> 21, 1023u) : 0u;
}

__global__ void __launch_bounds__(256) k_items_near_far(const ViewConst* __restrict__ views, const DrawItem* __restrict__ items, uint32_t itemCap,
    const uint32_t* __restrict__ itemCount, const float* __restrict__ occAabb6, unsigned char* __restrict__ itemPass, uint32_t* __restrict__ nearList,
    uint32_t* __restrict__ nearCount, uint32_t* __restrict__ farCount, uint32_t* __restrict__ farVisCount, uint32_t minNear, int shift)
{
    __shared__ uint32_t hist[1024];
    __shared__ uint32_t sCut, sNear, sFar;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const uint32_t n = min(itemCount[v], itemCap);
    const uint32_t target = max(minNear, n >> shift);
    const DrawItem* mine = items + (size_t)v * itemCap;
    unsigned char* pass = itemPass + (size_t)v * itemCap;
    uint32_t* list = nearList + (size_t)v * itemCap;
    if (tid == 0)
    {
        farVisCount[v] = 0;
        sNear = 0;
        sFar = 0;
    }
    if (n <= target || !views[v].useOcclusion)
    {
        for (uint32_t i = tid; i < n; i += 256)
        {
            pass[i] = kPassNear;
            list[i] = i;
        }
        if (tid == 0)
        {
            nearCount[v] = n;
            farCount[v] = 0;
        }
        return;
    }
    for (int i = tid; i < 1024; i += 256)
        hist[i] = 0;
    __syncthreads();
    const float* vp = views[v].vp;
    for (uint32_t i = tid; i < n; i += 256)
        atomicAdd(&hist[item_depth_bin(vp, occAabb6 + 6 * (size_t)mine[i].model)], 1u);
    __syncthreads();
    if (tid < 32)
    {
        // the first bin at which the running count reaches the target: lanes own 32 bins each
        uint32_t sum = 0;
        for (int k = 0; k < 32; ++k)
            sum += hist[lane * 32 + k];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        const uint32_t reached = __ballot_sync(0xffffffffu, incl >= target);
        const int owner = reached ? __ffs(reached) - 1 : 31;
        if (lane == owner)
        {
            uint32_t run = incl - sum;
            int k = 0;
            for (; k < 31; ++k)
            {
                run += hist[lane * 32 + k];
                if (run >= target)
                    break;
            }
            sCut = (uint32_t)(lane * 32 + k);
        }
    }
    __syncthreads();
    const uint32_t cut = sCut;
    for (uint32_t base = 0; base < n; base += 256)
    {
        const uint32_t i = base + tid;
        const bool near = i < n && item_depth_bin(vp, occAabb6 + 6 * (size_t)mine[i].model) <= cut;
        if (i < n)
            pass[i] = near ? kPassNear : kPassFarHidden;
        // the near items' positions, compacted (order is irrelevant: min-compositing)
        const uint32_t bal = __ballot_sync(0xffffffffu, near);
        uint32_t pos = 0;
        if (lane == 0 && bal)
            pos = atomicAdd(&sNear, (uint32_t)__popc(bal));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (near)
            list[pos + __popc(bal & ((1u << lane) - 1u))] = i;
        // ... and the far ones from the back of the same array (k_items_far_test walks them densely): list[itemCap - 1 - k]
        const uint32_t fbal = __ballot_sync(0xffffffffu, i < n && !near);
        uint32_t fpos = 0;
        if (lane == 0 && fbal)
            fpos = atomicAdd(&sFar, (uint32_t)__popc(fbal));
        fpos = __shfl_sync(0xffffffffu, fpos, 0);
        if (i < n && !near)
            list[itemCap - 1u - (fpos + __popc(fbal & ((1u << lane) - 1u)))] = i;
    }
    __syncthreads();
    if (tid == 0)
    {
        nearCount[v] = sNear;
        farCount[v] = sFar;
    }
}

__global__ void __launch_bounds__(256) k_items_far_test(BufGeom g, const ViewConst* __restrict__ views, const DrawItem* __restrict__ items, uint32_t itemCap,
    const uint32_t* __restrict__ nearList, const uint32_t* __restrict__ farCount, const float* __restrict__ occAabb6, const int* __restrict__ depthAll,
    const int2* __restrict__ mipAll, unsigned char* __restrict__ itemPass, uint32_t* __restrict__ farVisList, uint32_t* __restrict__ farVisCount)
{
    const int v = blockIdx.y;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= min(farCount[v], itemCap))
        return;
    const uint32_t i = nearList[(size_t)v * itemCap + (itemCap - 1u - k)]; // the far items sit at the back of the near list's array
    const float* b = occAabb6 + 6 * (size_t)items[(size_t)v * itemCap + i].model;
    BoxRect r;
    bool visible = ob_project(g, views[v].vp, b[0], b[1], b[2], b[3], b[4], b[5], r);
    if (!visible)
    {
        const int* depth = depthAll + (size_t)v * g.w * g.h;
        const int2* mips = mipAll + (size_t)v * g.mipTexels;
        uint32_t mask;
        const int state = walk_top(g, mips, r, mask);
        visible = state == 1;
        if (state == 2)
        {
            uint32_t stack[kWalkStack];
            int sp;
            // an unfinished walk counts as visible: drawing an occluder that would not have shown is only wasted work
            visible = walk_deep(g, depth, mips, r, mask, stack, sp, 64) != 0;
        }
    }
    if (visible)
    {
        itemPass[(size_t)v * itemCap + i] = kPassFarVisible;
        farVisList[(size_t)v * itemCap + atomicAdd(&farVisCount[v], 1u)] = i;
    }
}

void launch_items_near_far(const ViewConst* views, int numViews, const DrawItem* items, uint32_t itemCap, const uint32_t* itemCount, const float* occAabb6,
    unsigned char* itemPass, uint32_t* nearList, uint32_t* nearCount, uint32_t* farCount, uint32_t* farVisCount, uint32_t minNear, int shift,
    cudaStream_t s)
{
    if (numViews)
        k_items_near_far<<<numViews, 256, 0, s>>>(views, items, itemCap, itemCount, occAabb6, itemPass, nearList, nearCount, farCount, farVisCount, minNear,
            shift);
}

void launch_items_far_test(const BufGeom& g, const ViewConst* views, int numViews, const DrawItem* items, uint32_t itemCap, uint32_t maxItems,
    const uint32_t* nearList, const uint32_t* farCount, const float* occAabb6, const int* depth, const int2* mips, unsigned char* itemPass,
    uint32_t* farVisList, uint32_t* farVisCount, cudaStream_t s)
{
    if (!numViews || !maxItems)
        return;
    // 128-thread blocks over each view's compacted far list (a few hundred items of up to maxItems)
    dim3 grid((maxItems + 127) / 128, numViews);
    k_items_far_test<<<grid, 128, 0, s>>>(g, views, items, itemCap, nearList, farCount, occAabb6, depth, mips, itemPass, farVisList, farVisCount);
}

size_t dense_counter_bytes(int numViews, int numOctants)
{
    const int aliveWords = (numOctants + 1023) / 1024;
    return ((size_t)CTR_WORDS + (size_t)numViews * aliveWords) * 4;
}

static int g_denseGrid = 0;
static int g_denseTreeBlocks = 0; // "dense_tree_blocks": 256-thread units per SM of the persistent tree kernel (0 = as many as fit)
static int gridSms = 148;
static int g_frustumBulk = 0;  // "frustum_bulk": 1 = the drawable tests read their boxes through bulk asynchronous copies (TMA) into shared memory
static int g_denseTree = 1;   // "dense_tree": 1 = all octree levels in one persistent launch (k_dense_tree), 0 = one launch per level
static int g_denseWalker = 1; // "dense_walker": 1 = occlusion predicate with the warp walker, 0 = one walk per lane

void launch_cull_dense(const SceneDev& sc, const BufGeom& g, const ViewConst* views, int numViews, const int* depth, const int2* mips, int mipsValid,
    unsigned char* octState, int stage, uint32_t* outIdx, uint32_t outCap, uint32_t* outCounts, float* zRange, uint32_t* flags, const DenseDev& dd,
    cudaStream_t s, cudaEvent_t* phaseEvents, bool pipelined)
{
    if (!numViews)
        return;
    static int gridLevel = 0, gridTree = 0, gridOcc = 0, gridOccWalk = 0, gridFrustum = 0, gridLight = 0;
    if (!g_denseGrid)
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        gridSms = sms;
        // persistent grids: exactly the blocks that are resident at once (a second, partial wave would leave SMs idle at the end)
        auto fit = [&](auto kernel) {
            int perSm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kDenseThreads, 0);
            return sms * std::max(perSm, 1);
        };
        gridLevel = fit(k_dense_level);
        gridTree = fit(k_dense_tree);
        gridOcc = fit(k_dense_occlusion);
        gridOccWalk = fit(k_dense_occlusion_walk);
        gridFrustum = std::min(fit(k_dense_frustum<false>), fit(k_dense_frustum<true>));
        gridLight = sms * 8 * (256 / kDenseThreads);
        g_denseGrid = sms * 6;
    }
    const int grid = gridLight;
    // octants the walk never reaches must read as culled; counters and alive bits start at zero (the views' counts: k_dense_seed)
    cudaMemsetAsync(octState, 0, (size_t)numViews * dd.stateStride, s);
    cudaMemsetAsync(dd.ctr, 0, dense_counter_bytes(numViews, sc.numOctants), s);
    if (g_denseTree && numViews <= 65535)
    {
        const uint32_t epochTag = dd.epoch << 16; // 1..65535, new every run (capi.cu wipes the queue when it wraps)
        k_dense_tree_seed<<<(numViews + kDenseThreads - 1) / kDenseThreads, kDenseThreads, 0, s>>>(sc, views, numViews, octState, outCounts, dd, epochTag);
        // With other batches in flight on their own streams the persistent kernel takes one 256-thread unit per SM: it spends much of its
        // life waiting for the next octree level's pairs, and a full-size grid would keep everybody else off the GPU meanwhile (512 views,
        // 12 batches in flight: 407 k views/s full size, 483 k with one unit; 4096 views: the same either way).  Alone, it takes every slot.
        // (4096 views, 8 batches in flight: 685 k views/s with one unit, 677 k with two, 666 k with four.  In a serialised launch list -- ncu --
        // the one-unit kernel runs alone and looks twice as long as it costs in the live, overlapped run.)
        const int units = g_denseTreeBlocks > 0 ? g_denseTreeBlocks : (pipelined ? 1 : 0);
        const int blocks = units > 0 ? std::min(gridTree, gridSms * units * (256 / kDenseThreads)) : gridTree;
        k_dense_tree<<<blocks, kDenseThreads, 0, s>>>(sc, g, views, depth, mips, mipsValid, octState, dd, epochTag);
    }
    else
    {
        k_dense_seed<<<(numViews + kDenseThreads - 1) / kDenseThreads, kDenseThreads, 0, s>>>(sc, views, numViews, octState, outCounts, dd);
        for (int lv = 1; lv < sc.numLevels; ++lv)
            k_dense_level<<<gridLevel, kDenseThreads, 0, s>>>(sc, g, views, depth, mips, mipsValid, octState, dd, lv);
    }
    if (phaseEvents)
        cudaEventRecord(phaseEvents[0], s);
    k_dense_groups<<<grid, kDenseThreads, 0, s>>>(sc, numViews, octState, dd);
    k_dense_scan<<<(numViews * 32 + kDenseThreads - 1) / kDenseThreads, kDenseThreads, 0, s>>>(numViews, dd);
    k_dense_words<<<grid, kDenseThreads, 0, s>>>(sc, octState, dd);
    if (g_frustumBulk)
        k_dense_frustum<true><<<gridFrustum, kDenseThreads, 0, s>>>(sc, views, stage, outCounts, dd);
    else
        k_dense_frustum<false><<<gridFrustum, kDenseThreads, 0, s>>>(sc, views, stage, outCounts, dd);
    if (phaseEvents)
        cudaEventRecord(phaseEvents[1], s);
    if (stage >= STAGE_VISIBLE && depth)
    {
        if (g_denseWalker && mipsValid && g.nlev <= 16)
            k_dense_occlusion_walk<<<gridOccWalk, kDenseThreads, 0, s>>>(sc, g, views, depth, mips, dd);
        else
            k_dense_occlusion<<<gridOcc, kDenseThreads, 0, s>>>(sc, g, views, depth, mips, mipsValid, dd);
    }
    if (phaseEvents)
        cudaEventRecord(phaseEvents[2], s);
    {
        // few views: several blocks per view (the in-view stage reduces a Z range per view in one block)
        const uint32_t chunks = stage == STAGE_INVIEW ? 1u : std::min<uint32_t>(16u, std::max<uint32_t>(1u, (2048u + numViews - 1) / numViews));
        const uint32_t chunkWords = 0x7FFFFFFFu;
        k_dense_emit<<<dim3(chunks, numViews), kEmitThreads, 0, s>>>(sc, views, stage, outIdx, outCap, outCounts, zRange, flags, dd, chunkWords);
    }
}

int dense_launch_count(const SceneDev& sc)
{
    return g_denseTree ? 9 : 7 + (sc.numLevels > 1 ? sc.numLevels - 1 : 0);
}

void dense_set_walk_pool(int on)
{
    cudaMemcpyToSymbol(g_denseWalkPool, &on, sizeof(int));
}

void dense_prof_enable(int on)
{
    cudaMemcpyToSymbol(g_denseProfOn, &on, sizeof(int));
    unsigned long long init[8 + 4 * 32];
    for (int i = 0; i < 8 + 4 * 32; ++i)
        init[i] = (i & 3) == 0 ? ~0ull : 0ull;
    cudaMemcpyToSymbol(g_denseProf, init, sizeof(init));
}

void dense_prof_read(unsigned long long* out8)
{
    cudaMemcpyFromSymbol(out8, g_denseProf, 8 * sizeof(unsigned long long));
}

void dense_prof_read_levels(unsigned long long* out128)
{
    cudaMemcpyFromSymbol(out128, g_denseProf, 4 * 32 * sizeof(unsigned long long), 8 * sizeof(unsigned long long));
}

void dense_set_walk_budget(int busy, int idle)
{
    if (busy >= 0)
        cudaMemcpyToSymbol(g_walkBudgetBusy, &busy, sizeof(int));
    if (idle >= 0)
        cudaMemcpyToSymbol(g_walkBudgetIdle, &idle, sizeof(int));
}

void dense_set_frustum_bulk(int on)
{
    g_frustumBulk = on;
}

void dense_set_tree(int on)
{
    g_denseTree = on;
}

void dense_set_tree_blocks(int perSm)
{
    g_denseTreeBlocks = perSm;
}

void dense_set_walker(int on)
{
    g_denseWalker = on;
}

void dense_set_heavy_extent(int px)
{
    cudaMemcpyToSymbol(g_denseHeavyExtent, &px, sizeof(int));
}

} // namespace u3d
