// Octree query + occlusion culling for many views at once.  Replaces, per view:
//   Octree::GetDrawables(OctreeQuery&) -> Octant::GetDrawablesInternal        (Octree.cpp:495-499, 229-255)
//   FrustumOctreeQuery / OccludedFrustumOctreeQuery / ShadowCasterOctreeQuery (OctreeQuery.cpp:98-118, View.cpp:116-158, 59-84)
//   CheckVisibilityWork's occlusion predicate                                 (View.cpp:177)
//
// One CTA per view.  Phase 1 resolves octant states level by level (a child needs its parent's state: culled / visited /
// visited-inside).  Phase 2 sweeps the drawables of the surviving octants in DFS order, 32 bytes per AABB read as two float4,
// and emits the result by ballot-ranked ordered compaction, so the output sequence equals the reference's result_ order.
// Frustum survivors are queued in shared memory so that the expensive occlusion test always runs on dense warps.
#include "u3d_device.cuh"
#include "u3d_kernels.h"

namespace u3d
{

constexpr int kCullBigRing = 16384;      // queue entries of the one-CTA-per-SM launch shape
#ifndef U3D_WSTACK
#define U3D_WSTACK 128
#endif
constexpr int kWStackEntries = U3D_WSTACK; // per-warp texel stack of the cooperative range test
constexpr int kChunkWords = 256;         // chunk bitmaps: up to 256 * 32 chunks of >= 512 octants = 4M octants
constexpr int kHeavyExtentDefault = 96;
__device__ int g_heavyExtent = kHeavyExtentDefault; // rect extent (pixels) from which a box is walked by a whole warp (tuning hook)
static int g_cullRingHost = 4096;       // queue entries of the three-CTAs-per-SM shape (2048 / 4096 / 8192)
#ifndef U3D_WALK_POOL
#define U3D_WALK_POOL 0
#endif
#ifndef U3D_HEAVYCAP
#define U3D_HEAVYCAP 512
#endif
constexpr int kHeavyCap = U3D_HEAVYCAP;          // large-rect boxes parked per evaluation pass; each is then taken by one warp
// Tuning hook, off by default (measured slower: 8.1 vs 6.4 ms, the pass is latency- not issue-bound).  1: the boxes of a warp share one texel
// stack (range_visible_pool); 0: one walk per lane.  Only read when built with -DU3D_WALK_POOL=1.
__device__ int g_walkPool = 0;

/// T = threads per CTA, R = work-queue entries held in shared memory between evaluation passes.
template <int T, int R> struct CullShared
{
    ViewConst vc;
    uint32_t pre[T + 1];   // exclusive scan of drawable counts of the current octant chunk
    uint32_t first[T];     // first drawable of each octant of the chunk | inside << 31
    uint32_t ring[R];             // octants waiting for TestOctant / frustum survivors waiting for the occlusion test
    uint32_t visWord[R / 32];    // one ballot word per group of 32 queued drawables
    uint32_t groupOff[R / 32];
    uint32_t warpTot[T / 32];
    uint32_t wstack[T / 32][kWStackEntries];
    uint4 heavy[kHeavyCap];           // (left | top << 16, right | bottom << 16, z, queue index | result bits << 30)
    uint32_t heavyCount;
    uint32_t heavyNext;
    uint32_t bfsBits[kChunkWords];    // chunks (of T entries) of the level-sorted octant list that hold a child of a surviving octant
    uint32_t dfsBits[kChunkWords];    // chunks of the DFS octant list that hold a surviving octant
    uint32_t ringCount;
    uint32_t nextGroup;
    uint32_t outCursor;
    uint32_t queryCount;
};

/// Block-wide exclusive scan of one value per thread; returns the exclusive prefix, total through `total`.
template <int T> __device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t val, uint32_t* warpTot, uint32_t& total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o)
            x += y;
    }
    __syncthreads(); // warpTot may still be read from the previous use
    if (lane == 31)
        warpTot[wid] = x;
    __syncthreads();
    // every warp scans the (<= 32) warp totals with shuffles
    const uint32_t t = lane < T / 32 ? warpTot[lane] : 0;
    uint32_t y = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t z = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o)
            y += z;
    }
    total = __shfl_sync(0xffffffffu, y, 31);
    const uint32_t pre = __shfl_sync(0xffffffffu, y - t, wid);
    return pre + x - val;
}

/// Ordered block compaction of one predicate per thread: rank among the block's true predicates + their number.
template <int T> __device__ __forceinline__ uint32_t block_rank(bool pred, uint32_t* warpTot, uint32_t& total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, pred);
    __syncthreads();
    if (lane == 0)
        warpTot[wid] = __popc(bal);
    __syncthreads();
    const uint32_t t = lane < T / 32 ? warpTot[lane] : 0;
    uint32_t y = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t z = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o)
            y += z;
    }
    total = __shfl_sync(0xffffffffu, y, 31);
    const uint32_t pre = __shfl_sync(0xffffffffu, y - t, wid);
    return pre + __popc(bal & ((1u << lane) - 1u));
}

/// Stage A of an evaluation pass, one box per lane: project, answer small rects on the spot, park large rects in the CTA's heavy queue
/// (tag = queue index | caller bits).  Returns the lane's answer; `parked` tells that the answer will come from stage B instead.
/// If the queue is full the warp walks the box at once.  Every lane of the warp must call.
template <int T, int R> __device__ __forceinline__ bool visible_stage_a(const BufGeom& g, const float* vp, const int* __restrict__ depth, const int2* __restrict__ mips, bool useMips,
    bool active, const float4& mn, const float4& mx, uint32_t tag, CullShared<T, R>& sh, bool& parked)
{
    const int lane = threadIdx.x & 31;
    BoxRect r;
    r.left = r.top = r.right = r.bottom = r.z = 0;
    bool vis = false, heavy = false, pooled = false;
    parked = false;
#if U3D_WALK_POOL
    const bool pool = useMips && g_walkPool && g.nlev <= 8 && g.mw[0] < 4096;
#else
    const bool pool = false; // the pooled walk is compiled out by default (measured slower; build with -DU3D_WALK_POOL=1 to compare)
#endif
    if (active)
    {
        if (ob_project(g, vp, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, r))
            vis = true;
        else if (useMips && max(r.right - r.left, r.bottom - r.top) >= g_heavyExtent)
            heavy = true;
        else if (pool)
            pooled = true;
        else
            vis = range_visible_thread(g, depth, mips, useMips, r);
    }
#if U3D_WALK_POOL
    if (pool)
    {
        const bool pv = range_visible_pool(g, depth, mips, 0, r, pooled, sh.wstack[threadIdx.x >> 5], (int)(sizeof(sh.wstack[0]) / sizeof(uint32_t)));
        if (pooled)
            vis = pv;
    }
#endif
    uint32_t todo = __ballot_sync(0xffffffffu, heavy);
    if (todo)
    {
        uint32_t pos = 0;
        if (lane == 0)
            pos = atomicAdd(&sh.heavyCount, (uint32_t)__popc(todo));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (heavy)
        {
            const uint32_t mine = pos + __popc(todo & ((1u << lane) - 1u));
            if (mine < kHeavyCap)
            {
                sh.heavy[mine] = make_uint4((uint32_t)r.left | ((uint32_t)r.top << 16), (uint32_t)r.right | ((uint32_t)r.bottom << 16), (uint32_t)r.z, tag);
                parked = true;
                heavy = false;
            }
        }
        todo = __ballot_sync(0xffffffffu, heavy); // the ones that did not fit
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
            const bool v = range_visible_warp(g, depth, mips, rr, sh.wstack[threadIdx.x >> 5], (int)(sizeof(sh.wstack[0]) / sizeof(uint32_t)));
            if (lane == src)
                vis = v;
        }
    }
    return vis;
}

/// Stage B: warps take parked boxes one at a time and walk them cooperatively.  `apply(tag, visible)` is run by lane 0.
template <int T, int R, class Apply> __device__ __forceinline__ void visible_stage_b(const BufGeom& g, const int* __restrict__ depth, const int2* __restrict__ mips,
    CullShared<T, R>& sh, Apply apply)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n = min(sh.heavyCount, (uint32_t)kHeavyCap);
    for (;;)
    {
        uint32_t i = 0;
        if (lane == 0)
            i = atomicAdd(&sh.heavyNext, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n)
            break;
        const uint4 e = sh.heavy[i];
        BoxRect rr;
        rr.left = (int)(e.x & 0xFFFFu);
        rr.top = (int)(e.x >> 16);
        rr.right = (int)(e.y & 0xFFFFu);
        rr.bottom = (int)(e.y >> 16);
        rr.z = (int)e.z;
        const bool v = range_visible_warp(g, depth, mips, rr, sh.wstack[threadIdx.x >> 5], (int)(sizeof(sh.wstack[0]) / sizeof(uint32_t)));
        if (lane == 0)
            apply(e.w, v);
    }
}

// Optional in-kernel phase timers (test/diagnostic hook, off by default): clock64 deltas of thread 0 summed over all CTAs.
__device__ unsigned long long g_cullProf[8];
__device__ int g_cullProfOn;
#define CULL_PROF_MARK(slot)                                                                  \
    do                                                                                        \
    {                                                                                         \
        if (profOn && tid == 0)                                                               \
        {                                                                                     \
            const long long now_ = clock64();                                                 \
            atomicAdd(&g_cullProf[slot], (unsigned long long)(now_ - profT));                  \
            profT = now_;                                                                     \
        }                                                                                     \
    } while (0)

template <int T, int MINB, int R> __global__ void __launch_bounds__(T, MINB) k_cull_views(SceneDev sc, BufGeom g, const ViewConst* __restrict__ views,
    const int* __restrict__ depthAll, const int2* __restrict__ mipAll, int mipsValid, unsigned char* __restrict__ octStateAll, int stateStride, int stage,
    uint32_t* __restrict__ outIdx, uint32_t outCap, uint32_t* __restrict__ outCounts, float* __restrict__ zRangeAll, uint32_t* __restrict__ flags,
    const uint32_t* __restrict__ onlyIf)
{
    static_assert(R / 32 <= T && R % T == 0, "queue size must be a multiple of the block size and at most 32 entries per thread");
    if (onlyIf && !*onlyIf)
        return; // fallback launch behind the dense path: nothing overflowed
    extern __shared__ __align__(16) unsigned char cullSmem[];
    CullShared<T, R>& sh = *reinterpret_cast<CullShared<T, R>*>(cullSmem);
    const int v = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool profOn = g_cullProfOn != 0;
    long long profT = profOn ? clock64() : 0;

    // view constants -> shared
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(views + v);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.vc);
        for (int i = tid; i < (int)(sizeof(ViewConst) / 4); i += T)
            dst[i] = src[i];
        if (tid == 0)
        {
            sh.ringCount = 0;
            sh.nextGroup = 0;
            sh.outCursor = 0;
            sh.queryCount = 0;
            sh.heavyCount = 0;
            sh.heavyNext = 0;
        }
    }
    __syncthreads();
    const ViewConst& vc = sh.vc;
    const bool occl = vc.useOcclusion != 0;
    const bool occludedQuery = occl && vc.queryMode == QUERY_OCCLUDED;
    const bool useMips = mipsValid != 0;
    const int* depth = depthAll + (size_t)v * g.w * g.h;
    const int2* mips = mipAll + (size_t)v * g.mipTexels;
    unsigned char* state = octStateAll + (size_t)v * stateStride;
    uint32_t* out = outIdx + (size_t)v * outCap;

    CULL_PROF_MARK(0);
    // ---------------- phase 1: octant states (0 culled, 1 visited, 2 visited with inside == true) ----------------
    // Per level: a sweep queues the octants whose parent survived (dense queue, order irrelevant); warps then pull groups of 32 from
    // the queue and run TestOctant, so the expensive test never runs on mostly-idle warps.
    // A surviving octant flags the chunk it sits in (phase 2 skips chunks without survivors) and the chunk(s) its children sit in
    // (the next level's sweep skips chunks without candidates).  Skipped octants keep the 0 their state was cleared to.
    constexpr int kShift = T == 1024 ? 10 : 9; // log2(T)
    static_assert((1 << kShift) == T, "chunk size must equal the block size");
    auto mark_alive = [&](int o) {
        atomicOr(&sh.dfsBits[o >> (kShift + 5)], 1u << ((o >> kShift) & 31));
        const int2 ch = sc.octChild[o];
        if (ch.y)
        {
            const int c0 = ch.x >> kShift, c1 = (ch.x + ch.y - 1) >> kShift;
            atomicOr(&sh.bfsBits[c0 >> 5], 1u << (c0 & 31));
            if (c1 != c0)
                atomicOr(&sh.bfsBits[c1 >> 5], 1u << (c1 & 31));
        }
    };
    auto eval_octants = [&]() {
        __syncthreads();
        const uint32_t n = sh.ringCount;
        for (;;)
        {
            uint32_t grp = 0;
            if (lane == 0)
                grp = atomicAdd(&sh.nextGroup, 1u);
            grp = __shfl_sync(0xffffffffu, grp, 0);
            if (grp * 32 >= n)
                break;
            const uint32_t idx = grp * 32 + lane;
            int o = 0, res = OUTSIDE;
            float4 mn = make_float4(0.f, 0.f, 0.f, 0.f), mx = mn;
            if (idx < n)
            {
                const uint32_t e = sh.ring[idx];
                o = (int)(e & 0x7FFFFFFFu);
                mn = sc.octMin[o];
                mx = sc.octMax[o];
                if (e >> 31)
                    res = INSIDE; // FrustumOctreeQuery::TestOctant, OctreeQuery.cpp:98-104
                else
                    res = shape_test<false>(vc.queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z);
            }
            // OccludedFrustumOctreeQuery::TestOctant, View.cpp:128-139
            bool parked = false;
            if (occludedQuery)
            {
                const bool vis = visible_stage_a<T, R>(g, vc.vp, depth, mips, useMips, res != OUTSIDE, mn, mx, (uint32_t)o | ((uint32_t)res << 30), sh, parked);
                if (!vis)
                    res = OUTSIDE;
            }
            if (idx < n && !parked && res != OUTSIDE)
            {
                state[o] = res == INSIDE ? 2 : 1;
                mark_alive(o);
            }
        }
        __syncthreads();
        if (sh.heavyCount)
            visible_stage_b<T, R>(g, depth, mips, sh, [&](uint32_t tag, bool vis) {
                const int res = (int)(tag >> 30);
                if (vis)
                {
                    state[tag & 0x3FFFFFFFu] = res == INSIDE ? 2 : 1;
                    mark_alive((int)(tag & 0x3FFFFFFFu));
                }
            });
        __syncthreads();
        if (tid == 0)
        {
            sh.ringCount = 0;
            sh.nextGroup = 0;
            sh.heavyCount = 0;
            sh.heavyNext = 0;
        }
        __syncthreads();
    };

    for (int i = tid; i < kChunkWords; i += T)
    {
        sh.bfsBits[i] = 0;
        sh.dfsBits[i] = 0;
    }
    __syncthreads();
    if (tid == 0)
    {
        // the root is never tested by OctreeQuery walks (Octree.cpp:231); the ray walk tests every octant (Octree.cpp:257-261)
        bool rootAlive = true;
        if (vc.queryMode == QUERY_RAY || vc.queryMode == QUERY_RAY_CANDIDATES)
        {
            const float4 mn = sc.octMin[0], mx = sc.octMax[0];
            rootAlive = shape_test<false>(vc.queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z) != OUTSIDE;
        }
        if (rootAlive)
        {
            state[0] = 1;
            mark_alive(0);
        }
    }
    __syncthreads();
    for (int lv = 1; lv < sc.numLevels; ++lv)
    {
        const int a = sc.levelStart[lv], b = sc.levelStart[lv + 1];
        if (a == b)
            continue;
        int sweeps = 0;
        for (int chunk = a >> kShift; chunk <= ((b - 1) >> kShift); ++chunk)
        {
            if (!((sh.bfsBits[chunk >> 5] >> (chunk & 31)) & 1u))
                continue; // no surviving octant has a child in this chunk
            const int i = (chunk << kShift) + tid;
            bool need = false;
            uint32_t e = 0;
            if (i >= a && i < b)
            {
                const int o = sc.bfs[i];
                const unsigned char ps = state[__float_as_int(sc.octMin[o].w)];
                if (ps)
                {
                    need = true;
                    e = (uint32_t)o | (ps == 2 ? 0x80000000u : 0u);
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, need);
            if (bal)
            {
                uint32_t pos = 0;
                if (lane == 0)
                    pos = atomicAdd(&sh.ringCount, (uint32_t)__popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (need)
                    sh.ring[pos + __popc(bal & ((1u << lane) - 1u))] = e;
            }
            if (++sweeps == R / T) // the queue could be full after this many sweeps
            {
                CULL_PROF_MARK(1);
                eval_octants();
                CULL_PROF_MARK(2);
                sweeps = 0;
            }
        }
        CULL_PROF_MARK(1);
        eval_octants(); // children need these states
        CULL_PROF_MARK(lv <= 5 ? 6 : (lv == 6 ? 7 : 2));
    }

    // ---------------- phase 2: drawables of surviving octants, DFS order ----------------
    const bool shadowQuery = vc.queryMode == QUERY_SHADOWCASTER;
    const bool needOcclusion = stage >= STAGE_VISIBLE && occl;
    const bool inViewStage = stage == STAGE_INVIEW;
    const bool useQueue = needOcclusion || inViewStage; // survivors go through the evaluation pass
    float zMin = __int_as_float(0x7f800000), zMax = 0.0f; // PerThreadSceneResult::minZ_/maxZ_ start values, View.cpp:893-894

    // Occlusion test over the queued frustum survivors (warps pull groups of 32), then ordered emission.
    auto eval_drawables = [&]() {
        __syncthreads();
        const uint32_t n = sh.ringCount;
        for (;;)
        {
            uint32_t grp = 0;
            if (lane == 0)
                grp = atomicAdd(&sh.nextGroup, 1u);
            grp = __shfl_sync(0xffffffffu, grp, 0);
            if (grp * 32 >= n)
                break;
            const uint32_t idx = grp * 32 + lane;
            bool vis = false, test = false;
            float4 mn = make_float4(0.f, 0.f, 0.f, 0.f), mx = mn;
            if (idx < n)
            {
                const uint32_t d = sh.ring[idx];
                mn = sc.drwMin[d];
                mx = sc.drwMax[d];
                if (!needOcclusion || !(__float_as_uint(mn.w) & kMetaOccludee))
                    vis = true; // !buffer || !drawable->IsOccludee(), View.cpp:177
                else
                    test = true;
                if (inViewStage)
                {
                    // Drawable::UpdateBatches distance (Drawable.cpp:134-137) and the draw-distance test (View.cpp:179-186).
                    // Independent of the occlusion answer, so it is applied up front: a drawable beyond its draw distance is not in view.
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
                            // Abs((GetView() * worldPos).z_), Camera.cpp:495, with the operation order of the reference's default (URHO3D_SSE)
                            // build of Matrix3x4 * Vector3 (Matrix3x4.h:256-274): (m20*x + m22*z) + (m21*y + m23)
                            dist = fabsf((vc.viewZ[0] * cx + vc.viewZ[2] * cz) + (vc.viewZ[1] * cy + vc.viewZ[3]));
                        if (dist > maxDistance)
                        {
                            vis = false;
                            test = false;
                        }
                    }
                }
            }
            bool parked = false;
            if (visible_stage_a<T, R>(g, vc.vp, depth, mips, useMips, test, mn, mx, idx, sh, parked))
                vis = true;
            const uint32_t word = __ballot_sync(0xffffffffu, vis);
            if (lane == 0)
                sh.visWord[grp] = word;
        }
        __syncthreads();
        if (sh.heavyCount)
        {
            visible_stage_b<T, R>(g, depth, mips, sh, [&](uint32_t tag, bool vis) {
                if (vis)
                    atomicOr(&sh.visWord[tag >> 5], 1u << (tag & 31u));
            });
            __syncthreads();
        }
        const uint32_t nGroups = (n + 31) >> 5;
        // exclusive scan of the groups' popcounts (R / 32 <= T)
        uint32_t total;
        const uint32_t cnt = (uint32_t)tid < nGroups ? (uint32_t)__popc(sh.visWord[tid]) : 0u;
        const uint32_t off = block_exclusive_scan<T>(cnt, sh.warpTot, total);
        if ((uint32_t)tid < nGroups)
            sh.groupOff[tid] = off;
        __syncthreads();
        const uint32_t cursor = sh.outCursor;
        for (uint32_t grp = tid >> 5; grp < nGroups; grp += T / 32)
        {
            const uint32_t word = sh.visWord[grp];
            if ((word >> lane) & 1u)
            {
                const uint32_t d = sh.ring[grp * 32 + lane];
                const uint32_t pos = cursor + sh.groupOff[grp] + __popc(word & ((1u << lane) - 1u));
                if (pos < outCap)
                    out[pos] = sc.drwId[d];
                if (inViewStage)
                {
                    // view-space Z range of the geometries in view, View.cpp:198-211
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
        __syncthreads();
        if (tid == 0)
        {
            sh.outCursor = cursor + total;
            sh.ringCount = 0;
            sh.nextGroup = 0;
            sh.heavyCount = 0;
            sh.heavyNext = 0;
        }
        __syncthreads();
    };

    for (int chunk = 0; chunk < sc.numOctants; chunk += T)
    {
        if (!((sh.dfsBits[chunk >> (kShift + 5)] >> ((chunk >> kShift) & 31)) & 1u))
            continue; // no surviving octant in this chunk
        // per octant of the chunk: state, drawable count and first drawable (three independent loads), then a scan of the counts
        const int o = chunk + tid;
        uint32_t cnt = 0, firstWord = 0;
        if (o < sc.numOctants)
        {
            const unsigned char st = state[o];
            const uint32_t c = (uint32_t)sc.octCount[o];
            const uint32_t f = (uint32_t)__float_as_int(sc.octMax[o].w);
            cnt = st ? c : 0u;
            firstWord = f | (st == 2 ? 0x80000000u : 0u);
        }
        uint32_t total;
        const uint32_t pre = block_exclusive_scan<T>(cnt, sh.warpTot, total);
        sh.pre[tid] = pre;
        sh.first[tid] = firstWord;
        if (tid == 0)
            sh.pre[T] = total;
        __syncthreads();
        CULL_PROF_MARK(3);
        if (!total)
            continue;

        for (uint32_t base = 0; base < total; base += T)
        {
            const uint32_t j = base + tid;
            bool pass = false;
            uint32_t d = 0;
            if (j < total)
            {
                // owning octant: largest i with pre[i] <= j (octants with zero drawables share a prefix; take the last)
                uint32_t lo = 0, hi = T;
                while (hi - lo > 1)
                {
                    const uint32_t m = (lo + hi) >> 1;
                    if (sh.pre[m] <= j) lo = m; else hi = m;
                }
                const uint32_t fw = sh.first[lo];
                d = (fw & 0x7FFFFFFFu) + (j - sh.pre[lo]);
                const bool inside = (fw >> 31) != 0;
                const float4 mn = sc.drwMin[d];
                const float4 mx = sc.drwMax[d];
                const uint32_t meta = __float_as_uint(mn.w);
                // FrustumOctreeQuery::TestDrawables (OctreeQuery.cpp:106-118) / ShadowCasterOctreeQuery (View.cpp:70-82)
                // ZoneOccluderOctreeQuery (View.cpp:99-112) replaces the flag test: flags == DRAWABLE_ZONE, or == DRAWABLE_GEOMETRY with IsOccluder()
                const bool flagsOk = vc.queryMode != QUERY_ZONE_OCCLUDER ? (meta & 0xFFu & vc.drawableFlags) != 0
                                                                         : ((meta & 0xFFu) == 4u || ((meta & 0xFFu) == 1u && (meta & kMetaOccluder)));
                if (flagsOk && (__float_as_uint(mx.w) & vc.viewMask) && (!shadowQuery || (meta & kMetaCastShadows)))
                    pass = inside || shape_test<true>(vc.queryMode, vc.planes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z) != OUTSIDE;
            }
            uint32_t tot;
            const uint32_t rank = block_rank<T>(pass, sh.warpTot, tot);
            if (!useQueue)
            {
                const uint32_t ob = sh.outCursor;
                if (pass && ob + rank < outCap)
                    out[ob + rank] = sc.drwId[d];
                __syncthreads();
                if (tid == 0)
                {
                    sh.outCursor = ob + tot;
                    sh.queryCount += tot;
                }
                __syncthreads();
            }
            else
            {
                const uint32_t rc = sh.ringCount;
                if (pass)
                    sh.ring[rc + rank] = d;
                __syncthreads();
                if (tid == 0)
                {
                    sh.ringCount = rc + tot;
                    sh.queryCount += tot;
                }
                __syncthreads();
                if (sh.ringCount + T > R)
                {
                    CULL_PROF_MARK(4);
                    eval_drawables();
                    CULL_PROF_MARK(5);
                }
            }
        }
    }
    CULL_PROF_MARK(4);
    if (useQueue && sh.ringCount)
        eval_drawables();
    CULL_PROF_MARK(5);
    if (inViewStage && zRangeAll)
    {
        // min / max over the block (order-independent, so the result equals the reference's merge of per-thread results, View.cpp:926-951)
        for (int o = 16; o; o >>= 1)
        {
            const float a = __shfl_xor_sync(0xffffffffu, zMin, o), b = __shfl_xor_sync(0xffffffffu, zMax, o);
            zMin = a < zMin ? a : zMin;
            zMax = b > zMax ? b : zMax;
        }
        __syncthreads();
        float* red = reinterpret_cast<float*>(sh.ring);
        if (lane == 0)
        {
            red[2 * (tid >> 5)] = zMin;
            red[2 * (tid >> 5) + 1] = zMax;
        }
        __syncthreads();
        if (tid == 0)
        {
            for (int w = 1; w < T / 32; ++w)
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
    __syncthreads();
    if (tid == 0)
    {
        outCounts[2 * v] = sh.queryCount;
        outCounts[2 * v + 1] = sh.outCursor;
        if (sh.outCursor > outCap)
            atomicOr(flags, kFlagOutOverflow);
    }
}

__global__ void k_is_visible(BufGeom g, const ViewConst* __restrict__ view, const int* __restrict__ depth, const int2* __restrict__ mips, int mipsValid,
    const float* __restrict__ aabb6, int n, unsigned char* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* b = aabb6 + 6 * (size_t)i;
    out[i] = ob_is_visible(g, view->vp, depth, mips, mipsValid != 0, b[0], b[1], b[2], b[3], b[4], b[5]) ? 1 : 0;
}

/// Dynamic shared memory sizes of the cull kernel's instantiations; the attribute belongs to the current device's context, so this runs
/// once per u3d_ctx_create (not per launch).
cudaError_t init_cull_kernels()
{
    cudaError_t e = cudaFuncSetAttribute(k_cull_views<512, 3, 8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CullShared<512, 8192>));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_cull_views<512, 3, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CullShared<512, 4096>));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_cull_views<1024, 1, kCullBigRing>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CullShared<1024, kCullBigRing>));
    return e;
}

void launch_cull(const SceneDev& sc, const BufGeom& g, const ViewConst* views, int numViews, const int* depth, const int2* mips, int mipsValid,
    unsigned char* octState, int stateStride, int stage, uint32_t* outIdx, uint32_t outCap, uint32_t* outCounts, float* zRange, uint32_t* flags,
    bool pipelined, const uint32_t* onlyIf, cudaStream_t s)
{
    if (!numViews)
        return;
    if (!onlyIf)
        cudaMemsetAsync(octState, 0, (size_t)numViews * stateStride, s); // octants the sweeps skip must read as culled
    // Large batches: 512-thread CTAs, three per SM (best throughput).  Batches of a few views per SM: one 1024-thread CTA per SM, which
    // halves the latency of a view and avoids a nearly empty second wave (measured: 512 views 2.0 ms -> 1.46 ms; 4096 views 6.4 ms vs 9.2 ms).
    // When the caller keeps several batches in flight (`pipelined`) other launches fill that wave, and the first shape wins again
    // (3 x 512 views in flight: 1.57 -> 1.35 ms per batch).
    if (numViews > kCullSmallBatch || pipelined)
    {
        // queue entries per CTA (tuning hook "cull_ring"): a larger queue means fewer evaluation passes (each ends in block-wide barriers)
        if (g_cullRingHost == 8192)
        {
            k_cull_views<512, 3, 8192><<<numViews, 512, sizeof(CullShared<512, 8192>), s>>>(sc, g, views, depth, mips, mipsValid, octState, stateStride, stage, outIdx,
                outCap, outCounts, zRange, flags, onlyIf);
        }
        else if (g_cullRingHost == 4096)
        {
            k_cull_views<512, 3, 4096><<<numViews, 512, sizeof(CullShared<512, 4096>), s>>>(sc, g, views, depth, mips, mipsValid, octState, stateStride, stage, outIdx,
                outCap, outCounts, zRange, flags, onlyIf);
        }
        else
            k_cull_views<512, 3, 2048><<<numViews, 512, sizeof(CullShared<512, 2048>), s>>>(sc, g, views, depth, mips, mipsValid, octState, stateStride, stage, outIdx,
                outCap, outCounts, zRange, flags, onlyIf);
    }
    else
    {
        // one CTA per SM has the whole shared memory: a 16K-entry queue means one evaluation pass per ~16K survivors instead of one per 1-2K
        k_cull_views<1024, 1, kCullBigRing><<<numViews, 1024, sizeof(CullShared<1024, kCullBigRing>), s>>>(sc, g, views, depth, mips, mipsValid, octState,
            stateStride, stage, outIdx, outCap, outCounts, zRange, flags, onlyIf);
    }
}

/// World boxes of drawables that moved without leaving their octant: rewrite xyz of their two float4 records, keep the meta words.
__global__ void k_update_boxes(float4* __restrict__ drwMin, float4* __restrict__ drwMax, const uint32_t* __restrict__ slots, const float* __restrict__ boxes,
    uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t s = slots[i];
    const float* b = boxes + 6 * (size_t)i;
    float4 mn = drwMin[s], mx = drwMax[s];
    mn.x = b[0]; mn.y = b[1]; mn.z = b[2];
    mx.x = b[3]; mx.y = b[4]; mx.z = b[5];
    drwMin[s] = mn;
    drwMax[s] = mx;
}

void launch_update_boxes(float4* drwMin, float4* drwMax, const uint32_t* slots, const float* boxes, uint32_t n, cudaStream_t s)
{
    if (n)
        k_update_boxes<<<(n + 255) / 256, 256, 0, s>>>(drwMin, drwMax, slots, boxes, n);
}

/// Hit distances of the drawables a ray query emitted: one block per ray, Ray::HitDistance (Math/Ray.cpp:71-148) of each emitted drawable's
/// world box, in emission order (Drawable::ProcessRayQuery's distance_, Drawable.cpp:120; RaycastSingle's sort value, Octree.cpp:519-523).
__global__ void k_ray_distances(SceneDev sc, const ViewConst* __restrict__ views, const uint32_t* __restrict__ slotOfId, const uint32_t* __restrict__ outIdx,
    uint32_t outCap, const uint32_t* __restrict__ outCounts, float* __restrict__ outDist)
{
    const int v = blockIdx.x;
    const uint32_t n = min(outCounts[2 * v + 1], outCap);
    const float* ray = views[v].planes;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
        const uint32_t d = slotOfId[outIdx[(size_t)v * outCap + i]];
        const float4 mn = sc.drwMin[d], mx = sc.drwMax[d];
        outDist[(size_t)v * outCap + i] = ray_hit_distance(ray, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z);
    }
}

void launch_ray_distances(const SceneDev& sc, const ViewConst* views, int numRays, const uint32_t* slotOfId, const uint32_t* outIdx, uint32_t outCap,
    const uint32_t* outCounts, float* outDist, cudaStream_t s)
{
    if (numRays)
        k_ray_distances<<<numRays, 128, 0, s>>>(sc, views, slotOfId, outIdx, outCap, outCounts, outDist);
}

/// Drawable::SetOccluder for the drawables in `slots`: sets / clears kMetaOccluder in the meta word of their records.
__global__ void k_set_occluder_bits(float4* __restrict__ drwMin, const unsigned char* __restrict__ bySlot, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    uint32_t meta = __float_as_uint(drwMin[i].w);
    meta = bySlot[i] ? (meta | kMetaOccluder) : (meta & ~kMetaOccluder);
    drwMin[i].w = __uint_as_float(meta);
}

void launch_set_occluder_bits(float4* drwMin, const unsigned char* bySlot, uint32_t n, cudaStream_t s)
{
    if (n)
        k_set_occluder_bits<<<(n + 255) / 256, 256, 0, s>>>(drwMin, bySlot, n);
}

/// View::ProcessShadowCasters' per-drawable loop (View.cpp:2409-2450) over the list each view emitted: cast-shadow flag, shadow mask & light
/// mask, the cube-face frustum for point lights, shadow / draw distance against Drawable::UpdateBatches' distance to the MAIN camera, and
/// IsShadowCasterVisible.  One block per view (= shadow split); the list is compacted in place, order kept.
__global__ void __launch_bounds__(256) k_shadow_casters(SceneDev sc, const ShadowSplitDev* __restrict__ splits, const uint32_t* __restrict__ slotOfId,
    const uint32_t* __restrict__ shadowMask, const float* __restrict__ shadowDistance, const unsigned char* __restrict__ inView, uint32_t numInView,
    float4 mainPos /* xyz, w = ortho flag */, float4 mainViewZ, uint32_t* __restrict__ outIdx, uint32_t outCap, uint32_t* __restrict__ outCounts)
{
    __shared__ ShadowSplitDev sp;
    __shared__ uint32_t warpTot[8];
    __shared__ uint32_t sCursor;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < (int)(sizeof(ShadowSplitDev) / 4); i += 256)
        reinterpret_cast<uint32_t*>(&sp)[i] = reinterpret_cast<const uint32_t*>(splits + v)[i];
    if (tid == 0)
        sCursor = 0;
    __syncthreads();
    const uint32_t n = sp.skip ? 0u : min(outCounts[2 * v + 1], outCap);
    uint32_t* ids = outIdx + (size_t)v * outCap;
    const bool mainOrtho = mainPos.w != 0.0f;
    for (uint32_t base = 0; base < n; base += 256)
    {
        const uint32_t i = base + tid;
        bool keep = false;
        uint32_t id = 0;
        if (i < n)
        {
            id = ids[i];
            const uint32_t d = slotOfId[id];
            const float4 mn = sc.drwMin[d], mx = sc.drwMax[d];
            const uint32_t meta = __float_as_uint(mn.w);
            keep = (meta & kMetaCastShadows) != 0;
            if (keep && shadowMask)
                keep = (shadowMask[d] & sp.lightMask) != 0;
            if (keep && sp.lightType == 2)
                keep = frustum_test<true>(sp.shadowPlanes, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z) != OUTSIDE;
            if (keep)
            {
                float maxShadowDistance = shadowDistance ? shadowDistance[d] : 0.0f;
                const float drawDistance = sc.drwDrawDist[d];
                if (drawDistance > 0.0f && (maxShadowDistance <= 0.0f || drawDistance < maxShadowDistance))
                    maxShadowDistance = drawDistance;
                if (maxShadowDistance > 0.0f)
                {
                    const float cx = (mx.x + mn.x) * 0.5f, cy = (mx.y + mn.y) * 0.5f, cz = (mx.z + mn.z) * 0.5f;
                    float dist;
                    if (!mainOrtho)
                    {
                        const float dx = cx - mainPos.x, dy = cy - mainPos.y, dz = cz - mainPos.z;
                        dist = sqrtf(dx * dx + dy * dy + dz * dz);
                    }
                    else
                        dist = fabsf((mainViewZ.x * cx + mainViewZ.z * cz) + (mainViewZ.y * cy + mainViewZ.w));
                    if (dist > maxShadowDistance)
                        keep = false;
                }
            }
            if (keep)
                keep = shadow_caster_visible(sp, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, inView && id < numInView && inView[id]);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0)
            warpTot[wid] = __popc(bal);
        __syncthreads(); // every read of this chunk is done; positions written below are <= the positions read
        uint32_t pre = 0, tot = 0;
        for (int w = 0; w < 8; ++w)
        {
            if (w < wid)
                pre += warpTot[w];
            tot += warpTot[w];
        }
        if (keep)
            ids[sCursor + pre + __popc(bal & ((1u << lane) - 1u))] = id;
        __syncthreads();
        if (tid == 0)
            sCursor += tot;
        __syncthreads();
    }
    if (tid == 0)
        outCounts[2 * v + 1] = sCursor;
}

void launch_shadow_casters(const SceneDev& sc, const ShadowSplitDev* splits, int numViews, const uint32_t* slotOfId, const uint32_t* shadowMask,
    const float* shadowDistance, const unsigned char* inView, uint32_t numInView, float4 mainPos, float4 mainViewZ, uint32_t* outIdx, uint32_t outCap,
    uint32_t* outCounts, cudaStream_t s)
{
    if (numViews)
        k_shadow_casters<<<numViews, 256, 0, s>>>(sc, splits, slotOfId, shadowMask, shadowDistance, inView, numInView, mainPos, mainViewZ, outIdx, outCap,
            outCounts);
}

void cull_set_ring(int entries)
{
    g_cullRingHost = entries;
}

void cull_set_walk_pool(int on)
{
    cudaMemcpyToSymbol(g_walkPool, &on, sizeof(int));
}

void cull_set_heavy_extent(int px)
{
    cudaMemcpyToSymbol(g_heavyExtent, &px, sizeof(int));
}

void cull_prof_enable(int on)
{
    cudaMemcpyToSymbol(g_cullProfOn, &on, sizeof(int));
    unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_cullProf, zero, sizeof(zero));
}

void cull_prof_read(unsigned long long* out8)
{
    cudaMemcpyFromSymbol(out8, g_cullProf, 8 * sizeof(unsigned long long));
}

void launch_is_visible(const BufGeom& g, const ViewConst* view, const int* depth, const int2* mips, int mipsValid, const float* aabb6, int n,
    unsigned char* out, cudaStream_t s)
{
    if (!n)
        return;
    k_is_visible<<<(n + 127) / 128, 128, 0, s>>>(g, view, depth, mips, mipsValid, aabb6, n, out);
}

} // namespace u3d
