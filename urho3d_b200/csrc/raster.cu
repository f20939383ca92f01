// Occluder rasterization for the visibility path: occluder selection, triangle set-up (transform, clip, project, cull,
// edge set-up) and depth fill, plus the depth-hierarchy build.  Replaces OcclusionBuffer::DrawTriangles / DrawBatch /
// DrawTriangle / ClipVertices / DrawTriangle2D / BuildDepthHierarchy (OB.cpp:200-329, 464-1002) of the reference.
//
// The reference's depth test is a min (OB.cpp:909-910), so triangles may be drawn in any order and from any number of
// threads with atomicMin and still give a bit-identical plane (SURVEY hard part 1, measured).
#include "u3d_device.cuh"
#include "u3d_kernels.h"

#include <algorithm>

namespace u3d
{

// ---------------------------------------------------------------------------------------------------------------
// clear
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_clear_depth(int4* __restrict__ depth, size_t n4)
{
    const int4 v = make_int4(kClearDepth, kClearDepth, kClearDepth, kClearDepth);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        depth[i] = v;
}

__global__ void k_clear_tail(int* __restrict__ depth, int n)
{
    if ((int)threadIdx.x < n)
        depth[threadIdx.x] = kClearDepth;
}

// ---------------------------------------------------------------------------------------------------------------
// occluder selection (batched path): per view, every occluder whose world box passes frustum.IsInsideFast is submitted,
// as the threaded branch of View::DrawOccluders does (View.cpp:2258-2270) with SURVEY 8(d)'s selection rule.
// Order inside a view's item list is irrelevant (min-compositing).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_select_occluders(const ViewConst* __restrict__ views, int numOcc, const float* __restrict__ occAabb6,
    const uint32_t* __restrict__ occMesh, const uint32_t* __restrict__ occStart, const uint32_t* __restrict__ occCount,
    DrawItem* __restrict__ items, uint32_t itemCap, uint32_t* __restrict__ itemCount, uint32_t* __restrict__ submitted)
{
    const int v = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    __shared__ float sPlanes[24];
    if (threadIdx.x < 24)
        sPlanes[threadIdx.x] = views[v].planes[threadIdx.x];
    __syncthreads();
    uint32_t tris = 0;
    if (k < numOcc)
    {
        const float* b = occAabb6 + 6 * (size_t)k;
        if (frustum_test<true>(sPlanes, b[0], b[1], b[2], b[3], b[4], b[5]) != OUTSIDE)
            tris = occCount[k] / 3;
    }
    const uint32_t nsub = (tris + kItemTris - 1) / kItemTris;
    // one atomic per warp: inclusive scans of the item and triangle counts
    uint32_t sumItems = nsub, sumTris = tris;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t a = __shfl_up_sync(0xffffffffu, sumItems, o);
        const uint32_t c = __shfl_up_sync(0xffffffffu, sumTris, o);
        if (lane >= o)
        {
            sumItems += a;
            sumTris += c;
        }
    }
    const uint32_t totItems = __shfl_sync(0xffffffffu, sumItems, 31), totTris = __shfl_sync(0xffffffffu, sumTris, 31);
    if (!totItems)
        return;
    uint32_t base = 0;
    if (lane == 31)
    {
        base = atomicAdd(&itemCount[v], totItems);
        atomicAdd(&submitted[v], totTris);
    }
    base = __shfl_sync(0xffffffffu, base, 31) + sumItems - nsub;
    for (uint32_t s = 0; s < nsub; ++s)
    {
        if (base + s >= itemCap)
            break; // cannot happen: itemCap is the sum over all occluders
        DrawItem it;
        it.model = (uint32_t)k;
        it.mesh = occMesh[k];
        it.start = occStart[k] + 3 * s * kItemTris;
        it.count = 3 * min(tris - s * kItemTris, (uint32_t)kItemTris);
        items[(size_t)v * itemCap + base + s] = it;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Occluder selection with the reference's heuristics: View::UpdateOccluders (View.cpp:2171-2229) -- draw-distance test,
// screen-size threshold, density sort key, Sort() -- followed by the triangle budget of View::DrawOccluders' threaded branch
// (View.cpp:2258-2270: occluders are submitted in sorted order until one AddTriangles reports numTriangles_ > maxTriangles_; that
// occluder is still queued, OB.cpp:178-179).  One CTA per view; candidates = occluders passing IsInsideFast, in table order.
// The sort is a shared-memory bitonic sort on (key, candidate number).  The reference's Sort is not stable, but the depth buffer only
// depends on WHICH occluders are drawn, and that set is independent of the order of equal keys unless the budget cut separates two
// equal keys (or a key is NaN); only then one thread repeats the reference's exact Sort procedure (Container/Sort.h:70-141).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRankThreads = 256;

__device__ __forceinline__ uint32_t sortable_bits(float f)
{
    if (f == 0.0f)
        f = 0.0f; // -0 and +0 compare equal
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float from_sortable_bits(uint32_t s)
{
    return __uint_as_float((s & 0x80000000u) ? (s & 0x7FFFFFFFu) : ~s);
}

/// Block-wide bitonic sort of `P` (power of two) 64-bit keys in shared memory, ascending.
__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* a, int P)
{
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            __syncthreads();
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x)
            {
                const int i = 2 * t - (t & (j - 1)); // element with bit j clear
                const int l = i + j;
                const bool up = (i & k) == 0;
                const unsigned long long x = a[i], y = a[l];
                if ((x > y) == up)
                {
                    a[i] = y;
                    a[l] = x;
                }
            }
        }
    __syncthreads();
}

/// Container/Sort.h's procedure on (float key, payload) packed as key bits << 32 | payload, run by ONE thread (rare tie-at-the-cut path).
__device__ void exact_reference_sort(unsigned long long* a, int n)
{
    auto key = [&](int i) { return from_sortable_bits((uint32_t)(a[i] >> 32)); };
    int stackLo[32], stackHi[32];
    int sp = 0;
    stackLo[0] = 0;
    stackHi[0] = n;
    sp = 1;
    while (sp)
    {
        --sp;
        int lo = stackLo[sp], hi = stackHi[sp];
        while (hi - lo > 16)
        {
            int p = lo + (hi - lo) / 2;
            if (key(lo) < key(p) && key(hi - 1) < key(lo))
                p = lo;
            else if (key(hi - 1) < key(p) && key(lo) < key(hi - 1))
                p = hi - 1;
            const float pv = key(p);
            int i = lo - 1, j = hi;
            for (;;)
            {
                do
                    --j;
                while (pv < key(j));
                do
                    ++i;
                while (key(i) < pv);
                if (i >= j)
                    break;
                const unsigned long long t = a[i];
                a[i] = a[j];
                a[j] = t;
            }
            // the reference recurses into [lo, j + 1) and continues with [j + 1, hi): the two ranges are independent, so the order in which
            // they are processed does not change the result.  Push the smaller one (bounded stack), continue with the larger one.
            const int leftLen = j + 1 - lo, rightLen = hi - (j + 1);
            if (leftLen < rightLen)
            {
                if (sp < 32) { stackLo[sp] = lo; stackHi[sp] = j + 1; ++sp; }
                lo = j + 1;
            }
            else
            {
                if (sp < 32) { stackLo[sp] = j + 1; stackHi[sp] = hi; ++sp; }
                hi = j + 1;
            }
        }
    }
    for (int i = 1; i < n; ++i)
    {
        const unsigned long long t = a[i];
        const float tk = from_sortable_bits((uint32_t)(t >> 32));
        int j = i;
        for (; j > 0 && tk < key(j - 1); --j)
            a[j] = a[j - 1];
        a[j] = t;
    }
}

__global__ void __launch_bounds__(kRankThreads) k_select_ranked(const ViewConst* __restrict__ views, const float4* __restrict__ camParams, int numOcc,
    const float* __restrict__ occAabb6, const float* __restrict__ occDrawDist, const uint32_t* __restrict__ occMesh, const uint32_t* __restrict__ occStart,
    const uint32_t* __restrict__ occCount, float sizeThreshold, uint32_t maxTriangles, DrawItem* __restrict__ items, uint32_t itemCap,
    uint32_t* __restrict__ itemCount, uint32_t* __restrict__ submitted, uint32_t* __restrict__ flags, uint32_t* __restrict__ rankList,
    uint32_t* __restrict__ rankCount)
{
    extern __shared__ __align__(16) unsigned char rankSmem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(rankSmem); // kRankCap entries: sortable key bits << 32 | candidate number
    uint32_t* candOcc = reinterpret_cast<uint32_t*>(keys + kRankCap);           // candidate number -> occluder
    __shared__ ViewConst vc;
    __shared__ float4 cam;
    __shared__ uint32_t warpTot[kRankThreads / 32], warpTris[kRankThreads / 32];
    __shared__ uint32_t sCount, sCut, sCarryT, sCarryI, sAmbiguous;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < (int)(sizeof(ViewConst) / 4); i += kRankThreads)
        reinterpret_cast<uint32_t*>(&vc)[i] = reinterpret_cast<const uint32_t*>(views + v)[i];
    if (tid == 0)
    {
        cam = camParams[v];
        sCount = 0;
        sAmbiguous = 0;
    }
    __syncthreads();
    const float halfViewSize = cam.x, invOrthoSize = cam.y, farClip = cam.z;

    // ---- candidates in table order: frustum test, draw distance, size threshold, sort key ----
    for (int base = 0; base < numOcc; base += kRankThreads)
    {
        const int k = base + tid;
        bool keep = false;
        float sortValue = 0.0f;
        if (k < numOcc)
        {
            const float* b = occAabb6 + 6 * (size_t)k;
            const float mnx = b[0], mny = b[1], mnz = b[2], mxx = b[3], mxy = b[4], mxz = b[5];
            if (frustum_test<true>(vc.planes, mnx, mny, mnz, mxx, mxy, mxz) != OUTSIDE)
            {
                // Drawable::UpdateBatches distance (Drawable.cpp:134-138), Camera::GetDistance (Camera.cpp:487-496)
                const float cx = (mxx + mnx) * 0.5f, cy = (mxy + mny) * 0.5f, cz = (mxz + mnz) * 0.5f;
                float distance;
                if (!vc.ortho)
                {
                    const float dx = cx - vc.camPos[0], dy = cy - vc.camPos[1], dz = cz - vc.camPos[2];
                    distance = sqrtf(dx * dx + dy * dy + dz * dz);
                }
                else
                    distance = fabsf((vc.viewZ[0] * cx + vc.viewZ[2] * cz) + (vc.viewZ[1] * cy + vc.viewZ[3]));
                const float maxDistance = occDrawDist ? occDrawDist[k] : 0.0f;
                if (maxDistance <= 0.0f || distance <= maxDistance) // View.cpp:2187-2188
                {
                    const float sx = mxx - mnx, sy = mxy - mny, sz = mxz - mnz;
                    const float diagonal = sqrtf(sx * sx + sy * sy + sz * sz); // box.Size().Length()
                    float compare;
                    if (!vc.ortho)
                    {
                        const float cameraMaxDistanceFraction = distance / farClip;
                        compare = diagonal * halfViewSize / (distance * cameraMaxDistanceFraction);
                        const float px = vc.camPos[0], py = vc.camPos[1], pz = vc.camPos[2];
                        if (!(px < mnx || px > mxx || py < mny || py > mxy || pz < mnz || pz > mxz)) // box.IsInside(cameraPos)
                            compare *= diagonal;
                    }
                    else
                        compare = diagonal * invOrthoSize;
                    if (!(compare < sizeThreshold)) // View.cpp:2208-2209
                    {
                        const float density = (float)(occCount[k] / 3) / diagonal;
                        sortValue = density / compare;
                        keep = true;
                    }
                }
            }
        }
        // ordered compaction
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0)
            warpTot[wid] = __popc(bal);
        __syncthreads();
        uint32_t pre = 0, tot = 0;
        for (int w = 0; w < kRankThreads / 32; ++w)
        {
            if (w < wid)
                pre += warpTot[w];
            tot += warpTot[w];
        }
        const uint32_t pos = sCount + pre + __popc(bal & ((1u << lane) - 1u));
        if (keep)
        {
            if (pos < (uint32_t)kRankCap)
            {
                keys[pos] = ((unsigned long long)sortable_bits(sortValue) << 32) | pos;
                candOcc[pos] = (uint32_t)k;
            }
            if (sortValue != sortValue)
                sAmbiguous = 1; // NaN keys: only the exact procedure tells where they end up
        }
        __syncthreads();
        if (tid == 0)
            sCount += tot;
        __syncthreads();
    }
    uint32_t n = sCount;
    if (n > (uint32_t)kRankCap)
    {
        if (tid == 0)
            atomicOr(flags, kFlagRankOverflow);
        n = kRankCap;
    }
    int P = 1;
    while (P < (int)n)
        P <<= 1;
    for (int i = n + tid; i < P; i += kRankThreads)
        keys[i] = ~0ull;
    bitonic_sort_u64(keys, P);

    if (rankList)
    {
        // Non-threaded branch (View.cpp:2236-2256): the draw loop itself decides what is submitted, and it depends on the ORDER of the
        // list, so equal keys (or NaN keys) always take the reference's exact Sort procedure.
        __shared__ uint32_t sTie;
        if (tid == 0)
            sTie = sAmbiguous;
        __syncthreads();
        for (uint32_t i = tid; i + 1 < n; i += kRankThreads)
            if ((uint32_t)(keys[i] >> 32) == (uint32_t)(keys[i + 1] >> 32))
                sTie = 1;
        __syncthreads();
        if (sTie)
        {
            for (uint32_t i = tid; i < n; i += kRankThreads)
                keys[i] = (keys[i] << 32) | (keys[i] >> 32);
            bitonic_sort_u64(keys, P);
            for (uint32_t i = tid; i < n; i += kRankThreads)
                keys[i] = (keys[i] << 32) | (keys[i] >> 32);
            __syncthreads();
            if (tid == 0)
                exact_reference_sort(keys, (int)n);
            __syncthreads();
        }
        for (uint32_t i = tid; i < n; i += kRankThreads)
            rankList[(size_t)v * kRankCap + i] = candOcc[(uint32_t)keys[i]];
        if (tid == 0)
        {
            rankCount[v] = n;
            itemCount[v] = 0;
            submitted[v] = 0;
        }
        return;
    }

    // ---- budget: occluder i is submitted iff the triangles submitted before it do not exceed maxTriangles ----
    auto find_cut = [&]() {
        if (tid == 0)
        {
            sCarryT = 0;
            sCut = n;
        }
        __syncthreads();
        for (uint32_t base = 0; base < n; base += kRankThreads)
        {
            const uint32_t i = base + tid;
            const uint32_t t = i < n ? occCount[candOcc[(uint32_t)keys[i]]] / 3 : 0;
            uint32_t x = t;
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o)
                    x += y;
            }
            if (lane == 31)
                warpTot[wid] = x;
            __syncthreads();
            uint32_t pre = 0, tot = 0;
            for (int w = 0; w < kRankThreads / 32; ++w)
            {
                if (w < wid)
                    pre += warpTot[w];
                tot += warpTot[w];
            }
            const unsigned long long before = (unsigned long long)sCarryT + pre + x - t; // exclusive prefix
            if (i < n && before > maxTriangles)
                atomicMin(&sCut, i);
            __syncthreads();
            if (tid == 0)
                sCarryT = (uint32_t)min((unsigned long long)sCarryT + tot, 0xFFFFFFFFull);
            __syncthreads();
        }
    };
    find_cut();
    if (tid == 0 && sCut > 0 && sCut < n)
    {
        const float a = from_sortable_bits((uint32_t)(keys[sCut - 1] >> 32)), b = from_sortable_bits((uint32_t)(keys[sCut] >> 32));
        if (a == b)
            sAmbiguous = 1;
    }
    __syncthreads();
    if (sAmbiguous)
    {
        // back to candidate order (sort by the low word), then the reference's own procedure on one thread
        for (uint32_t i = tid; i < n; i += kRankThreads)
            keys[i] = (keys[i] << 32) | (keys[i] >> 32);
        bitonic_sort_u64(keys, P); // padding (~0) stays at the end
        for (uint32_t i = tid; i < n; i += kRankThreads)
            keys[i] = (keys[i] << 32) | (keys[i] >> 32);
        __syncthreads();
        if (tid == 0)
            exact_reference_sort(keys, (int)n);
        __syncthreads();
        find_cut();
    }
    const uint32_t cut = sCut;

    // ---- draw items of the submitted occluders, in sorted order ----
    if (tid == 0)
    {
        sCarryT = 0;
        sCarryI = 0;
    }
    __syncthreads();
    for (uint32_t base = 0; base < cut; base += kRankThreads)
    {
        const uint32_t i = base + tid;
        uint32_t k = 0, tris = 0;
        if (i < cut)
        {
            k = candOcc[(uint32_t)keys[i]];
            tris = occCount[k] / 3;
        }
        const uint32_t nsub = (tris + kItemTris - 1) / kItemTris;
        uint32_t x = nsub, y = tris;
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t a = __shfl_up_sync(0xffffffffu, x, o), c = __shfl_up_sync(0xffffffffu, y, o);
            if (lane >= o)
            {
                x += a;
                y += c;
            }
        }
        if (lane == 31)
        {
            warpTot[wid] = x;
            warpTris[wid] = y;
        }
        __syncthreads();
        uint32_t pre = 0, tot = 0, totT = 0;
        for (int w = 0; w < kRankThreads / 32; ++w)
        {
            if (w < wid)
                pre += warpTot[w];
            tot += warpTot[w];
            totT += warpTris[w];
        }
        const uint32_t first = sCarryI + pre + x - nsub;
        for (uint32_t sidx = 0; sidx < nsub; ++sidx)
        {
            if (first + sidx >= itemCap)
                break;
            DrawItem it;
            it.model = k;
            it.mesh = occMesh[k];
            it.start = occStart[k] + 3 * sidx * kItemTris;
            it.count = 3 * min(tris - sidx * kItemTris, (uint32_t)kItemTris);
            items[(size_t)v * itemCap + first + sidx] = it;
        }
        __syncthreads();
        if (tid == 0)
        {
            sCarryI += tot;
            sCarryT += totT;
        }
        __syncthreads();
    }
    if (tid == 0)
    {
        itemCount[v] = min(sCarryI, itemCap);
        submitted[v] = sCarryT;
    }
}

/// Per-view triangle regions inside the pool: exclusive scan of (submitted + slack).  One block.
__global__ void k_scan_regions(int numViews, const uint32_t* __restrict__ submitted, uint32_t slack, uint64_t poolCap,
    uint64_t* __restrict__ triBase, uint32_t* __restrict__ triCap, uint32_t* __restrict__ flags)
{
    __shared__ uint64_t carry;
    __shared__ uint64_t warpSum[32];
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int base = 0; base < numViews; base += blockDim.x)
    {
        const int v = base + threadIdx.x;
        uint64_t c = v < numViews ? (uint64_t)submitted[v] + slack : 0;
        uint64_t x = c;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1)
        {
            uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o)
                x += y;
        }
        if (lane == 31)
            warpSum[wid] = x;
        __syncthreads();
        uint64_t pre = 0;
        for (int i = 0; i < wid; ++i)
            pre += warpSum[i];
        uint64_t tot = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i)
            tot += warpSum[i];
        const uint64_t start = carry + pre + x - c;
        if (v < numViews)
        {
            uint32_t cap = (uint32_t)c;
            if (start + c > poolCap)
            {
                cap = start < poolCap ? (uint32_t)(poolCap - start) : 0;
                atomicOr(flags, kFlagPoolOverflow);
            }
            triBase[v] = start < poolCap ? start : poolCap;
            triCap[v] = cap;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            carry += tot;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// triangle set-up
// ---------------------------------------------------------------------------------------------------------------

/// Edge constructor, OB.cpp:792-803.
struct EdgeI { int x, xs, z, zs; };
__device__ __forceinline__ EdgeI make_edge(float dzdx, float dzdy, const V3& top, const V3& bot, int topY)
{
    EdgeI e;
    float height = bot.y - top.y;
    float slope = (height != 0.0f) ? (bot.x - top.x) / height : 0.0f;
    float yPre = (float)(topY + 1) - top.y;
    float xPre = slope * yPre;
    e.x = round_x86((xPre + top.x) * kXScale);
    e.xs = round_x86(slope * kXScale);
    e.z = round_x86(top.z + xPre * dzdx + yPre * dzdy);
    e.zs = round_x86(slope * dzdx + dzdy);
    return e;
}

/// Where a set-up thread puts its records: the view's region of the triangle pool plus (tile path) the view's bins.
struct EmitCtx
{
    TriSetup* region;
    uint32_t cap;
    uint32_t* counter;       // triCount[v]
    uint32_t* flags;
    // tile path (binIdx == nullptr: general path only)
    uint32_t* binIdx;        // numTiles x cap entries for this view
    uint32_t* binCount;      // numTiles counters for this view
    unsigned long long* irrList;
    uint32_t irrCap;
    uint32_t* irrTotal;
    uint32_t* irrOfView;
    uint32_t view;
    int w, h, tileW, tilesX, tilesY, tileHLog;
    int forceIrregular;
    int laneAreaMax, warpAreaMax;
    int countOnly;           // decide drawOk only: nothing is emitted
    // k_clip_triangles: triangles that cover many tiles are not binned by the thread that clipped them but queued here, and the whole
    // warp bins them afterwards (BinRequest; nullptr elsewhere)
    struct BinRequest* defer;
    uint32_t* deferCount;
};

struct BinRequest { uint32_t view; int txa, tya, nx, n; uint32_t cls, entry, pad; };
constexpr int kDeferCap = 160; // requests per warp between two drains; beyond it the thread bins alone

/// Tiles first + k * stride (k = 0, 1, ...) of a triangle's n tiles, four at a time: the four atomics are in flight together; a thread that
/// waits for each position before it asks for the next pays a round trip per tile.
__device__ __forceinline__ void bin_tiles(uint32_t* binCount, uint32_t* binIdx, uint32_t cap, int tilesX, int txa, int tya, int nx, int n, uint32_t c, uint32_t e,
    int first, int stride)
{
    for (int k0 = first; k0 < n; k0 += 4 * stride)
    {
        uint32_t pos[4];
        int tile[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const int k = k0 + j * stride;
            // buffers up to 256 wide have one tile per row of tiles: no division
            tile[j] = k >= n ? -1 : nx == 1 ? (tya + k) * tilesX + txa : (tya + k / nx) * tilesX + txa + k % nx;
            pos[j] = tile[j] >= 0 ? atomicAdd(&binCount[tile[j] * kNumCls + c], 1u) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (pos[j] < cap)
                binIdx[((size_t)tile[j] * kNumCls + c) * cap + pos[j]] = e;
    }
}

/// Append one record to the view's region; on the tile path also classify it and bin it.
/// "Regular" = every span of the triangle lies inside its own row of the buffer (0 <= row < h, 0 <= xl, xr <= w for all rows),
/// which the closed form makes checkable at the end rows of each half because x is affine in the row.  Anything else
/// (the reference's "3D clipping is not exact" case, OB.cpp:87-90, or float garbage) goes to the general path.
__device__ __forceinline__ void emit_tri(const EmitCtx& ec, const TriSetup& ts)
{
    const uint32_t slot = atomicAdd(ec.counter, 1u);
    if (slot >= ec.cap)
    {
        atomicOr(ec.flags, kFlagTriOverflow);
        return;
    }
    int4* dst = reinterpret_cast<int4*>(ec.region + slot);
    dst[0] = make_int4(ts.y0, ts.y1, ts.y2, ts.dzdx);
    dst[1] = make_int4(ts.tlx, ts.tlxs, ts.tlz, ts.tlzs);
    dst[2] = make_int4(ts.trx, ts.trxs, ts.blx, ts.blxs);
    dst[3] = make_int4(ts.blz, ts.blzs, ts.brx, ts.brxs);
    if (!ec.binIdx)
        return;

    bool regular = !ec.forceIrregular && ts.y0 >= 0 && ts.y2 <= ec.h && ts.y0 < ts.y2 && ts.y1 >= ts.y0 && ts.y1 <= ts.y2;
    const long long xLimit = ((long long)ec.w << 16) + 65535;
    long long xmin = xLimit, xmax = 0;
    auto edge_range = [&](int x, int xs, int rows) {
        const long long a = x, b = (long long)x + (long long)(rows - 1) * xs;
        if (a < 0 || a > xLimit || b < 0 || b > xLimit)
            regular = false;
        xmin = min(xmin, min(a, b));
        xmax = max(xmax, max(a, b));
    };
    if (regular)
    {
        const int nt = ts.y1 - ts.y0, nb = ts.y2 - ts.y1;
        if (nt > 0)
        {
            edge_range(ts.tlx, ts.tlxs, nt);
            edge_range(ts.trx, ts.trxs, nt);
        }
        if (nb > 0)
        {
            edge_range(ts.blx, ts.blxs, nb);
            edge_range(ts.brx, ts.brxs, nb);
        }
    }
    if (!regular)
    {
        const uint32_t n = atomicAdd(ec.irrTotal, 1u);
        if (n < ec.irrCap)
        {
            ec.irrList[n] = ((unsigned long long)ec.view << 32) | slot;
            atomicAdd(ec.irrOfView, 1u);
        }
        else
            atomicOr(ec.flags, kFlagIrrOverflow);
        return;
    }
    const int x0 = (int)(xmin >> 16), x1 = (int)(xmax >> 16); // pixels touched lie in [x0, x1)
    if (x1 <= x0)
        return; // no pixel in any row
    const int width = x1 - x0, height = ts.y2 - ts.y0;
    // class = who rasterises it in the tile kernel; hint (top bits of the entry) = lane layout for the warp class
    uint32_t cls, hint = 0;
    const long long area = (long long)width * height;
    if (area <= ec.laneAreaMax)
        cls = kClsLane;
    else if (area <= ec.warpAreaMax)
    {
        // lanes per row: each lane walks about 8 pixels of its row; the rest of the warp covers more rows at once
        cls = kClsWarp;
        while (hint < 5 && (8 << hint) < width)
            ++hint;
    }
    else
        cls = kClsCta;
    const int txa = x0 / ec.tileW, txb = min((x1 - 1) / ec.tileW, ec.tilesX - 1);
    const int tya = ts.y0 >> ec.tileHLog, tyb = min((ts.y2 - 1) >> ec.tileHLog, ec.tilesY - 1);
    const int nx = txb - txa + 1, n = nx * (tyb - tya + 1);
    const uint32_t entry = slot | (hint << 29);
    constexpr int kBinAlone = 16;
    if (n > kBinAlone && ec.defer)
    {
        const uint32_t q = atomicAdd(ec.deferCount, 1u);
        if (q < (uint32_t)kDeferCap)
        {
            BinRequest rq;
            rq.view = ec.view; rq.txa = txa; rq.tya = tya; rq.nx = nx; rq.n = n; rq.cls = cls; rq.entry = entry; rq.pad = 0;
            ec.defer[q] = rq;
            return;
        }
    }
    // A triangle that covers many tiles (a ground plane over a 2048 x 2048 buffer: hundreds) is binned by all lanes that arrived here
    // together.  Lanes of one warp may belong to different views, so the bins' addresses travel with the triangle.
    const uint32_t active = __activemask();
    uint32_t big = __ballot_sync(active, n > kBinAlone);
    if (n <= kBinAlone)
        bin_tiles(ec.binCount, ec.binIdx, ec.cap, ec.tilesX, txa, tya, nx, n, cls, entry, 0, 1);
    if (!big)
        return;
    const int lane = threadIdx.x & 31;
    const int rank = __popc(active & ((1u << lane) - 1u)), nActive = __popc(active);
    const unsigned long long pc = (unsigned long long)ec.binCount, pi = (unsigned long long)ec.binIdx;
    while (big)
    {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        uint32_t* sCount = (uint32_t*)__shfl_sync(active, pc, src);
        uint32_t* sIdx = (uint32_t*)__shfl_sync(active, pi, src);
        const uint32_t sCap = __shfl_sync(active, ec.cap, src), sCls = __shfl_sync(active, cls, src), sEntry = __shfl_sync(active, entry, src);
        const int sTilesX = __shfl_sync(active, ec.tilesX, src), sTxa = __shfl_sync(active, txa, src), sTya = __shfl_sync(active, tya, src);
        const int sNx = __shfl_sync(active, nx, src), sN = __shfl_sync(active, n, src);
        bin_tiles(sCount, sIdx, sCap, sTilesX, sTxa, sTya, sNx, sN, sCls, sEntry, rank, nActive);
    }
}

/// ViewportTransform x3 + SignedArea + cull test (OB.cpp:629-634).  Returns true when the reference would call DrawTriangle2D.
__device__ __forceinline__ bool project_and_cull(const BufGeom& g, int cullMode, const V4& c0, const V4& c1, const V4& c2, V3& p0, V3& p1, V3& p2,
    bool& clockwise)
{
    p0 = viewport_transform(g, c0);
    p1 = viewport_transform(g, c1);
    p2 = viewport_transform(g, c2);
    const float aX = p0.x - p1.x, aY = p0.y - p1.y, bX = p2.x - p1.x, bY = p2.y - p1.y;
    clockwise = (aX * bY - aY * bX) < 0.0f; // SignedArea, OB.cpp:569-576
    return cullMode == CULL_NONE || (cullMode == CULL_CCW && clockwise) || (cullMode == CULL_CW && !clockwise);
}

/// DrawTriangle2D's set-up part (OB.cpp:815-899) for a triangle that passed the cull test; emits the record if it has rows.
__device__ __forceinline__ void raster_setup(const V3& p0, const V3& p1, const V3& p2, bool clockwise, const EmitCtx& ec)
{
    // Y sort, OB.cpp:820-872 (written with selects so the vertices stay in registers)
    V3 vt, vm, vb;
    bool midRight;
    if (p0.y < p1.y)
    {
        if (p2.y < p0.y) { vt = p2; vm = p0; vb = p1; midRight = true; }
        else if (p1.y < p2.y) { vt = p0; vm = p1; vb = p2; midRight = true; }
        else { vt = p0; vm = p2; vb = p1; midRight = false; }
    }
    else
    {
        if (p2.y < p1.y) { vt = p2; vm = p1; vb = p0; midRight = false; }
        else if (p0.y < p2.y) { vt = p1; vm = p0; vb = p2; midRight = false; }
        else { vt = p1; vm = p2; vb = p0; midRight = true; }
    }
    const int topY = trunc_x86(vt.y), midY = trunc_x86(vm.y), botY = trunc_x86(vb.y);
    if (topY == botY)
        return; // degenerate: the reference returns from DrawTriangle2D (drawOk is already set by the caller)
    if (!clockwise)
        midRight = !midRight;

    // Gradients, OB.cpp:762-778
    const float invdX = 1.0f / (((p1.x - p2.x) * (p0.y - p2.y)) - ((p0.x - p2.x) * (p1.y - p2.y)));
    const float invdY = -invdX;
    const float dzdx = invdX * (((p1.z - p2.z) * (p0.y - p2.y)) - ((p0.z - p2.z) * (p1.y - p2.y)));
    const float dzdy = invdY * (((p1.z - p2.z) * (p0.x - p2.x)) - ((p0.z - p2.z) * (p1.x - p2.x)));

    TriSetup out;
    out.y0 = topY;
    out.y1 = midY;
    out.y2 = botY;
    out.dzdx = trunc_x86(dzdx);
    const EdgeI lng = make_edge(dzdx, dzdy, vt, vb, topY);
    // long edge advanced to the first row of the bottom half (it is stepped once per top-half row, OB.cpp:915-916)
    // (no top-half rows -> no steps: the reference's row loop simply does not run when middleY <= topY)
    const uint32_t kTop = midY > topY ? (uint32_t)midY - (uint32_t)topY : 0u;
    const int lngBx = (int)((uint32_t)lng.x + kTop * (uint32_t)lng.xs);
    const int lngBz = (int)((uint32_t)lng.z + kTop * (uint32_t)lng.zs);
    EdgeI et = {0, 0, 0, 0}, eb = {0, 0, 0, 0};
    if (topY != midY)
        et = make_edge(dzdx, dzdy, vt, vm, topY);
    if (midY != botY)
        eb = make_edge(dzdx, dzdy, vm, vb, midY);
    if (midRight)
    {
        out.tlx = lng.x; out.tlxs = lng.xs; out.tlz = lng.z; out.tlzs = lng.zs; out.trx = et.x; out.trxs = et.xs;
        out.blx = lngBx; out.blxs = lng.xs; out.blz = lngBz; out.blzs = lng.zs; out.brx = eb.x; out.brxs = eb.xs;
    }
    else
    {
        out.tlx = et.x; out.tlxs = et.xs; out.tlz = et.z; out.tlzs = et.zs; out.trx = lng.x; out.trxs = lng.xs;
        out.blx = eb.x; out.blxs = eb.xs; out.blz = eb.z; out.blzs = eb.zs; out.brx = lngBx; out.brxs = lng.xs;
    }
    emit_tri(ec, out);
}

__device__ __forceinline__ bool setup_triangle(const BufGeom& g, int cullMode, const V4& c0, const V4& c1, const V4& c2, const EmitCtx& ec)
{
    V3 p0, p1, p2;
    bool clockwise;
    if (!project_and_cull(g, cullMode, c0, c1, c2, p0, p1, p2, clockwise))
        return false;
    if (!ec.countOnly)
        raster_setup(p0, p1, p2, clockwise, ec);
    return true;
}

__device__ __forceinline__ V4 clip_edge(const V4& v0, const V4& v1, float d0, float d1) // ClipEdge, OB.cpp:563-567
{
    const float t = d0 / (d0 - d1);
    V4 r;
    r.x = v0.x + t * (v1.x - v0.x);
    r.y = v0.y + t * (v1.y - v0.y);
    r.z = v0.z + t * (v1.z - v0.z);
    r.w = v0.w + t * (v1.w - v0.w);
    return r;
}

/// ClipVertices, OB.cpp:685-753.
__device__ __noinline__ void clip_plane(float px, float py, float pz, float pw, V4* vv, uint64_t& live, unsigned& count)
{
    const unsigned n = count;
    for (unsigned i = 0; i < n; ++i)
    {
        if (!((live >> i) & 1))
            continue;
        V4* t = vv + 3 * i;
        const float d0 = px * t[0].x + py * t[0].y + pz * t[0].z + pw * t[0].w;
        const float d1 = px * t[1].x + py * t[1].y + pz * t[1].z + pw * t[1].w;
        const float d2 = px * t[2].x + py * t[2].y + pz * t[2].z + pw * t[2].w;
        const bool b0 = d0 < 0.0f, b1 = d1 < 0.0f, b2 = d2 < 0.0f;
        if (b0 && b1 && b2)
            live &= ~(1ull << i);
        else if (b0 && b1)
        {
            t[0] = clip_edge(t[0], t[2], d0, d2);
            t[1] = clip_edge(t[1], t[2], d1, d2);
        }
        else if (b0 && b2)
        {
            t[0] = clip_edge(t[0], t[1], d0, d1);
            t[2] = clip_edge(t[2], t[1], d2, d1);
        }
        else if (b1 && b2)
        {
            t[1] = clip_edge(t[1], t[0], d1, d0);
            t[2] = clip_edge(t[2], t[0], d2, d0);
        }
        else if (b0 || b1 || b2)
        {
            if (count >= 64)
                continue; // the reference's scratch holds 64 triangles (OB.cpp:476); 6 planes cannot exceed it
            V4* nt = vv + 3 * count;
            live |= 1ull << count;
            ++count;
            if (b0)
            {
                nt[0] = clip_edge(t[0], t[2], d0, d2);
                nt[1] = t[0] = clip_edge(t[0], t[1], d0, d1);
                nt[2] = t[2];
            }
            else if (b1)
            {
                nt[1] = clip_edge(t[1], t[0], d1, d0);
                nt[2] = t[1] = clip_edge(t[1], t[2], d1, d2);
                nt[0] = t[0];
            }
            else
            {
                nt[2] = clip_edge(t[2], t[1], d2, d1);
                nt[0] = t[2] = clip_edge(t[2], t[0], d2, d0);
                nt[1] = t[1];
            }
        }
    }
}

/// Slow path of DrawTriangle (OB.cpp:640-679): clip against the flagged planes, then set up every surviving piece.
__device__ __noinline__ bool clip_and_setup(const BufGeom& g, int cullMode, unsigned clipMask, V4 c0, V4 c1, V4 c2, const EmitCtx& ec)
{
    V4 vv[64 * 3];
    vv[0] = c0;
    vv[1] = c1;
    vv[2] = c2;
    uint64_t live = 1;
    unsigned count = 1;
    if (clipMask & 1) clip_plane(-1.0f, 0.0f, 0.0f, 1.0f, vv, live, count);
    if (clipMask & 2) clip_plane(1.0f, 0.0f, 0.0f, 1.0f, vv, live, count);
    if (clipMask & 4) clip_plane(0.0f, -1.0f, 0.0f, 1.0f, vv, live, count);
    if (clipMask & 8) clip_plane(0.0f, 1.0f, 0.0f, 1.0f, vv, live, count);
    if (clipMask & 16) clip_plane(0.0f, 0.0f, -1.0f, 1.0f, vv, live, count);
    if (clipMask & 32) clip_plane(0.0f, 0.0f, 1.0f, 0.0f, vv, live, count);
    bool drawOk = false;
    for (unsigned i = 0; i < count; ++i)
    {
        if (!((live >> i) & 1))
            continue;
        if (setup_triangle(g, cullMode, vv[3 * i], vv[3 * i + 1], vv[3 * i + 2], ec))
            drawOk = true;
    }
    return drawOk;
}

__device__ __forceinline__ EmitCtx make_emit_ctx(const BufGeom& g, const TileGeom& tg, const RasterDev& rd, int v)
{
    EmitCtx ec;
    ec.region = rd.pool + rd.triBase[v];
    ec.cap = rd.triCap[v];
    ec.counter = rd.triCount + v;
    ec.flags = rd.flags;
    ec.binIdx = rd.binIdx ? rd.binIdx + rd.triBase[v] * (size_t)tg.numTiles * kNumCls : nullptr;
    ec.binCount = rd.binCount ? rd.binCount + (size_t)v * tg.numTiles * kNumCls : nullptr;
    ec.irrList = rd.irrList;
    ec.irrCap = rd.irrCap;
    ec.irrTotal = rd.irrTotal;
    ec.irrOfView = rd.irrOfView ? rd.irrOfView + v : nullptr;
    ec.view = (uint32_t)v;
    ec.w = g.w;
    ec.h = g.h;
    ec.tileW = tg.tileW;
    ec.tilesX = tg.tilesX;
    ec.tilesY = tg.tilesY;
    ec.tileHLog = tg.tileHLog;
    ec.forceIrregular = rd.forceIrregular;
    ec.laneAreaMax = rd.laneAreaMax;
    ec.warpAreaMax = rd.warpAreaMax;
    ec.countOnly = rd.countOnly;
    ec.defer = nullptr;
    ec.deferCount = nullptr;
    return ec;
}

/// The clipper as its own launch: one thread per triangle that crosses a clip plane (queued by k_setup_triangles), dense over all views,
/// so that the slow path (<= 64 triangles of scratch in local memory) never stalls the set-up blocks.  DrawTriangle's else-branch, OB.cpp:640-679.
__global__ void __launch_bounds__(128) k_clip_triangles(BufGeom g, TileGeom tg, const ViewConst* __restrict__ views, RasterDev rd)
{
    __shared__ BinRequest sDefer[4][kDeferCap];
    __shared__ uint32_t sDeferCount[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t n = min(*rd.clipCount, rd.clipCap);
    if (lane == 0)
        sDeferCount[wid] = 0;
    __syncwarp();
    // every warp takes 32 queue entries at a time, a triangle per lane; what the clipped pieces cover in many tiles is binned by the
    // whole warp afterwards (one thread alone would pay an atomic's round trip per tile and piece: 0.6 ms for a 2048 x 2048 city view)
    for (uint32_t base = (blockIdx.x * 4u + (uint32_t)wid) * 32u; base < n; base += gridDim.x * 128u)
    {
        const uint32_t i = base + (uint32_t)lane;
        if (i < n)
        {
            const float4* d = reinterpret_cast<const float4*>(rd.clipQueue + (size_t)i * kClipEntryFloats);
            const float4 a = d[0], b = d[1], c = d[2], m = d[3];
            const int v = (int)__float_as_uint(m.y);
            EmitCtx ec = make_emit_ctx(g, tg, rd, v);
            if (ec.binIdx)
            {
                ec.defer = sDefer[wid];
                ec.deferCount = &sDeferCount[wid];
            }
            const V4 c0 = {a.x, a.y, a.z, a.w}, c1 = {b.x, b.y, b.z, b.w}, c2 = {c.x, c.y, c.z, c.w};
            if (clip_and_setup(g, views[v].cullMode, __float_as_uint(m.x), c0, c1, c2, ec))
                atomicAdd(&rd.drawnSources[v], 1u);
        }
        __syncwarp();
        const uint32_t nReq = min(sDeferCount[wid], (uint32_t)kDeferCap);
        for (uint32_t r = 0; r < nReq; ++r)
        {
            const BinRequest rq = sDefer[wid][r];
            const EmitCtx ec = make_emit_ctx(g, tg, rd, (int)rq.view);
            bin_tiles(ec.binCount, ec.binIdx, ec.cap, ec.tilesX, rq.txa, rq.tya, rq.nx, rq.n, rq.cls, rq.entry, lane, 32);
        }
        __syncwarp();
        if (lane == 0)
            sDeferCount[wid] = 0;
        __syncwarp();
    }
}

/// Blocks of a view sweep its draw items in slices of kSetupItems; inside a slice the threads sweep the flat triangle index
/// space of those items.  DrawBatch + DrawTriangle (OB.cpp:464-541, 589-683).
#ifndef U3D_SETUP_MINB
#define U3D_SETUP_MINB 4 // 64 registers (24 bytes of spills), 4 blocks per SM; A/B on the bench: set-up 0.39 ms at 106 registers / 2 blocks, 0.31 at
                         // 79 / 3, 0.28 at 64 / 4 (the kernel waits on barriers and dependent loads, more resident warps hide them)
#endif
__global__ void __launch_bounds__(kSetupThreads, U3D_SETUP_MINB) k_setup_triangles(BufGeom g, TileGeom tg, const ViewConst* __restrict__ views,
    const DrawItem* __restrict__ items, uint32_t itemCap, const uint32_t* __restrict__ itemCount, const float* __restrict__ models,
    const MeshDev* __restrict__ meshes, RasterDev rd, uint32_t sliceItems /* <= kSetupItems: items a block takes at a time */)
{
    const int v = blockIdx.y;
    const uint32_t nItems = rd.itemList ? min(rd.listCount[v], itemCap) : itemCount[v];

    __shared__ float sMvp[kSetupItems][16];
    __shared__ DrawItem sItem[kSetupItems];
    __shared__ uint32_t sPre[kSetupItems + 1];
    // queues between the stages (dynamic shared memory): unclipped survivors of stage 1 (three screen vertices + winding, odd stride
    // against bank conflicts) and triangles for the clipper (three clip-space vertices + plane mask).  Each holds 2 x blockDim entries
    // and is drained 256 at a time, so stages 2 and 3 always run on full warps.
    extern __shared__ float sQueues[];
    float (*sSurv)[11] = reinterpret_cast<float (*)[11]>(sQueues);
    __shared__ uint32_t sCount[2];
    if (threadIdx.x == 0)
    {
        sCount[0] = 0;
        sCount[1] = 0;
    }

    const ViewConst& vc = views[v];
    const int cullMode = vc.cullMode;
    const EmitCtx ec = make_emit_ctx(g, tg, rd, v);


    uint32_t drawn = 0;
    // stage 2: edge set-up + emit for the first n queued survivors (n <= blockDim), then close the gap.  Every unclipped survivor
    // reaches DrawTriangle2D, i.e. counts as drawn (drawOk).
    auto drain_survivors = [&](uint32_t n) {
        const uint32_t total = sCount[0];
        float keep[11];
        const bool mover = threadIdx.x + n < total;
        if (threadIdx.x < n)
        {
            const float* d = sSurv[threadIdx.x];
            const V3 a = {d[0], d[1], d[2]}, b = {d[3], d[4], d[5]}, c = {d[6], d[7], d[8]};
            raster_setup(a, b, c, d[9] != 0.0f, ec);
            ++drawn;
        }
        if (mover)
            for (int i = 0; i < 10; ++i)
                keep[i] = sSurv[threadIdx.x + n][i];
        __syncthreads();
        if (mover)
            for (int i = 0; i < 10; ++i)
                sSurv[threadIdx.x][i] = keep[i];
        if (threadIdx.x == 0)
            sCount[0] = total - n;
        __syncthreads();
    };
    for (uint32_t itemBase = blockIdx.x * sliceItems; itemBase < nItems; itemBase += gridDim.x * sliceItems)
    {
        const uint32_t nLocal = min(sliceItems, nItems - itemBase);
        __syncthreads(); // previous slice fully consumed
        if (threadIdx.x < nLocal)
        {
            DrawItem it;
            if (rd.itemList)
                it = items[(size_t)v * itemCap + rd.itemList[(size_t)v * itemCap + itemBase + threadIdx.x]];
            else
            {
                it = items[(size_t)v * itemCap + itemBase + threadIdx.x];
                if (rd.itemPass && rd.itemPass[(size_t)v * itemCap + itemBase + threadIdx.x] != rd.pass)
                    it.count = 0; // another pass's item (two-phase occluder drawing)
            }
            sItem[threadIdx.x] = it;
        }
        __syncthreads();
        // modelViewProj = viewProj_ * model  (Matrix4 * Matrix3x4, Matrix4.cpp:43-63)
        for (uint32_t e = threadIdx.x; e < nLocal * 16; e += blockDim.x)
        {
            const uint32_t it = e >> 4, r = (e >> 2) & 3, c = e & 3;
            if (!sItem[it].count)
                continue;
            const float* a = vc.vp + 4 * r;
            const float* b = models + 12 * (size_t)sItem[it].model;
            float val;
            if (c < 3)
                val = a[0] * b[c] + a[1] * b[4 + c] + a[2] * b[8 + c];
            else
                val = a[0] * b[3] + a[1] * b[7] + a[2] * b[11] + a[3];
            sMvp[it][4 * r + c] = val;
        }
        if (threadIdx.x < 32)
        {
            // exclusive scan of the items' triangle counts (kSetupItems <= 64: two elements per lane)
            uint32_t acc = 0;
            for (uint32_t base = 0; base < kSetupItems; base += 32)
            {
                const uint32_t i = base + threadIdx.x;
                const uint32_t c = i < nLocal ? sItem[i].count / 3 : 0;
                uint32_t x = c;
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                    if ((int)threadIdx.x >= o)
                        x += y;
                }
                if (i < kSetupItems)
                    sPre[i] = acc + x - c;
                acc += __shfl_sync(0xffffffffu, x, 31);
            }
            if (threadIdx.x == 0)
                sPre[kSetupItems] = acc;
        }
        __syncthreads();

        const uint32_t total = sPre[kSetupItems];
        for (uint32_t tBase = 0; tBase < total; tBase += blockDim.x)
        {
            // ---- stage 1: fetch, transform, clip masks, project, cull.  Survivors are compacted into shared memory so that the
            //      expensive edge set-up (stage 2) and the rare clipper (stage 3) run on dense warps ----
            const uint32_t t = tBase + threadIdx.x;
            int kind = 0; // 0 nothing, 1 unclipped survivor, 2 needs clipping
            V4 c0, c1, c2;
            V3 p0, p1, p2;
            bool clockwise = false;
            unsigned clipMask = 0;
            if (t < total)
            {
                // owning item: largest i < nLocal with sPre[i] <= t
                uint32_t lo = 0, hi = nLocal;
                while (hi - lo > 1)
                {
                    const uint32_t m = (lo + hi) >> 1;
                    if (sPre[m] <= t) lo = m; else hi = m;
                }
                const DrawItem& it = sItem[lo];
                const MeshDev mesh = meshes[it.mesh];
                const uint32_t first = it.start + 3 * (t - sPre[lo]);
                uint32_t i0, i1, i2;
                if (mesh.idx == nullptr)
                {
                    i0 = first; i1 = first + 1; i2 = first + 2;
                }
                else if (mesh.idxSize == 2)
                {
                    const unsigned short* ix = (const unsigned short*)mesh.idx;
                    i0 = ix[first]; i1 = ix[first + 1]; i2 = ix[first + 2];
                }
                else
                {
                    const unsigned* ix = (const unsigned*)mesh.idx;
                    i0 = ix[first]; i1 = ix[first + 1]; i2 = ix[first + 2];
                }
                const float* q0 = (const float*)(mesh.vtx + (size_t)i0 * mesh.stride);
                const float* q1 = (const float*)(mesh.vtx + (size_t)i1 * mesh.stride);
                const float* q2 = (const float*)(mesh.vtx + (size_t)i2 * mesh.stride);
                const float* mvp = sMvp[lo];
                c0 = model_transform(mvp, q0[0], q0[1], q0[2]);
                c1 = model_transform(mvp, q1[0], q1[1], q1[2]);
                c2 = model_transform(mvp, q2[0], q2[1], q2[2]);

                // clip masks, OB.cpp:596-620
                unsigned m0 = 0, m1 = 0, m2 = 0;
                if (c0.x > c0.w) m0 |= 1; if (c0.x < -c0.w) m0 |= 2; if (c0.y > c0.w) m0 |= 4; if (c0.y < -c0.w) m0 |= 8; if (c0.z > c0.w) m0 |= 16; if (c0.z < 0.0f) m0 |= 32;
                if (c1.x > c1.w) m1 |= 1; if (c1.x < -c1.w) m1 |= 2; if (c1.y > c1.w) m1 |= 4; if (c1.y < -c1.w) m1 |= 8; if (c1.z > c1.w) m1 |= 16; if (c1.z < 0.0f) m1 |= 32;
                if (c2.x > c2.w) m2 |= 1; if (c2.x < -c2.w) m2 |= 2; if (c2.y > c2.w) m2 |= 4; if (c2.y < -c2.w) m2 |= 8; if (c2.z > c2.w) m2 |= 16; if (c2.z < 0.0f) m2 |= 32;
                if (!(m0 & m1 & m2))
                {
                    clipMask = m0 | m1 | m2;
                    if (clipMask)
                        kind = 2;
                    else if (project_and_cull(g, cullMode, c0, c1, c2, p0, p1, p2, clockwise))
                        kind = 1;
                }
                if (kind == 1 && rd.countOnly)
                {
                    ++drawn; // DrawTriangle2D would run: drawOk
                    kind = 0;
                }
            }
            {
                const int lane = threadIdx.x & 31;
                const uint32_t b1 = __ballot_sync(0xffffffffu, kind == 1), b2 = __ballot_sync(0xffffffffu, kind == 2);
                uint32_t pos1 = 0, pos2 = 0;
                if (lane == 0)
                {
                    if (b1) pos1 = atomicAdd(&sCount[0], (uint32_t)__popc(b1));
                    if (b2) pos2 = atomicAdd(rd.clipCount, (uint32_t)__popc(b2)); // global queue for the clipper kernel
                }
                pos1 = __shfl_sync(0xffffffffu, pos1, 0);
                pos2 = __shfl_sync(0xffffffffu, pos2, 0);
                const uint32_t below = (1u << lane) - 1u;
                if (kind == 1)
                {
                    float* d = sSurv[pos1 + __popc(b1 & below)];
                    d[0] = p0.x; d[1] = p0.y; d[2] = p0.z; d[3] = p1.x; d[4] = p1.y; d[5] = p1.z; d[6] = p2.x; d[7] = p2.y; d[8] = p2.z;
                    d[9] = clockwise ? 1.0f : 0.0f;
                }
                else if (kind == 2)
                {
                    const uint32_t slot = pos2 + __popc(b2 & below);
                    if (slot < rd.clipCap)
                    {
                        float4* d = reinterpret_cast<float4*>(rd.clipQueue + (size_t)slot * kClipEntryFloats);
                        d[0] = make_float4(c0.x, c0.y, c0.z, c0.w);
                        d[1] = make_float4(c1.x, c1.y, c1.z, c1.w);
                        d[2] = make_float4(c2.x, c2.y, c2.z, c2.w);
                        d[3] = make_float4(__uint_as_float(clipMask), __uint_as_float((uint32_t)v), 0.0f, 0.0f);
                    }
                    else
                        atomicOr(rd.flags, kFlagClipOverflow);
                }
            }
            __syncthreads();
            // ---- stage 2 whenever the queue holds a full block of work ----
            if (sCount[0] >= kSetupThreads)
                drain_survivors(kSetupThreads);
        }
    }
    __syncthreads();
    if (sCount[0])
        drain_survivors(sCount[0]);
    // ++numTriangles_ per source triangle that drew anything (OB.cpp:681-682)
    for (int o = 16; o; o >>= 1)
        drawn += __shfl_xor_sync(0xffffffffu, drawn, o);
    if ((threadIdx.x & 31) == 0 && drawn)
        atomicAdd(&rd.drawnSources[v], drawn);
}

// ---------------------------------------------------------------------------------------------------------------
// fill, general path: one warp per set-up triangle, lanes across the span, atomicMin straight into the view's plane in
// global memory with the reference's linear addressing (row + x may spill into the neighbouring row, SURVEY hard part 4).
// Handles every record, including the irregular ones the tile path refuses.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fill_half_global(int* __restrict__ depth, int w, int h, int ya, int yb, int lx, int lxs, int lz, int lzs, int rx, int rxs,
    int dzdx, int lane, int rowOff = 0, int rowStep = 1, bool preRead = true)
{
    // rows whose linear span could reach [0, w*h): |x>>16| <= 32768
    const long long guard = 32768 / w + 2;
    const long long lo = max((long long)ya, -guard), hi = min((long long)yb, (long long)h + guard);
    const long long npix = (long long)w * h;
    for (long long y = lo + rowOff; y < hi; y += rowStep)
    {
        const uint32_t k = (uint32_t)((long long)y - ya);
        const int xl = (int)((uint32_t)lx + k * (uint32_t)lxs) >> 16;
        const int xr = (int)((uint32_t)rx + k * (uint32_t)rxs) >> 16;
        const uint32_t z0 = (uint32_t)lz + k * (uint32_t)lzs;
        const long long row = y * w;
        for (int x = xl + lane; x < xr; x += 32)
        {
            const long long off = row + x;
            if (off >= 0 && off < npix)
            {
                const int z = (int)(z0 + (uint32_t)(x - xl) * (uint32_t)dzdx);
                // preRead = false: fire-and-forget reductions, no load latency per pixel (pays off when most pixels do get written)
                if (!preRead || z < depth[off])
                    atomicMin(depth + off, z);
            }
        }
    }
}

/// rowOff / rowStep: this warp takes every rowStep-th row of each half starting at rowOff (several warps share one triangle).
__device__ __forceinline__ void fill_tri_global(int* __restrict__ depth, int w, int h, const TriSetup& ts, int lane, int rowOff = 0, int rowStep = 1,
    bool preRead = true)
{
    if (ts.y0 != ts.y1)
        fill_half_global(depth, w, h, ts.y0, ts.y1, ts.tlx, ts.tlxs, ts.tlz, ts.tlzs, ts.trx, ts.trxs, ts.dzdx, lane, rowOff, rowStep, preRead);
    if (ts.y1 != ts.y2)
        fill_half_global(depth, w, h, ts.y1, ts.y2, ts.blx, ts.blxs, ts.blz, ts.blzs, ts.brx, ts.brxs, ts.dzdx, lane, rowOff, rowStep, preRead);
}

__device__ __forceinline__ TriSetup load_tri_global(const TriSetup* p)
{
    const int4* q = reinterpret_cast<const int4*>(p);
    const int4 a = q[0], b = q[1], c = q[2], d = q[3];
    TriSetup t;
    t.y0 = a.x; t.y1 = a.y; t.y2 = a.z; t.dzdx = a.w;
    t.tlx = b.x; t.tlxs = b.y; t.tlz = b.z; t.tlzs = b.w;
    t.trx = c.x; t.trxs = c.y; t.blx = c.z; t.blxs = c.w;
    t.blz = d.x; t.blzs = d.y; t.brx = d.z; t.brxs = d.w;
    return t;
}

/// Every record of every view (the OcclusionBuffer drop-in path, which accumulates into an existing plane).
__global__ void __launch_bounds__(256) k_fill_global(BufGeom g, int* __restrict__ depthAll, const TriSetup* __restrict__ pool,
    const uint64_t* __restrict__ triBase, const uint32_t* __restrict__ triCap, const uint32_t* __restrict__ triCount)
{
    const int v = blockIdx.y;
    const uint32_t n = min(triCount[v], triCap[v]);
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerBlock = blockDim.x >> 5;
    int* depth = depthAll + (size_t)v * g.w * g.h;
    const TriSetup* region = pool + triBase[v];
    for (uint32_t t = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); t < n; t += gridDim.x * warpsPerBlock)
    {
        const TriSetup ts = region[t];
        fill_tri_global(depth, g.w, g.h, ts, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// View::DrawOccluders, NON-threaded branch (View.cpp:2236-2256; the engine's default, Renderer::threadedOcclusion_ = false): occluders are
// drawn one by one in sorted order, every occluder after the first is first tested with OcclusionBuffer::IsVisible against what has been
// drawn so far (hierarchy dirty -> pixel scan, OB.cpp:404, 440-455), and the triangle budget sees numTriangles_ = submitted triangles +
// source triangles that drew something (OB.cpp:196-197, 681-682).  The loop is inherently serial per view, so one CTA walks one view's
// list: block-wide pixel test, then transform / clip / set-up of the occluder's triangles into a small scratch region, then one warp per
// set-up triangle fills the view's plane in global memory (general path).  Views run in parallel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSerialThreads = 256;

__global__ void __launch_bounds__(kSerialThreads) k_draw_occluders_serial(BufGeom g, const ViewConst* __restrict__ views, const uint32_t* __restrict__ rankList,
    const uint32_t* __restrict__ rankCount, const float* __restrict__ occAabb6, const float* __restrict__ models, const uint32_t* __restrict__ occMesh,
    const uint32_t* __restrict__ occStart, const uint32_t* __restrict__ occCount, const MeshDev* __restrict__ meshes, uint32_t maxTriangles,
    int* __restrict__ depthAll, TriSetup* __restrict__ scratchAll, uint32_t scratchCap, uint32_t* __restrict__ submitted, uint32_t* __restrict__ drawnSources,
    uint32_t* __restrict__ activeOccluders, uint32_t* __restrict__ flags)
{
    __shared__ ViewConst vc;
    __shared__ float sMvp[16];
    __shared__ uint32_t sRecords, sDrawn;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < (int)(sizeof(ViewConst) / 4); i += kSerialThreads)
        reinterpret_cast<uint32_t*>(&vc)[i] = reinterpret_cast<const uint32_t*>(views + v)[i];
    if (tid == 0)
    {
        sRecords = 0;
        sDrawn = 0;
    }
    __syncthreads();
    int* depth = depthAll + (size_t)v * g.w * g.h;
    TriSetup* scratch = scratchAll + (size_t)v * scratchCap;
    EmitCtx ec{};
    ec.region = scratch;
    ec.cap = scratchCap;
    ec.counter = &sRecords;
    ec.flags = flags;
    ec.binIdx = nullptr; // general path only
    ec.view = (uint32_t)v;
    ec.w = g.w;
    ec.h = g.h;
    const uint32_t n = rankCount[v];
    uint32_t numTriangles = 0, numSubmitted = 0, active = 0; // uniform across the block
    __shared__ uint32_t sGate;
    constexpr uint32_t kGate = 2 * (kSerialThreads / 32); // occluders tested per gate round, two per warp
    for (uint32_t i = 0; i < n;)
    {
        if (i > 0)
        {
            // buffer->IsVisible(occluder->GetWorldBoundingBox()) with the hierarchy dirty: every pixel of the rect (OB.cpp:440-455).
            // The buffer only changes when an occluder is drawn, so the next kGate occluders of the list are tested at once (one warp
            // each, lanes across the rect's pixels); everything before the first visible one is rejected exactly as the serial loop would.
            if (tid == 0)
                sGate = 0xFFFFFFFFu;
            __syncthreads();
            for (uint32_t c = wid; c < kGate; c += kSerialThreads / 32)
            {
                const uint32_t j = i + c;
                if (j >= n || sGate < c)
                    continue; // an earlier occluder of this round is already known to be visible
                const float* b = occAabb6 + 6 * (size_t)rankList[(size_t)v * kRankCap + j];
                BoxRect r;
                bool vis = ob_project(g, vc.vp, b[0], b[1], b[2], b[3], b[4], b[5], r);
                if (!vis)
                {
                    const int rw = r.right - r.left + 1, area = rw * (r.bottom - r.top + 1);
                    // 8 independent L2 loads per lane and vote (the loads bypass L1: atomics update the plane), 256 pixels per step
                    for (int base = 0; base < area; base += 256)
                    {
                        bool hit = false;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                        {
                            const int p = base + 32 * u + lane;
                            const int pc = min(p, area - 1);
                            const int z = __ldcg(depth + (r.top + pc / rw) * g.w + r.left + pc % rw);
                            hit |= p < area && r.z <= z;
                        }
                        if (__any_sync(0xffffffffu, hit))
                        {
                            vis = true;
                            break;
                        }
                    }
                }
                if (vis && lane == 0)
                    atomicMin(&sGate, c);
            }
            __syncthreads();
            const uint32_t first = sGate;
            __syncthreads();
            if (first == 0xFFFFFFFFu)
            {
                i += kGate;
                continue;
            }
            i += first;
        }
        const uint32_t k = rankList[(size_t)v * kRankCap + i];
        ++i;
        ++active;
        const uint32_t tris = occCount[k] / 3;
        numTriangles += tris; // AddTriangles, OB.cpp:196
        numSubmitted += tris;
        const bool success = numTriangles <= maxTriangles;
        // DrawTriangles -> DrawBatch: modelViewProj = viewProj_ * model (Matrix4 * Matrix3x4, Matrix4.cpp:43-63)
        if (tid < 16)
        {
            const int r = tid >> 2, c = tid & 3;
            const float* a = vc.vp + 4 * r;
            const float* m = models + 12 * (size_t)k;
            sMvp[tid] = c < 3 ? a[0] * m[c] + a[1] * m[4 + c] + a[2] * m[8 + c] : a[0] * m[3] + a[1] * m[7] + a[2] * m[11] + a[3];
        }
        __syncthreads();
        const MeshDev mesh = meshes[occMesh[k]];
        const uint32_t start = occStart[k];
        for (uint32_t tBase = 0; tBase < tris; tBase += kSerialThreads)
        {
            const uint32_t t = tBase + tid;
            bool drew = false;
            if (t < tris)
            {
                const uint32_t first = start + 3 * t;
                uint32_t i0, i1, i2;
                if (mesh.idx == nullptr)
                {
                    i0 = first; i1 = first + 1; i2 = first + 2;
                }
                else if (mesh.idxSize == 2)
                {
                    const unsigned short* ix = (const unsigned short*)mesh.idx;
                    i0 = ix[first]; i1 = ix[first + 1]; i2 = ix[first + 2];
                }
                else
                {
                    const unsigned* ix = (const unsigned*)mesh.idx;
                    i0 = ix[first]; i1 = ix[first + 1]; i2 = ix[first + 2];
                }
                const float* q0 = (const float*)(mesh.vtx + (size_t)i0 * mesh.stride);
                const float* q1 = (const float*)(mesh.vtx + (size_t)i1 * mesh.stride);
                const float* q2 = (const float*)(mesh.vtx + (size_t)i2 * mesh.stride);
                const V4 c0 = model_transform(sMvp, q0[0], q0[1], q0[2]);
                const V4 c1 = model_transform(sMvp, q1[0], q1[1], q1[2]);
                const V4 c2 = model_transform(sMvp, q2[0], q2[1], q2[2]);
                unsigned m0 = 0, m1 = 0, m2 = 0; // clip masks, OB.cpp:596-620
                if (c0.x > c0.w) m0 |= 1; if (c0.x < -c0.w) m0 |= 2; if (c0.y > c0.w) m0 |= 4; if (c0.y < -c0.w) m0 |= 8; if (c0.z > c0.w) m0 |= 16; if (c0.z < 0.0f) m0 |= 32;
                if (c1.x > c1.w) m1 |= 1; if (c1.x < -c1.w) m1 |= 2; if (c1.y > c1.w) m1 |= 4; if (c1.y < -c1.w) m1 |= 8; if (c1.z > c1.w) m1 |= 16; if (c1.z < 0.0f) m1 |= 32;
                if (c2.x > c2.w) m2 |= 1; if (c2.x < -c2.w) m2 |= 2; if (c2.y > c2.w) m2 |= 4; if (c2.y < -c2.w) m2 |= 8; if (c2.z > c2.w) m2 |= 16; if (c2.z < 0.0f) m2 |= 32;
                if (!(m0 & m1 & m2))
                {
                    const unsigned clipMask = m0 | m1 | m2;
                    drew = clipMask ? clip_and_setup(g, vc.cullMode, clipMask, c0, c1, c2, ec) : setup_triangle(g, vc.cullMode, c0, c1, c2, ec);
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, drew);
            if (lane == 0 && bal)
                atomicAdd(&sDrawn, (uint32_t)__popc(bal));
            __syncthreads();
            // fill what was set up so far (the scratch region holds at least one chunk's worth).  The occluders that pass the gate are the
            // near, large ones: with few records every triangle's rows are spread over all warps, otherwise one warp takes one record.
            const uint32_t nrec = min(sRecords, scratchCap);
            if (nrec <= 32)
                for (uint32_t rIdx = 0; rIdx < nrec; ++rIdx)
                {
                    const TriSetup ts = load_tri_global(scratch + rIdx);
                    fill_tri_global(depth, g.w, g.h, ts, lane, wid, kSerialThreads / 32, false);
                }
            else
                for (uint32_t rIdx = wid; rIdx < nrec; rIdx += kSerialThreads / 32)
                {
                    const TriSetup ts = load_tri_global(scratch + rIdx);
                    fill_tri_global(depth, g.w, g.h, ts, lane);
                }
            __threadfence_block();
            __syncthreads();
            if (tid == 0)
                sRecords = 0;
            __syncthreads();
        }
        numTriangles += sDrawn; // ++numTriangles_ per source triangle that drew something, OB.cpp:681-682
        __syncthreads();
        if (tid == 0)
            sDrawn = 0;
        __syncthreads();
        if (!success)
            break;
    }
    if (tid == 0)
    {
        submitted[v] = numSubmitted;
        drawnSources[v] = numTriangles - numSubmitted;
        activeOccluders[v] = active;
    }
}

/// Tile path, pre-pass 1: views that own irregular triangles get their plane cleared in global memory (the tile kernel then starts
/// from that plane instead of from the clear value).
__global__ void k_irregular_clear(BufGeom g, int* __restrict__ depthAll, const uint32_t* __restrict__ irrOfView)
{
    const int v = blockIdx.y;
    if (!irrOfView[v])
        return;
    int* depth = depthAll + (size_t)v * g.w * g.h;
    const int n = g.w * g.h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        depth[i] = kClearDepth;
}

/// Tile path, pre-pass 2: the irregular triangles themselves, through the general path.
__global__ void __launch_bounds__(256) k_irregular_fill(BufGeom g, int* __restrict__ depthAll, const TriSetup* __restrict__ pool,
    const uint64_t* __restrict__ triBase, const unsigned long long* __restrict__ irrList, uint32_t irrCap, const uint32_t* __restrict__ irrTotal)
{
    const uint32_t n = min(*irrTotal, irrCap);
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerBlock = blockDim.x >> 5;
    for (uint32_t i = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); i < n; i += gridDim.x * warpsPerBlock)
    {
        const unsigned long long e = irrList[i];
        const uint32_t v = (uint32_t)(e >> 32), slot = (uint32_t)e;
        const TriSetup ts = pool[triBase[v] + slot];
        fill_tri_global(depthAll + (size_t)v * g.w * g.h, g.w, g.h, ts, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fill, tile path: one CTA per (view, screen tile).  The tile's integer depth lives in shared memory; the CTA walks the tile's
// bin of regular triangles, composites with shared-memory atomicMin, then writes the tile to HBM once and reduces it to the
// first `fused` levels of the depth hierarchy while it is still on chip (BuildDepthHierarchy, OB.cpp:234-329).
// ---------------------------------------------------------------------------------------------------------------
struct HalfRow { int xl, xr; uint32_t z0; };

/// Span of row y of a regular triangle (closed form, see TriSetup).
__device__ __forceinline__ HalfRow tri_row(const TriSetup& ts, int y)
{
    HalfRow r;
    if (y < ts.y1)
    {
        const uint32_t k = (uint32_t)(y - ts.y0);
        r.xl = (int)((uint32_t)ts.tlx + k * (uint32_t)ts.tlxs) >> 16;
        r.xr = (int)((uint32_t)ts.trx + k * (uint32_t)ts.trxs) >> 16;
        r.z0 = (uint32_t)ts.tlz + k * (uint32_t)ts.tlzs;
    }
    else
    {
        const uint32_t k = (uint32_t)(y - ts.y1);
        r.xl = (int)((uint32_t)ts.blx + k * (uint32_t)ts.blxs) >> 16;
        r.xr = (int)((uint32_t)ts.brx + k * (uint32_t)ts.brxs) >> 16;
        r.z0 = (uint32_t)ts.blz + k * (uint32_t)ts.blzs;
    }
    return r;
}

__device__ __forceinline__ TriSetup load_tri(const TriSetup* p)
{
    const int4* q = reinterpret_cast<const int4*>(p);
    const int4 a = q[0], b = q[1], c = q[2], d = q[3];
    TriSetup ts;
    ts.y0 = a.x; ts.y1 = a.y; ts.y2 = a.z; ts.dzdx = a.w;
    ts.tlx = b.x; ts.tlxs = b.y; ts.tlz = b.z; ts.tlzs = b.w;
    ts.trx = c.x; ts.trxs = c.y; ts.blx = c.z; ts.blxs = c.w;
    ts.blz = d.x; ts.blzs = d.y; ts.brx = d.z; ts.brxs = d.w;
    return ts;
}

__device__ __forceinline__ TriSetup bcast_tri(const TriSetup& t, int src)
{
    TriSetup r;
    r.y0 = __shfl_sync(0xffffffffu, t.y0, src); r.y1 = __shfl_sync(0xffffffffu, t.y1, src);
    r.y2 = __shfl_sync(0xffffffffu, t.y2, src); r.dzdx = __shfl_sync(0xffffffffu, t.dzdx, src);
    r.tlx = __shfl_sync(0xffffffffu, t.tlx, src); r.tlxs = __shfl_sync(0xffffffffu, t.tlxs, src);
    r.tlz = __shfl_sync(0xffffffffu, t.tlz, src); r.tlzs = __shfl_sync(0xffffffffu, t.tlzs, src);
    r.trx = __shfl_sync(0xffffffffu, t.trx, src); r.trxs = __shfl_sync(0xffffffffu, t.trxs, src);
    r.blx = __shfl_sync(0xffffffffu, t.blx, src); r.blxs = __shfl_sync(0xffffffffu, t.blxs, src);
    r.blz = __shfl_sync(0xffffffffu, t.blz, src); r.blzs = __shfl_sync(0xffffffffu, t.blzs, src);
    r.brx = __shfl_sync(0xffffffffu, t.brx, src); r.brxs = __shfl_sync(0xffffffffu, t.brxs, src);
    return r;
}

#ifndef U3D_FILL_BULK_STORE
#define U3D_FILL_BULK_STORE 1 // A/B on a B200 (round 2, C3 bench): fill phase 2.708 ms with int4 stores, 2.627 ms with the bulk copy
#endif

__device__ __forceinline__ void tile_min(int* tile, int idx, int z)
{
    if (z < tile[idx])
        atomicMin(tile + idx, z);
}

#ifndef U3D_FILL_SHARE_SHIFT
#define U3D_FILL_SHARE_SHIFT 0 // lanes per lane-class triangle ~ (512 << shift) / triangles in the bin ...
#define U3D_FILL_SHARE_MAX 3   // ... up to 1 << max
#endif
#ifndef U3D_FILL_WARP_DIV
#define U3D_FILL_WARP_DIV 128  // warp-class claims of a short bin: bin / this many triangles each (A/B, fill: 16 -> 1.29 ms, 32 -> 1.23, 64 -> 1.18, 128 -> 1.16)
#endif
#ifndef U3D_TILE_MINB
#define U3D_TILE_MINB 3 // three 72 KB tiles fit an SM's shared memory; 40 registers cost 52 bytes of spills and win (A/B: fill 1.15 -> 1.06 ms)
#endif
__global__ void __launch_bounds__(kTileThreads, U3D_TILE_MINB) k_fill_tiles(BufGeom g, TileGeom tg, int* __restrict__ depthAll, int2* __restrict__ mipAll, RasterDev rd)
{
    extern __shared__ int tile[]; // tileW x tileH ints, later reused for the mip levels
    __shared__ uint32_t sNext[2];
    constexpr int kCtaStage = kTileThreads / 4; // CTA-class records staged per batch: one int4 per thread
    __shared__ int4 sCtaTri[kCtaStage][4];

    const int v = blockIdx.y;
    const int tIdx = blockIdx.x;
    const int tileH = tg.tileH;
    const int tx0 = (tIdx % tg.tilesX) * tg.tileW, ty0 = (tIdx / tg.tilesX) * tileH;
    const int tw = tg.tileW, twLog = tg.tileWLog, tw4Log = twLog - 2;
    const int rows = min(tileH, g.h - ty0); // valid rows of this tile
    const int tid = threadIdx.x, lane = tid & 31;
    int* depth = depthAll + (size_t)v * g.w * g.h;
    // second pass of the two-phase occluder drawing: composite on top of what the first pass left in HBM; tiles without new triangles
    // keep their depth and their hierarchy texels
    const bool fromGlobal = rd.secondPass || rd.irrOfView[v] != 0;

    // ---- tiles nothing was binned to: constant depth and constant mips, no shared-memory round trip ----
    {
        const uint32_t* bc = rd.binCount + ((size_t)v * tg.numTiles + tIdx) * kNumCls;
        if (rd.secondPass && rd.irrOfView[v] == 0 && (bc[0] | bc[1] | bc[2]) == 0)
            return;
        if (!fromGlobal && (bc[0] | bc[1] | bc[2]) == 0)
        {
            const int tw4 = tw >> 2;
            const int4 c4 = make_int4(kClearDepth, kClearDepth, kClearDepth, kClearDepth);
            for (int i = tid; i < tw4 * rows; i += kTileThreads)
                *reinterpret_cast<int4*>(depth + (size_t)(ty0 + (i >> tw4Log)) * g.w + tx0 + ((i & (tw4 - 1)) << 2)) = c4;
            if (tg.fused && rd.writeMips)
            {
                int2* mips = mipAll + (size_t)v * g.mipTexels;
                const int2 c2 = make_int2(kClearDepth, kClearDepth);
                for (int lv = 0; lv < tg.fused; ++lv)
                {
                    const int lw = tw >> (lv + 1);
                    const int gy0 = ty0 >> (lv + 1), gx0 = tx0 >> (lv + 1);
                    const int lrows = min(g.mh[lv] - gy0, tileH >> (lv + 1));
                    for (int i = tid; i < lw * lrows; i += kTileThreads)
                        mips[g.moff[lv] + (size_t)(gy0 + (i >> (twLog - lv - 1))) * g.mw[lv] + gx0 + (i & (lw - 1))] = c2;
                }
            }
            return;
        }
    }

    // ---- initialise ----
    {
        const int n4 = (tw * tileH) >> 2;
        int4* t4 = reinterpret_cast<int4*>(tile);
        const int tw4 = tw >> 2;
        for (int i = tid; i < n4; i += kTileThreads)
        {
            int4 val = make_int4(kClearDepth, kClearDepth, kClearDepth, kClearDepth);
            const int r = i >> tw4Log;
            if (fromGlobal && r < rows)
                val = *reinterpret_cast<const int4*>(depth + (size_t)(ty0 + r) * g.w + tx0 + ((i & (tw4 - 1)) << 2));
            t4[i] = val;
        }
        if (tid == 0)
        {
            sNext[0] = 0;
            sNext[1] = 0;
        }
    }
    __syncthreads();

    // ---- composite the three bins of this tile ----
    const uint32_t cap = rd.triCap[v];
    const uint32_t* binCount = rd.binCount + ((size_t)v * tg.numTiles + tIdx) * kNumCls;
    const uint32_t* binBase = rd.binIdx + (rd.triBase[v] * (size_t)tg.numTiles + (size_t)tIdx * cap) * kNumCls;
    const TriSetup* region = rd.pool + rd.triBase[v];
    const int ty1 = ty0 + rows, tx1 = tx0 + tw;
    const uint32_t nLane = min(binCount[kClsLane], cap), nWarp = min(binCount[kClsWarp], cap), nCta = min(binCount[kClsCta], cap);

    // (1) large triangles first: the whole CTA, and nothing else touches the tile meanwhile.  Row y of the tile belongs to warp y mod 16 for
    //     EVERY triangle of this class, so a warp is the only writer of its rows: plain read-modify-write, no atomics, and each lane takes
    //     4 neighbouring pixels with one 16-byte load and one 16-byte store (the tile's rows are 16-byte aligned, tw is a multiple of 4).
    {
        const uint32_t* bin = binBase + (size_t)kClsCta * cap;
        const int wid = tid >> 5;
        constexpr int kWarps = kTileThreads / 32;
        // The records are staged in shared memory 128 at a time by all threads (one coalesced 16-byte load each): every warp walks every
        // record of this class, and a 64-byte L2 round trip per record and warp is what the tile would otherwise wait for.
        for (uint32_t base = 0; base < nCta; base += kCtaStage)
        {
            const uint32_t nStage = min((uint32_t)kCtaStage, nCta - base);
            __syncthreads(); // the previous batch has been walked by every warp
            if ((uint32_t)(tid >> 2) < nStage)
                sCtaTri[tid >> 2][tid & 3] = reinterpret_cast<const int4*>(region + (bin[base + (tid >> 2)] & 0x1FFFFFFFu))[tid & 3];
            __syncthreads();
        for (uint32_t i = 0; i < nStage; ++i)
        {
            TriSetup ts;
            const int4 a = sCtaTri[i][0];
            ts.y0 = a.x; ts.y1 = a.y; ts.y2 = a.z; ts.dzdx = a.w;
            const int ya = max(ts.y0, ty0), yb = min(ts.y2, ty1);
            // first row >= ya owned by this warp (rows are counted from the tile's first row)
            const int first = ya + ((wid - (ya - ty0)) & (kWarps - 1));
            if (first >= yb)
                continue; // none of this warp's rows: the rest of the record is not even read
            {
                const int4 b = sCtaTri[i][1], c = sCtaTri[i][2], d = sCtaTri[i][3];
                ts.tlx = b.x; ts.tlxs = b.y; ts.tlz = b.z; ts.tlzs = b.w;
                ts.trx = c.x; ts.trxs = c.y; ts.blx = c.z; ts.blxs = c.w;
                ts.blz = d.x; ts.blzs = d.y; ts.brx = d.z; ts.brxs = d.w;
            }
            // Spans of this class are ~40 pixels on the benchmark scene: each HALF of the warp takes one of the warp's rows (16 lanes x 4
            // pixels = 64 pixels per step), two rows per iteration.
            for (int y = first + kWarps * (lane >> 4); y < yb; y += 2 * kWarps)
            {
                const HalfRow r = tri_row(ts, y);
                const int xa = max(r.xl, tx0), xb = min(r.xr, tx1);
                if (xa >= xb)
                    continue;
                int* trow = tile + (y - ty0) * tw - tx0;
                if ((tw & 3) == 0)
                {
                    for (int x4 = ((xa - tx0) & ~3) + tx0 + 4 * (lane & 15); x4 < xb; x4 += 64)
                    {
                        int4 d = *reinterpret_cast<int4*>(trow + x4);
                        const uint32_t z0 = r.z0 + (uint32_t)(x4 - r.xl) * (uint32_t)ts.dzdx;
                        const int za = (int)z0, zb = (int)(z0 + (uint32_t)ts.dzdx), zc = (int)(z0 + 2u * (uint32_t)ts.dzdx), zd = (int)(z0 + 3u * (uint32_t)ts.dzdx);
                        if (x4 >= xa && x4 < xb && za < d.x) d.x = za;
                        if (x4 + 1 >= xa && x4 + 1 < xb && zb < d.y) d.y = zb;
                        if (x4 + 2 >= xa && x4 + 2 < xb && zc < d.z) d.z = zc;
                        if (x4 + 3 >= xa && x4 + 3 < xb && zd < d.w) d.w = zd;
                        *reinterpret_cast<int4*>(trow + x4) = d;
                    }
                }
                else
                    for (int x = xa + (lane & 15); x < xb; x += 16)
                    {
                        const int z = (int)(r.z0 + (uint32_t)(x - r.xl) * (uint32_t)ts.dzdx);
                        if (z < trow[x])
                            trow[x] = z;
                    }
            }
        }
        }
    }
    __syncthreads(); // the other classes composite with atomics, from any warp
    // (2) small triangles (bbox area <= 256 px^2, kLaneAreaMax; 86 % of the triangles of the benchmark scene have <= 64): one per lane, groups of 32 handed out
    //     dynamically
    //     -- and when the bin is short (the first pass of the two-phase drawing leaves ~100 per tile), 2, 4 or 8 neighbouring lanes share a
    //     triangle's rows so that the bin still spreads over all 16 warps instead of keeping 3 of them busy while 13 wait at the barrier
    {
        const uint32_t* bin = binBase + (size_t)kClsLane * cap;
        const uint32_t nl = nLane >> U3D_FILL_SHARE_SHIFT;
        const int shareLog = min(U3D_FILL_SHARE_MAX, nl > 256u ? 0 : nl > 128u ? 1 : nl > 64u ? 2 : nl > 32u ? 3 : nl > 16u ? 4 : 5);
        const uint32_t perClaim = 32u >> shareLog;
        const int sub = lane & ((1 << shareLog) - 1), step = 1 << shareLog;
        for (;;)
        {
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(&sNext[0], perClaim);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= nLane)
                break;
            const uint32_t mine = base + (uint32_t)(lane >> shareLog);
            if (mine < nLane)
            {
                const TriSetup ts = load_tri(region + (bin[mine] & 0x1FFFFFFFu));
                const int ya = max(ts.y0, ty0), yb = min(ts.y2, ty1);
                for (int y = ya + sub; y < yb; y += step)
                {
                    const HalfRow r = tri_row(ts, y);
                    const int xa = max(r.xl, tx0), xb = min(r.xr, tx1);
                    int* trow = tile + (y - ty0) * tw - tx0;
                    for (int x = xa; x < xb; ++x)
                        tile_min(trow, x, (int)(r.z0 + (uint32_t)(x - r.xl) * (uint32_t)ts.dzdx));
                }
            }
        }
    }
    // (3) medium triangles: one warp each.  The warp is laid out as (32 / L) rows x L lanes per row, L chosen at set-up so that every
    //     lane walks about 8 pixels of its row; groups of 8 triangles are handed out so that short bins still use all warps.
    {
        const uint32_t* bin = binBase + (size_t)kClsWarp * cap;
        const uint32_t warpGroup = min((uint32_t)kWarpGroup, max(1u, (nWarp + U3D_FILL_WARP_DIV - 1) / U3D_FILL_WARP_DIV)); // short bins: one claim per warp
        for (;;)
        {
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(&sNext[1], warpGroup);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= nWarp)
                break;
            const uint32_t cnt = min(warpGroup, nWarp - base);
            // lanes 0..cnt-1 fetch one record each (one L2 round trip per group); the warp then walks them from registers
            uint32_t mine = 0;
            TriSetup my;
            if (lane < cnt)
            {
                mine = bin[base + lane];
                my = load_tri(region + (mine & 0x1FFFFFFFu));
            }
            for (uint32_t k = 0; k < cnt; ++k)
            {
                const TriSetup ts = bcast_tri(my, (int)k);
                const int lanesLog = (int)(__shfl_sync(0xffffffffu, mine, (int)k) >> 29);
                const int perRow = 1 << lanesLog, rowsPer = 32 >> lanesLog;
                const int lr = lane >> lanesLog, lc = lane & (perRow - 1);
                const int ya = max(ts.y0, ty0), yb = min(ts.y2, ty1);
                for (int yb0 = ya; yb0 < yb; yb0 += rowsPer)
                {
                    const int y = yb0 + lr;
                    if (y < yb)
                    {
                        const HalfRow r = tri_row(ts, y);
                        const int xa = max(r.xl, tx0), xb = min(r.xr, tx1);
                        int* trow = tile + (y - ty0) * tw - tx0;
                        for (int x = xa + lc; x < xb; x += perRow)
                            tile_min(trow, x, (int)(r.z0 + (uint32_t)(x - r.xl) * (uint32_t)ts.dzdx));
                    }
                }
            }
        }
    }
#if U3D_FILL_BULK_STORE
    // When the tile spans whole buffer rows (w <= 256, the benchmark's shape) its rows are contiguous in HBM too, so the finished tile
    // leaves shared memory as ONE bulk asynchronous copy (TMA, UBLKCP.G.S) issued by one thread while all threads go on to reduce the
    // hierarchy from the same shared memory (-DU3D_FILL_BULK_STORE=0 restores the int4 stores).  The writers' generic-proxy stores are fenced
    // towards the async proxy before the barrier; the copy's read of shared memory is waited for before level 0 overwrites the tile in place.
    const bool bulkStore = tw == g.w;
    if (bulkStore)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#else
    constexpr bool bulkStore = false;
#endif
    __syncthreads();

    // ---- write the tile's depth to HBM ----
#if U3D_FILL_BULK_STORE
    if (bulkStore)
    {
        if (tid == 0)
        {
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(depth + (size_t)ty0 * g.w), "r"(src), "r"(rows * tw * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    else
#endif
    {
        const int tw4 = tw >> 2;
        const int n4 = tw4 * rows;
        const int4* t4 = reinterpret_cast<const int4*>(tile);
        for (int i = tid; i < n4; i += kTileThreads)
        {
            const int r = i >> tw4Log, c4 = i & (tw4 - 1);
            *reinterpret_cast<int4*>(depth + (size_t)(ty0 + r) * g.w + tx0 + (c4 << 2)) = t4[i];
        }
    }
    if (!tg.fused || !rd.writeMips)
    {
#if U3D_FILL_BULK_STORE
        if (bulkStore && tid == 0)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the copy's read
#endif
        return;
    }

    // ---- fused depth hierarchy ----
    // Level 0 is computed in place: every thread first pulls its inputs into registers, then (after a barrier) the (min,max) pairs
    // overwrite the front of the tile.  Further levels go to fresh regions behind it.
    int2* mips = mipAll + (size_t)v * g.mipTexels;
    int2* lv0 = reinterpret_cast<int2*>(tile);
    constexpr int kPerThread = (256 / 2) * (kTileH / 2) / kTileThreads; // level-0 texels per thread for the widest tile
    {
        const int lw = tw >> 1;                          // texels per row of this tile at level 0
        const int lrows = (rows + 1) >> 1;               // texel rows of this tile at level 0
        const int n = lw * lrows;
        int2 acc[kPerThread];
#pragma unroll
        for (int j = 0; j < kPerThread; ++j)
        {
            const int i = tid + j * kTileThreads;
            if (i < n)
            {
                const int x = i & (lw - 1), y = i >> (twLog - 1);
                const int2 a = *reinterpret_cast<const int2*>(tile + (2 * y) * tw + 2 * x);
                int mn = min(a.x, a.y), mx = max(a.x, a.y);
                if (2 * y + 1 < rows) // the buffer's last row pair may be a single row only when h is odd; heights are even, kept for safety
                {
                    const int2 b = *reinterpret_cast<const int2*>(tile + (2 * y + 1) * tw + 2 * x);
                    mn = min(mn, min(b.x, b.y));
                    mx = max(mx, max(b.x, b.y));
                }
                acc[j] = make_int2(mn, mx);
            }
        }
#if U3D_FILL_BULK_STORE
        if (bulkStore && tid == 0)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the tile is overwritten in place below
#endif
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPerThread; ++j)
        {
            const int i = tid + j * kTileThreads;
            if (i < n)
            {
                const int x = i & (lw - 1), y = i >> (twLog - 1);
                lv0[i] = acc[j];
                mips[g.moff[0] + (size_t)((ty0 >> 1) + y) * g.mw[0] + (tx0 >> 1) + x] = acc[j];
            }
        }
        __syncthreads();
    }
    int2* prev = lv0;
    int prevW = tw >> 1, prevRows = (rows + 1) >> 1;
    int2* next = lv0 + (tw >> 1) * (tileH >> 1);
    for (int lv = 1; lv < tg.fused; ++lv)
    {
        const int lw = tw >> (lv + 1);
        const int gy0 = ty0 >> (lv + 1), gx0 = tx0 >> (lv + 1);
        // rows of this level that belong to the tile: those whose children rows exist in prev
        const int lrows = (prevRows + 1) >> 1;
        const int prevGlobalH = g.mh[lv - 1], prevGy0 = ty0 >> lv;
        const int n = lw * lrows;
        for (int i = tid; i < n; i += kTileThreads)
        {
            const int x = i & (lw - 1), y = i >> (twLog - lv - 1);
            const int2 a = prev[(2 * y) * prevW + 2 * x], b = prev[(2 * y) * prevW + 2 * x + 1];
            int mn = min(a.x, b.x), mx = max(a.y, b.y);
            if (prevGy0 + 2 * y + 1 < prevGlobalH) // odd previous height: the last row reduces a single row (OB.cpp:297-324)
            {
                const int2 c = prev[(2 * y + 1) * prevW + 2 * x], d = prev[(2 * y + 1) * prevW + 2 * x + 1];
                mn = min(mn, min(c.x, d.x));
                mx = max(mx, max(c.y, d.y));
            }
            const int2 o = make_int2(mn, mx);
            next[i] = o;
            mips[g.moff[lv] + (size_t)(gy0 + y) * g.mw[lv] + gx0 + x] = o;
        }
        __syncthreads();
        prev = next;
        prevW = lw;
        prevRows = lrows;
        next = next + lw * (tileH >> (lv + 1));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// depth hierarchy (general path): one launch per level, one thread per output texel.  BuildDepthHierarchy, OB.cpp:234-329.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_mip_level(BufGeom g, int lv, const int* __restrict__ depthAll, int2* __restrict__ mipAll)
{
    const int v = blockIdx.y;
    const int w = g.mw[lv], h = g.mh[lv];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h)
        return;
    const int x = i % w, y = i / w;
    int2* mips = mipAll + (size_t)v * g.mipTexels;
    int mn, mx;
    if (lv == 0)
    {
        const int* src = depthAll + (size_t)v * g.w * g.h + (size_t)(2 * y) * g.w + 2 * x;
        mn = min(src[0], src[1]);
        mx = max(src[0], src[1]);
        if (2 * y + 1 < g.h)
        {
            mn = min(mn, min(src[g.w], src[g.w + 1]));
            mx = max(mx, max(src[g.w], src[g.w + 1]));
        }
    }
    else
    {
        const int pw = g.mw[lv - 1], ph = g.mh[lv - 1];
        const int2* src = mips + g.moff[lv - 1] + (size_t)(2 * y) * pw + 2 * x;
        mn = min(src[0].x, src[1].x);
        mx = max(src[0].y, src[1].y);
        if (2 * y + 1 < ph)
        {
            mn = min(mn, min(src[pw].x, src[pw + 1].x));
            mx = max(mx, max(src[pw].y, src[pw + 1].y));
        }
    }
    mips[g.moff[lv] + i] = make_int2(mn, mx);
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
TileGeom make_tile_geom(const BufGeom& g, uint32_t maxViews, uint32_t targetTiles)
{
    TileGeom t{};
    t.tileW = g.w < 256 ? g.w : 256;
    t.tilesX = g.w / t.tileW;
    // Tile height: kTileH rows, halved (down to 16: a row per warp for the CTA-class triangles) while a full batch would launch fewer tiles
    // than `targetTiles` (three per SM, one wave of k_fill_tiles): a single 2048 x 2048 view gets 512 tiles of 256 x 32 instead of 256,
    // which also splits the rows a city's horizon crowds with triangles over twice as many CTAs (C5: fill 1.01 -> 0.54 ms; 16 rows: 0.51,
    // but twice the bin inserts in set-up).
    t.tileH = kTileH;
    while (t.tileH > 16 && (uint64_t)std::max<uint32_t>(maxViews, 1u) * t.tilesX * ((g.h + t.tileH - 1) / t.tileH) < targetTiles)
        t.tileH >>= 1;
    t.tileHLog = 0;
    while ((1 << t.tileHLog) < t.tileH)
        ++t.tileHLog;
    t.tilesY = (g.h + t.tileH - 1) / t.tileH;
    t.numTiles = t.tilesX * t.tilesY;
    // levels that can be reduced inside one tile: both tile dimensions must cover whole texels of the level
    int f = 0;
    while (f < g.nlev && (2 << f) <= t.tileW && (2 << f) <= t.tileH)
        ++f;
    // the in-place level-0 pass keeps its inputs in registers: needs tileW >= 4 for the int4 paths
    t.fused = t.tileW >= 4 ? f : 0;
    t.tileWLog = 0;
    while ((1 << t.tileWLog) < t.tileW)
        ++t.tileWLog;
    return t;
}

void launch_clear(int* depth, size_t nInts, cudaStream_t s)
{
    if (!nInts)
        return;
    const size_t n4 = nInts / 4;
    if (n4)
    {
        size_t nb = (n4 + 255) / 256;
        if (nb > 148 * 16) nb = 148 * 16;
        const int blocks = (int)nb;
        k_clear_depth<<<blocks, 256, 0, s>>>((int4*)depth, n4);
    }
    if (nInts & 3)
        k_clear_tail<<<1, 32, 0, s>>>(depth + 4 * n4, (int)(nInts & 3));
}

void launch_select(const ViewConst* views, int numViews, int numOcc, const float* occAabb6, const uint32_t* occMesh, const uint32_t* occStart,
    const uint32_t* occCount, DrawItem* items, uint32_t itemCap, uint32_t* itemCount, uint32_t* submitted, cudaStream_t s)
{
    if (!numOcc || !numViews)
        return;
    dim3 grid((numOcc + 127) / 128, numViews);
    k_select_occluders<<<grid, 128, 0, s>>>(views, numOcc, occAabb6, occMesh, occStart, occCount, items, itemCap, itemCount, submitted);
}

cudaError_t init_rank_kernel()
{
    return cudaFuncSetAttribute(k_select_ranked, cudaFuncAttributeMaxDynamicSharedMemorySize, kRankCap * 12);
}

void launch_select_ranked(const ViewConst* views, const float4* camParams, int numViews, int numOcc, const float* occAabb6, const float* occDrawDist,
    const uint32_t* occMesh, const uint32_t* occStart, const uint32_t* occCount, float sizeThreshold, uint32_t maxTriangles, DrawItem* items,
    uint32_t itemCap, uint32_t* itemCount, uint32_t* submitted, uint32_t* flags, uint32_t* rankList, uint32_t* rankCount, cudaStream_t s)
{
    if (!numOcc || !numViews)
        return;
    k_select_ranked<<<numViews, kRankThreads, kRankCap * 12, s>>>(views, camParams, numOcc, occAabb6, occDrawDist, occMesh, occStart, occCount, sizeThreshold,
        maxTriangles, items, itemCap, itemCount, submitted, flags, rankList, rankCount);
}

void launch_draw_occluders_serial(const BufGeom& g, const ViewConst* views, int numViews, const uint32_t* rankList, const uint32_t* rankCount,
    const float* occAabb6, const float* models, const uint32_t* occMesh, const uint32_t* occStart, const uint32_t* occCount, const MeshDev* meshes,
    uint32_t maxTriangles, int* depth, TriSetup* scratch, uint32_t scratchCap, uint32_t* submitted, uint32_t* drawnSources, uint32_t* activeOccluders,
    uint32_t* flags, cudaStream_t s)
{
    if (numViews)
        k_draw_occluders_serial<<<numViews, kSerialThreads, 0, s>>>(g, views, rankList, rankCount, occAabb6, models, occMesh, occStart, occCount, meshes,
            maxTriangles, depth, scratch, scratchCap, submitted, drawnSources, activeOccluders, flags);
}

void launch_scan_regions(int numViews, const uint32_t* submitted, uint32_t slack, uint64_t poolCap, uint64_t* triBase, uint32_t* triCap, uint32_t* flags,
    cudaStream_t s)
{
    k_scan_regions<<<1, 1024, 0, s>>>(numViews, submitted, slack, poolCap, triBase, triCap, flags);
}

void launch_setup(const BufGeom& g, const TileGeom& tg, const ViewConst* views, int numViews, const DrawItem* items, uint32_t itemCap, uint32_t maxItems,
    const uint32_t* itemCount, const float* models, const MeshDev* meshes, const RasterDev& rd, cudaStream_t s)
{
    if (!maxItems || !numViews)
        return;
    // A few blocks per view sweep its items in slices.  Many views: slices of kSetupItems (64 boxes = 768 triangles = 3 full sweeps of the
    // block).  Few views (one 250k-triangle mesh = 977 items of 256 triangles): smaller slices, so that the launch still has a few blocks per SM.
    const uint32_t wantBlocks = numViews >= 148 ? 8u : std::max<uint32_t>(8u, (148u * 4u + numViews - 1) / numViews);
    uint32_t sliceItems = kSetupItems;
    while (sliceItems > 1 && (maxItems + sliceItems - 1) / sliceItems < wantBlocks)
        sliceItems >>= 1;
    const uint32_t slices = (maxItems + sliceItems - 1) / sliceItems;
    dim3 grid(std::min(slices, std::max(wantBlocks, 1u)), numViews);
    k_setup_triangles<<<grid, kSetupThreads, kSetupQueueBytes, s>>>(g, tg, views, items, itemCap, itemCount, models, meshes, rd, sliceItems);
    k_clip_triangles<<<148 * 8, 128, 0, s>>>(g, tg, views, rd);
}

void launch_fill_global(const BufGeom& g, int numViews, int* depth, const TriSetup* pool, const uint64_t* triBase, const uint32_t* triCap,
    const uint32_t* triCount, int blocksPerView, cudaStream_t s)
{
    if (!numViews)
        return;
    dim3 grid(blocksPerView, numViews);
    k_fill_global<<<grid, 256, 0, s>>>(g, depth, pool, triBase, triCap, triCount);
}

cudaError_t init_tile_kernel()
{
    cudaError_t e = cudaFuncSetAttribute(k_setup_triangles, cudaFuncAttributeMaxDynamicSharedMemorySize, kSetupQueueBytes);
    if (e != cudaSuccess)
        return e;
    return cudaFuncSetAttribute(k_fill_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * kTileH * 4);
}

void launch_fill_tiles(const BufGeom& g, const TileGeom& tg, int numViews, int* depth, int2* mips, const RasterDev& rd, cudaStream_t s)
{
    if (!numViews)
        return;
    if (!rd.secondPass)
        k_irregular_clear<<<dim3(4, numViews), 256, 0, s>>>(g, depth, rd.irrOfView);
    k_irregular_fill<<<148, 256, 0, s>>>(g, depth, rd.pool, rd.triBase, rd.irrList, rd.irrCap, rd.irrTotal);
    k_fill_tiles<<<dim3(tg.numTiles, numViews), kTileThreads, (size_t)tg.tileW * tg.tileH * 4, s>>>(g, tg, depth, mips, rd);
}

void launch_hierarchy(const BufGeom& g, int numViews, const int* depth, int2* mips, int firstLevel, cudaStream_t s)
{
    if (!numViews)
        return;
    for (int lv = firstLevel; lv < g.nlev; ++lv)
    {
        const int n = g.mw[lv] * g.mh[lv];
        dim3 grid((n + 255) / 256, numViews);
        k_mip_level<<<grid, 256, 0, s>>>(g, lv, depth, mips);
    }
}

} // namespace u3d
