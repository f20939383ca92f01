// Occluder rasterization for the visibility path: occluder selection, triangle set-up (transform, clip, project, cull,
// edge set-up) and depth fill, plus the depth-hierarchy build.  Replaces OcclusionBuffer::DrawTriangles / DrawBatch /
// DrawTriangle / ClipVertices / DrawTriangle2D / BuildDepthHierarchy (OB.cpp:200-329, 464-1002) of the reference.
//
// The reference's depth test is a min (OB.cpp:909-910), so triangles may be drawn in any order and from any number of
// threads with atomicMin and still give a bit-identical plane (SURVEY hard part 1, measured).
#include "u3d_device.cuh"
#include "u3d_kernels.h"

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
    if (k >= numOcc)
        return;
    const float* b = occAabb6 + 6 * (size_t)k;
    if (frustum_test<true>(views[v].planes, b[0], b[1], b[2], b[3], b[4], b[5]) == OUTSIDE)
        return;
    const uint32_t tris = occCount[k] / 3;
    if (!tris)
        return;
    const uint32_t nsub = (tris + kItemTris - 1) / kItemTris;
    const uint32_t base = atomicAdd(&itemCount[v], nsub);
    atomicAdd(&submitted[v], tris);
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

/// ViewportTransform x3 + SignedArea + cull test + DrawTriangle2D's set-up part (OB.cpp:629-637, 815-899).
/// Returns true when the reference would have called DrawTriangle2D (drawOk), and writes the record if the triangle has rows.
__device__ __forceinline__ bool setup_triangle(const BufGeom& g, int cullMode, const V4& c0, const V4& c1, const V4& c2, TriSetup& out, bool& hasRows)
{
    V3 p[3];
    p[0] = viewport_transform(g, c0);
    p[1] = viewport_transform(g, c1);
    p[2] = viewport_transform(g, c2);
    const float aX = p[0].x - p[1].x, aY = p[0].y - p[1].y, bX = p[2].x - p[1].x, bY = p[2].y - p[1].y;
    const bool clockwise = (aX * bY - aY * bX) < 0.0f; // SignedArea, OB.cpp:569-576
    hasRows = false;
    if (!(cullMode == CULL_NONE || (cullMode == CULL_CCW && clockwise) || (cullMode == CULL_CW && !clockwise)))
        return false;

    // Y sort, OB.cpp:820-872
    int top, mid, bot;
    bool midRight;
    if (p[0].y < p[1].y)
    {
        if (p[2].y < p[0].y) { top = 2; mid = 0; bot = 1; midRight = true; }
        else
        {
            top = 0;
            if (p[1].y < p[2].y) { mid = 1; bot = 2; midRight = true; }
            else { mid = 2; bot = 1; midRight = false; }
        }
    }
    else
    {
        if (p[2].y < p[1].y) { top = 2; mid = 1; bot = 0; midRight = false; }
        else
        {
            top = 1;
            if (p[0].y < p[2].y) { mid = 0; bot = 2; midRight = false; }
            else { mid = 2; bot = 0; midRight = true; }
        }
    }
    const V3 vt = p[top], vm = p[mid], vb = p[bot];
    const int topY = trunc_x86(vt.y), midY = trunc_x86(vm.y), botY = trunc_x86(vb.y);
    if (topY == botY)
        return true; // degenerate: reference returns from DrawTriangle2D but drawOk is already set
    if (!clockwise)
        midRight = !midRight;

    // Gradients, OB.cpp:762-778
    const float invdX = 1.0f / (((p[1].x - p[2].x) * (p[0].y - p[2].y)) - ((p[0].x - p[2].x) * (p[1].y - p[2].y)));
    const float invdY = -invdX;
    const float dzdx = invdX * (((p[1].z - p[2].z) * (p[0].y - p[2].y)) - ((p[0].z - p[2].z) * (p[1].y - p[2].y)));
    const float dzdy = invdY * (((p[1].z - p[2].z) * (p[0].x - p[2].x)) - ((p[0].z - p[2].z) * (p[1].x - p[2].x)));

    out.y0 = topY;
    out.y1 = midY;
    out.y2 = botY;
    out.dzdx = trunc_x86(dzdx);
    const EdgeI lng = make_edge(dzdx, dzdy, vt, vb, topY);
    // long edge advanced to the first row of the bottom half (it is stepped once per top-half row, OB.cpp:915-916)
    const uint32_t kTop = (uint32_t)midY - (uint32_t)topY;
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
    hasRows = true;
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

/// Append one record to the view's region (warp-uncoordinated; callers are divergent).
__device__ __forceinline__ void emit_tri(const TriSetup& ts, TriSetup* __restrict__ region, uint32_t cap, uint32_t* __restrict__ counter, uint32_t* __restrict__ flags)
{
    const uint32_t slot = atomicAdd(counter, 1u);
    if (slot < cap)
        region[slot] = ts;
    else
        atomicOr(flags, kFlagTriOverflow);
}

/// Slow path of DrawTriangle (OB.cpp:640-679): clip against the flagged planes, then set up every surviving piece.
__device__ __noinline__ bool clip_and_setup(const BufGeom& g, int cullMode, unsigned clipMask, V4 c0, V4 c1, V4 c2, TriSetup* region,
    uint32_t cap, uint32_t* counter, uint32_t* flags)
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
        TriSetup ts;
        bool rows;
        if (setup_triangle(g, cullMode, vv[3 * i], vv[3 * i + 1], vv[3 * i + 2], ts, rows))
        {
            drawOk = true;
            if (rows)
                emit_tri(ts, region, cap, counter, flags);
        }
    }
    return drawOk;
}

/// One block = kSetupItems draw items of one view; threads sweep the flat triangle index space of those items.
/// DrawBatch + DrawTriangle (OB.cpp:464-541, 589-683).
__global__ void __launch_bounds__(kSetupThreads) k_setup_triangles(BufGeom g, const ViewConst* __restrict__ views,
    const DrawItem* __restrict__ items, uint32_t itemCap, const uint32_t* __restrict__ itemCount, const float* __restrict__ models,
    const MeshDev* __restrict__ meshes, TriSetup* __restrict__ pool, const uint64_t* __restrict__ triBase, const uint32_t* __restrict__ triCap,
    uint32_t* __restrict__ triCount, uint32_t* __restrict__ drawnSources, uint32_t* __restrict__ flags)
{
    const int v = blockIdx.y;
    const uint32_t nItems = itemCount[v];
    const uint32_t itemBase = blockIdx.x * kSetupItems;
    if (itemBase >= nItems)
        return;
    const uint32_t nLocal = min((uint32_t)kSetupItems, nItems - itemBase);

    __shared__ float sMvp[kSetupItems][16];
    __shared__ DrawItem sItem[kSetupItems];
    __shared__ uint32_t sPre[kSetupItems + 1];

    const ViewConst& vc = views[v];
    if (threadIdx.x < nLocal)
        sItem[threadIdx.x] = items[(size_t)v * itemCap + itemBase + threadIdx.x];
    __syncthreads();
    // modelViewProj = viewProj_ * model  (Matrix4 * Matrix3x4, Matrix4.cpp:43-63)
    for (uint32_t e = threadIdx.x; e < nLocal * 16; e += blockDim.x)
    {
        const uint32_t it = e >> 4, r = (e >> 2) & 3, c = e & 3;
        const float* a = vc.vp + 4 * r;
        const float* b = models + 12 * (size_t)sItem[it].model;
        float val;
        if (c < 3)
            val = a[0] * b[c] + a[1] * b[4 + c] + a[2] * b[8 + c];
        else
            val = a[0] * b[3] + a[1] * b[7] + a[2] * b[11] + a[3];
        sMvp[it][4 * r + c] = val;
    }
    if (threadIdx.x == 0)
    {
        uint32_t acc = 0;
        for (uint32_t i = 0; i < nLocal; ++i)
        {
            sPre[i] = acc;
            acc += sItem[i].count / 3;
        }
        sPre[nLocal] = acc;
    }
    __syncthreads();

    const uint32_t total = sPre[nLocal];
    TriSetup* region = pool + triBase[v];
    const uint32_t cap = triCap[v];
    const int cullMode = vc.cullMode;
    uint32_t drawn = 0;
    for (uint32_t t = threadIdx.x; t < total; t += blockDim.x)
    {
        // owning item: largest i with sPre[i] <= t
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
        const float* p0 = (const float*)(mesh.vtx + (size_t)i0 * mesh.stride);
        const float* p1 = (const float*)(mesh.vtx + (size_t)i1 * mesh.stride);
        const float* p2 = (const float*)(mesh.vtx + (size_t)i2 * mesh.stride);
        const float* mvp = sMvp[lo];
        const V4 c0 = model_transform(mvp, p0[0], p0[1], p0[2]);
        const V4 c1 = model_transform(mvp, p1[0], p1[1], p1[2]);
        const V4 c2 = model_transform(mvp, p2[0], p2[1], p2[2]);

        // clip masks, OB.cpp:596-620
        unsigned m0 = 0, m1 = 0, m2 = 0;
        if (c0.x > c0.w) m0 |= 1; if (c0.x < -c0.w) m0 |= 2; if (c0.y > c0.w) m0 |= 4; if (c0.y < -c0.w) m0 |= 8; if (c0.z > c0.w) m0 |= 16; if (c0.z < 0.0f) m0 |= 32;
        if (c1.x > c1.w) m1 |= 1; if (c1.x < -c1.w) m1 |= 2; if (c1.y > c1.w) m1 |= 4; if (c1.y < -c1.w) m1 |= 8; if (c1.z > c1.w) m1 |= 16; if (c1.z < 0.0f) m1 |= 32;
        if (c2.x > c2.w) m2 |= 1; if (c2.x < -c2.w) m2 |= 2; if (c2.y > c2.w) m2 |= 4; if (c2.y < -c2.w) m2 |= 8; if (c2.z > c2.w) m2 |= 16; if (c2.z < 0.0f) m2 |= 32;
        if (m0 & m1 & m2)
            continue;
        const unsigned clipMask = m0 | m1 | m2;
        bool drawOk;
        if (!clipMask)
        {
            TriSetup ts;
            bool rows;
            drawOk = setup_triangle(g, cullMode, c0, c1, c2, ts, rows);
            if (drawOk && rows)
                emit_tri(ts, region, cap, &triCount[v], flags);
        }
        else
            drawOk = clip_and_setup(g, cullMode, clipMask, c0, c1, c2, region, cap, &triCount[v], flags);
        drawn += drawOk ? 1u : 0u;
    }
    // ++numTriangles_ per source triangle that drew anything (OB.cpp:681-682)
    for (int o = 16; o; o >>= 1)
        drawn += __shfl_xor_sync(0xffffffffu, drawn, o);
    if ((threadIdx.x & 31) == 0 && drawn)
        atomicAdd(&drawnSources[v], drawn);
}

// ---------------------------------------------------------------------------------------------------------------
// fill (general path): one warp per set-up triangle, lanes across the span, atomicMin straight into the view's plane in
// global memory with the reference's linear addressing (row + x may spill into the neighbouring row, SURVEY hard part 4).
// Handles every record, including the irregular ones the tile path refuses.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fill_half_global(int* __restrict__ depth, int w, int h, int ya, int yb, int lx, int lxs, int lz, int lzs, int rx, int rxs,
    int dzdx, int lane)
{
    // rows whose linear span could reach [0, w*h): |x>>16| <= 32768
    const long long guard = 32768 / w + 2;
    const long long lo = max((long long)ya, -guard), hi = min((long long)yb, (long long)h + guard);
    const long long npix = (long long)w * h;
    for (long long y = lo; y < hi; ++y)
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
                if (z < depth[off])
                    atomicMin(depth + off, z);
            }
        }
    }
}

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
        if (ts.y0 != ts.y1)
            fill_half_global(depth, g.w, g.h, ts.y0, ts.y1, ts.tlx, ts.tlxs, ts.tlz, ts.tlzs, ts.trx, ts.trxs, ts.dzdx, lane);
        if (ts.y1 != ts.y2)
            fill_half_global(depth, g.w, g.h, ts.y1, ts.y2, ts.blx, ts.blxs, ts.blz, ts.blzs, ts.brx, ts.brxs, ts.dzdx, lane);
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

void launch_scan_regions(int numViews, const uint32_t* submitted, uint32_t slack, uint64_t poolCap, uint64_t* triBase, uint32_t* triCap, uint32_t* flags,
    cudaStream_t s)
{
    k_scan_regions<<<1, 1024, 0, s>>>(numViews, submitted, slack, poolCap, triBase, triCap, flags);
}

void launch_setup(const BufGeom& g, const ViewConst* views, int numViews, const DrawItem* items, uint32_t itemCap, uint32_t maxItems, const uint32_t* itemCount,
    const float* models, const MeshDev* meshes, TriSetup* pool, const uint64_t* triBase, const uint32_t* triCap, uint32_t* triCount,
    uint32_t* drawnSources, uint32_t* flags, cudaStream_t s)
{
    if (!maxItems || !numViews)
        return;
    dim3 grid((maxItems + kSetupItems - 1) / kSetupItems, numViews);
    k_setup_triangles<<<grid, kSetupThreads, 0, s>>>(g, views, items, itemCap, itemCount, models, meshes, pool, triBase, triCap, triCount, drawnSources, flags);
}

void launch_fill_global(const BufGeom& g, int numViews, int* depth, const TriSetup* pool, const uint64_t* triBase, const uint32_t* triCap,
    const uint32_t* triCount, int blocksPerView, cudaStream_t s)
{
    if (!numViews)
        return;
    dim3 grid(blocksPerView, numViews);
    k_fill_global<<<grid, 256, 0, s>>>(g, depth, pool, triBase, triCap, triCount);
}

void launch_hierarchy(const BufGeom& g, int numViews, const int* depth, int2* mips, cudaStream_t s)
{
    if (!numViews)
        return;
    for (int lv = 0; lv < g.nlev; ++lv)
    {
        const int n = g.mw[lv] * g.mh[lv];
        dim3 grid((n + 255) / 256, numViews);
        k_mip_level<<<grid, 256, 0, s>>>(g, lv, depth, mips);
    }
}

} // namespace u3d
