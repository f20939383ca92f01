// Device-side types and exact-arithmetic helpers shared by the sm_100a kernels of the visibility path.
//
// Float contract (SURVEY.md hard parts 1-3): this translation unit is compiled with -fmad=false -prec-div=true
// -prec-sqrt=true -ftz=false, every expression keeps the reference's left-to-right evaluation order, and float->int
// conversions reproduce x86 cvttss2si (out of range / NaN -> INT_MIN).  "OB.cpp" below is the reference's
// Source/Urho3D/Graphics/OcclusionBuffer.cpp.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace u3d
{

constexpr float kZScale = 16777216.0f;   // OCCLUSION_Z_SCALE      OcclusionBuffer.h:86
constexpr float kXScale = 65536.0f;      // OCCLUSION_X_SCALE      OcclusionBuffer.h:85
constexpr int kMinMip = 8;               // OCCLUSION_MIN_SIZE     OcclusionBuffer.h:83
constexpr int kFixedBias = 16;           // OCCLUSION_FIXED_BIAS   OcclusionBuffer.h:87
constexpr float kRelBias = 0.00001f;     // OCCLUSION_RELATIVE_BIAS OcclusionBuffer.h:88
constexpr int kClearDepth = 16777216;    // (int)OCCLUSION_Z_SCALE, OB.cpp:1035
constexpr int kMaxLevels = 16;

enum : int { CULL_NONE = 0, CULL_CCW = 1, CULL_CW = 2 };
enum : int { OUTSIDE = 0, INTERSECTS = 1, INSIDE = 2 };
enum : int { QUERY_FRUSTUM = 0, QUERY_OCCLUDED = 1, QUERY_SHADOWCASTER = 2 };
enum : int { STAGE_QUERY = 0, STAGE_VISIBLE = 1 };

// Per-drawable meta bits packed into the first pad word of the 32-byte AABB record (BoundingBox.h:326-330 layout).
constexpr uint32_t kMetaOccludee = 1u << 8;
constexpr uint32_t kMetaCastShadows = 1u << 9;

/// Geometry of one occlusion buffer and its mip chain (OB.cpp:61-113); identical for every view of a batch.
struct BufGeom
{
    int w, h, nlev;
    int mw[kMaxLevels], mh[kMaxLevels];
    int moff[kMaxLevels];   // texel offset of each level inside one view's mip block
    int mipTexels;          // texels (int2 min,max) per view over all levels
    float sx, sy, ox, oy;   // CalculateViewport, OB.cpp:578-587
};

/// Per-view constants (OcclusionBuffer::SetView OB.cpp:115-127 + the query's frustum/flags).
struct ViewConst
{
    float vp[16];        // viewProj_ = projection * view, row-major
    float planes[24];    // 6 x (nx, ny, nz, d): NEAR, LEFT, RIGHT, UP, DOWN, FAR (Frustum.h:37-45)
    uint32_t drawableFlags;
    uint32_t viewMask;
    int cullMode;        // effective mode after the reverseCulling flip (OB.cpp:134-144)
    int queryMode;       // QUERY_*
    int useOcclusion;    // 0: no buffer (frustum only)
    int pad[3];
};

/// One set-up triangle, self-contained per half (closed form of the incremental loops OB.cpp:897-1001):
/// rows [y0,y1) use edge set T, rows [y1,y2) use edge set B.  For row y of a half starting at ys, k = y - ys:
///   xl = (lx + k*lxs) >> 16, xr = (rx + k*rxs) >> 16, depth at pixel x = lz + k*lzs + (x - xl)*dzdx,
/// all in wrapping 32-bit arithmetic, written to linear offset y*width + x when inside [0, w*h).
struct TriSetup
{
    int y0, y1, y2, dzdx;
    int tlx, tlxs, tlz, tlzs, trx, trxs;
    int blx, blxs, blz, blzs, brx, brxs;
};
static_assert(sizeof(TriSetup) == 64, "TriSetup must be 64 bytes");

struct MeshDev
{
    const unsigned char* vtx;
    const void* idx;      // nullptr: non-indexed
    uint32_t stride;
    uint32_t idxSize;     // 2 or 4 (0 when non-indexed)
};

struct DrawItem
{
    uint32_t model;   // index into the model-matrix array (12 floats each)
    uint32_t mesh;    // index into the MeshDev table
    uint32_t start;   // first index (or vertex when non-indexed)
    uint32_t count;   // number of indices/vertices; floor(count/3) triangles
};

struct V4 { float x, y, z, w; };
struct V3 { float x, y, z; };

// ---------------------------------------------------------------------------------------------------------------
// exact helpers
// ---------------------------------------------------------------------------------------------------------------

/// (int)x as x86 does it: truncation, INT_MIN for NaN / out of range.
__device__ __forceinline__ int trunc_x86(float v)
{
    return (v >= -2147483648.0f && v < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
}

/// RoundToInt (MathDefs.h:238): round half away from zero, then x86 conversion.
__device__ __forceinline__ int round_x86(float v)
{
    return trunc_x86(roundf(v));
}

/// ModelTransform, OB.cpp:543-551.
__device__ __forceinline__ V4 model_transform(const float* m, float x, float y, float z)
{
    V4 r;
    r.x = m[0] * x + m[1] * y + m[2] * z + m[3];
    r.y = m[4] * x + m[5] * y + m[6] * z + m[7];
    r.z = m[8] * x + m[9] * y + m[10] * z + m[11];
    r.w = m[12] * x + m[13] * y + m[14] * z + m[15];
    return r;
}

/// ViewportTransform, OB.cpp:553-561.
__device__ __forceinline__ V3 viewport_transform(const BufGeom& g, const V4& v)
{
    float invW = 1.0f / v.w;
    V3 r;
    r.x = invW * v.x * g.sx + g.ox;
    r.y = invW * v.y * g.sy + g.oy;
    r.z = invW * v.z * kZScale;
    return r;
}

/// Frustum::IsInside / IsInsideFast, Math/Frustum.h:123-159.  absNormal_ == |normal_| (Plane.h:83-88).
template <bool FAST> __device__ __forceinline__ int frustum_test(const float* planes, float mnx, float mny, float mnz, float mxx, float mxy, float mxz)
{
    float cx = (mxx + mnx) * 0.5f, cy = (mxy + mny) * 0.5f, cz = (mxz + mnz) * 0.5f;
    float ex = cx - mnx, ey = cy - mny, ez = cz - mnz;
    bool allInside = true;
#pragma unroll
    for (int i = 0; i < 6; ++i)
    {
        const float nx = planes[4 * i], ny = planes[4 * i + 1], nz = planes[4 * i + 2], d = planes[4 * i + 3];
        float dist = nx * cx + ny * cy + nz * cz + d;
        float absDist = fabsf(nx) * ex + fabsf(ny) * ey + fabsf(nz) * ez;
        if (dist < -absDist)
            return OUTSIDE;
        if (!FAST && dist < absDist)
            allInside = false;
    }
    return (FAST || allInside) ? INSIDE : INTERSECTS;
}

/// OcclusionBuffer::IsVisible, OB.cpp:336-456, as the exact predicate it computes (SURVEY hard part 5):
/// after the near-plane / off-screen early-outs, "visible" <=> exists pixel p in the clamped rect with z <= depth[p].
/// The reference walks every mip texel overlapping the rect level by level; here a quadtree descent visits only texels
/// whose [min,max] straddles z.  useMips = false reproduces the depthHierarchyDirty_ path (pixel scan, OB.cpp:440-455).
__device__ __forceinline__ bool ob_is_visible(const BufGeom& g, const float* vp, const int* __restrict__ depth, const int2* __restrict__ mips,
    bool useMips, float mnx, float mny, float mnz, float mxx, float mxy, float mxz)
{
    float minX = 0.f, maxX = 0.f, minY = 0.f, maxY = 0.f, minZ = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        V4 v = model_transform(vp, (i & 1) ? mxx : mnx, (i & 2) ? mxy : mny, (i & 4) ? mxz : mnz);
        v.z -= kRelBias;
        if (v.z <= 0.0f)
            return true;
        V3 s = viewport_transform(g, v);
        if (i == 0)
        {
            minX = maxX = s.x;
            minY = maxY = s.y;
            minZ = s.z;
        }
        else
        {
            if (s.x < minX) minX = s.x;
            if (s.x > maxX) maxX = s.x;
            if (s.y < minY) minY = s.y;
            if (s.y > maxY) maxY = s.y;
            if (s.z < minZ) minZ = s.z;
        }
    }
    int left = trunc_x86(minX - 1.5f), top = trunc_x86(minY - 1.5f), right = round_x86(maxX), bottom = round_x86(maxY);
    if (right < 0 || bottom < 0)
        return true;
    if (left >= g.w || top >= g.h)
        return true;
    if (left < 0) left = 0;
    if (top < 0) top = 0;
    if (right >= g.w) right = g.w - 1;
    if (bottom >= g.h) bottom = g.h - 1;
    const int z = (int)((uint32_t)round_x86(minZ) - (uint32_t)kFixedBias);

    if (!useMips)
    {
        for (int y = top; y <= bottom; ++y)
            for (int x = left; x <= right; ++x)
                if (z <= depth[y * g.w + x])
                    return true;
        return false;
    }

    // Quadtree descent.  A texel is conclusive unless min < z <= max ("mixed").  A mixed texel that lies completely inside the rect
    // proves visibility (its max is a pixel of the rect); only mixed texels cut by the rect's border are refined, so the walk
    // costs O(perimeter) instead of the reference's O(area) per level.  Stack entries: level (5 bits) | x (13 bits) | y (14 bits).
    uint32_t stack[3 * kMaxLevels + 4];
    // Start at the level whose texels are about the size of the rect (at most 2 x 2 texels overlap it) instead of at the coarsest
    // one: any set of texels covering the rect evaluates the same predicate, and small boxes skip the useless coarse levels.
    const int extent = max(right - left, bottom - top) + 1;
    int top_lv = extent <= 2 ? 0 : (32 - __clz(extent - 1)) - 1;
    top_lv = min(top_lv, g.nlev - 1);
    const int sh0 = top_lv + 1;
    for (int ry = top >> sh0; ry <= (bottom >> sh0); ++ry)
        for (int rx = left >> sh0; rx <= (right >> sh0); ++rx)
        {
            int sp = 0;
            stack[sp++] = ((uint32_t)top_lv << 27) | ((uint32_t)rx << 14) | (uint32_t)ry;
            while (sp)
            {
                const uint32_t e = stack[--sp];
                const int lv = (int)(e >> 27), x = (int)((e >> 14) & 0x1FFF), y = (int)(e & 0x3FFF);
                const int2 t = mips[g.moff[lv] + y * g.mw[lv] + x];
                if (z <= t.x)
                    return true;      // every pixel under this texel is at least as far as z, and the texel overlaps the rect
                if (z > t.y)
                    continue;         // every pixel under this texel is nearer than z
                // pixels covered by this texel: [x << s, ((x+1) << s) - 1] x [y << s, ...] clipped to the buffer, s = lv + 1
                const int s = lv + 1;
                const int px0 = x << s, py0 = y << s;
                const int px1 = min(((x + 1) << s) - 1, g.w - 1), py1 = min(((y + 1) << s) - 1, g.h - 1);
                if (px0 >= left && px1 <= right && py0 >= top && py1 <= bottom)
                    return true;      // mixed and fully inside the rect: the pixel holding max is in the rect and z <= max
                if (lv == 0)
                {
                    const int x0 = max(px0, left), x1 = min(px1, right);
                    const int y0 = max(py0, top), y1 = min(py1, bottom);
                    for (int py = y0; py <= y1; ++py)
                        for (int px = x0; px <= x1; ++px)
                            if (z <= depth[py * g.w + px])
                                return true;
                }
                else
                {
                    const int cl = lv - 1, sh = lv; // child level texel = pixel >> (cl + 1) = pixel >> lv
                    const int x0 = max(2 * x, left >> sh), x1 = min(min(2 * x + 1, right >> sh), g.mw[cl] - 1);
                    const int y0 = max(2 * y, top >> sh), y1 = min(min(2 * y + 1, bottom >> sh), g.mh[cl] - 1);
                    for (int cy = y0; cy <= y1; ++cy)
                        for (int cx = x0; cx <= x1; ++cx)
                            stack[sp++] = ((uint32_t)cl << 27) | ((uint32_t)cx << 14) | (uint32_t)cy;
                }
            }
        }
    return false;
}

} // namespace u3d
