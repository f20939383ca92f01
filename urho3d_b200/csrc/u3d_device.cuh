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
enum : int { QUERY_FRUSTUM = 0, QUERY_OCCLUDED = 1, QUERY_SHADOWCASTER = 2, QUERY_SPHERE = 3, QUERY_BOX = 4, QUERY_POINT = 5, QUERY_RAY = 6,
    QUERY_RAY_CANDIDATES = 7, QUERY_ZONE_OCCLUDER = 8 };
enum : int { STAGE_QUERY = 0, STAGE_VISIBLE = 1, STAGE_INVIEW = 2 };

// Per-drawable meta bits packed into the first pad word of the 32-byte AABB record (BoundingBox.h:326-330 layout).
constexpr uint32_t kMetaOccludee = 1u << 8;
constexpr uint32_t kMetaCastShadows = 1u << 9;
constexpr uint32_t kMetaOccluder = 1u << 10;   // Drawable::IsOccluder() (ZoneOccluderOctreeQuery, View.cpp:106)

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
    int ortho;           // Camera::IsOrthographic()
    float viewZ[4];      // third row of the view matrix (m20, m21, m22, m23): view-space Z of a point
    float camPos[3];     // camera world position
    int pad;
};

/// One shadow split of View::ProcessShadowCasters (View.cpp:2379-2407); same layout as u3d_shadow_split (include/u3d_gpu.h).
struct ShadowSplitDev
{
    float lightView[12];           // shadowCamera->GetView()
    float shadowPlanes[24];        // shadowCamera->GetFrustum(): tested for point lights (View.cpp:2422)
    float lvfPlanes[24];           // lightViewFrustum (scene split frustum in light view space)
    float lvfMaxZ;                 // BoundingBox(lightViewFrustum).max_.z_
    float shadowFarClip;           // shadowCamera->GetFarClip()
    int shadowOrtho;               // shadowCamera->IsOrthographic()
    int lightType;                 // 0 directional, 1 spot, 2 point
    uint32_t lightMask;
    int skip;                      // degenerate split frustum: no shadow casters
    int reserved[2];
};
static_assert(sizeof(ShadowSplitDev) == 272, "ShadowSplitDev layout");

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
    // All six planes, no early return: a warp runs every plane anyway as long as one of its lanes is still inside, and the reference's
    // answer depends only on whether SOME plane rejects the box / cuts it (the comparisons are the same, NaNs compare false on both sides).
    bool outside = false, allInside = true;
#pragma unroll
    for (int i = 0; i < 6; ++i)
    {
        const float nx = planes[4 * i], ny = planes[4 * i + 1], nz = planes[4 * i + 2], d = planes[4 * i + 3];
        float dist = nx * cx + ny * cy + nz * cz + d;
        float absDist = fabsf(nx) * ex + fabsf(ny) * ey + fabsf(nz) * ez;
        outside |= dist < -absDist;
        if (!FAST)
            allInside &= !(dist < absDist);
    }
    return outside ? OUTSIDE : (FAST || allInside) ? INSIDE : INTERSECTS;
}

/// Ray::HitDistance(const BoundingBox&), Math/Ray.cpp:71-148: 0 when the origin is inside, else the nearest entry through one of the
/// (at most three) faces turned towards the origin, +inf for a miss or an undefined box.  ray = origin.xyz, direction.xyz.
__device__ __forceinline__ float ray_hit_distance(const float* ray, float mnx, float mny, float mnz, float mxx, float mxy, float mxz)
{
    const float inf = __int_as_float(0x7f800000);
    const float ox = ray[0], oy = ray[1], oz = ray[2], dx = ray[3], dy = ray[4], dz = ray[5];
    if (mnx == inf)
        return inf;
    if (!(ox < mnx || ox > mxx || oy < mny || oy > mxy || oz < mnz || oz > mxz))
        return 0.0f;
    float dist = inf;
    // x faces
    if ((ox < mnx && dx > 0.0f) || (ox > mxx && dx < 0.0f))
    {
        const float x = ((ox < mnx ? mnx : mxx) - ox) / dx;
        if (x < dist)
        {
            const float py = oy + x * dy, pz = oz + x * dz;
            if (py >= mny && py <= mxy && pz >= mnz && pz <= mxz)
                dist = x;
        }
    }
    // y faces
    if ((oy < mny && dy > 0.0f) || (oy > mxy && dy < 0.0f))
    {
        const float x = ((oy < mny ? mny : mxy) - oy) / dy;
        if (x < dist)
        {
            const float px = ox + x * dx, pz = oz + x * dz;
            if (px >= mnx && px <= mxx && pz >= mnz && pz <= mxz)
                dist = x;
        }
    }
    // z faces
    if ((oz < mnz && dz > 0.0f) || (oz > mxz && dz < 0.0f))
    {
        const float x = ((oz < mnz ? mnz : mxz) - oz) / dz;
        if (x < dist)
        {
            const float px = ox + x * dx, py = oy + x * dy;
            if (px >= mnx && px <= mxx && py >= mny && py <= mxy)
                dist = x;
        }
    }
    return dist;
}

/// View::IsShadowCasterVisible (View.cpp:2452-2489) after BoundingBox::Transformed(lightView) (BoundingBox.cpp:131-156, the default URHO3D_SSE
/// build: each row is summed as (m0*cx + m2*cz) + (m1*cy + m3), the extents as (|m0*hx| + |m2*hz|) + (|m1*hy| + 0)).
__device__ __forceinline__ bool shadow_caster_visible(const ShadowSplitDev& sp, float mnx, float mny, float mnz, float mxx, float mxy, float mxz, bool inView)
{
    const float cx = (mnx + mxx) * 0.5f, cy = (mny + mxy) * 0.5f, cz = (mnz + mxz) * 0.5f;
    const float hx = cx - mnx, hy = cy - mny, hz = cz - mnz;
    float lo[3], hi[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
    {
        const float* row = sp.lightView + 4 * r;
        const float c = (row[0] * cx + row[2] * cz) + (row[1] * cy + row[3]);
        const float d = (fabsf(row[0] * hx) + fabsf(row[2] * hz)) + (fabsf(row[1] * hy) + 0.0f);
        lo[r] = c - d;
        hi[r] = c + d;
    }
    if (sp.shadowOrtho)
    {
        // extrude up to the far edge of the frustum's light-space box
        hi[2] = hi[2] > sp.lvfMaxZ ? hi[2] : sp.lvfMaxZ;
        return frustum_test<true>(sp.lvfPlanes, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]) != OUTSIDE;
    }
    if (inView)
        return true; // a visible object's shadow is visible too (View.cpp:2464-2465)
    const float ccx = (hi[0] + lo[0]) * 0.5f, ccy = (hi[1] + lo[1]) * 0.5f, ccz = (hi[2] + lo[2]) * 0.5f;
    const float len2 = ccx * ccx + ccy * ccy + ccz * ccz;
    float dx = ccx, dy = ccy, dz = ccz; // Ray(center, center).direction_ = center.Normalized() (Vector3.h:425-435)
    if (!(fabsf(len2 - 1.0f) <= 0.000001f) && len2 > 0.0f)
    {
        const float inv = 1.0f / sqrtf(len2);
        dx = ccx * inv;
        dy = ccy * inv;
        dz = ccz * inv;
    }
    const float extrusion = sp.shadowFarClip;
    float original = sqrtf(len2); // Clamp(center.Length(), M_EPSILON, extrusionDistance)
    if (original < 0.000001f)
        original = 0.000001f;
    else if (original > extrusion)
        original = extrusion;
    const float factor = extrusion / original;
    const float ncx = dx * extrusion, ncy = dy * extrusion, ncz = dz * extrusion;
    const float ex = (hi[0] - lo[0]) * factor * 0.5f, ey = (hi[1] - lo[1]) * factor * 0.5f, ez = (hi[2] - lo[2]) * factor * 0.5f;
    const float e0 = ncx - ex, e1 = ncy - ey, e2 = ncz - ez, e3 = ncx + ex, e4 = ncy + ey, e5 = ncz + ez;
    lo[0] = lo[0] < e0 ? lo[0] : e0; // BoundingBox::Merge, SSE min / max semantics (BoundingBox.h:201-206)
    lo[1] = lo[1] < e1 ? lo[1] : e1;
    lo[2] = lo[2] < e2 ? lo[2] : e2;
    hi[0] = hi[0] > e3 ? hi[0] : e3;
    hi[1] = hi[1] > e4 ? hi[1] : e4;
    hi[2] = hi[2] > e5 ? hi[2] : e5;
    return frustum_test<true>(sp.lvfPlanes, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]) != OUTSIDE;
}

/// Octant / drawable test of every GPU-evaluable OctreeQuery.  FAST = the drawable form (IsInsideFast), else the 3-way octant form.
/// Frustum-type queries (QUERY_FRUSTUM / OCCLUDED / SHADOWCASTER) use the 6 planes; the shape queries carry their parameters in the
/// same 24 floats: sphere = (centre.xyz, radius)  Sphere::IsInside / IsInsideFast (Math/Sphere.cpp:136-256);
/// box = (min.xyz, max.xyz)  BoundingBox::IsInside / IsInsideFast (BoundingBox.h:295-315) with the QUERY's box as `this`;
/// point = (xyz)  BoundingBox::IsInside(point) with the octant / drawable box as `this` (BoundingBox.h:285-292; OctreeQuery.cpp:32-52);
/// ray = (origin.xyz, direction.xyz, maxDistance)  RayOctreeQuery: HitDistance(box) < maxDistance for octants (Octree.cpp:257-261, never
/// "inside") and, in QUERY_RAY, for drawables (Drawable::ProcessRayQuery, Drawable.cpp:118-121); QUERY_RAY_CANDIDATES passes every drawable
/// of a touched octant (GetDrawablesOnlyInternal, Octree.cpp:284-309: RaycastSingle's candidate pass).
template <bool FAST> __device__ __forceinline__ int shape_test(int mode, const float* p, float mnx, float mny, float mnz, float mxx, float mxy, float mxz)
{
    if (mode <= QUERY_SHADOWCASTER || mode == QUERY_ZONE_OCCLUDER)
        return frustum_test<FAST>(p, mnx, mny, mnz, mxx, mxy, mxz);
    if (mode == QUERY_SPHERE)
    {
        const float cx = p[0], cy = p[1], cz = p[2];
        const float radiusSquared = p[3] * p[3];
        float distSquared = 0.0f, temp;
        if (cx < mnx) { temp = cx - mnx; distSquared += temp * temp; }
        else if (cx > mxx) { temp = cx - mxx; distSquared += temp * temp; }
        if (cy < mny) { temp = cy - mny; distSquared += temp * temp; }
        else if (cy > mxy) { temp = cy - mxy; distSquared += temp * temp; }
        if (cz < mnz) { temp = cz - mnz; distSquared += temp * temp; }
        else if (cz > mxz) { temp = cz - mxz; distSquared += temp * temp; }
        if (distSquared >= radiusSquared)
            return OUTSIDE;
        if (FAST)
            return INSIDE;
        const float ax = mnx - cx, ay = mny - cy, az = mnz - cz, bx = mxx - cx, by = mxy - cy, bz = mxz - cz;
        // corner order of the reference: ---, +--, ++-, -+-, -++, --+, +-+, +++
        if (ax * ax + ay * ay + az * az >= radiusSquared) return INTERSECTS;
        if (bx * bx + ay * ay + az * az >= radiusSquared) return INTERSECTS;
        if (bx * bx + by * by + az * az >= radiusSquared) return INTERSECTS;
        if (ax * ax + by * by + az * az >= radiusSquared) return INTERSECTS;
        if (ax * ax + by * by + bz * bz >= radiusSquared) return INTERSECTS;
        if (ax * ax + ay * ay + bz * bz >= radiusSquared) return INTERSECTS;
        if (bx * bx + ay * ay + bz * bz >= radiusSquared) return INTERSECTS;
        if (bx * bx + by * by + bz * bz >= radiusSquared) return INTERSECTS;
        return INSIDE;
    }
    if (mode == QUERY_BOX)
    {
        if (mxx < p[0] || mnx > p[3] || mxy < p[1] || mny > p[4] || mxz < p[2] || mnz > p[5])
            return OUTSIDE;
        if (!FAST && (mnx < p[0] || mxx > p[3] || mny < p[1] || mxy > p[4] || mnz < p[2] || mxz > p[5]))
            return INTERSECTS;
        return INSIDE;
    }
    if (mode == QUERY_RAY || mode == QUERY_RAY_CANDIDATES)
    {
        if (FAST && mode == QUERY_RAY_CANDIDATES)
            return INSIDE;
        const float d = ray_hit_distance(p, mnx, mny, mnz, mxx, mxy, mxz);
        // octants: `if (octantDist >= maxDistance_) return;` (Octree.cpp:259-261); drawables: `if (distance < maxDistance_)` (Drawable.cpp:120-121).
        // The two differ when maxDistance is NaN: every octant is walked and no drawable is a hit.
        if (FAST ? !(d < p[6]) : d >= p[6])
            return OUTSIDE;
        return FAST ? INSIDE : INTERSECTS;
    }
    // QUERY_POINT
    if (p[0] < mnx || p[0] > mxx || p[1] < mny || p[1] > mxy || p[2] < mnz || p[2] > mxz)
        return OUTSIDE;
    return INSIDE;
}

/// Screen rect + biased integer depth of a box (the float part of OcclusionBuffer::IsVisible, OB.cpp:341-402).
struct BoxRect
{
    int left, top, right, bottom; // clamped, inclusive
    int z;
};

/// Returns true when the reference answers "visible" before touching the buffer (near-plane crossing OB.cpp:359/370, rect off screen
/// :386-389); otherwise fills `r`.
__device__ __forceinline__ bool ob_project(const BufGeom& g, const float* vp, float mnx, float mny, float mnz, float mxx, float mxy, float mxz, BoxRect& r)
{
    // The reference's `if (s < min) min = s` chain (OB.cpp:373-380) skips NaN corners and never leaves a NaN first corner: fminf / fmaxf
    // (one FMNMX each, they return the other operand for a NaN) do the same for every corner but the first, which is put back below.
    // (They may pick -0 where the chain keeps +0; both become the same integer further down.)
    float minX = 0.f, maxX = 0.f, minY = 0.f, maxY = 0.f, minZ = 0.f;
    float firstX = 0.f, firstY = 0.f, firstZ = 0.f;
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
            minX = maxX = firstX = s.x;
            minY = maxY = firstY = s.y;
            minZ = firstZ = s.z;
        }
        else
        {
            minX = fminf(minX, s.x);
            maxX = fmaxf(maxX, s.x);
            minY = fminf(minY, s.y);
            maxY = fmaxf(maxY, s.y);
            minZ = fminf(minZ, s.z);
        }
    }
    if (firstX != firstX) minX = maxX = firstX;
    if (firstY != firstY) minY = maxY = firstY;
    if (firstZ != firstZ) minZ = firstZ;
    int left = trunc_x86(minX - 1.5f), top = trunc_x86(minY - 1.5f), right = round_x86(maxX), bottom = round_x86(maxY);
    if (right < 0 || bottom < 0)
        return true;
    if (left >= g.w || top >= g.h)
        return true;
    if (left < 0) left = 0;
    if (top < 0) top = 0;
    if (right >= g.w) right = g.w - 1;
    if (bottom >= g.h) bottom = g.h - 1;
    r.left = left;
    r.top = top;
    r.right = right;
    r.bottom = bottom;
    r.z = (int)((uint32_t)round_x86(minZ) - (uint32_t)kFixedBias);
    return false;
}

/// Level at which at most 2 x 2 texels overlap the rect (texel size >= rect extent), clamped to the coarsest level.
__device__ __forceinline__ int start_level(const BufGeom& g, const BoxRect& r)
{
    const int extent = max(r.right - r.left, r.bottom - r.top) + 1;
    const int lv = extent <= 2 ? 0 : (32 - __clz(extent - 1)) - 1;
    return min(lv, g.nlev - 1);
}

__device__ __forceinline__ uint32_t pack_texel(int lv, int x, int y) { return ((uint32_t)lv << 27) | ((uint32_t)x << 14) | (uint32_t)y; }

/// One texel of the walk.  Returns 1 = proves visible, 0 = nothing visible under it (inside the rect), 2 = must be refined.
/// A texel is conclusive unless min < z <= max ("mixed"); a mixed texel that lies completely inside the rect proves visibility
/// (the pixel holding its max is in the rect), so only mixed texels cut by the rect's border are refined.
__device__ __forceinline__ int classify_texel(const BufGeom& g, const int2* __restrict__ mips, const BoxRect& r, int lv, int x, int y)
{
    const int2 t = mips[g.moff[lv] + y * g.mw[lv] + x];
    if (r.z <= t.x)
        return 1;
    if (r.z > t.y)
        return 0;
    const int s = lv + 1;
    const int px0 = x << s, py0 = y << s;
    const int px1 = min(((x + 1) << s) - 1, g.w - 1), py1 = min(((y + 1) << s) - 1, g.h - 1);
    if (px0 >= r.left && px1 <= r.right && py0 >= r.top && py1 <= r.bottom)
        return 1;
    return 2;
}

/// Pixels of a level-0 texel that lie in the rect.
__device__ __forceinline__ bool texel_pixels_visible(const BufGeom& g, const int* __restrict__ depth, const BoxRect& r, int x, int y)
{
    const int x0 = max(2 * x, r.left), x1 = min(min(2 * x + 1, r.right), g.w - 1);
    const int y0 = max(2 * y, r.top), y1 = min(min(2 * y + 1, r.bottom), g.h - 1);
    for (int py = y0; py <= y1; ++py)
        for (int px = x0; px <= x1; ++px)
            if (r.z <= depth[py * g.w + px])
                return true;
    return false;
}

/// The buffer part of OcclusionBuffer::IsVisible (OB.cpp:404-455) as the exact predicate it computes (SURVEY hard part 5):
/// "visible" <=> exists pixel p in the rect with z <= depth[p].  The reference rescans every overlapping texel level by level;
/// here a depth-first walk refines only border-cut mixed texels, starting at the level matched to the rect's size.
/// useMips = false reproduces the depthHierarchyDirty_ path (pixel scan, OB.cpp:440-455).
__device__ __forceinline__ bool range_visible_thread(const BufGeom& g, const int* __restrict__ depth, const int2* __restrict__ mips, bool useMips,
    const BoxRect& r)
{
    if (!useMips)
    {
        for (int y = r.top; y <= r.bottom; ++y)
            for (int x = r.left; x <= r.right; ++x)
                if (r.z <= depth[y * g.w + x])
                    return true;
        return false;
    }
    uint32_t stack[3 * kMaxLevels + 4];
    const int top_lv = start_level(g, r);
    const int sh0 = top_lv + 1;
    for (int ry = r.top >> sh0; ry <= (r.bottom >> sh0); ++ry)
        for (int rx = r.left >> sh0; rx <= (r.right >> sh0); ++rx)
        {
            int sp = 0;
            stack[sp++] = pack_texel(top_lv, rx, ry);
            while (sp)
            {
                const uint32_t e = stack[--sp];
                const int lv = (int)(e >> 27), x = (int)((e >> 14) & 0x1FFF), y = (int)(e & 0x3FFF);
                const int c = classify_texel(g, mips, r, lv, x, y);
                if (c == 1)
                    return true;
                if (c == 0)
                    continue;
                if (lv == 0)
                {
                    if (texel_pixels_visible(g, depth, r, x, y))
                        return true;
                }
                else
                {
                    const int cl = lv - 1, sh = lv; // child level texel = pixel >> (cl + 1) = pixel >> lv
                    const int x0 = max(2 * x, r.left >> sh), x1 = min(min(2 * x + 1, r.right >> sh), g.mw[cl] - 1);
                    const int y0 = max(2 * y, r.top >> sh), y1 = min(min(2 * y + 1, r.bottom >> sh), g.mh[cl] - 1);
                    for (int cy = y0; cy <= y1; ++cy)
                        for (int cx = x0; cx <= x1; ++cx)
                            stack[sp++] = pack_texel(cl, cx, cy);
                }
            }
        }
    return false;
}

__device__ __forceinline__ bool range_visible_warp_from(const BufGeom& g, const int* __restrict__ depth, const int2* __restrict__ mips, const BoxRect& r,
    uint32_t* wstack, int n, int cap);

/// Same predicate evaluated by a whole warp for ONE rect (all lanes pass the same r): the pending texels sit on a per-warp stack in
/// shared memory and are classified 32 at a time, so a box whose border cuts hundreds of mixed texels costs tens of steps instead
/// of hundreds of dependent loads on one lane.  `wstack` has `cap` entries; if it would overflow, the surplus subtrees are walked
/// by single lanes.  Must be called by all 32 lanes.
__device__ __forceinline__ bool range_visible_warp(const BufGeom& g, const int* __restrict__ depth, const int2* __restrict__ mips, const BoxRect& r,
    uint32_t* wstack, int cap)
{
    const int lane = threadIdx.x & 31;
    const int top_lv = start_level(g, r);
    const int sh0 = top_lv + 1;
    const int rx0 = r.left >> sh0, ry0 = r.top >> sh0;
    const int nx = (r.right >> sh0) - rx0 + 1, ny = (r.bottom >> sh0) - ry0 + 1;
    const int n = nx * ny; // <= 64: the coarsest level is at most 8 x 8 texels, finer start levels give at most 2 x 2
    __syncwarp();
    for (int i = lane; i < n; i += 32)
        wstack[i] = pack_texel(top_lv, rx0 + i % nx, ry0 + i / nx);
    __syncwarp();
    return range_visible_warp_from(g, depth, mips, r, wstack, n, cap);
}

/// The cooperative walk from `n` pending texels already on the warp's stack (an entry with bit 31 set is known to be mixed and cut by
/// the rect's border: it is refined without being classified again).  Must be called by all 32 lanes with the same arguments.
__device__ __forceinline__ bool range_visible_warp_from(const BufGeom& g, const int* __restrict__ depth, const int2* __restrict__ mips, const BoxRect& r,
    uint32_t* wstack, int n, int cap)
{
    const int lane = threadIdx.x & 31;
    while (n > 0)
    {
        const int take = min(n, 32);
        n -= take;
        uint32_t e = 0;
        int c = 0, lv = 0, x = 0, y = 0;
        if (lane < take)
        {
            e = wstack[n + lane];
            lv = (int)((e >> 27) & 15u);
            x = (int)((e >> 14) & 0x1FFF);
            y = (int)(e & 0x3FFF);
            c = (e >> 31) ? 2 : classify_texel(g, mips, r, lv, x, y);
            if (c == 2 && lv == 0)
                c = texel_pixels_visible(g, depth, r, x, y) ? 1 : 0;
        }
        if (__any_sync(0xffffffffu, c == 1))
            return true;
        // children of the texels that must be refined
        int x0 = 0, x1 = -1, y0 = 0, y1 = -1, cl = 0;
        if (c == 2)
        {
            cl = lv - 1;
            const int sh = lv;
            x0 = max(2 * x, r.left >> sh);
            x1 = min(min(2 * x + 1, r.right >> sh), g.mw[cl] - 1);
            y0 = max(2 * y, r.top >> sh);
            y1 = min(min(2 * y + 1, r.bottom >> sh), g.mh[cl] - 1);
        }
        const int kids = c == 2 ? (x1 - x0 + 1) * (y1 - y0 + 1) : 0;
        int off = kids;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, off, o);
            if (lane >= o)
                off += t;
        }
        const int total = __shfl_sync(0xffffffffu, off, 31);
        off -= kids;
        __syncwarp();
        bool spill = false;
        if (kids)
        {
            if (n + off + kids <= cap)
            {
                int k = n + off;
                for (int cy = y0; cy <= y1; ++cy)
                    for (int cx = x0; cx <= x1; ++cx)
                        wstack[k++] = pack_texel(cl, cx, cy);
            }
            else
                spill = true;
        }
        // lanes whose children did not fit walk their subtree alone (rare: only perimeters far beyond the stack size)
        bool lonely = false;
        if (spill)
        {
            BoxRect sub = r;
            // restrict the rect to this texel's footprint: same predicate over the part of the rect under the texel
            const int s = lv + 1;
            sub.left = max(r.left, x << s);
            sub.top = max(r.top, y << s);
            sub.right = min(r.right, ((x + 1) << s) - 1);
            sub.bottom = min(r.bottom, ((y + 1) << s) - 1);
            lonely = range_visible_thread(g, depth, mips, true, sub);
        }
        if (__any_sync(0xffffffffu, lonely))
            return true;
        // stack grows by the children that fitted: those of lanes in prefix order up to the first spilling lane are contiguous only
        // if no earlier lane spilled; recompute the kept total
        const uint32_t spillMask = __ballot_sync(0xffffffffu, spill);
        if (spillMask == 0)
            n += total;
        else
        {
            // compact: lanes that kept their children re-push them densely
            int keep = spill ? 0 : kids;
            int koff = keep;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int t = __shfl_up_sync(0xffffffffu, koff, o);
                if (lane >= o)
                    koff += t;
            }
            const int ktotal = __shfl_sync(0xffffffffu, koff, 31);
            koff -= keep;
            __syncwarp();
            if (keep)
            {
                int k = n + koff;
                for (int cy = y0; cy <= y1; ++cy)
                    for (int cx = x0; cx <= x1; ++cx)
                        wstack[k++] = pack_texel(cl, cx, cy);
            }
            n += ktotal;
        }
        __syncwarp();
    }
    return false;
}

__device__ __forceinline__ uint32_t pack_pool(int box, int lv, int x, int y)
{
    return ((uint32_t)box << 27) | ((uint32_t)lv << 24) | ((uint32_t)x << 12) | (uint32_t)y;
}

/// Same predicate for up to 32 rects at once, one per lane (`need` = this lane has a rect to test): the pending texels of ALL the
/// warp's boxes share one per-warp stack in shared memory and are classified 32 at a time, each lane fetching the rect of the box
/// its texel belongs to by shuffle.  Most boxes are decided by their first <= 4 texels, a few have long border walks: pooled, both
/// kinds run on dense warps, whereas one walk per lane leaves ~3/4 of the lanes idle.  Entries of boxes that were decided in the
/// meantime are dropped when popped.  Children that do not fit on the stack are walked by their lane alone.
/// Requires level < 8 and mip dimensions < 4096 (buffers up to 8192 x 8192).  Must be called by all 32 lanes.
/// Every lane may look into another view's buffer: depthAll / mipAll are the batch's planes, `view` this lane's view index.
__device__ __forceinline__ bool range_visible_pool(const BufGeom& g, const int* __restrict__ depthAll, const int2* __restrict__ mipAll, int view, const BoxRect& r,
    bool need, uint32_t* wstack, int cap)
{
    const int* depth = depthAll + (size_t)view * g.w * g.h;
    const int2* mips = mipAll + (size_t)view * g.mipTexels;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t visMask = 0;
    const uint32_t needMask = __ballot_sync(0xffffffffu, need);
    if (!needMask)
        return false;
    int n = 0;
    __syncwarp();
    // top-level texels of every box (at most 2 x 2 unless the start level was clamped to the coarsest one)
    {
        const int top_lv = start_level(g, r);
        const int sh0 = top_lv + 1;
        const int rx0 = r.left >> sh0, ry0 = r.top >> sh0;
        const int nx = (r.right >> sh0) - rx0 + 1, ny = (r.bottom >> sh0) - ry0 + 1;
        const int kids = need ? nx * ny : 0;
        int off = kids;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, off, o);
            if (lane >= o)
                off += t;
        }
        const int total = __shfl_sync(0xffffffffu, off, 31);
        off -= kids;
        if (total <= cap)
        {
            if (kids)
            {
                int k = off;
                for (int cy = 0; cy < ny; ++cy)
                    for (int cx = 0; cx < nx; ++cx)
                        wstack[k++] = pack_pool(lane, top_lv, rx0 + cx, ry0 + cy);
            }
            n = total;
        }
        else
        {
            // cannot happen with the queue sizes in use (32 boxes x <= 4 x 4 texels below the heavy threshold); walk per lane
            const bool v = need && range_visible_thread(g, depth, mips, true, r);
            return v;
        }
    }
    __syncwarp();
    while (n > 0)
    {
        const int take = min(n, 32);
        n -= take;
        uint32_t e = 0;
        if (lane < take)
            e = wstack[n + lane];
        const int b = (int)(e >> 27);
        BoxRect rb;
        rb.left = __shfl_sync(0xffffffffu, r.left, b);
        rb.top = __shfl_sync(0xffffffffu, r.top, b);
        rb.right = __shfl_sync(0xffffffffu, r.right, b);
        rb.bottom = __shfl_sync(0xffffffffu, r.bottom, b);
        rb.z = __shfl_sync(0xffffffffu, r.z, b);
        const int vb = __shfl_sync(0xffffffffu, view, b);
        const int* depthB = depthAll + (size_t)vb * g.w * g.h;
        const int2* mipsB = mipAll + (size_t)vb * g.mipTexels;
        const int lv = (int)((e >> 24) & 7u), x = (int)((e >> 12) & 0xFFFu), y = (int)(e & 0xFFFu);
        int c = 0;
        if (lane < take && !((visMask >> b) & 1u))
        {
            c = classify_texel(g, mipsB, rb, lv, x, y);
            if (c == 2 && lv == 0)
                c = texel_pixels_visible(g, depthB, rb, x, y) ? 1 : 0;
        }
        visMask |= __reduce_or_sync(0xffffffffu, c == 1 ? (1u << b) : 0u);
        if ((visMask & needMask) == needMask)
            break;
        // children of the texels that must be refined (unless another texel of the same box has just proved it visible)
        int x0 = 0, x1 = -1, y0 = 0, y1 = -1, cl = 0;
        const bool refine = c == 2 && !((visMask >> b) & 1u);
        if (refine)
        {
            cl = lv - 1;
            const int sh = lv;
            x0 = max(2 * x, rb.left >> sh);
            x1 = min(min(2 * x + 1, rb.right >> sh), g.mw[cl] - 1);
            y0 = max(2 * y, rb.top >> sh);
            y1 = min(min(2 * y + 1, rb.bottom >> sh), g.mh[cl] - 1);
        }
        const int kids = refine ? (x1 - x0 + 1) * (y1 - y0 + 1) : 0;
        const uint32_t anyKids = __ballot_sync(0xffffffffu, kids != 0);
        if (!anyKids)
            continue;
        int off = kids;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, off, o);
            if (lane >= o)
                off += t;
        }
        off -= kids;
        // lanes whose children fit (in prefix order) push them; the first lane that does not fit and all later ones walk alone
        const bool fits = n + off + kids <= cap;
        const uint32_t spillMask = __ballot_sync(0xffffffffu, kids != 0 && !fits);
        __syncwarp();
        if (kids && fits)
        {
            int k = n + off;
            for (int cy = y0; cy <= y1; ++cy)
                for (int cx = x0; cx <= x1; ++cx)
                    wstack[k++] = pack_pool(b, cl, cx, cy);
        }
        if (!spillMask)
            n += __shfl_sync(0xffffffffu, off + kids, 31);
        else
        {
            // prefix order: every lane before the first spilling one fitted, so the kept entries are contiguous up to that lane's offset
            const int firstSpill = __ffs(spillMask) - 1;
            const int keptEnd = __shfl_sync(0xffffffffu, off, firstSpill);
            // later lanes that "fit" by the test above wrote beyond keptEnd only if an earlier lane spilled: treat them as spilled too
            bool lonely = false;
            if (kids && lane >= firstSpill)
            {
                BoxRect sub = rb;
                const int s = lv + 1;
                sub.left = max(rb.left, x << s);
                sub.top = max(rb.top, y << s);
                sub.right = min(rb.right, ((x + 1) << s) - 1);
                sub.bottom = min(rb.bottom, ((y + 1) << s) - 1);
                lonely = range_visible_thread(g, depthB, mipsB, true, sub);
            }
            visMask |= __reduce_or_sync(0xffffffffu, lonely ? (1u << b) : 0u);
            n += keptEnd;
        }
        __syncwarp();
    }
    (void)lt;
    return need && ((visMask >> lane) & 1u);
}

/// OcclusionBuffer::IsVisible, OB.cpp:336-456, evaluated by one thread.
__device__ __forceinline__ bool ob_is_visible(const BufGeom& g, const float* vp, const int* __restrict__ depth, const int2* __restrict__ mips,
    bool useMips, float mnx, float mny, float mnz, float mxx, float mxy, float mxz)
{
    BoxRect r;
    if (ob_project(g, vp, mnx, mny, mnz, mxx, mxy, mxz, r))
        return true;
    return range_visible_thread(g, depth, mips, useMips, r);
}

} // namespace u3d
