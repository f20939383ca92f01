// GPUCompute -- C++ host module that puts the reference's own class API on top of the C-ABI of libu3dgpu.so (include/u3d_gpu.h).
//
// Drop-in intent: the classes below keep the method names, argument meaning and return conventions of
//   OcclusionBuffer                    Source/Urho3D/Graphics/OcclusionBuffer.h:91-226
//   Octree::GetDrawables(OctreeQuery&) Source/Urho3D/Graphics/Octree.h:190, Octree.cpp:495-499
//   OctreeQuery / FrustumOctreeQuery   Source/Urho3D/Graphics/OctreeQuery.h:40-70, 140-158
//   OccludedFrustumOctreeQuery         Source/Urho3D/Graphics/View.cpp:116-158 (file-local there; promoted here)
//   ShadowCasterOctreeQuery            Source/Urho3D/Graphics/View.cpp:59-84
// so that View::DrawOccluders / View::GetDrawables / CheckVisibilityWork can be re-pointed with the edits shown in INTEGRATION.md.
// The module is header-only and engine-agnostic: matrices, boxes and frusta are passed as plain float arrays (Matrix3x4::Data(),
// Matrix4::Data(), BoundingBox min_/max_, Frustum::planes_); when compiled inside the engine (URHO3D_API defined) thin overloads taking
// Camera*, Matrix3x4, BoundingBox and Frustum are enabled.
//
// There is no CPU fallback: every compute call ends in a CUDA kernel launch or returns false / throws GpuError at construction.
#pragma once

#include "../../include/u3d_gpu.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace Urho3D
{
namespace GPUCompute
{

/// Raised only by constructors (no usable device, out of memory); the per-frame API reports through return values like the reference.
struct GpuError : std::runtime_error
{
    explicit GpuError(const std::string& what) : std::runtime_error(what) {}
};

inline void Check(int rc, const char* what)
{
    if (rc < 0)
        throw GpuError(std::string(what) + ": " + u3d_last_error());
}

/// One CUDA device.  Create one per GPU (one process per GPU in the multi-GPU layout).
class GpuContext
{
public:
    explicit GpuContext(int device = 0)
    {
        if (u3d_abi_version() != U3D_ABI_VERSION)
            throw std::runtime_error("libu3dgpu.so was built from another version of u3d_gpu.h: rebuild it");
        Check(u3d_ctx_create(device, &ctx_), "u3d_ctx_create");
    }
    ~GpuContext() { u3d_ctx_destroy(ctx_); }
    GpuContext(const GpuContext&) = delete;
    GpuContext& operator=(const GpuContext&) = delete;
    u3d_ctx* Handle() const { return ctx_; }

private:
    u3d_ctx* ctx_{};
};

/// CullMode, GraphicsDefs.h
enum GpuCullMode { GPU_CULL_NONE = U3D_CULL_NONE, GPU_CULL_CCW = U3D_CULL_CCW, GPU_CULL_CW = U3D_CULL_CW };

/// Occluder geometry kept resident in HBM (extension: avoids re-uploading vertex/index buffers every frame).
class GpuMesh
{
public:
    GpuMesh(GpuContext& ctx, const void* vertexData, unsigned vertexSize, unsigned vertexCount, const void* indexData, unsigned indexSize,
        unsigned indexCount)
    {
        Check(u3d_mesh_create(ctx.Handle(), vertexData, vertexSize, vertexCount, indexData, indexSize, indexCount, &mesh_), "u3d_mesh_create");
    }
    ~GpuMesh() { u3d_mesh_destroy(mesh_); }
    GpuMesh(const GpuMesh&) = delete;
    GpuMesh& operator=(const GpuMesh&) = delete;
    u3d_mesh* Handle() const { return mesh_; }

private:
    u3d_mesh* mesh_{};
};

/// Drop-in for class OcclusionBuffer (OcclusionBuffer.h:91-226).  Same call sequence as View::DrawOccluders (View.cpp:2231-2274).
class GpuOcclusionBuffer
{
public:
    explicit GpuOcclusionBuffer(GpuContext& ctx) { Check(u3d_ob_create(ctx.Handle(), &ob_), "u3d_ob_create"); }
    ~GpuOcclusionBuffer() { u3d_ob_destroy(ob_); }
    GpuOcclusionBuffer(const GpuOcclusionBuffer&) = delete;
    GpuOcclusionBuffer& operator=(const GpuOcclusionBuffer&) = delete;

    /// Set occlusion buffer size (OcclusionBuffer.cpp:61-113): odd heights are rounded up, non power-of-two widths are refused.
    bool SetSize(int width, int height, bool threaded)
    {
        const bool ok = u3d_ob_set_size(ob_, width, height, threaded ? 1 : 0) > 0;
        if (ok)
            host_.assign((size_t)GetWidth() * GetHeight(), 0);
        return ok;
    }
    /// Set camera view to render from (OcclusionBuffer.cpp:115-127), from the values the Camera getters return.
    void SetView(const float* view3x4, const float* projection4x4, float nearClip, float farClip, bool reverseCulling)
    {
        std::memcpy(view_, view3x4, sizeof(view_));
        std::memcpy(projection_, projection4x4, sizeof(projection_));
        u3d_ob_set_view(ob_, view3x4, projection4x4, nearClip, farClip, reverseCulling ? 1 : 0);
    }
    void SetMaxTriangles(unsigned triangles) { u3d_ob_set_max_triangles(ob_, triangles); }
    void SetCullMode(GpuCullMode mode) { u3d_ob_set_cull_mode(ob_, (int)mode); }
    void Reset() { u3d_ob_reset(ob_); }
    void Clear() { u3d_ob_clear(ob_); }
    /// Submit non-indexed geometry.  Return true if the triangle budget was not exceeded (the batch is queued either way, as in the reference).
    bool AddTriangles(const float* model3x4, const void* vertexData, unsigned vertexSize, unsigned vertexStart, unsigned vertexCount)
    {
        return u3d_ob_add_triangles(ob_, model3x4, vertexData, vertexSize, vertexStart, vertexCount) > 0;
    }
    /// Submit indexed geometry.
    bool AddTriangles(const float* model3x4, const void* vertexData, unsigned vertexSize, const void* indexData, unsigned indexSize,
        unsigned indexStart, unsigned indexCount)
    {
        return u3d_ob_add_triangles_indexed(ob_, model3x4, vertexData, vertexSize, indexData, indexSize, indexStart, indexCount) > 0;
    }
    /// Submit geometry that already lives on the GPU.
    bool AddTriangles(const float* model3x4, const GpuMesh& mesh, unsigned start, unsigned count)
    {
        return u3d_ob_add_triangles_mesh(ob_, model3x4, mesh.Handle(), start, count) > 0;
    }
    /// Draw submitted batches.
    void DrawTriangles() { u3d_ob_draw_triangles(ob_); }
    /// Build reduced size mip levels.
    void BuildDepthHierarchy() { u3d_ob_build_depth_hierarchy(ob_); }
    void ResetUseTimer() { useTimer_ = std::chrono::steady_clock::now(); }
    /// Return highest level depth values (a host copy refreshed by this call; the reference returns its own storage).
    int* GetBuffer()
    {
        if (host_.empty() || u3d_ob_get_buffer(ob_, host_.data()) < 0)
            return nullptr;
        return host_.data();
    }
    const float* GetView() const { return view_; }
    const float* GetProjection() const { return projection_; }
    int GetWidth() const { return u3d_ob_get_width(ob_); }
    int GetHeight() const { return u3d_ob_get_height(ob_); }
    unsigned GetNumTriangles() const { return u3d_ob_get_num_triangles(ob_); }
    unsigned GetMaxTriangles() const { return u3d_ob_get_max_triangles(ob_); }
    GpuCullMode GetCullMode() const { return (GpuCullMode)u3d_ob_get_cull_mode(ob_); }
    bool IsThreaded() const { return u3d_ob_is_threaded(ob_) != 0; }
    /// Test a bounding box (min.xyz, max.xyz) for visibility.  One launch + synchronisation per call: prefer the batched forms.
    bool IsVisible(const float* boxMinMax6) const { return u3d_ob_is_visible(ob_, boxMinMax6) != 0; }
    /// Many boxes at once (extension).
    bool IsVisible(const float* boxesMinMax6, unsigned count, unsigned char* out) const { return u3d_ob_is_visible_many(ob_, boxesMinMax6, count, out) >= 0; }
    unsigned GetUseTimer() const
    {
        return (unsigned)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - useTimer_).count();
    }
    /// Mip level `level` as (min,max) pairs, for tests.
    bool GetMip(int level, std::vector<int>& out, int& width, int& height)
    {
        if (u3d_ob_mip_size(ob_, level, &width, &height) < 0)
            return false;
        out.resize((size_t)width * height * 2);
        return u3d_ob_get_mip(ob_, level, out.data()) >= 0;
    }
    int GetNumLevels() const { return u3d_ob_num_levels(ob_); }
    u3d_ob* Handle() const { return ob_; }

#ifdef URHO3D_API
    // In-engine overloads with the reference's exact signatures.
    void SetView(Camera* camera)
    {
        if (!camera)
            return;
        SetView(camera->GetView().Data(), camera->GetProjection().Data(), camera->GetNearClip(), camera->GetFarClip(), camera->GetReverseCulling());
    }
    void SetCullMode(CullMode mode) { SetCullMode((GpuCullMode)mode); }
    bool AddTriangles(const Matrix3x4& model, const void* vertexData, unsigned vertexSize, unsigned vertexStart, unsigned vertexCount)
    {
        return AddTriangles(model.Data(), vertexData, vertexSize, vertexStart, vertexCount);
    }
    bool AddTriangles(const Matrix3x4& model, const void* vertexData, unsigned vertexSize, const void* indexData, unsigned indexSize,
        unsigned indexStart, unsigned indexCount)
    {
        return AddTriangles(model.Data(), vertexData, vertexSize, indexData, indexSize, indexStart, indexCount);
    }
    bool IsVisible(const BoundingBox& worldSpaceBox) const
    {
        const float b[6] = {worldSpaceBox.min_.x_, worldSpaceBox.min_.y_, worldSpaceBox.min_.z_, worldSpaceBox.max_.x_, worldSpaceBox.max_.y_,
            worldSpaceBox.max_.z_};
        return IsVisible(b);
    }
#endif

private:
    u3d_ob* ob_{};
    std::vector<int> host_;
    float view_[12]{};
    float projection_[16]{};
    std::chrono::steady_clock::time_point useTimer_{std::chrono::steady_clock::now()};
};

/// Base of the GPU-evaluable queries (OctreeQuery.h:40-70).  result_ receives drawable ids in the reference's result_ order.
struct GpuOctreeQuery
{
    GpuOctreeQuery(std::vector<unsigned>& result, unsigned char drawableFlags, unsigned viewMask) :
        result_(result), drawableFlags_(drawableFlags), viewMask_(viewMask)
    {
    }
    virtual ~GpuOctreeQuery() = default;
    virtual int Mode() const = 0;
    virtual const GpuOcclusionBuffer* Buffer() const { return nullptr; }

    std::vector<unsigned>& result_;
    unsigned char drawableFlags_;
    unsigned viewMask_;
    float planes_[24]{}; // Frustum::planes_: 6 x (normal.xyz, d)
};

/// FrustumOctreeQuery (OctreeQuery.h:140-158).
struct GpuFrustumOctreeQuery : GpuOctreeQuery
{
    GpuFrustumOctreeQuery(std::vector<unsigned>& result, const float* planes6x4, unsigned char drawableFlags = 0xFF, unsigned viewMask = 0xFFFFFFFFu) :
        GpuOctreeQuery(result, drawableFlags, viewMask)
    {
        std::memcpy(planes_, planes6x4, sizeof(planes_));
    }
    int Mode() const override { return U3D_QUERY_FRUSTUM; }
};

/// OccludedFrustumOctreeQuery (View.cpp:116-158).
struct GpuOccludedFrustumOctreeQuery : GpuFrustumOctreeQuery
{
    GpuOccludedFrustumOctreeQuery(std::vector<unsigned>& result, const float* planes6x4, GpuOcclusionBuffer* buffer, unsigned char drawableFlags = 0xFF,
        unsigned viewMask = 0xFFFFFFFFu) :
        GpuFrustumOctreeQuery(result, planes6x4, drawableFlags, viewMask), buffer_(buffer)
    {
    }
    int Mode() const override { return U3D_QUERY_OCCLUDED; }
    const GpuOcclusionBuffer* Buffer() const override { return buffer_; }
    GpuOcclusionBuffer* buffer_;
};

/// ShadowCasterOctreeQuery (View.cpp:59-84).
struct GpuShadowCasterOctreeQuery : GpuFrustumOctreeQuery
{
    using GpuFrustumOctreeQuery::GpuFrustumOctreeQuery;
    int Mode() const override { return U3D_QUERY_SHADOWCASTER; }
};

/// ZoneOccluderOctreeQuery (View.cpp:87-113): zones and occluders inside a frustum; drawableFlags_ is not consulted (the reference compares the
/// drawable's flags with DRAWABLE_ZONE / DRAWABLE_GEOMETRY directly).  IsOccluder() comes from GpuOctree::SetOccluder.
struct GpuZoneOccluderOctreeQuery : GpuFrustumOctreeQuery
{
    using GpuFrustumOctreeQuery::GpuFrustumOctreeQuery;
    int Mode() const override { return U3D_QUERY_ZONE_OCCLUDER; }
};

/// RayQueryResult (OctreeQuery.h:189-224) as Drawable::ProcessRayQuery fills it at RAY_AABB level (Drawable.cpp:118-132).
using GpuRayQueryResult = u3d_ray_hit;

/// RayOctreeQuery (OctreeQuery.h:226-260) with level_ fixed to RAY_AABB (per-triangle levels need the drawables' geometry: out of scope).
/// origin / direction are Ray::origin_ / Ray::direction_ (the caller normalises, as Ray's constructor does).
struct GpuRayOctreeQuery
{
    GpuRayOctreeQuery(std::vector<GpuRayQueryResult>& result, const float* origin3, const float* direction3, float maxDistance = INFINITY,
        unsigned char drawableFlags = 0xFF, unsigned viewMask = 0xFFFFFFFFu) :
        result_(result), drawableFlags_(drawableFlags), viewMask_(viewMask), maxDistance_(maxDistance)
    {
        std::memcpy(origin_, origin3, sizeof(origin_));
        std::memcpy(direction_, direction3, sizeof(direction_));
    }
    std::vector<GpuRayQueryResult>& result_;
    float origin_[3], direction_[3];
    unsigned char drawableFlags_;
    unsigned viewMask_;
    float maxDistance_;
};

/// Mirror of the reference Octree for the query path: host structure kept by the reference's insertion rules, flattened into a
/// GPU-resident SoA on Sync().  Insert / move / remove stay on the host exactly as SURVEY 2a prescribes.
class GpuOctree
{
public:
    GpuOctree(GpuContext& ctx, const float* boxMin3, const float* boxMax3, unsigned numLevels) : ctx_(ctx)
    {
        Check(u3d_octree_create(boxMin3, boxMax3, numLevels, &tree_), "u3d_octree_create");
    }
    ~GpuOctree()
    {
        u3d_scene_destroy(scene_);
        u3d_octree_destroy(tree_);
    }
    GpuOctree(const GpuOctree&) = delete;
    GpuOctree& operator=(const GpuOctree&) = delete;

    /// Octree::InsertDrawable (Octree.cpp:134-166).  Returns the drawable id used in query results.
    unsigned InsertDrawable(const float* worldBoxMinMax6, unsigned char drawableFlags, unsigned viewMask, bool occludee, bool castShadows)
    {
        const uint8_t oc = occludee, cs = castShadows;
        uint32_t id = 0;
        Check(u3d_octree_add(tree_, 1, worldBoxMinMax6, &drawableFlags, &viewMask, &oc, &cs, &id), "u3d_octree_add");
        dirty_ = true;
        return id;
    }
    /// Many drawables at once (arrays may be null for the defaults).
    unsigned InsertDrawables(unsigned count, const float* worldBoxes, const unsigned char* drawableFlags, const unsigned* viewMasks,
        const unsigned char* occludee, const unsigned char* castShadows)
    {
        uint32_t id = 0;
        Check(u3d_octree_add(tree_, count, worldBoxes, drawableFlags, viewMasks, occludee, castShadows, &id), "u3d_octree_add");
        dirty_ = true;
        return id;
    }
    /// Drawable::SetOccluder (used by GpuZoneOccluderOctreeQuery).  Takes effect at the next Sync (the scene is re-flattened).
    void SetOccluder(unsigned id, bool enable)
    {
        const uint8_t f = enable;
        Check(u3d_octree_set_occluder_flags(tree_, id, 1, &f), "u3d_octree_set_occluder_flags");
        dirty_ = true;
    }
    /// A drawable moved or was resized: the reinsertion rule of Octree::Update (Octree.cpp:440-472).
    void UpdateDrawable(unsigned id, const float* worldBoxMinMax6)
    {
        if (u3d_octree_update(tree_, id, worldBoxMinMax6) == 1 && !dirty_)
        {
            // stayed in its octant: only its box has to be refreshed in HBM (several moves of one drawable before the next Sync: the
            // last box wins -- u3d_scene_update_boxes folds duplicate ids in call order)
            pendingIds_.push_back(id);
            pendingBoxes_.insert(pendingBoxes_.end(), worldBoxMinMax6, worldBoxMinMax6 + 6);
        }
        else
            dirty_ = true;
    }
    void RemoveDrawable(unsigned id)
    {
        u3d_octree_remove(tree_, id);
        dirty_ = true;
    }
    /// Re-flatten and upload if anything changed since the last query (the analogue of Octree::Update being called once per frame,
    /// Renderer.cpp:1475-1524).
    void Sync()
    {
        if (!dirty_ && scene_)
        {
            if (!pendingIds_.empty())
            {
                Check(u3d_scene_update_boxes(scene_, (uint32_t)pendingIds_.size(), pendingIds_.data(), pendingBoxes_.data(), nullptr), "u3d_scene_update_boxes");
                pendingIds_.clear();
                pendingBoxes_.clear();
            }
            return;
        }
        pendingIds_.clear();
        pendingBoxes_.clear();
        u3d_scene_destroy(scene_);
        scene_ = nullptr;
        Check(u3d_scene_create(ctx_.Handle(), tree_, &scene_), "u3d_scene_create");
        u3d_scene_counts(scene_, &numOctants_, &numDrawables_);
        dirty_ = false;
        ++generation_; // every holder of the old u3d_scene* (GpuViewBatch) must rebind: that object is gone
    }
    /// Octree::GetDrawables(OctreeQuery&) (Octree.cpp:495-499): clears query.result_, then fills it in pre-order DFS order.
    void GetDrawables(GpuOctreeQuery& query)
    {
        Sync();
        query.result_.clear();
        query.result_.resize(numDrawables_ ? numDrawables_ : 1);
        uint32_t queryCount = 0, count = 0;
        const GpuOcclusionBuffer* buffer = query.Buffer();
        const int rc = u3d_octree_query(scene_, buffer ? buffer->Handle() : nullptr, query.planes_, query.drawableFlags_, query.viewMask_, query.Mode(),
            U3D_STAGE_QUERY, query.result_.data(), (uint32_t)query.result_.size(), &queryCount, &count);
        query.result_.resize(rc < 0 ? 0 : count);
    }
    /// GetDrawables followed by CheckVisibilityWork's occlusion predicate (View.cpp:177) in one pass: what View::GetDrawables needs.
    void GetVisibleDrawables(GpuOctreeQuery& query, unsigned* queryResultSize = nullptr)
    {
        Sync();
        query.result_.clear();
        query.result_.resize(numDrawables_ ? numDrawables_ : 1);
        uint32_t queryCount = 0, count = 0;
        const GpuOcclusionBuffer* buffer = query.Buffer();
        const int rc = u3d_octree_query(scene_, buffer ? buffer->Handle() : nullptr, query.planes_, query.drawableFlags_, query.viewMask_, query.Mode(),
            U3D_STAGE_VISIBLE, query.result_.data(), (uint32_t)query.result_.size(), &queryCount, &count);
        query.result_.resize(rc < 0 ? 0 : count);
        if (queryResultSize)
            *queryResultSize = queryCount;
    }
    /// Octree::Raycast(RayOctreeQuery&) (Octree.cpp:501-508): every hit, sorted by distance with the reference's Sort procedure.
    void Raycast(GpuRayOctreeQuery& query) { RaycastImpl(query, false); }
    /// Octree::RaycastSingle(RayOctreeQuery&) (Octree.cpp:510-548): the closest hit only.
    void RaycastSingle(GpuRayOctreeQuery& query) { RaycastImpl(query, true); }
    unsigned GetNumOctants() const { return numOctants_; }
    unsigned GetNumDrawables() const { return numDrawables_; }
    u3d_scene* Scene()
    {
        Sync();
        return scene_;
    }
    u3d_octree* Tree() const { return tree_; }
    /// Bumped whenever Sync() replaced the device scene (structural change of the octree).
    unsigned long long Generation() const { return generation_; }

private:
    void RaycastImpl(GpuRayOctreeQuery& query, bool single)
    {
        Sync();
        const uint32_t cap = numDrawables_ ? numDrawables_ : 1;
        query.result_.clear();
        query.result_.resize(cap);
        const float ray[7] = {query.origin_[0], query.origin_[1], query.origin_[2], query.direction_[0], query.direction_[1], query.direction_[2],
            query.maxDistance_};
        uint32_t count = 0;
        const int rc = u3d_scene_raycast(scene_, 1, ray, query.drawableFlags_, query.viewMask_, single ? 1 : 0, cap, &count, query.result_.data(), nullptr);
        query.result_.resize(rc < 0 ? 0 : count);
    }

    GpuContext& ctx_;
    u3d_octree* tree_{};
    u3d_scene* scene_{};
    uint32_t numOctants_{}, numDrawables_{};
    bool dirty_{true};
    unsigned long long generation_{0};
    std::vector<uint32_t> pendingIds_;
    std::vector<float> pendingBoxes_;
};

/// The batched many-view entry point (new): V views per call through raster + hierarchy + octree query + visibility.
class GpuViewBatch
{
public:
    GpuViewBatch(GpuContext& ctx, GpuOctree& octree, u3d_occluders* occluders, int width, int height, unsigned maxViews, unsigned outCapacity,
        unsigned long long triPool = 0)
        : octree_(octree)
    {
        Check(u3d_batch_create(ctx.Handle(), octree.Scene(), occluders, width, height, maxViews, outCapacity, triPool, &batch_), "u3d_batch_create");
        generation_ = octree.Generation();
        counts_.resize((size_t)maxViews * 2);
        offsets_.resize((size_t)maxViews + 1);
    }
    ~GpuViewBatch() { u3d_batch_destroy(batch_); }
    GpuViewBatch(const GpuViewBatch&) = delete;
    GpuViewBatch& operator=(const GpuViewBatch&) = delete;

    /// Visible drawable ids of every view, packed; Offsets()[v]..Offsets()[v+1] delimit view v.  Returns false on failure (u3d_last_error()).
    bool CullViews(const u3d_view* views, unsigned count, std::vector<unsigned>& ids, bool visibilityPredicate = true)
    {
        const int stage = visibilityPredicate ? U3D_STAGE_VISIBLE : U3D_STAGE_QUERY;
        if (!Rebind())
            return false;
        if (u3d_batch_set_views(batch_, views, count, nullptr) < 0)
            return false;
        if (policy_ && camera3_.size() < 3 * (size_t)count)
            return false;
        if (u3d_batch_set_occluder_policy(batch_, policy_ ? (threaded_ ? 1 : 2) : 0, sizeThreshold_, maxTriangles_, policy_ ? camera3_.data() : nullptr, nullptr) < 0)
            return false;
        if (u3d_batch_run(batch_, stage, nullptr) < 0 || u3d_batch_read_counts(batch_, counts_.data(), nullptr) < 0)
            return false;
        unsigned long long total = 0;
        for (unsigned v = 0; v < count; ++v)
            total += counts_[2 * v + 1];
        ids.resize(total ? total : 1);
        if (u3d_batch_read_packed(batch_, offsets_.data(), ids.data(), ids.size(), nullptr) < 0)
            return false;
        ids.resize(total);
        return true;
    }
    /// Occluder selection of View::UpdateOccluders + the budget of View::DrawOccluders' threaded branch (View.cpp:2171-2229, 2258-2270) for the
    /// views given to the next CullViews call.  camera3 = per view (Camera::GetHalfViewSize(), GetOrthoSize(), GetFarClip());
    /// sizeThreshold = Renderer::GetOccluderSizeThreshold(), maxTriangles = Renderer::GetMaxOccluderTriangles().
    void SetOccluderPolicy(float sizeThreshold, unsigned maxTriangles, const float* camera3, unsigned count)
    {
        policy_ = true;
        sizeThreshold_ = sizeThreshold;
        maxTriangles_ = maxTriangles;
        camera3_.assign(camera3, camera3 + 3 * (size_t)count);
    }
    void ClearOccluderPolicy() { policy_ = false; }
    /// threaded = false selects View::DrawOccluders' non-threaded branch (the engine default, Renderer::SetThreadedOcclusion(false)).
    void SetThreadedOcclusion(bool threaded) { threaded_ = threaded; }
    /// View::ProcessShadowCasters over the lists of the last CullViews call (one u3d_shadow_split per view); see include/u3d_gpu.h.
    bool ProcessShadowCasters(const u3d_shadow_split* splits, const float* mainView12, const float* mainCameraPos3, bool mainOrthographic,
        const unsigned char* inView, unsigned numInView, unsigned count, std::vector<unsigned>& ids)
    {
        if (!Rebind())
            return false;
        if (u3d_batch_process_shadow_casters(batch_, splits, mainView12, mainCameraPos3, mainOrthographic ? 1 : 0, inView, numInView, nullptr) < 0 ||
            u3d_batch_read_counts(batch_, counts_.data(), nullptr) < 0)
            return false;
        unsigned long long total = 0;
        for (unsigned v = 0; v < count; ++v)
            total += counts_[2 * v + 1];
        ids.resize(total ? total : 1);
        if (u3d_batch_read_packed(batch_, offsets_.data(), ids.data(), ids.size(), nullptr) < 0)
            return false;
        ids.resize(total);
        return true;
    }
    const std::vector<unsigned>& Offsets() const { return offsets_; }
    const std::vector<unsigned>& Counts() const { return counts_; }
    u3d_batch* Handle() const { return batch_; }

private:
    /// GpuOctree::Sync() replaces its u3d_scene after every structural change (insert / remove / a move across octants); the batch is then
    /// re-pointed at the new one before it runs, instead of launching kernels on the freed scene.
    bool Rebind()
    {
        u3d_scene* scene = octree_.Scene(); // syncs
        if (generation_ != octree_.Generation())
        {
            if (u3d_batch_set_scene(batch_, scene) < 0)
                return false;
            generation_ = octree_.Generation();
        }
        return true;
    }

    GpuOctree& octree_;
    unsigned long long generation_{};
    u3d_batch* batch_{};
    std::vector<unsigned> counts_, offsets_;
    bool policy_{};
    bool threaded_{true};
    float sizeThreshold_{};
    unsigned maxTriangles_{};
    std::vector<float> camera3_;
};

/// Result exchange between the GPUs of one box: every rank's per-view counts and visible sets land in every rank's receive block,
/// stored there by the producing rank's own kernel through NVLink peer memory (u3d_gather_*).  One object per batch slot and rank.
class GpuResultExchange
{
public:
    GpuResultExchange(GpuContext& ctx, int world, int rank, unsigned viewsPerRank, unsigned long long idsCapacityPerRank)
    {
        Check(u3d_gather_create(ctx.Handle(), world, rank, viewsPerRank, idsCapacityPerRank, &gather_), "u3d_gather_create");
    }
    ~GpuResultExchange() { u3d_gather_destroy(gather_); }
    GpuResultExchange(const GpuResultExchange&) = delete;
    GpuResultExchange& operator=(const GpuResultExchange&) = delete;

    /// 64-byte handle of this rank's block for the other processes; Connect takes all ranks' handles in rank order.
    void Export(void* handle64) { Check(u3d_gather_export(gather_, handle64), "u3d_gather_export"); }
    void Connect(const void* handlesOfAllRanks) { Check(u3d_gather_connect(gather_, handlesOfAllRanks), "u3d_gather_connect"); }
    /// Ranks living in one process (one host thread per GPU).
    void ConnectLocal(GpuResultExchange* const* allRanks, int world)
    {
        std::vector<u3d_gather*> h((size_t)world);
        for (int r = 0; r < world; ++r)
            h[(size_t)r] = allRanks[r]->gather_;
        Check(u3d_gather_connect_local(gather_, h.data()), "u3d_gather_connect_local");
    }
    /// After the batch's run, on its stream: pack + store into every rank's block + signal.
    bool Push(GpuViewBatch& batch, void* stream = nullptr) { return u3d_batch_push_results(batch.Handle(), gather_, stream) >= 0; }
    /// Device-side wait for every rank's results / acknowledgement that they were consumed.
    bool Wait(void* stream = nullptr) { return u3d_gather_wait(gather_, stream) >= 0; }
    bool Release(void* stream = nullptr) { return u3d_gather_release(gather_, stream) >= 0; }
    u3d_gather* Handle() const { return gather_; }

private:
    u3d_gather* gather_{};
};

/// Which rank (GPU) culls which view: equal view counts, every per-view cost column (previous frame's submitted triangles, query sizes,
/// visible counts ...) shared out evenly (u3d_balance_views).  costs is numViews x numCosts, row-major.  Returns the owner rank per view.
inline std::vector<unsigned> BalanceViews(const std::vector<double>& costs, unsigned numViews, unsigned numCosts, unsigned world)
{
    std::vector<unsigned> owner(numViews);
    if (costs.size() < (size_t)numViews * numCosts)
        throw GpuError("BalanceViews: costs array too small");
    Check(u3d_balance_views(numViews, numCosts, costs.data(), world, owner.data()), "u3d_balance_views");
    return owner;
}

#ifdef URHO3D_API
// ---------------------------------------------------------------------------------------------------------------------------------
// In-engine helpers (compiled only inside the engine, i.e. after the reference's own headers: Camera.h, Frustum.h, BoundingBox.h).
// They turn the engine objects the call sites of View.cpp hold into the plain arrays the classes above take, so that the edited call
// sites of INTEGRATION.md read like the originals.  Compiled and run against the reference's headers by tests/test_engine_types.py.
// ---------------------------------------------------------------------------------------------------------------------------------

/// Frustum::planes_ (Frustum.h:37-45, 196) as 6 x (normal.xyz, d) in the reference's plane order.
inline void PlanesOf(const Frustum& frustum, float* planes6x4)
{
    for (unsigned i = 0; i < NUM_FRUSTUM_PLANES; ++i)
    {
        const Plane& p = frustum.planes_[i];
        planes6x4[4 * i + 0] = p.normal_.x_;
        planes6x4[4 * i + 1] = p.normal_.y_;
        planes6x4[4 * i + 2] = p.normal_.z_;
        planes6x4[4 * i + 3] = p.d_;
    }
}

/// BoundingBox as (min.xyz, max.xyz).
inline void BoxOf(const BoundingBox& box, float* minMax6)
{
    minMax6[0] = box.min_.x_; minMax6[1] = box.min_.y_; minMax6[2] = box.min_.z_;
    minMax6[3] = box.max_.x_; minMax6[4] = box.max_.y_; minMax6[5] = box.max_.z_;
}

/// One view of the batched entry point from the Camera a View culls with: what OcclusionBuffer::SetView (OcclusionBuffer.cpp:115-127) and
/// the octree query of View::GetDrawables (View.cpp:868-878) read from it.
inline u3d_view MakeView(Camera* camera, unsigned char drawableFlags, unsigned viewMask, int queryMode = U3D_QUERY_OCCLUDED, bool useOcclusion = true)
{
    u3d_view v;
    std::memset(&v, 0, sizeof(v));
    std::memcpy(v.view, camera->GetView().Data(), sizeof(v.view));
    const Matrix4 projection = camera->GetProjection();
    std::memcpy(v.projection, projection.Data(), sizeof(v.projection));
    PlanesOf(camera->GetFrustum(), v.planes);
    v.near_clip = camera->GetNearClip();
    v.far_clip = camera->GetFarClip();
    v.drawable_flags = drawableFlags;
    v.view_mask = viewMask;
    v.reverse_culling = camera->GetReverseCulling() ? 1 : 0;
    v.query_mode = queryMode;
    v.use_occlusion = useOcclusion ? 1 : 0;
    v.orthographic = camera->IsOrthographic() ? 1 : 0;
    const Vector3 position = camera->GetNode() ? camera->GetNode()->GetWorldPosition() : Vector3::ZERO;
    v.camera_position[0] = position.x_; v.camera_position[1] = position.y_; v.camera_position[2] = position.z_;
    return v;
}

/// The reference's constructor signatures: FrustumOctreeQuery(result, frustum, drawableFlags, viewMask) (OctreeQuery.h:144-149),
/// OccludedFrustumOctreeQuery(result, frustum, buffer, drawableFlags, viewMask) (View.cpp:120-125), ShadowCasterOctreeQuery (View.cpp:63-67),
/// ZoneOccluderOctreeQuery (View.cpp:91-95).  result receives drawable ids of the mirror instead of Drawable pointers.
struct FrustumPlanes
{
    explicit FrustumPlanes(const Frustum& frustum) { PlanesOf(frustum, data); }
    float data[24];
};
inline GpuFrustumOctreeQuery MakeFrustumQuery(std::vector<unsigned>& result, const Frustum& frustum, unsigned char drawableFlags = DRAWABLE_ANY,
    unsigned viewMask = DEFAULT_VIEWMASK)
{
    return GpuFrustumOctreeQuery(result, FrustumPlanes(frustum).data, drawableFlags, viewMask);
}
inline GpuOccludedFrustumOctreeQuery MakeOccludedFrustumQuery(std::vector<unsigned>& result, const Frustum& frustum, GpuOcclusionBuffer* buffer,
    unsigned char drawableFlags = DRAWABLE_ANY, unsigned viewMask = DEFAULT_VIEWMASK)
{
    return GpuOccludedFrustumOctreeQuery(result, FrustumPlanes(frustum).data, buffer, drawableFlags, viewMask);
}
inline GpuShadowCasterOctreeQuery MakeShadowCasterQuery(std::vector<unsigned>& result, const Frustum& frustum, unsigned char drawableFlags = DRAWABLE_ANY,
    unsigned viewMask = DEFAULT_VIEWMASK)
{
    return GpuShadowCasterOctreeQuery(result, FrustumPlanes(frustum).data, drawableFlags, viewMask);
}
inline GpuZoneOccluderOctreeQuery MakeZoneOccluderQuery(std::vector<unsigned>& result, const Frustum& frustum, unsigned char drawableFlags = DRAWABLE_ANY,
    unsigned viewMask = DEFAULT_VIEWMASK)
{
    return GpuZoneOccluderOctreeQuery(result, FrustumPlanes(frustum).data, drawableFlags, viewMask);
}

/// Octree::InsertDrawable for a real Drawable: its world box, flags, view mask, occludee / cast-shadows / occluder state.
inline unsigned InsertDrawable(GpuOctree& octree, Drawable* drawable)
{
    float box[6];
    BoxOf(drawable->GetWorldBoundingBox(), box);
    const unsigned id = octree.InsertDrawable(box, drawable->GetDrawableFlags(), drawable->GetViewMask(), drawable->IsOccludee(), drawable->GetCastShadows());
    if (drawable->IsOccluder())
        octree.SetOccluder(id, true);
    return id;
}
#endif // URHO3D_API

} // namespace GPUCompute
} // namespace Urho3D
