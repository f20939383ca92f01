// TEST INFRASTRUCTURE ONLY (oracle/): a C-ABI harness that drives the UNMODIFIED reference
// classes compiled straight out of /root/reference (see oracle/build_ref.sh).  Nothing in the
// product path may load the library this file builds into (oracle/_ref/libu3dref.so).
//
// It exposes, for ctypes:
//   * the reference Octree (Octree.cpp:134-190 insertion, :229-255 DFS query) over drawables whose
//     world AABBs are supplied directly (a test Drawable subclass),
//   * the reference Camera (Camera.cpp:284-308 frustum, :585-595 view, :648-691 projection),
//   * the reference OcclusionBuffer (OcclusionBuffer.h:101-158),
//   * the reference's OWN text of ShadowCasterOctreeQuery / ZoneOccluderOctreeQuery / OccludedFrustumOctreeQuery (file-local classes of
//     View.cpp:58-157), extracted by the build recipe into a generated include and compiled here,
//   * the reference's OWN text of View::UpdateOccluders (View.cpp:2171-2229) and View::IsShadowCasterVisible (View.cpp:2455-2489), extracted
//     the same way and compiled as members of a stand-in class that holds the two View members they touch (renderer_, frame_), and of the
//     free function CheckVisibilityWork (View.cpp:160-227) with that stand-in as the View it works for, View::DrawOccluders (View.cpp:2231-2275:
//     budget, IsVisible gating, activeOccluders_) over Drawables whose DrawOcclusion submits the test mesh, and View::ProcessShadowCasters
//     (View.cpp:2379-2453) with stand-ins for the Light and the LightQueryResult it reads,
//   * the CheckVisibilityWork occlusion predicate (View.cpp:177), split over the WorkQueue the way
//     View.cpp:897-920 does when worker threads exist.
//
// The private-member peeking (#define protected public) is needed only to dump the octant tree.

#include <cstdint>
#include <cstring>
#include <chrono>
#include <vector>

#define protected public
#define private public
#include <Urho3D/Graphics/Octree.h>
#include <Urho3D/Graphics/OcclusionBuffer.h>
#undef protected
#undef private

#include <Urho3D/Core/Context.h>
#include <Urho3D/Core/WorkQueue.h>
#include <Urho3D/Graphics/Camera.h>
#include <Urho3D/Graphics/Drawable.h>
#include <Urho3D/Graphics/OctreeQuery.h>
#include <Urho3D/Scene/Node.h>
#include <Urho3D/Scene/Scene.h>
#include <Urho3D/Math/Sphere.h>
#include <Urho3D/Math/Ray.h>
#include <Urho3D/Container/Sort.h>
#include <Urho3D/Graphics/Light.h>
#include <Urho3D/Graphics/Zone.h>
#include <cstdlib>

// The reference's own text of View.cpp:58-157 (ShadowCasterOctreeQuery, ZoneOccluderOctreeQuery, OccludedFrustumOctreeQuery), extracted at
// build time by oracle/build_ref.sh into oracle/_ref/view_queries.inc: the classes are file-local in View.cpp, which cannot be compiled here.
namespace Urho3D
{
#include "view_queries.inc"
}

// The reference's own text of View::UpdateOccluders (View.cpp:2171-2229) and View::IsShadowCasterVisible (View.cpp:2455-2489), extracted the
// same way and compiled as members of a stand-in for View that holds the two members they touch: renderer_ (asked for
// GetOccluderSizeThreshold()) and frame_.  Everything else they use -- Drawable, Camera, BoundingBox, Frustum, Ray, Sort, CompareDrawables -- is
// the reference's real code.
namespace Urho3D
{
struct RefRendererShim
{
    float occluderSizeThreshold_{};
    float GetOccluderSizeThreshold() const { return occluderSizeThreshold_; }
};

/// Data holder with the fields of View.h's PerThreadSceneResult that CheckVisibilityWork fills (View.h:97-109).
struct PerThreadSceneResult
{
    PODVector<Drawable*> geometries_;
    PODVector<Light*> lights_;
    float minZ_;
    float maxZ_;
};

/// What View::ProcessShadowCasters asks a Light for (Light.h: GetLightMask is Drawable's, GetLightType, GetShadowFocus().focus_).
struct RefFocusShim { bool focus_{}; };
struct RefLightShim
{
    unsigned lightMask_{};
    LightType lightType_{};
    RefFocusShim shadowFocus_;
    unsigned GetLightMask() const { return lightMask_; }
    LightType GetLightType() const { return lightType_; }
    const RefFocusShim& GetShadowFocus() const { return shadowFocus_; }
};

/// Data holder with the members of View.h's LightQueryResult that ProcessShadowCasters touches (View.h:55-77).
struct RefLightQueryShim
{
    RefLightShim* light_{};
    PODVector<Drawable*> shadowCasters_;
    Camera* shadowCameras_[MAX_LIGHT_SPLITS];
    unsigned shadowCasterEnd_[MAX_LIGHT_SPLITS];
    BoundingBox shadowCasterBox_[MAX_LIGHT_SPLITS];
    float shadowNearSplits_[MAX_LIGHT_SPLITS];
    float shadowFarSplits_[MAX_LIGHT_SPLITS];
};

class RefViewShim
{
public:
    void ProcessShadowCasters(RefLightQueryShim& query, const PODVector<Drawable*>& drawables, unsigned splitIndex);
    /// View::GetShadowMask (View.h:297-300) without a zone: zones are outside this path, their mask is all ones.
    unsigned GetShadowMask(Drawable* drawable) { return drawable->GetShadowMask(); }
    float minZ_{};
    float maxZ_{};
    void UpdateOccluders(PODVector<Drawable*>& occluders, Camera* camera);
    void DrawOccluders(OcclusionBuffer* buffer, const PODVector<Drawable*>& occluders);
    int maxOccluderTriangles_{};
    unsigned activeOccluders_{};
    bool IsShadowCasterVisible(Drawable* drawable, BoundingBox lightViewBox, Camera* shadowCamera, const Matrix3x4& lightView,
        const Frustum& lightViewFrustum, const BoundingBox& lightViewFrustumBox);
    /// Never reached: the harness runs CheckVisibilityWork with cameraZoneOverride_ set (zones are outside this path).
    void FindZone(Drawable*) { abort(); }
    RefRendererShim* renderer_{};
    FrameInfo frame_{};
    // what CheckVisibilityWork reads through its View pointer
    OcclusionBuffer* occlusionBuffer_{};
    Camera* cullCamera_{};
    bool cameraZoneOverride_{};
    Vector<PerThreadSceneResult> sceneResults_;
};

#define View RefViewShim
#include "view_update_occluders.inc"
#include "view_draw_occluders.inc"
#include "view_shadow_caster_visible.inc"
#include "view_check_visibility_work.inc"
#define Light RefLightShim
#define LightQueryResult RefLightQueryShim
#include "view_process_shadow_casters.inc"
#undef LightQueryResult
#undef Light
#undef View
}

using namespace Urho3D;

namespace
{
using ZoneOccluderQuery = ZoneOccluderOctreeQuery;
using OccludedQuery = OccludedFrustumOctreeQuery;
using ShadowCasterQuery = ShadowCasterOctreeQuery;

/// Drawable whose world AABB is handed in by the test, and whose DrawOcclusion does what
/// StaticModel::DrawOcclusion does (StaticModel.cpp:189-228): SetCullMode + indexed AddTriangles.
class TestDrawable : public Drawable
{
    URHO3D_OBJECT(TestDrawable, Drawable);

public:
    explicit TestDrawable(Context* context) : Drawable(context, DRAWABLE_GEOMETRY) {}

    void Setup(const BoundingBox& worldBox, unsigned char flags, unsigned viewMask, bool occludee, bool castShadows, int id)
    {
        fixedWorldBox_ = worldBox;
        drawableFlags_ = flags;
        viewMask_ = viewMask;
        occludee_ = occludee;
        castShadows_ = castShadows;
        id_ = id;
        worldBoundingBoxDirty_ = true;
    }

    void OnWorldBoundingBoxUpdate() override { worldBoundingBox_ = fixedWorldBox_; }
    unsigned GetNumOccluderTriangles() override { return numOccluderTriangles_; }

    /// The calls StaticModel::DrawOcclusion makes per geometry (StaticModel.cpp:189-228): the material's cull mode, then the indexed triangles.
    void SetOcclusionMesh(const Matrix3x4& model, const void* vtx, unsigned vtxSize, const void* idx, unsigned idxSize, unsigned idxCount, CullMode cull)
    {
        occModel_ = model; occVtx_ = vtx; occVtxSize_ = vtxSize; occIdx_ = idx; occIdxSize_ = idxSize; occIdxCount_ = idxCount; occCull_ = cull;
    }
    bool DrawOcclusion(OcclusionBuffer* buffer) override
    {
        buffer->SetCullMode(occCull_);
        return buffer->AddTriangles(occModel_, occVtx_, occVtxSize_, occIdx_, occIdxSize_, 0, occIdxCount_);
    }

    BoundingBox fixedWorldBox_;
    int id_{};
    unsigned numOccluderTriangles_{};
    Matrix3x4 occModel_;
    const void* occVtx_{};
    const void* occIdx_{};
    unsigned occVtxSize_{}, occIdxSize_{}, occIdxCount_{};
    CullMode occCull_{CULL_CCW};
};

/// Occluder Drawables for View::DrawOccluders (the reference's text, RefViewShim): one per selected table entry, in order.
struct OccluderSet
{
    std::vector<SharedPtr<Node> > nodes_;
    PODVector<Drawable*> list_;
    OccluderSet(Context* c, int numSelected, const int* selected, const float* occAabb6, const float* occModel12, const void* vtx, unsigned vtxSize,
        const void* idx, unsigned idxSize, unsigned idxCount, int cullMode)
    {
        for (int i = 0; i < numSelected; ++i)
        {
            SharedPtr<Node> node(new Node(c));
            auto* d = node->CreateComponent<TestDrawable>();
            BoundingBox box;
            if (occAabb6)
            {
                const float* b = occAabb6 + 6 * (size_t)selected[i];
                box = BoundingBox(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]));
            }
            d->Setup(box, DRAWABLE_GEOMETRY, DEFAULT_VIEWMASK, true, false, selected[i]);
            d->SetOcclusionMesh(Matrix3x4(occModel12 + 12 * (size_t)selected[i]), vtx, vtxSize, idx, idxSize, idxCount, (CullMode)cullMode);
            nodes_.push_back(node);
            list_.Push(d);
        }
    }
};

struct Ref
{
    SharedPtr<Context> context_;
    SharedPtr<Scene> scene_;
    Octree* octree_{};
    SharedPtr<Node> cameraNode_;
    Camera* camera_{};
    SharedPtr<OcclusionBuffer> buffer_;
    std::vector<TestDrawable*> drawables_;
    PODVector<Drawable*> result_;
    std::vector<unsigned char> visFlags_;
    /// When set (ref_set_frustum_raw), queries use this frustum instead of the camera's.
    bool rawFrustum_{};
    Frustum frustum_;
    const Frustum& QueryFrustum() const { return rawFrustum_ ? frustum_ : camera_->GetFrustum(); }
};

struct VisWork
{
    OcclusionBuffer* buffer_;
    unsigned char* flagsBase_;
    Drawable** base_;
};

/// The occlusion predicate of CheckVisibilityWork (View.cpp:177) over a slice.
void CheckVisWork(const WorkItem* item, unsigned)
{
    auto* w = reinterpret_cast<VisWork*>(item->aux_);
    auto** start = reinterpret_cast<Drawable**>(item->start_);
    auto** end = reinterpret_cast<Drawable**>(item->end_);
    while (start != end)
    {
        Drawable* drawable = *start;
        bool vis = !w->buffer_ || !drawable->IsOccludee() || w->buffer_->IsVisible(drawable->GetWorldBoundingBox());
        w->flagsBase_[start - w->base_] = vis ? 1 : 0;
        ++start;
    }
}

double Now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void FlattenRec(const Octant* o, int parent, std::vector<int>& parents, std::vector<int>& levels, std::vector<float>& boxes,
    std::vector<int>& firsts, std::vector<int>& counts, std::vector<int>& order)
{
    int me = (int)parents.size();
    parents.push_back(parent);
    levels.push_back((int)o->level_);
    const BoundingBox& b = o->cullingBox_;
    boxes.push_back(b.min_.x_); boxes.push_back(b.min_.y_); boxes.push_back(b.min_.z_);
    boxes.push_back(b.max_.x_); boxes.push_back(b.max_.y_); boxes.push_back(b.max_.z_);
    firsts.push_back((int)order.size());
    counts.push_back((int)o->drawables_.Size());
    for (unsigned i = 0; i < o->drawables_.Size(); ++i)
        order.push_back(static_cast<TestDrawable*>(o->drawables_[i])->id_);
    for (unsigned i = 0; i < NUM_OCTANTS; ++i)
        if (o->children_[i])
            FlattenRec(o->children_[i], me, parents, levels, boxes, firsts, counts, order);
}

}

extern "C"
{

void* ref_create(int numWorkerThreads)
{
    auto* r = new Ref();
    r->context_ = new Context();
    Context* c = r->context_;
    auto* queue = new WorkQueue(c);
    c->RegisterSubsystem(queue);
    if (numWorkerThreads > 0)
        queue->CreateThreads((unsigned)numWorkerThreads);
    RegisterSceneLibrary(c);
    Octree::RegisterObject(c);
    Camera::RegisterObject(c);
    c->RegisterFactory<TestDrawable>();
    r->scene_ = new Scene(c);
    r->octree_ = r->scene_->CreateComponent<Octree>();
    r->cameraNode_ = new Node(c);
    r->camera_ = r->cameraNode_->CreateComponent<Camera>();
    r->buffer_ = new OcclusionBuffer(c);
    return r;
}

void ref_destroy(void* h)
{
    auto* r = reinterpret_cast<Ref*>(h);
    r->buffer_.Reset();
    r->cameraNode_.Reset();
    r->scene_.Reset();
    r->context_.Reset();
    delete r;
}

int ref_num_threads(void* h)
{
    return (int)reinterpret_cast<Ref*>(h)->context_->GetSubsystem<WorkQueue>()->GetNumThreads();
}

void ref_octree_set_size(void* h, const float* mn, const float* mx, unsigned levels)
{
    reinterpret_cast<Ref*>(h)->octree_->SetSize(BoundingBox(Vector3(mn), Vector3(mx)), levels);
}

/// Create n drawables (in order) and insert each once, from the root, with its final world box.
void ref_add_drawables(void* h, int n, const float* aabb6, const unsigned char* flags, const unsigned* viewMask,
    const unsigned char* occludee, const unsigned char* castShadows)
{
    auto* r = reinterpret_cast<Ref*>(h);
    Context* c = r->context_;
    for (int i = 0; i < n; ++i)
    {
        Node* node = r->scene_->CreateChild(String::EMPTY, LOCAL);
        auto* d = new TestDrawable(c);
        const float* b = aabb6 + 6 * (size_t)i;
        d->Setup(BoundingBox(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5])), flags ? flags[i] : (unsigned char)DRAWABLE_GEOMETRY,
            viewMask ? viewMask[i] : DEFAULT_VIEWMASK, occludee ? occludee[i] != 0 : true, castShadows ? castShadows[i] != 0 : false,
            (int)r->drawables_.size());
        r->drawables_.push_back(d);
        node->AddComponent(d, 0, LOCAL); // -> OnSceneSet -> AddToOctree -> Octree::InsertDrawable (Drawable.cpp:366-405)
    }
}

int ref_num_drawables(void* h) { return (int)reinterpret_cast<Ref*>(h)->drawables_.size(); }

/// Drawables moved / resized: what Drawable::OnMarkedDirty does (Drawable.cpp:374-383: world box dirty + Octree::QueueUpdate), followed by
/// the per-frame Octree::Update (Octree.cpp:364-475), whose reinsertion loop decides which drawables change octant.
void ref_move_drawables(void* h, int n, const int* ids, const float* aabb6)
{
    auto* r = reinterpret_cast<Ref*>(h);
    for (int i = 0; i < n; ++i)
    {
        TestDrawable* d = r->drawables_[ids[i]];
        if (!d)
            continue;
        const float* b = aabb6 + 6 * (size_t)i;
        d->fixedWorldBox_ = BoundingBox(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]));
        d->worldBoundingBoxDirty_ = true;
        if (!d->updateQueued_)
            r->octree_->QueueUpdate(d);
    }
    FrameInfo frame;
    frame.frameNumber_ = 1;
    frame.timeStep_ = 0.0f;
    frame.camera_ = nullptr;
    r->octree_->Update(frame);
}

/// A drawable leaves the scene: Node::Remove -> Drawable::OnSceneSet(nullptr) -> RemoveFromOctree (Drawable.cpp:366-420).
void ref_remove_drawable(void* h, int id)
{
    auto* r = reinterpret_cast<Ref*>(h);
    TestDrawable* d = r->drawables_[id];
    if (!d)
        return;
    r->drawables_[id] = nullptr;
    d->GetNode()->Remove();
}

int ref_octree_num_octants(void* h)
{
    auto* r = reinterpret_cast<Ref*>(h);
    std::vector<int> p, l, f, cnt, o; std::vector<float> b;
    FlattenRec(r->octree_, -1, p, l, b, f, cnt, o);
    return (int)p.size();
}

/// DFS pre-order dump (the order Octant::GetDrawablesInternal visits, Octree.cpp:229-255).
void ref_octree_flatten(void* h, int* parent, int* level, float* cullingBox6, int* first, int* count, int* order)
{
    auto* r = reinterpret_cast<Ref*>(h);
    std::vector<int> p, l, f, cnt, o; std::vector<float> b;
    FlattenRec(r->octree_, -1, p, l, b, f, cnt, o);
    memcpy(parent, p.data(), p.size() * sizeof(int));
    memcpy(level, l.data(), l.size() * sizeof(int));
    memcpy(cullingBox6, b.data(), b.size() * sizeof(float));
    memcpy(first, f.data(), f.size() * sizeof(int));
    memcpy(count, cnt.data(), cnt.size() * sizeof(int));
    memcpy(order, o.data(), o.size() * sizeof(int));
}

/// Camera set-up.  rot = quaternion (w,x,y,z).
void ref_camera_set(void* h, const float* pos, const float* rot, float fov, float aspect, float nearClip, float farClip,
    int ortho, float orthoSize, float zoom, int flipVertical)
{
    auto* r = reinterpret_cast<Ref*>(h);
    r->cameraNode_->SetTransform(Vector3(pos), Quaternion(rot[0], rot[1], rot[2], rot[3]));
    Camera* cam = r->camera_;
    cam->SetOrthographic(ortho != 0);
    cam->SetFov(fov);
    cam->SetNearClip(nearClip);
    cam->SetFarClip(farClip);
    if (ortho)
        cam->SetOrthoSize(orthoSize);
    cam->SetAspectRatio(aspect);
    cam->SetZoom(zoom);
    cam->SetFlipVertical(flipVertical != 0);
}

/// view (3x4 row-major), projection (4x4 row-major, as Camera::GetProjection incl. flip), 6 planes (nx,ny,nz,d),
/// absNormals (6x3), near/far (Camera::GetNearClip/GetFarClip), reverseCulling.
void ref_camera_get(void* h, float* view12, float* proj16, float* planes24, float* absNormals18, float* nearFar2, int* reverseCulling)
{
    auto* r = reinterpret_cast<Ref*>(h);
    Camera* cam = r->camera_;
    memcpy(view12, cam->GetView().Data(), 12 * sizeof(float));
    Matrix4 p = cam->GetProjection();
    memcpy(proj16, p.Data(), 16 * sizeof(float));
    const Frustum& f = cam->GetFrustum();
    for (unsigned i = 0; i < NUM_FRUSTUM_PLANES; ++i)
    {
        planes24[i * 4 + 0] = f.planes_[i].normal_.x_;
        planes24[i * 4 + 1] = f.planes_[i].normal_.y_;
        planes24[i * 4 + 2] = f.planes_[i].normal_.z_;
        planes24[i * 4 + 3] = f.planes_[i].d_;
        absNormals18[i * 3 + 0] = f.planes_[i].absNormal_.x_;
        absNormals18[i * 3 + 1] = f.planes_[i].absNormal_.y_;
        absNormals18[i * 3 + 2] = f.planes_[i].absNormal_.z_;
    }
    nearFar2[0] = cam->GetNearClip();
    nearFar2[1] = cam->GetFarClip();
    *reverseCulling = cam->GetReverseCulling() ? 1 : 0;
}

/// Inject the frustum planes directly (normal xyz + d per plane, order NEAR,LEFT,RIGHT,UP,DOWN,FAR; Frustum.h:37-45).
/// Plane::Define(Vector4) (Plane.h:83-88) derives absNormal_ exactly as the reference does.  planes24 == NULL -> camera frustum again.
void ref_set_frustum_raw(void* h, const float* planes24)
{
    auto* r = reinterpret_cast<Ref*>(h);
    r->rawFrustum_ = planes24 != nullptr;
    if (planes24)
        for (unsigned i = 0; i < NUM_FRUSTUM_PLANES; ++i)
            r->frustum_.planes_[i].Define(Vector4(planes24[i * 4], planes24[i * 4 + 1], planes24[i * 4 + 2], planes24[i * 4 + 3]));
}

/// Inject view/projection directly: exactly the assignments of OcclusionBuffer::SetView (OcclusionBuffer.cpp:115-127)
/// with the Camera getters replaced by caller-supplied values, so the same bytes can be fed to every implementation.
void ref_ob_set_view_raw(void* h, const float* view12, const float* proj16, float nearClip, float farClip, int reverseCulling)
{
    auto* r = reinterpret_cast<Ref*>(h);
    OcclusionBuffer* ob = r->buffer_;
    ob->view_ = Matrix3x4(view12);
    ob->projection_ = Matrix4(proj16);
    ob->viewProj_ = ob->projection_ * ob->view_;
    ob->nearClip_ = nearClip;
    ob->farClip_ = farClip;
    ob->reverseCulling_ = reverseCulling != 0;
    ob->CalculateViewport();
}

void ref_ob_get_viewproj(void* h, float* out16)
{
    memcpy(out16, reinterpret_cast<Ref*>(h)->buffer_->viewProj_.Data(), 16 * sizeof(float));
}

int ref_frustum_is_inside(void* h, const float* b, int fast)
{
    auto* r = reinterpret_cast<Ref*>(h);
    BoundingBox box(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]));
    return (int)(fast ? r->QueryFrustum().IsInsideFast(box) : r->QueryFrustum().IsInside(box));
}

int ref_ob_set_size(void* h, int w, int ht, int threaded)
{
    return reinterpret_cast<Ref*>(h)->buffer_->SetSize(w, ht, threaded != 0) ? 1 : 0;
}
int ref_ob_width(void* h) { return reinterpret_cast<Ref*>(h)->buffer_->GetWidth(); }
int ref_ob_height(void* h) { return reinterpret_cast<Ref*>(h)->buffer_->GetHeight(); }
void ref_ob_set_view(void* h) { auto* r = reinterpret_cast<Ref*>(h); r->buffer_->SetView(r->camera_); }
void ref_ob_set_max_triangles(void* h, unsigned n) { reinterpret_cast<Ref*>(h)->buffer_->SetMaxTriangles(n); }
void ref_ob_set_cull_mode(void* h, int mode) { reinterpret_cast<Ref*>(h)->buffer_->SetCullMode((CullMode)mode); }
int ref_ob_get_cull_mode(void* h) { return (int)reinterpret_cast<Ref*>(h)->buffer_->GetCullMode(); }
void ref_ob_clear(void* h) { reinterpret_cast<Ref*>(h)->buffer_->Clear(); }
void ref_ob_reset(void* h) { reinterpret_cast<Ref*>(h)->buffer_->Reset(); }

int ref_ob_add_triangles(void* h, const float* model12, const void* vtx, unsigned vtxSize, const void* idx, unsigned idxSize,
    unsigned start, unsigned count)
{
    auto* r = reinterpret_cast<Ref*>(h);
    Matrix3x4 model(model12);
    bool ok = idx ? r->buffer_->AddTriangles(model, vtx, vtxSize, idx, idxSize, start, count)
                  : r->buffer_->AddTriangles(model, vtx, vtxSize, start, count);
    return ok ? 1 : 0;
}

void ref_ob_draw(void* h) { reinterpret_cast<Ref*>(h)->buffer_->DrawTriangles(); }
void ref_ob_build_hierarchy(void* h) { reinterpret_cast<Ref*>(h)->buffer_->BuildDepthHierarchy(); }
unsigned ref_ob_num_triangles(void* h) { return reinterpret_cast<Ref*>(h)->buffer_->GetNumTriangles(); }

int ref_ob_is_visible(void* h, const float* b)
{
    return reinterpret_cast<Ref*>(h)->buffer_->IsVisible(BoundingBox(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]))) ? 1 : 0;
}

void ref_ob_is_visible_many(void* h, int n, const float* aabb6, unsigned char* out)
{
    auto* r = reinterpret_cast<Ref*>(h);
    for (int i = 0; i < n; ++i)
    {
        const float* b = aabb6 + 6 * (size_t)i;
        out[i] = r->buffer_->IsVisible(BoundingBox(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]))) ? 1 : 0;
    }
}

void ref_ob_read_depth(void* h, int* out)
{
    auto* r = reinterpret_cast<Ref*>(h);
    memcpy(out, r->buffer_->GetBuffer(), sizeof(int) * (size_t)r->buffer_->GetWidth() * r->buffer_->GetHeight());
}

int ref_ob_num_mips(void* h) { return (int)reinterpret_cast<Ref*>(h)->buffer_->mipBuffers_.Size(); }

/// Level `level` as interleaved (min,max) ints; dims follow OcclusionBuffer.cpp:97-106.
void ref_ob_read_mip(void* h, int level, int* out)
{
    auto* r = reinterpret_cast<Ref*>(h);
    int w = r->buffer_->GetWidth(), ht = r->buffer_->GetHeight();
    for (int i = 0; i <= level; ++i) { w = (w + 1) / 2; ht = (ht + 1) / 2; }
    memcpy(out, r->buffer_->mipBuffers_[level].Get(), sizeof(DepthValue) * (size_t)w * ht);
}

/// Octree::GetDrawables with a frustum query built from the current camera.
/// mode 0 = FrustumOctreeQuery, 1 = OccludedFrustumOctreeQuery (current buffer), 2 = ShadowCasterOctreeQuery.
int ref_query(void* h, int mode, unsigned flags, unsigned viewMask, int* outIdx, int cap)
{
    auto* r = reinterpret_cast<Ref*>(h);
    const Frustum& fr = r->QueryFrustum();
    if (mode == 1)
    {
        OccludedQuery q(r->result_, fr, r->buffer_, (unsigned char)flags, viewMask);
        r->octree_->GetDrawables(q);
    }
    else if (mode == 2)
    {
        ShadowCasterQuery q(r->result_, fr, (unsigned char)flags, viewMask);
        r->octree_->GetDrawables(q);
    }
    else
    {
        FrustumOctreeQuery q(r->result_, fr, (unsigned char)flags, viewMask);
        r->octree_->GetDrawables(q);
    }
    int n = (int)r->result_.Size();
    for (int i = 0; i < n && i < cap; ++i)
        outIdx[i] = static_cast<TestDrawable*>(r->result_[i])->id_;
    return n;
}

/// Octree::GetDrawables with the reference's own SphereOctreeQuery (shape 3: params = centre.xyz, radius), BoxOctreeQuery (4: min.xyz, max.xyz)
/// or PointOctreeQuery (5: xyz), OctreeQuery.h:72-138.
int ref_query_shape(void* h, int shape, const float* params, unsigned flags, unsigned viewMask, int* outIdx, int cap)
{
    auto* r = reinterpret_cast<Ref*>(h);
    if (shape == 3)
    {
        SphereOctreeQuery q(r->result_, Sphere(Vector3(params[0], params[1], params[2]), params[3]), (unsigned char)flags, viewMask);
        r->octree_->GetDrawables(q);
    }
    else if (shape == 4)
    {
        BoxOctreeQuery q(r->result_, BoundingBox(Vector3(params[0], params[1], params[2]), Vector3(params[3], params[4], params[5])), (unsigned char)flags,
            viewMask);
        r->octree_->GetDrawables(q);
    }
    else
    {
        PointOctreeQuery q(r->result_, Vector3(params[0], params[1], params[2]), (unsigned char)flags, viewMask);
        r->octree_->GetDrawables(q);
    }
    int n = (int)r->result_.Size();
    for (int i = 0; i < n && i < cap; ++i)
        outIdx[i] = static_cast<TestDrawable*>(r->result_[i])->id_;
    return n;
}

/// CheckVisibilityWork's occlusion predicate over the last ref_query() result, work split as View.cpp:897-920.
/// Output keeps query order (set-equal to what the reference's per-thread lists hold).
int ref_check_visibility(void* h, int useBuffer, int* outIdx, int cap)
{
    auto* r = reinterpret_cast<Ref*>(h);
    auto* queue = r->context_->GetSubsystem<WorkQueue>();
    unsigned n = r->result_.Size();
    r->visFlags_.assign(n, 0);
    if (!n)
        return 0;
    VisWork w{useBuffer ? r->buffer_.Get() : nullptr, r->visFlags_.data(), &r->result_[0]};

    int numWorkItems = (int)queue->GetNumThreads() + 1;
    int drawablesPerItem = (int)n / numWorkItems;
    Drawable** start = &r->result_[0];
    Drawable** endAll = start + n;
    for (int i = 0; i < numWorkItems; ++i)
    {
        SharedPtr<WorkItem> item = queue->GetFreeItem();
        item->priority_ = M_MAX_UNSIGNED;
        item->workFunction_ = CheckVisWork;
        item->aux_ = &w;
        Drawable** end = endAll;
        if (i < numWorkItems - 1 && end - start > drawablesPerItem)
            end = start + drawablesPerItem;
        item->start_ = start;
        item->end_ = end;
        queue->AddWorkItem(item);
        start = end;
    }
    queue->Complete(M_MAX_UNSIGNED);

    int m = 0;
    for (unsigned i = 0; i < n; ++i)
        if (r->visFlags_[i])
        {
            if (m < cap)
                outIdx[m] = static_cast<TestDrawable*>(r->result_[i])->id_;
            ++m;
        }
    return m;
}

/// The part of CheckVisibilityWork (View.cpp:160-227) that decides what is in view and the Z range, over the last ref_query() result, one
/// thread, with the reference's own math types.  The view matrix and camera position are injected (the harness runs without a Camera node
/// in raw mode): distance = Camera::GetDistance (Camera.cpp:487-496) evaluated on those values.  Returns the number in view.
int ref_check_in_view(void* h, int useBuffer, const float* view12, const float* camPos3, int ortho, const float* drawDistance /* by id */,
    int* outIdx, int cap, float* minmax2)
{
    auto* r = reinterpret_cast<Ref*>(h);
    OcclusionBuffer* buffer = useBuffer ? r->buffer_.Get() : nullptr;
    Matrix3x4 viewMatrix(view12);
    Vector3 cameraPos(camPos3);
    Vector3 viewZ = Vector3(viewMatrix.m20_, viewMatrix.m21_, viewMatrix.m22_);
    Vector3 absViewZ = viewZ.Abs();
    float minZ = M_INFINITY, maxZ = 0.0f;
    int m = 0;
    for (unsigned i = 0; i < r->result_.Size(); ++i)
    {
        auto* drawable = static_cast<TestDrawable*>(r->result_[i]);
        if (!buffer || !drawable->IsOccludee() || buffer->IsVisible(drawable->GetWorldBoundingBox()))
        {
            const BoundingBox& geomBox = drawable->GetWorldBoundingBox();
            float maxDistance = drawDistance ? drawDistance[drawable->id_] : 0.0f;
            if (maxDistance > 0.0f)
            {
                Vector3 worldPos = geomBox.Center();
                float distance = !ortho ? (worldPos - cameraPos).Length() : Abs((viewMatrix * worldPos).z_);
                if (distance > maxDistance)
                    continue;
            }
            if (drawable->GetDrawableFlags() & DRAWABLE_GEOMETRY)
            {
                Vector3 center = geomBox.Center();
                Vector3 edge = geomBox.Size() * 0.5f;
                if (edge.LengthSquared() < M_LARGE_VALUE * M_LARGE_VALUE)
                {
                    float viewCenterZ = viewZ.DotProduct(center) + viewMatrix.m23_;
                    float viewEdgeZ = absViewZ.DotProduct(edge);
                    minZ = Min(minZ, viewCenterZ - viewEdgeZ);
                    maxZ = Max(maxZ, viewCenterZ + viewEdgeZ);
                }
            }
            if (m < cap)
                outIdx[m] = drawable->id_;
            ++m;
        }
    }
    if (minZ == M_INFINITY)
        minZ = 0.0f;
    minmax2[0] = minZ;
    minmax2[1] = maxZ;
    return m;
}

/// CheckVisibilityWork as the reference wrote it (the extracted text of View.cpp:160-227, one work item on thread 0) over the last ref_query()
/// result, with the harness's real Camera (ref_camera_set) as the cull camera: Drawable::UpdateBatches computes distance_, MarkInView marks,
/// the Z range accumulates in a PerThreadSceneResult initialised and merged as View::GetDrawables does (View.cpp:890-893, 926-951).
/// Draw distances are set on the Drawables (by id).  Returns the drawables marked in view, in result order.  Pins ref_check_in_view.
int ref_check_in_view_text(void* h, int useBuffer, const float* drawDistance /* by id */, int* outIdx, int cap, float* minmax2)
{
    auto* r = reinterpret_cast<Ref*>(h);
    static unsigned frameNumber = 9000;
    RefViewShim view;
    view.occlusionBuffer_ = useBuffer ? r->buffer_.Get() : nullptr;
    view.cullCamera_ = r->camera_;
    view.cameraZoneOverride_ = true;
    view.frame_.frameNumber_ = ++frameNumber;
    view.frame_.camera_ = r->camera_;
    view.sceneResults_.Resize(1);
    PerThreadSceneResult& result = view.sceneResults_[0];
    result.geometries_.Clear();
    result.lights_.Clear();
    result.minZ_ = M_INFINITY;
    result.maxZ_ = 0.0f;
    for (unsigned i = 0; i < r->result_.Size(); ++i)
    {
        auto* d = static_cast<TestDrawable*>(r->result_[i]);
        d->SetDrawDistance(drawDistance ? drawDistance[d->id_] : 0.0f);
    }
    if (r->result_.Size())
    {
        WorkItem item;
        item.aux_ = &view;
        item.start_ = &r->result_[0];
        item.end_ = &r->result_[0] + r->result_.Size();
        CheckVisibilityWork(&item, 0);
    }
    int m = 0;
    for (unsigned i = 0; i < r->result_.Size(); ++i)
        if (r->result_[i]->IsInView(view.frame_))
        {
            if (m < cap)
                outIdx[m] = static_cast<TestDrawable*>(r->result_[i])->id_;
            ++m;
        }
    float minZ = result.minZ_, maxZ = result.maxZ_;
    if (minZ == M_INFINITY)
        minZ = 0.0f;
    minmax2[0] = minZ;
    minmax2[1] = maxZ;
    return m;
}

/// One whole view, the way View::GetDrawables runs it with threaded occlusion (View.cpp:2258-2273, :873-878, :885-921):
/// select occluders by frustum.IsInsideFast(worldBox) in scene order, Clear, AddTriangles each,
/// DrawTriangles, BuildDepthHierarchy, occluded query, visibility predicate.  Camera must be set already.
/// times4 (seconds): raster(incl. select+submit), hierarchy, query, visibility.  Returns visible count.
int ref_run_view(void* h, int numOccluders, const float* occAabb6, const float* occModel12, const void* vtx, unsigned vtxSize,
    const void* idx, unsigned idxSize, unsigned idxCount, int cullMode, unsigned flags, unsigned viewMask, int* outVisible, int cap,
    int* outCounts3 /* submittedTris, candidates, visible */, double* times4)
{
    auto* r = reinterpret_cast<Ref*>(h);
    OcclusionBuffer* ob = r->buffer_;
    double t0 = Now();
    if (!r->rawFrustum_)
        ob->SetView(r->camera_); // raw mode: caller already injected the view with ref_ob_set_view_raw
    ob->SetMaxTriangles(1u << 30);
    ob->Clear();
    const Frustum& fr = r->QueryFrustum();
    unsigned submitted = 0;
    for (int i = 0; i < numOccluders; ++i)
    {
        const float* b = occAabb6 + 6 * (size_t)i;
        if (fr.IsInsideFast(BoundingBox(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]))) == OUTSIDE)
            continue;
        ob->SetCullMode((CullMode)cullMode);
        ob->AddTriangles(Matrix3x4(occModel12 + 12 * (size_t)i), vtx, vtxSize, idx, idxSize, 0, idxCount);
        submitted += idxCount / 3;
    }
    ob->DrawTriangles();
    double t1 = Now();
    ob->BuildDepthHierarchy();
    double t2 = Now();
    OccludedQuery q(r->result_, fr, ob, (unsigned char)flags, viewMask);
    r->octree_->GetDrawables(q);
    double t3 = Now();
    int nvis = ref_check_visibility(h, 1, outVisible, cap);
    double t4 = Now();
    if (outCounts3)
    {
        outCounts3[0] = (int)submitted;
        outCounts3[1] = (int)r->result_.Size();
        outCounts3[2] = nvis;
    }
    if (times4)
    {
        times4[0] = t1 - t0; times4[1] = t2 - t1; times4[2] = t3 - t2; times4[3] = t4 - t3;
    }
    return nvis;
}

/// Octree::Raycast (single == 0) / Octree::RaycastSingle (single != 0) with the reference's own RayOctreeQuery at RAY_AABB level
/// (Octree.cpp:501-548; the test drawable keeps Drawable::ProcessRayQuery, Drawable.cpp:118-132).  ray7 = origin, direction (used as given: no
/// normalisation, Ray::direction_ is assigned directly), maxDistance.  Returns the result size; ids / distances / positions in result_ order.
int ref_raycast(void* h, const float* ray7, unsigned flags, unsigned viewMask, int single, int* outIdx, float* outDist, float* outPos3, int cap)
{
    auto* r = reinterpret_cast<Ref*>(h);
    Ray ray;
    ray.origin_ = Vector3(ray7[0], ray7[1], ray7[2]);
    ray.direction_ = Vector3(ray7[3], ray7[4], ray7[5]);
    PODVector<RayQueryResult> results;
    RayOctreeQuery q(results, ray, RAY_AABB, ray7[6], (unsigned char)flags, viewMask);
    if (single)
        r->octree_->RaycastSingle(q);
    else
        r->octree_->Raycast(q);
    int n = (int)results.Size();
    for (int i = 0; i < n && i < cap; ++i)
    {
        outIdx[i] = static_cast<TestDrawable*>(results[i].drawable_)->id_;
        outDist[i] = results[i].distance_;
        if (outPos3)
        {
            outPos3[3 * i] = results[i].position_.x_;
            outPos3[3 * i + 1] = results[i].position_.y_;
            outPos3[3 * i + 2] = results[i].position_.z_;
        }
    }
    return n;
}

/// The reference's Sort (Container/Sort.h) over float keys with an index payload, compare = key less-than (as CompareDrawables /
/// CompareRayQueryResults do).  Used to pin the restated sort.
namespace
{
struct KeyIdx { float key; int idx; };
bool KeyLess(const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; }
}
void ref_sort_by_key(int n, float* keys, int* vals)
{
    PODVector<KeyIdx> v;
    v.Resize((unsigned)n);
    for (int i = 0; i < n; ++i)
    {
        v[i].key = keys[i];
        v[i].idx = vals[i];
    }
    if (n)
        Sort(v.Begin(), v.End(), KeyLess);
    for (int i = 0; i < n; ++i)
    {
        keys[i] = v[i].key;
        vals[i] = v[i].idx;
    }
}

/// View::UpdateOccluders (View.cpp:2171-2229) restated on the reference's own math types, BoundingBox / Vector3 / Matrix3x4 arithmetic and Sort
/// (View.cpp itself needs the GL back-end and cannot be compiled here).  Candidates = table entries whose box passes the query frustum's
/// IsInsideFast, in table order.  Camera quantities are injected like in ref_check_in_view: distance_ = Camera::GetDistance (Camera.cpp:487-496).
/// Returns the number of occluders kept; outIdx / outKeys = occluder indices and sort values in sorted order.
namespace
{
struct OccKey { float sortValue; int index; };
bool OccKeyLess(const OccKey& a, const OccKey& b) { return a.sortValue < b.sortValue; } // CompareDrawables, Drawable.h:419-422
}
int ref_select_occluders(void* h, int numOccluders, const float* occAabb6, const unsigned* numTriangles, const float* drawDistance, const float* view12,
    const float* camPos3, int ortho, float halfViewSize, float orthoSize, float farClip, float occluderSizeThreshold, int* outIdx, float* outKeys, int cap)
{
    auto* r = reinterpret_cast<Ref*>(h);
    const Frustum& fr = r->QueryFrustum();
    Matrix3x4 viewMatrix(view12);
    Vector3 cameraPos(camPos3);
    float invOrthoSize = 1.0f / orthoSize;
    PODVector<OccKey> occluders;
    for (int i = 0; i < numOccluders; ++i)
    {
        const float* b = occAabb6 + 6 * (size_t)i;
        BoundingBox box(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]));
        if (fr.IsInsideFast(box) == OUTSIDE)
            continue;
        Vector3 worldPos = box.Center();
        float distance = !ortho ? (worldPos - cameraPos).Length() : Abs((viewMatrix * worldPos).z_);
        float maxDistance = drawDistance ? drawDistance[i] : 0.0f;
        if (maxDistance <= 0.0f || distance <= maxDistance)
        {
            float diagonal = box.Size().Length();
            float compare;
            if (!ortho)
            {
                float cameraMaxDistanceFraction = distance / farClip;
                compare = diagonal * halfViewSize / (distance * cameraMaxDistanceFraction);
                if (box.IsInside(cameraPos))
                    compare *= diagonal;
            }
            else
                compare = diagonal * invOrthoSize;
            if (compare < occluderSizeThreshold)
                continue;
            float density = numTriangles[i] / diagonal;
            OccKey k;
            k.sortValue = density / compare;
            k.index = i;
            occluders.Push(k);
        }
    }
    if (occluders.Size())
        Sort(occluders.Begin(), occluders.End(), OccKeyLess);
    int n = (int)occluders.Size();
    for (int i = 0; i < n && i < cap; ++i)
    {
        outIdx[i] = occluders[i].index;
        outKeys[i] = occluders[i].sortValue;
    }
    return n;
}

/// View::UpdateOccluders as the reference wrote it: the extracted text of View.cpp:2171-2229 (RefViewShim above) over real Drawables, one per
/// table entry whose box passes the query frustum's IsInsideFast, with the harness's real Camera (ref_camera_set) as the cull camera.
/// Drawable::UpdateBatches (Drawable.cpp:118-132) computes distance_ with Camera::GetDistance, CompareDrawables + Sort order the list.
/// Pins ref_select_occluders (the restatement with injected camera quantities) to the reference's own lines.
int ref_select_occluders_text(void* h, int numOccluders, const float* occAabb6, const unsigned* numTriangles, const float* drawDistance,
    float occluderSizeThreshold, int* outIdx, float* outKeys, int cap)
{
    auto* r = reinterpret_cast<Ref*>(h);
    const Frustum& fr = r->QueryFrustum();
    static unsigned frameNumber = 1000;
    RefRendererShim renderer;
    renderer.occluderSizeThreshold_ = occluderSizeThreshold;
    RefViewShim view;
    view.renderer_ = &renderer;
    view.frame_.frameNumber_ = ++frameNumber;
    view.frame_.timeStep_ = 0.0f;
    view.frame_.viewSize_ = IntVector2(1024, 768);
    view.frame_.camera_ = r->camera_;
    std::vector<SharedPtr<Node> > nodes;
    PODVector<Drawable*> occluders;
    for (int i = 0; i < numOccluders; ++i)
    {
        const float* b = occAabb6 + 6 * (size_t)i;
        BoundingBox box(Vector3(b[0], b[1], b[2]), Vector3(b[3], b[4], b[5]));
        if (fr.IsInsideFast(box) == OUTSIDE)
            continue;
        SharedPtr<Node> node(new Node(r->context_));
        auto* d = node->CreateComponent<TestDrawable>();
        d->Setup(box, DRAWABLE_GEOMETRY, DEFAULT_VIEWMASK, true, false, i);
        d->numOccluderTriangles_ = numTriangles[i];
        d->SetDrawDistance(drawDistance ? drawDistance[i] : 0.0f);
        nodes.push_back(node);
        occluders.Push(d);
    }
    view.UpdateOccluders(occluders, r->camera_);
    int n = (int)occluders.Size();
    for (int i = 0; i < n && i < cap; ++i)
    {
        outIdx[i] = static_cast<TestDrawable*>(occluders[i])->id_;
        outKeys[i] = occluders[i]->GetSortValue();
    }
    return n;
}

/// View::DrawOccluders (the reference's own text, View.cpp:2231-2275; threaded branch), over an already sorted occluder list, then the query and the
/// visibility predicate as ref_run_view.  outCounts4 = submitted triangles, query result size, visible, activeOccluders_.
int ref_run_view_selected(void* h, int numSelected, const int* selected, const float* occModel12, const void* vtx, unsigned vtxSize, const void* idx,
    unsigned idxSize, unsigned idxCount, int cullMode, unsigned maxTriangles, unsigned flags, unsigned viewMask, int* outVisible, int cap, int* outCounts4)
{
    auto* r = reinterpret_cast<Ref*>(h);
    OcclusionBuffer* ob = r->buffer_;
    if (!r->rawFrustum_)
        ob->SetView(r->camera_);
    // View::DrawOccluders, the reference's own text, threaded branch.  IsThreaded() is "more than one work buffer" (OcclusionBuffer.h), which
    // SetSize(.., threaded) only allocates when the WorkQueue has worker threads; a harness without workers gets the same layout by adding the
    // second work buffer the way SetSize does (OB.cpp:83-92): the queued batches then run on the calling thread and MergeBuffers folds them.
    const bool single = ob->buffers_.Size() == 1;
    if (single)
    {
        ob->buffers_.Resize(2);
        OcclusionBufferData& extra = ob->buffers_[1];
        extra.dataWithSafety_ = new int[ob->width_ * (ob->height_ + 2) + 2];
        extra.data_ = extra.dataWithSafety_.Get() + ob->width_ + 1;
        extra.used_ = false;
    }
    OccluderSet occ(r->context_, numSelected, selected, nullptr, occModel12, vtx, vtxSize, idx, idxSize, idxCount, cullMode);
    RefViewShim view;
    view.maxOccluderTriangles_ = (int)maxTriangles;
    view.DrawOccluders(ob, occ.list_);
    if (single)
        ob->buffers_.Resize(1);
    const unsigned active = view.activeOccluders_, submitted = active * (idxCount / 3);
    const Frustum& fr = r->QueryFrustum();
    OccludedQuery q(r->result_, fr, ob, (unsigned char)flags, viewMask);
    r->octree_->GetDrawables(q);
    int nvis = ref_check_visibility(h, 1, outVisible, cap);
    if (outCounts4)
    {
        outCounts4[0] = (int)submitted;
        outCounts4[1] = (int)r->result_.Size();
        outCounts4[2] = nvis;
        outCounts4[3] = (int)active;
    }
    return nvis;
}

/// Drawable::SetOccluder for every test drawable (by id), then Octree::GetDrawables with the zone / occluder query of View.cpp:803-807.
int ref_query_zone_occluder(void* h, const unsigned char* isOccluder, unsigned viewMask, int* outIdx, int cap)
{
    auto* r = reinterpret_cast<Ref*>(h);
    for (unsigned i = 0; i < r->drawables_.size(); ++i)
        r->drawables_[i]->SetOccluder(isOccluder[i] != 0);
    ZoneOccluderQuery q(r->result_, r->QueryFrustum(), DRAWABLE_GEOMETRY | DRAWABLE_ZONE, viewMask);
    r->octree_->GetDrawables(q);
    int n = (int)r->result_.Size();
    for (int i = 0; i < n && i < cap; ++i)
        outIdx[i] = static_cast<TestDrawable*>(r->result_[i])->id_;
    return n;
}

/// Inputs of View::ProcessShadowCasters for one shadow split, computed with the reference's own Camera / Frustum / BoundingBox code
/// (View.cpp:2385-2407): the cull camera's split frustum clamped to [nearZ, farZ], transformed into the shadow camera's view space.
/// cam = pos3, rot4 (w,x,y,z), fov, aspect, near, far, ortho, orthoSize  (10 + 2 floats: see refbind.shadow_split_setup).
/// Outputs: shadow camera view (12), shadow camera frustum planes (24), light-view frustum planes (24), light-view frustum box max z,
/// shadow camera far clip; returns 1 when the split frustum is degenerate (vertices_[0] == vertices_[4], View.cpp:2406-2407).
int ref_shadow_split_setup(void* h, const float* mainCam12, const float* shadowCam12, float nearZ, float farZ, float* lightView12,
    float* shadowPlanes24, float* lightViewFrustumPlanes24, float* lightViewFrustumMaxZ, float* shadowFarClip)
{
    auto* r = reinterpret_cast<Ref*>(h);
    Context* c = r->context_;
    auto make = [&](const float* p, SharedPtr<Node>& node) -> Camera* {
        node = new Node(c);
        node->SetTransform(Vector3(p), Quaternion(p[3], p[4], p[5], p[6]));
        Camera* cam = node->CreateComponent<Camera>();
        cam->SetOrthographic(p[10] != 0.0f);
        cam->SetFov(p[7]);
        cam->SetNearClip(p[8]);
        cam->SetFarClip(p[9]);
        if (p[10] != 0.0f)
            cam->SetOrthoSize(p[11]);
        cam->SetAspectRatio(1.0f);
        return cam;
    };
    SharedPtr<Node> mainNode, shadowNode;
    Camera* cullCamera = make(mainCam12, mainNode);
    Camera* shadowCamera = make(shadowCam12, shadowNode);
    const Matrix3x4& lightView = shadowCamera->GetView();
    Frustum lightViewFrustum = cullCamera->GetSplitFrustum(nearZ, farZ).Transformed(lightView);
    BoundingBox lightViewFrustumBox(lightViewFrustum);
    memcpy(lightView12, lightView.Data(), 12 * sizeof(float));
    const Frustum& sf = shadowCamera->GetFrustum();
    for (unsigned i = 0; i < NUM_FRUSTUM_PLANES; ++i)
    {
        shadowPlanes24[4 * i] = sf.planes_[i].normal_.x_;
        shadowPlanes24[4 * i + 1] = sf.planes_[i].normal_.y_;
        shadowPlanes24[4 * i + 2] = sf.planes_[i].normal_.z_;
        shadowPlanes24[4 * i + 3] = sf.planes_[i].d_;
        lightViewFrustumPlanes24[4 * i] = lightViewFrustum.planes_[i].normal_.x_;
        lightViewFrustumPlanes24[4 * i + 1] = lightViewFrustum.planes_[i].normal_.y_;
        lightViewFrustumPlanes24[4 * i + 2] = lightViewFrustum.planes_[i].normal_.z_;
        lightViewFrustumPlanes24[4 * i + 3] = lightViewFrustum.planes_[i].d_;
    }
    *lightViewFrustumMaxZ = lightViewFrustumBox.max_.z_;
    *shadowFarClip = shadowCamera->GetFarClip();
    return lightViewFrustum.vertices_[0] == lightViewFrustum.vertices_[4] ? 1 : 0;
}

/// The per-drawable part of View::ProcessShadowCasters (View.cpp:2409-2452) restated on the reference's math types (BoundingBox::Transformed,
/// Frustum::IsInsideFast), with View::IsShadowCasterVisible (View.cpp:2455-2489) called as the reference wrote it (extracted text, RefViewShim),
/// over `numIn` drawable ids in order.  Per-drawable state the engine keeps on
/// the Drawable is injected: shadow mask (View::GetShadowMask), shadow / draw distance, in-view flag (Drawable::IsInView(frame_)); distance_ is
/// Camera::GetDistance of the cull camera (mainView12 / mainPos3 / mainOrtho, as ref_check_in_view).  Returns the shadow casters kept, in order.
int ref_process_shadow_casters(void* h, int numIn, const int* inIds, const float* lightView12, const float* shadowPlanes24,
    const float* lightViewFrustumPlanes24, float lightViewFrustumMaxZ, float shadowFarClip, int shadowOrtho, int lightType, unsigned lightMask,
    const unsigned* shadowMask, const float* shadowDistance, const float* drawDistanceArr, const unsigned char* inView, const float* mainView12,
    const float* mainPos3, int mainOrtho, int* outIds)
{
    auto* r = reinterpret_cast<Ref*>(h);
    Matrix3x4 lightView(lightView12);
    Matrix3x4 mainView(mainView12);
    Vector3 cameraPos(mainPos3);
    Frustum shadowCameraFrustum, lightViewFrustum;
    for (unsigned i = 0; i < NUM_FRUSTUM_PLANES; ++i)
    {
        shadowCameraFrustum.planes_[i].Define(Vector4(shadowPlanes24[4 * i], shadowPlanes24[4 * i + 1], shadowPlanes24[4 * i + 2], shadowPlanes24[4 * i + 3]));
        lightViewFrustum.planes_[i].Define(Vector4(lightViewFrustumPlanes24[4 * i], lightViewFrustumPlanes24[4 * i + 1], lightViewFrustumPlanes24[4 * i + 2],
            lightViewFrustumPlanes24[4 * i + 3]));
    }
    // what IsShadowCasterVisible asks its arguments and frame_ for: the shadow camera's projection type and far clip, the light-view frustum
    // box's far side, and whether the drawable was marked in view this frame for the cull camera
    static unsigned frameNumber = 5000;
    SharedPtr<Node> shadowNode(new Node(r->context_));
    Camera* shadowCamera = shadowNode->CreateComponent<Camera>();
    shadowCamera->SetOrthographic(shadowOrtho != 0);
    shadowCamera->SetFarClip(shadowFarClip);
    BoundingBox lightViewFrustumBox(Vector3(0.0f, 0.0f, 0.0f), Vector3(0.0f, 0.0f, lightViewFrustumMaxZ));
    RefViewShim view;
    view.frame_.frameNumber_ = ++frameNumber;
    view.frame_.camera_ = r->camera_;
    int n = 0;
    for (int k = 0; k < numIn; ++k)
    {
        const int id = inIds[k];
        TestDrawable* drawable = r->drawables_[id];
        if (!drawable->GetCastShadows())
            continue;
        if (!(shadowMask[id] & lightMask))
            continue;
        if (lightType == 2 /* LIGHT_POINT */ && shadowCameraFrustum.IsInsideFast(drawable->GetWorldBoundingBox()) == OUTSIDE)
            continue;
        float maxShadowDistance = shadowDistance[id];
        float drawDistance = drawDistanceArr[id];
        if (drawDistance > 0.0f && (maxShadowDistance <= 0.0f || drawDistance < maxShadowDistance))
            maxShadowDistance = drawDistance;
        Vector3 worldPos = drawable->GetWorldBoundingBox().Center();
        float distance = !mainOrtho ? (worldPos - cameraPos).Length() : Abs((mainView * worldPos).z_);
        if (maxShadowDistance > 0.0f && distance > maxShadowDistance)
            continue;
        BoundingBox lightViewBox = drawable->GetWorldBoundingBox().Transformed(lightView);
        // View::IsShadowCasterVisible: the reference's own text (View.cpp:2455-2489, RefViewShim above)
        if (inView[id])
            drawable->MarkInView(view.frame_);
        if (view.IsShadowCasterVisible(drawable, lightViewBox, shadowCamera, lightView, lightViewFrustum, lightViewFrustumBox))
            outIds[n++] = id;
    }
    return n;
}

/// View::ProcessShadowCasters as the reference wrote it (the extracted text of View.cpp:2379-2453, which calls the extracted
/// IsShadowCasterVisible) for one split: real cull and shadow Cameras built from the 12-float descriptions of ref_shadow_split_setup, the split
/// frustum computed inside the text (minZ_ / maxZ_ = nearZ / farZ, the directional splits wide open, so both light kinds use [nearZ, farZ]), real
/// Drawables carrying their shadow mask / shadow distance / draw distance; drawables flagged in view have had UpdateBatches + MarkInView for this
/// frame, as CheckVisibilityWork leaves them.  Returns the shadow casters kept, in order.  Pins ref_shadow_split_setup + ref_process_shadow_casters.
int ref_process_shadow_casters_text(void* h, int numIn, const int* inIds, const float* mainCam12, const float* shadowCam12, float nearZ, float farZ,
    int lightType, unsigned lightMask, const unsigned* shadowMask, const float* shadowDistance, const float* drawDistanceArr,
    const unsigned char* inView, int* outIds)
{
    auto* r = reinterpret_cast<Ref*>(h);
    Context* c = r->context_;
    auto make = [&](const float* p, SharedPtr<Node>& node) -> Camera* {
        node = new Node(c);
        node->SetTransform(Vector3(p), Quaternion(p[3], p[4], p[5], p[6]));
        Camera* cam = node->CreateComponent<Camera>();
        cam->SetOrthographic(p[10] != 0.0f);
        cam->SetFov(p[7]);
        cam->SetNearClip(p[8]);
        cam->SetFarClip(p[9]);
        if (p[10] != 0.0f)
            cam->SetOrthoSize(p[11]);
        cam->SetAspectRatio(1.0f);
        return cam;
    };
    SharedPtr<Node> mainNode, shadowNode;
    Camera* cullCamera = make(mainCam12, mainNode);
    Camera* shadowCamera = make(shadowCam12, shadowNode);
    static unsigned frameNumber = 20000;
    RefViewShim view;
    view.cullCamera_ = cullCamera;
    view.minZ_ = nearZ;
    view.maxZ_ = farZ;
    view.frame_.frameNumber_ = ++frameNumber;
    view.frame_.camera_ = cullCamera;
    RefLightShim light;
    light.lightMask_ = lightMask;
    light.lightType_ = (LightType)lightType;
    RefLightQueryShim query;
    query.light_ = &light;
    for (int i = 0; i < MAX_LIGHT_SPLITS; ++i)
    {
        query.shadowCameras_[i] = shadowCamera;
        query.shadowNearSplits_[i] = 0.0f;
        query.shadowFarSplits_[i] = M_INFINITY;
        query.shadowCasterEnd_[i] = 0;
    }
    PODVector<Drawable*> drawables;
    for (int k = 0; k < numIn; ++k)
    {
        const int id = inIds[k];
        TestDrawable* d = r->drawables_[id];
        d->SetShadowMask(shadowMask[id]);
        d->SetShadowDistance(shadowDistance[id]);
        d->SetDrawDistance(drawDistanceArr[id]);
        if (inView[id])
        {
            d->UpdateBatches(view.frame_);
            d->MarkInView(view.frame_);
        }
        drawables.Push(d);
    }
    view.ProcessShadowCasters(query, drawables, 0);
    const int n = (int)query.shadowCasters_.Size();
    for (int i = 0; i < n; ++i)
        outIds[i] = static_cast<TestDrawable*>(query.shadowCasters_[i])->id_;
    return n;
}

/// View::DrawOccluders (the reference's own text, View.cpp:2231-2275; NON-threaded branch) with the reference's own OcclusionBuffer (created non-threaded by
/// ref_ob_set_size(.., threaded = 0)) over an already sorted occluder list, then the query and the visibility predicate as ref_run_view.
/// outCounts4 = numTriangles_, query result size, visible, activeOccluders_.
int ref_run_view_serial(void* h, int numSelected, const int* selected, const float* occAabb6, const float* occModel12, const void* vtx, unsigned vtxSize,
    const void* idx, unsigned idxSize, unsigned idxCount, int cullMode, unsigned maxTriangles, unsigned flags, unsigned viewMask, int* outVisible, int cap,
    int* outCounts4)
{
    auto* r = reinterpret_cast<Ref*>(h);
    OcclusionBuffer* buffer = r->buffer_;
    if (buffer->IsThreaded())
        return -1;
    if (!r->rawFrustum_)
        buffer->SetView(r->camera_);
    // View::DrawOccluders, the reference's own text (non-threaded branch: IsVisible-gated, one occluder at a time)
    OccluderSet occ(r->context_, numSelected, selected, occAabb6, occModel12, vtx, vtxSize, idx, idxSize, idxCount, cullMode);
    RefViewShim view;
    view.maxOccluderTriangles_ = (int)maxTriangles;
    view.DrawOccluders(buffer, occ.list_);
    const unsigned active = view.activeOccluders_;
    const Frustum& fr = r->QueryFrustum();
    OccludedQuery q(r->result_, fr, buffer, (unsigned char)flags, viewMask);
    r->octree_->GetDrawables(q);
    int nvis = ref_check_visibility(h, 1, outVisible, cap);
    if (outCounts4)
    {
        outCounts4[0] = (int)buffer->GetNumTriangles();
        outCounts4[1] = (int)r->result_.Size();
        outCounts4[2] = nvis;
        outCounts4[3] = (int)active;
    }
    return nvis;
}

}
