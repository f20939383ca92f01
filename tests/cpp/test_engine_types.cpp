// TEST: the in-engine side of urho3d_b200/GPUCompute/GPUCompute.h (the URHO3D_API branch) compiled against the reference's OWN headers and
// linked with the reference's own objects (oracle/_ref/libu3dref.so) and with libu3dgpu.so.  The same call-site code -- written like
// View::DrawOccluders' threaded branch (View.cpp:2258-2273) around StaticModel::DrawOcclusion (StaticModel.cpp:189-228) and like
// View::GetDrawables (View.cpp:868-878) -- is instantiated once with the reference's OcclusionBuffer and once with GpuOcclusionBuffer.
//
// Without a device (the CPU suite) the program checks the host-only helpers (MakeView, PlanesOf, BoxOf against the engine objects), runs the
// call-site code on the reference class, and expects GpuContext to fail loudly.  With a device it also runs the call-site code on the GPU
// class and compares depth plane, numTriangles_ and IsVisible answers with the reference's, in the same process.
// Built only where /root/reference exists (tests/test_engine_types.py); prints "OK" on success.
#include <Urho3D/Urho3D.h>

#include <Urho3D/Core/Context.h>
#include <Urho3D/Core/WorkQueue.h>
#include <Urho3D/Graphics/Camera.h>
#include <Urho3D/Graphics/Drawable.h>
#include <Urho3D/Graphics/OcclusionBuffer.h>
#include <Urho3D/Graphics/OctreeQuery.h>
#include <Urho3D/Scene/Node.h>
#include <Urho3D/Scene/Scene.h>

#include "../../urho3d_b200/GPUCompute/GPUCompute.h"
#include "../../urho3d_b200/GPUCompute/GpuOcclusionBufferObject.h"
#include "../../urho3d_b200/GPUCompute/GpuOctreeMirror.h"
#include <Urho3D/Graphics/Octree.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace Urho3D;

#define REQUIRE(cond)                                                          \
    do                                                                         \
    {                                                                          \
        if (!(cond))                                                           \
        {                                                                      \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            std::exit(1);                                                      \
        }                                                                      \
    } while (0)

namespace
{

// Box.mdl-like occluder: 8 corners, 12 triangles, 16-bit indices, 12-byte stride.
const float kBoxVertices[8 * 3] = {-0.5f, -0.5f, -0.5f, 0.5f, -0.5f, -0.5f, 0.5f, 0.5f, -0.5f, -0.5f, 0.5f, -0.5f,
    -0.5f, -0.5f, 0.5f, 0.5f, -0.5f, 0.5f, 0.5f, 0.5f, 0.5f, -0.5f, 0.5f, 0.5f};
const unsigned short kBoxIndices[36] = {0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7, 0, 1, 5, 0, 5, 4, 3, 6, 2, 3, 7, 6, 0, 4, 7, 0, 7, 3, 1, 2, 6, 1, 6, 5};

/// A Drawable whose world box is handed in (no Model / GL needed); inserted into the real Octree by Node::AddComponent -> AddToOctree.
class BoxDrawable : public Drawable
{
    URHO3D_OBJECT(BoxDrawable, Drawable);

public:
    explicit BoxDrawable(Context* context) : Drawable(context, DRAWABLE_GEOMETRY) {}
    void Setup(const BoundingBox& worldBox, unsigned char flags, unsigned viewMask, bool occludee, bool castShadows)
    {
        fixedWorldBox_ = worldBox;
        drawableFlags_ = flags;
        viewMask_ = viewMask;
        occludee_ = occludee;
        castShadows_ = castShadows;
        worldBoundingBoxDirty_ = true;
    }
    void OnWorldBoundingBoxUpdate() override { worldBoundingBox_ = fixedWorldBox_; }
    BoundingBox fixedWorldBox_;
};

/// OccludedFrustumOctreeQuery and ShadowCasterOctreeQuery are file-local in View.cpp (116-158, 59-84): restated on the real base class so that
/// the real Octree can run them.
class OccludedQuery : public FrustumOctreeQuery
{
public:
    OccludedQuery(PODVector<Drawable*>& result, const Frustum& frustum, OcclusionBuffer* buffer, unsigned char flags, unsigned mask) :
        FrustumOctreeQuery(result, frustum, flags, mask), buffer_(buffer) {}
    Intersection TestOctant(const BoundingBox& box, bool inside) override
    {
        if (inside)
            return buffer_->IsVisible(box) ? INSIDE : OUTSIDE;
        Intersection result = frustum_.IsInside(box);
        if (result != OUTSIDE && !buffer_->IsVisible(box))
            result = OUTSIDE;
        return result;
    }
    OcclusionBuffer* buffer_;
};

class ShadowCasterQuery : public FrustumOctreeQuery
{
public:
    ShadowCasterQuery(PODVector<Drawable*>& result, const Frustum& frustum, unsigned char flags, unsigned mask) :
        FrustumOctreeQuery(result, frustum, flags, mask) {}
    void TestDrawables(Drawable** start, Drawable** end, bool inside) override
    {
        while (start != end)
        {
            Drawable* drawable = *start++;
            if (drawable->GetCastShadows() && (drawable->GetDrawableFlags() & drawableFlags_) && (drawable->GetViewMask() & viewMask_) &&
                (inside || frustum_.IsInsideFast(drawable->GetWorldBoundingBox())))
                result_.Push(drawable);
        }
    }
};

bool SameSequence(const PODVector<Drawable*>& a, const PODVector<Drawable*>& b)
{
    if (a.Size() != b.Size())
        return false;
    for (unsigned i = 0; i < a.Size(); ++i)
        if (a[i] != b[i])
            return false;
    return true;
}

/// The call-site code, generic in the buffer class: what View::DrawOccluders (threaded branch) and the occluders' DrawOcclusion do.
template <class Buffer> bool DrawOccluders(Buffer* buffer, Camera* camera, const Matrix3x4* worldTransforms, unsigned numOccluders, unsigned maxTriangles)
{
    buffer->SetView(camera);
    buffer->SetMaxTriangles(maxTriangles);
    buffer->Clear();
    bool withinBudget = true;
    for (unsigned i = 0; i < numOccluders; ++i)
    {
        buffer->SetCullMode(i & 1 ? CULL_CW : CULL_CCW);
        if (!buffer->AddTriangles(worldTransforms[i], kBoxVertices, 3 * sizeof(float), kBoxIndices, sizeof(unsigned short), 0, 36))
            withinBudget = false;
    }
    // non-indexed submission as well (OcclusionBuffer.h:113)
    buffer->SetCullMode(CULL_NONE);
    buffer->AddTriangles(worldTransforms[0], kBoxVertices, 3 * sizeof(float), 0, 6);
    buffer->DrawTriangles();
    buffer->BuildDepthHierarchy();
    buffer->ResetUseTimer();
    return withinBudget;
}

template <class Buffer> void Answers(Buffer* buffer, const BoundingBox* boxes, unsigned count, std::vector<unsigned char>& out)
{
    out.resize(count);
    for (unsigned i = 0; i < count; ++i)
        out[i] = buffer->IsVisible(boxes[i]) ? 1 : 0;
}

// Compile-time proof that GpuOcclusionBufferObject has OcclusionBuffer's public member signatures (OcclusionBuffer.h:101-158): taking the address
// of an (overloaded) member with a static_cast to the exact pointer-to-member type only compiles when such a member exists.
#define SAME_MEMBER(ret, name, args, cv)                                                            \
    (void)static_cast<ret(OcclusionBuffer::*) args cv>(&OcclusionBuffer::name);                     \
    (void)static_cast<ret(GpuOcclusionBufferObject::*) args cv>(&GpuOcclusionBufferObject::name)

void SignatureParity()
{
    SAME_MEMBER(bool, SetSize, (int, int, bool), );
    SAME_MEMBER(void, SetView, (Camera*), );
    SAME_MEMBER(void, SetMaxTriangles, (unsigned), );
    SAME_MEMBER(void, SetCullMode, (CullMode), );
    SAME_MEMBER(void, Reset, (), );
    SAME_MEMBER(void, Clear, (), );
    SAME_MEMBER(bool, AddTriangles, (const Matrix3x4&, const void*, unsigned, unsigned, unsigned), );
    SAME_MEMBER(bool, AddTriangles, (const Matrix3x4&, const void*, unsigned, const void*, unsigned, unsigned, unsigned), );
    SAME_MEMBER(void, DrawTriangles, (), );
    SAME_MEMBER(void, BuildDepthHierarchy, (), );
    SAME_MEMBER(void, ResetUseTimer, (), );
    SAME_MEMBER(int*, GetBuffer, (), const);
    SAME_MEMBER(const Matrix3x4&, GetView, (), const);
    SAME_MEMBER(const Matrix4&, GetProjection, (), const);
    SAME_MEMBER(int, GetWidth, (), const);
    SAME_MEMBER(int, GetHeight, (), const);
    SAME_MEMBER(unsigned, GetNumTriangles, (), const);
    SAME_MEMBER(unsigned, GetMaxTriangles, (), const);
    SAME_MEMBER(CullMode, GetCullMode, (), const);
    SAME_MEMBER(bool, IsThreaded, (), const);
    SAME_MEMBER(bool, IsVisible, (const BoundingBox&), const);
    SAME_MEMBER(unsigned, GetUseTimer, (), );
}

} // namespace

int main()
{
    SignatureParity();
    SharedPtr<Context> context(new Context());
    context->RegisterSubsystem(new WorkQueue(context));
    RegisterSceneLibrary(context);
    Camera::RegisterObject(context);

    SharedPtr<Node> cameraNode(new Node(context));
    Camera* camera = cameraNode->CreateComponent<Camera>();
    cameraNode->SetTransform(Vector3(1.5f, 2.0f, -12.0f), Quaternion(3.0f, 6.0f, 0.0f));
    camera->SetFov(50.0f);
    camera->SetNearClip(0.1f);
    camera->SetFarClip(300.0f);
    camera->SetAspectRatio(256.0f / 160.0f);

    // ---- host-only helpers against the engine objects ----
    const u3d_view v = GPUCompute::MakeView(camera, DRAWABLE_GEOMETRY | DRAWABLE_LIGHT, camera->GetViewMask());
    REQUIRE(std::memcmp(v.view, camera->GetView().Data(), sizeof(v.view)) == 0);
    const Matrix4 projection = camera->GetProjection();
    REQUIRE(std::memcmp(v.projection, projection.Data(), sizeof(v.projection)) == 0);
    const Frustum& frustum = camera->GetFrustum();
    for (unsigned i = 0; i < NUM_FRUSTUM_PLANES; ++i)
    {
        const Vector4 p = frustum.planes_[i].ToVector4();
        REQUIRE(std::memcmp(&v.planes[4 * i], p.Data(), 4 * sizeof(float)) == 0);
    }
    REQUIRE(v.near_clip == camera->GetNearClip() && v.far_clip == camera->GetFarClip());
    REQUIRE(v.drawable_flags == (DRAWABLE_GEOMETRY | DRAWABLE_LIGHT) && v.view_mask == DEFAULT_VIEWMASK);
    REQUIRE(v.reverse_culling == 0 && v.orthographic == 0 && v.use_occlusion == 1 && v.query_mode == U3D_QUERY_OCCLUDED);
    REQUIRE(Vector3(v.camera_position) == cameraNode->GetWorldPosition());
    // the view a reference OcclusionBuffer captures from the same camera (OcclusionBuffer.cpp:115-127) is the one MakeView hands over
    SharedPtr<OcclusionBuffer> reference(new OcclusionBuffer(context));
    REQUIRE(reference->SetSize(256, 160, false));
    reference->SetView(camera);
    REQUIRE(std::memcmp(reference->GetView().Data(), v.view, sizeof(v.view)) == 0);
    REQUIRE(std::memcmp(reference->GetProjection().Data(), v.projection, sizeof(v.projection)) == 0);
    float box6[6];
    GPUCompute::BoxOf(BoundingBox(Vector3(-1.0f, -2.0f, -3.0f), Vector3(4.0f, 5.0f, 6.0f)), box6);
    REQUIRE(box6[0] == -1.0f && box6[2] == -3.0f && box6[3] == 4.0f && box6[5] == 6.0f);
    // queries with the reference's constructor arguments
    std::vector<unsigned> ids;
    PODVector<Drawable*> drawables;
    FrustumOctreeQuery referenceQuery(drawables, frustum, DRAWABLE_GEOMETRY, 0x7u);
    GPUCompute::GpuFrustumOctreeQuery gpuQuery = GPUCompute::MakeFrustumQuery(ids, frustum, DRAWABLE_GEOMETRY, 0x7u);
    REQUIRE(gpuQuery.drawableFlags_ == referenceQuery.drawableFlags_ && gpuQuery.viewMask_ == referenceQuery.viewMask_);
    for (unsigned i = 0; i < NUM_FRUSTUM_PLANES; ++i)
        REQUIRE(gpuQuery.planes_[4 * i + 3] == referenceQuery.frustum_.planes_[i].d_ && gpuQuery.planes_[4 * i] == referenceQuery.frustum_.planes_[i].normal_.x_);
    REQUIRE(GPUCompute::MakeShadowCasterQuery(ids, frustum).Mode() == U3D_QUERY_SHADOWCASTER);
    REQUIRE(GPUCompute::MakeZoneOccluderQuery(ids, frustum).Mode() == U3D_QUERY_ZONE_OCCLUDER);
    REQUIRE(GPUCompute::MakeOccludedFrustumQuery(ids, frustum, nullptr).Mode() == U3D_QUERY_OCCLUDED);

    // multi-GPU host logic is host code as well: 10 views of very different cost over 2 ranks
    {
        const std::vector<double> costs = {100.0, 1.0, 1.0, 1.0, 1.0, 96.0, 1.0, 1.0, 1.0, 1.0};
        const std::vector<unsigned> owner = GPUCompute::BalanceViews(costs, 10, 1, 2);
        unsigned held[2] = {0, 0};
        double load[2] = {0.0, 0.0};
        for (unsigned i = 0; i < 10; ++i)
        {
            REQUIRE(owner[i] < 2);
            ++held[owner[i]];
            load[owner[i]] += costs[i];
        }
        REQUIRE(held[0] == 5 && held[1] == 5 && owner[0] != owner[5] && load[0] + load[1] == 204.0 && load[0] >= 100.0 && load[1] >= 100.0);
    }

    // ---- the call-site code on the reference class ----
    const unsigned numOccluders = 24;
    std::vector<Matrix3x4> transforms;
    for (unsigned i = 0; i < numOccluders; ++i)
        transforms.push_back(Matrix3x4(Vector3(-14.0f + 2.6f * (float)(i % 12), 1.0f + 2.5f * (float)(i / 12), 6.0f + 1.7f * (float)(i % 5)),
            Quaternion(0.0f, 6.0f * ((float)(i % 7) - 3.0f), 0.0f), Vector3(2.0f + 0.1f * (float)i, 2.5f, 0.4f)));
    std::vector<BoundingBox> boxes;
    for (int i = 0; i < 200; ++i)
    {
        const Vector3 c(-16.0f + 0.17f * (float)i, 0.5f + 0.03f * (float)(i % 40), i % 3 ? 18.0f + 0.23f * (float)((i * 7) % 100) : -6.0f + 0.11f * (float)(i % 90));
        boxes.push_back(BoundingBox(c - Vector3(0.4f, 0.4f, 0.4f), c + Vector3(0.4f, 0.4f, 0.4f)));
    }
    const bool referenceBudget = DrawOccluders(reference.Get(), camera, transforms.data(), numOccluders, 250);
    REQUIRE(!referenceBudget); // 24 x 12 + 2 triangles against a budget of 250: AddTriangles reports it, the batches stay queued
    std::vector<unsigned char> referenceAnswers;
    Answers(reference.Get(), boxes.data(), (unsigned)boxes.size(), referenceAnswers);
    unsigned visible = 0;
    for (unsigned char a : referenceAnswers)
        visible += a;
    std::printf("reference: %u triangles, %u of %u boxes visible\n", reference->GetNumTriangles(), visible, (unsigned)boxes.size());
    REQUIRE(visible > 10 && visible < boxes.size() - 10); // the scene exercises both answers

    // ---- the real Octree with real Drawables: the lists Octree::GetDrawables gives for the query classes of View.cpp ----
    Octree::RegisterObject(context);
    SharedPtr<Scene> scene(new Scene(context));
    Octree* octree = scene->CreateComponent<Octree>();
    PODVector<Drawable*> sceneDrawables;
    unsigned seed = 12345u;
    auto rnd = [&seed]() { seed = seed * 1664525u + 1013904223u; return (float)(seed >> 8) / 16777216.0f; };
    for (unsigned i = 0; i < 1500; ++i)
    {
        const Vector3 c(-70.0f + 140.0f * rnd(), 0.5f + 3.0f * rnd(), -30.0f + 140.0f * rnd());
        const Vector3 e(0.3f + 1.5f * rnd(), 0.3f + 1.5f * rnd(), 0.3f + 1.5f * rnd());
        BoxDrawable* d = new BoxDrawable(context);
        d->Setup(BoundingBox(c - e, c + e), i % 11 == 0 ? DRAWABLE_LIGHT : DRAWABLE_GEOMETRY, i % 7 == 0 ? 0x2u : DEFAULT_VIEWMASK, i % 50 != 0, i % 2 == 0);
        scene->CreateChild(String::EMPTY, LOCAL)->AddComponent(d, 0, LOCAL); // -> Octree::InsertDrawable
        sceneDrawables.Push(d);
    }
    PODVector<Drawable*> wantFrustum[2], wantShadow[2], wantOccluded[2], wantVisible[2];
    for (unsigned pass = 0; pass < 2; ++pass)
    {
        const unsigned char flags = pass ? (unsigned char)DRAWABLE_GEOMETRY : (unsigned char)DRAWABLE_ANY;
        const unsigned mask = pass ? 0x1u : DEFAULT_VIEWMASK;
        FrustumOctreeQuery frustumQuery(wantFrustum[pass], frustum, flags, mask);
        octree->GetDrawables(frustumQuery);
        ShadowCasterQuery shadowQuery(wantShadow[pass], frustum, flags, mask);
        octree->GetDrawables(shadowQuery);
        OccludedQuery occludedQuery(wantOccluded[pass], frustum, reference.Get(), flags, mask);
        octree->GetDrawables(occludedQuery);
        for (unsigned i = 0; i < wantOccluded[pass].Size(); ++i) // CheckVisibilityWork's predicate over the query result (View.cpp:177)
            if (!wantOccluded[pass][i]->IsOccludee() || reference->IsVisible(wantOccluded[pass][i]->GetWorldBoundingBox()))
                wantVisible[pass].Push(wantOccluded[pass][i]);
        std::printf("real Octree, pass %u: frustum %u, shadow casters %u, occluded query %u, visible %u\n", pass, wantFrustum[pass].Size(),
            wantShadow[pass].Size(), wantOccluded[pass].Size(), wantVisible[pass].Size());
        REQUIRE(wantFrustum[pass].Size() > 20 && wantShadow[pass].Size() > 10 && wantOccluded[pass].Size() > 10);
        REQUIRE(wantOccluded[pass].Size() < wantFrustum[pass].Size() && wantVisible[pass].Size() < wantOccluded[pass].Size());
    }

    // ---- the GPU class: loud failure without a device, bit parity with one ----
    try
    {
        GPUCompute::GpuContext gpu(0);
        GPUCompute::GpuOcclusionBuffer buffer(gpu);
        REQUIRE(buffer.SetSize(256, 160, true));
        REQUIRE(!buffer.SetSize(100, 64, true)); // non power-of-two width is refused, OcclusionBuffer.cpp:70-77
        REQUIRE(buffer.SetSize(256, 160, true));
        const bool gpuBudget = DrawOccluders(&buffer, camera, transforms.data(), numOccluders, 250);
        REQUIRE(gpuBudget == referenceBudget);
        REQUIRE(buffer.GetNumTriangles() == reference->GetNumTriangles());
        REQUIRE(buffer.GetWidth() == reference->GetWidth() && buffer.GetHeight() == reference->GetHeight());
        const int* gpuDepth = buffer.GetBuffer();
        REQUIRE(gpuDepth && std::memcmp(gpuDepth, reference->GetBuffer(), sizeof(int) * (size_t)reference->GetWidth() * reference->GetHeight()) == 0);
        std::vector<unsigned char> gpuAnswers;
        Answers(&buffer, boxes.data(), (unsigned)boxes.size(), gpuAnswers);
        REQUIRE(gpuAnswers == referenceAnswers);
        // the Object wrapper Renderer's pool would hold: same template, same answers
        SharedPtr<GpuOcclusionBufferObject> pooled(new GpuOcclusionBufferObject(context, gpu));
        REQUIRE(pooled->SetSize(256, 160, true) && pooled->IsThreaded());
        REQUIRE(DrawOccluders(pooled.Get(), camera, transforms.data(), numOccluders, 250) == referenceBudget);
        REQUIRE(pooled->GetNumTriangles() == reference->GetNumTriangles() && pooled->GetMaxTriangles() == 250 && pooled->GetCullMode() == CULL_NONE);
        REQUIRE(pooled->GetView().Equals(reference->GetView()) && pooled->GetProjection().Equals(reference->GetProjection()));
        REQUIRE(std::memcmp(pooled->GetBuffer(), reference->GetBuffer(), sizeof(int) * (size_t)reference->GetWidth() * reference->GetHeight()) == 0);
        Answers(pooled.Get(), boxes.data(), (unsigned)boxes.size(), gpuAnswers);
        REQUIRE(gpuAnswers == referenceAnswers);
        // ---- Octree::GetDrawables: the mirror against the lists the real Octree gave above, as PODVector<Drawable*> in sequence ----
        GpuOctreeMirror mirror(gpu, BoundingBox(-1000.0f, 1000.0f), 8); // Octree's defaults, Octree.cpp:46-47
        for (unsigned i = 0; i < sceneDrawables.Size(); ++i)
            mirror.InsertDrawable(sceneDrawables[i]);
        PODVector<Drawable*> got;
        unsigned compared = 0;
        for (unsigned pass = 0; pass < 2; ++pass)
        {
            const unsigned char flags = pass ? (unsigned char)DRAWABLE_GEOMETRY : (unsigned char)DRAWABLE_ANY;
            const unsigned mask = pass ? 0x1u : DEFAULT_VIEWMASK;
            mirror.GetDrawables(got, frustum, flags, mask);
            REQUIRE(SameSequence(wantFrustum[pass], got));
            mirror.GetShadowCasters(got, frustum, flags, mask);
            REQUIRE(SameSequence(wantShadow[pass], got));
            mirror.GetDrawables(got, frustum, pooled.Get(), flags, mask);
            REQUIRE(SameSequence(wantOccluded[pass], got));
            mirror.GetVisibleDrawables(got, frustum, pooled.Get(), flags, mask);
            REQUIRE(SameSequence(wantVisible[pass], got));
            compared += wantFrustum[pass].Size() + wantShadow[pass].Size() + wantOccluded[pass].Size() + wantVisible[pass].Size();
        }
        std::printf("Octree::GetDrawables compared with the GPU mirror in-process: %u drawables over the query classes\n", compared);
        std::printf("device path compared with the reference OcclusionBuffer in-process: %u triangles, %u of %u boxes visible\n",
            buffer.GetNumTriangles(), visible, (unsigned)boxes.size());
    }
    catch (const GPUCompute::GpuError& e)
    {
        if (std::getenv("U3D_EXPECT_DEVICE"))
        {
            std::printf("FAILED: a device was expected: %s\n", e.what());
            return 1;
        }
        std::printf("no device: GpuContext failed loudly (%s)\n", e.what());
    }
    std::printf("OK\n");
    return 0;
}
