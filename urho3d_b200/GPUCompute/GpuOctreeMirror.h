// GpuOctreeMirror -- in-engine face of the GPU octree: keeps the Drawable* <-> id table next to GPUCompute::GpuOctree and answers the query
// classes of View.cpp into a PODVector<Drawable*>, cleared first and filled in the reference's result_ order, exactly what
// Octree::GetDrawables(OctreeQuery&) leaves behind (Octree.cpp:495-499).  Engine-only header (needs Drawable.h, Frustum.h, PODVector).
//
// The engine's own Octree stays the owner of the structure: a Drawable is mirrored where Octree::InsertDrawable /
// Octree::Update / Octree::RemoveDrawable handle it (Octree.cpp:134-166, 440-472; Octree.h:56-136).
#pragma once

#ifndef URHO3D_API
#error "GpuOctreeMirror.h is the in-engine wrapper: include it after Urho3D's own headers"
#endif

#include "GPUCompute.h"
#include "GpuOcclusionBufferObject.h"

namespace Urho3D
{

class GpuOctreeMirror
{
public:
    /// Same bounds and level count as the Octree it mirrors (Octree::SetSize, Octree.cpp:344-353).
    GpuOctreeMirror(GPUCompute::GpuContext& gpu, const BoundingBox& box, unsigned numLevels) :
        octree_(gpu, box.min_.Data(), box.max_.Data(), numLevels)
    {
    }

    /// Octree::InsertDrawable.  Returns the drawable's id in the mirror.
    unsigned InsertDrawable(Drawable* drawable)
    {
        const unsigned id = GPUCompute::InsertDrawable(octree_, drawable);
        if (id >= drawables_.Size())
            drawables_.Resize(id + 1, nullptr);
        drawables_[id] = drawable;
        return id;
    }
    /// The drawable moved or was resized (what Octree::Update's reinsertion loop handles).
    void UpdateDrawable(unsigned id)
    {
        float box[6];
        GPUCompute::BoxOf(drawables_[id]->GetWorldBoundingBox(), box);
        octree_.UpdateDrawable(id, box);
    }
    void RemoveDrawable(unsigned id)
    {
        octree_.RemoveDrawable(id);
        drawables_[id] = nullptr;
    }

    /// FrustumOctreeQuery(result, frustum, drawableFlags, viewMask) + Octree::GetDrawables.
    void GetDrawables(PODVector<Drawable*>& result, const Frustum& frustum, unsigned char drawableFlags = DRAWABLE_ANY, unsigned viewMask = DEFAULT_VIEWMASK)
    {
        GPUCompute::GpuFrustumOctreeQuery query = GPUCompute::MakeFrustumQuery(ids_, frustum, drawableFlags, viewMask);
        octree_.GetDrawables(query);
        Resolve(result);
    }
    /// OccludedFrustumOctreeQuery(result, frustum, buffer, drawableFlags, viewMask) (View.cpp:116-158) + Octree::GetDrawables.
    void GetDrawables(PODVector<Drawable*>& result, const Frustum& frustum, GpuOcclusionBufferObject* buffer, unsigned char drawableFlags = DRAWABLE_ANY,
        unsigned viewMask = DEFAULT_VIEWMASK)
    {
        GPUCompute::GpuOccludedFrustumOctreeQuery query =
            GPUCompute::MakeOccludedFrustumQuery(ids_, frustum, buffer ? buffer->Gpu() : nullptr, drawableFlags, viewMask);
        octree_.GetDrawables(query);
        Resolve(result);
    }
    /// The same query followed by CheckVisibilityWork's occlusion predicate (View.cpp:177) in one pass: the list View::GetDrawables keeps.
    void GetVisibleDrawables(PODVector<Drawable*>& result, const Frustum& frustum, GpuOcclusionBufferObject* buffer,
        unsigned char drawableFlags = DRAWABLE_ANY, unsigned viewMask = DEFAULT_VIEWMASK)
    {
        GPUCompute::GpuOccludedFrustumOctreeQuery query =
            GPUCompute::MakeOccludedFrustumQuery(ids_, frustum, buffer ? buffer->Gpu() : nullptr, drawableFlags, viewMask);
        octree_.GetVisibleDrawables(query);
        Resolve(result);
    }
    /// ShadowCasterOctreeQuery(result, frustum, drawableFlags, viewMask) (View.cpp:59-84) + Octree::GetDrawables.
    void GetShadowCasters(PODVector<Drawable*>& result, const Frustum& frustum, unsigned char drawableFlags = DRAWABLE_ANY, unsigned viewMask = DEFAULT_VIEWMASK)
    {
        GPUCompute::GpuShadowCasterOctreeQuery query = GPUCompute::MakeShadowCasterQuery(ids_, frustum, drawableFlags, viewMask);
        octree_.GetDrawables(query);
        Resolve(result);
    }

    Drawable* GetDrawable(unsigned id) const { return id < drawables_.Size() ? drawables_[id] : nullptr; }
    GPUCompute::GpuOctree& Gpu() { return octree_; }

private:
    void Resolve(PODVector<Drawable*>& result)
    {
        result.Clear();
        for (size_t i = 0; i < ids_.size(); ++i)
            result.Push(drawables_[ids_[i]]);
    }

    GPUCompute::GpuOctree octree_;
    PODVector<Drawable*> drawables_;
    std::vector<unsigned> ids_;
};

} // namespace Urho3D
