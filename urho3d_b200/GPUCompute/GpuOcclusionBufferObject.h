// GpuOcclusionBufferObject -- the type Renderer's occlusion-buffer pool holds in a GPU build (Renderer.cpp:253, 1103-1121, 1533-1542):
// an Urho3D::Object with EXACTLY the public member signatures of class OcclusionBuffer (OcclusionBuffer.h:101-158), so that
// View::DrawOccluders / View::GetDrawables / CheckVisibilityWork and every Drawable::DrawOcclusion override compile against it unchanged
// (tests/cpp/test_engine_types.cpp checks each signature at compile time and runs the same call-site template on both classes).
// Engine-only header: include it after the engine's Camera.h / OcclusionBuffer.h (for CullMode, Matrix3x4, BoundingBox, Object).
//
// Not carried over: DrawBatch(const OcclusionBatch&, unsigned threadIndex) -- "called internally" by the CPU class's worker threads
// (OcclusionBuffer.cpp:47-52); the GPU class draws every queued batch in DrawTriangles().
#pragma once

#ifndef URHO3D_API
#error "GpuOcclusionBufferObject.h is the in-engine wrapper: include it after Urho3D's own headers"
#endif

#include "GPUCompute.h"

namespace Urho3D
{

class GpuOcclusionBufferObject : public Object
{
    URHO3D_OBJECT(GpuOcclusionBufferObject, Object);

public:
    /// Construct on the engine's GPU context (one per device).
    GpuOcclusionBufferObject(Context* context, GPUCompute::GpuContext& gpu) : Object(context), buffer_(gpu) {}

    /// Set occlusion buffer size.  `threaded` is kept for IsThreaded(); the device path always behaves like the threaded branch (submit all, draw once).
    bool SetSize(int width, int height, bool threaded) { return buffer_.SetSize(width, height, threaded); }
    /// Set camera view to render from (OcclusionBuffer.cpp:115-127).
    void SetView(Camera* camera)
    {
        if (!camera)
            return;
        view_ = camera->GetView();
        projection_ = camera->GetProjection();
        buffer_.SetView(camera);
    }
    /// Set maximum triangles to render.
    void SetMaxTriangles(unsigned triangles) { buffer_.SetMaxTriangles(triangles); }
    /// Set culling mode.
    void SetCullMode(CullMode mode) { buffer_.SetCullMode(mode); }
    /// Reset number of triangles.
    void Reset() { buffer_.Reset(); }
    /// Clear the buffer.
    void Clear() { buffer_.Clear(); }
    /// Submit non-indexed geometry.  Return true if did not overflow the allowed triangle count.
    bool AddTriangles(const Matrix3x4& model, const void* vertexData, unsigned vertexSize, unsigned vertexStart, unsigned vertexCount)
    {
        return buffer_.AddTriangles(model, vertexData, vertexSize, vertexStart, vertexCount);
    }
    /// Submit indexed geometry.  Return true if did not overflow the allowed triangle count.
    bool AddTriangles(const Matrix3x4& model, const void* vertexData, unsigned vertexSize, const void* indexData, unsigned indexSize,
        unsigned indexStart, unsigned indexCount)
    {
        return buffer_.AddTriangles(model, vertexData, vertexSize, indexData, indexSize, indexStart, indexCount);
    }
    /// Draw submitted batches.
    void DrawTriangles() { buffer_.DrawTriangles(); }
    /// Build reduced size mip levels.
    void BuildDepthHierarchy() { buffer_.BuildDepthHierarchy(); }
    /// Reset last used timer.
    void ResetUseTimer() { buffer_.ResetUseTimer(); }

    /// Return highest level depth values (a host copy refreshed by this call).
    int* GetBuffer() const { return buffer_.GetBuffer(); }
    /// Return view transform matrix.
    const Matrix3x4& GetView() const { return view_; }
    /// Return projection matrix.
    const Matrix4& GetProjection() const { return projection_; }
    /// Return buffer width.
    int GetWidth() const { return buffer_.GetWidth(); }
    /// Return buffer height.
    int GetHeight() const { return buffer_.GetHeight(); }
    /// Return number of rendered triangles.
    unsigned GetNumTriangles() const { return buffer_.GetNumTriangles(); }
    /// Return maximum number of triangles.
    unsigned GetMaxTriangles() const { return buffer_.GetMaxTriangles(); }
    /// Return culling mode.
    CullMode GetCullMode() const { return (CullMode)buffer_.GetCullMode(); }
    /// Return whether the threaded branch of View::DrawOccluders is taken (View.cpp:2236).
    bool IsThreaded() const { return buffer_.IsThreaded(); }
    /// Test a bounding box for visibility.
    bool IsVisible(const BoundingBox& worldSpaceBox) const { return buffer_.IsVisible(worldSpaceBox); }
    /// Return time since last use in milliseconds.
    unsigned GetUseTimer() { return buffer_.GetUseTimer(); }

    /// The device object, for the GPU query classes (GPUCompute::MakeOccludedFrustumQuery) and the batched entry point.
    GPUCompute::GpuOcclusionBuffer* Gpu() { return &buffer_; }

private:
    mutable GPUCompute::GpuOcclusionBuffer buffer_;
    Matrix3x4 view_;
    Matrix4 projection_;
};

} // namespace Urho3D
