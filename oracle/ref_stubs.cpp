// TEST INFRASTRUCTURE: link-time stand-ins for symbols that the reference's Core/IO/Resource/Scene/Graphics
// translation units mention but the visibility path never calls (SDL window/message-box, debug drawing,
// ETC texture decode, GL vertex-buffer helpers).  Every one aborts if it is ever reached, so a silent
// behavioural substitution is impossible.  OUR glue for oracle/_ref; contains no reference code.
#include <cstdio>
#include <cstdlib>
#include <Urho3D/Graphics/DebugRenderer.h>
#include <Urho3D/Graphics/Geometry.h>
#include <Urho3D/Graphics/Light.h>
#include <Urho3D/Graphics/VertexBuffer.h>

#define U3D_STUB_DIE() do { fprintf(stderr, "oracle/_ref stub reached: %s\n", __func__); abort(); } while (0)

extern "C"
{
struct SDL_Surface;
int SDL_Init(unsigned) { return 0; }
int SDL_InitSubSystem(unsigned) { return 0; }
void SDL_Quit() {}
void SDL_QuitSubSystem(unsigned) {}
unsigned SDL_WasInit(unsigned) { return 0; }
const char* SDL_GetError() { return ""; }
void SDL_free(void* p) { free(p); }
char* SDL_GetPrefPath(const char*, const char*) { return nullptr; }
int SDL_ShowSimpleMessageBox(unsigned, const char*, const char*, void*) { return 0; }
SDL_Surface* SDL_CreateRGBSurface(unsigned, int, int, int, unsigned, unsigned, unsigned, unsigned) { U3D_STUB_DIE(); }
int SDL_LockSurface(SDL_Surface*) { U3D_STUB_DIE(); }
void SDL_UnlockSurface(SDL_Surface*) { U3D_STUB_DIE(); }
}

void decompressBlockAlphaC(unsigned char*, unsigned char*, int, int, int, int, int) { U3D_STUB_DIE(); }
void decompressBlockETC2c(unsigned int, unsigned int, unsigned char*, int, int, int, int, int) { U3D_STUB_DIE(); }
void setupAlphaTable() {}

namespace Urho3D
{
void DebugRenderer::AddBoundingBox(const BoundingBox&, const Color&, bool, bool) { U3D_STUB_DIE(); }
void DebugRenderer::AddFrustum(const Frustum&, const Color&, bool) { U3D_STUB_DIE(); }
void DebugRenderer::AddLine(const Vector3&, const Vector3&, const Color&, bool) { U3D_STUB_DIE(); }
void DebugRenderer::AddNode(Node*, float, bool) { U3D_STUB_DIE(); }
bool DebugRenderer::IsInside(const BoundingBox&) const { U3D_STUB_DIE(); }
void Geometry::GetRawData(const unsigned char*&, unsigned&, const unsigned char*&, unsigned&, const PODVector<VertexElement>*&) const { U3D_STUB_DIE(); }
void Light::SetIntensitySortValue(const BoundingBox&) { U3D_STUB_DIE(); }
Color Light::GetEffectiveColor() const { U3D_STUB_DIE(); } // CheckVisibilityWork's light branch: the harness has no Light drawables
unsigned VertexBuffer::GetElementOffset(const PODVector<VertexElement>&, VertexElementType, VertexElementSemantic, unsigned char) { U3D_STUB_DIE(); }
bool VertexBuffer::HasElement(const PODVector<VertexElement>&, VertexElementType, VertexElementSemantic, unsigned char) { U3D_STUB_DIE(); }
}
