/* Stand-in for the export header that the reference's CMake normally generates
 * (Source/Urho3D/CMakeLists.txt writes Urho3D.h; it is included from
 * Container/HashBase.h:28).  Static, headless, GL-flavoured, threading on.
 * This file is OUR build glue for oracle/_ref; it contains no reference code. */
#ifndef URHO3D_API_H
#define URHO3D_API_H
#define URHO3D_STATIC_DEFINE
#define URHO3D_API
#define URHO3D_NO_EXPORT
#define URHO3D_DEPRECATED
#define URHO3D_OPENGL
#define URHO3D_THREADING
#define URHO3D_LOGGING
#define URHO3D_SSE
#endif
