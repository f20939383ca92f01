/* Minimal SDL_config.h so that <SDL/SDL.h> parses when the reference's Core/IO
 * sources are compiled for the CPU oracle (only SDL_config.h.cmake ships).
 * No SDL code is built or linked; the few SDL symbols referenced are stubbed
 * in ref_stubs.cpp.  OUR build glue, not reference code. */
#ifndef SDL_config_h_
#define SDL_config_h_
#include "SDL_platform.h"
#define SIZEOF_VOIDP 8
#define HAVE_GCC_ATOMICS 1
#define HAVE_LIBC 1
#define STDC_HEADERS 1
#define HAVE_ALLOCA_H 1
#define HAVE_CTYPE_H 1
#define HAVE_FLOAT_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_LIMITS_H 1
#define HAVE_MATH_H 1
#define HAVE_MEMORY_H 1
#define HAVE_SIGNAL_H 1
#define HAVE_STDARG_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDIO_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_STRINGS_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_WCHAR_H 1
#define HAVE_MALLOC 1
#define HAVE_CALLOC 1
#define HAVE_REALLOC 1
#define HAVE_FREE 1
#define HAVE_ALLOCA 1
#define SDL_AUDIO_DISABLED 1
#define SDL_JOYSTICK_DISABLED 1
#define SDL_HAPTIC_DISABLED 1
#define SDL_SENSOR_DISABLED 1
#define SDL_LOADSO_DISABLED 1
#define SDL_THREADS_DISABLED 1
#define SDL_TIMERS_DISABLED 1
#define SDL_VIDEO_DISABLED 1
#define SDL_FILESYSTEM_DUMMY 1
#endif
