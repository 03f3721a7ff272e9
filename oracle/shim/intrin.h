// oracle/shim/intrin.h -- TEST INFRASTRUCTURE ONLY.
// GCC spellings of the MSVC intrinsics the reference uses
// (LinesOfActionZasady.cpp:54,183; object_pool_multisize.h:43,64,111; GraWPanaZasadyV2.cpp:379,458).
#pragma once
#include <cstdint>
#include <cassert>
static inline unsigned char _BitScanForward64(unsigned long* idx, unsigned long long v)
{ if (!v) return 0; *idx = (unsigned long)__builtin_ctzll(v); return 1; }
static inline unsigned char _BitScanReverse64(unsigned long* idx, unsigned long long v)
{ if (!v) return 0; *idx = 63ul - (unsigned long)__builtin_clzll(v); return 1; }
static inline unsigned char _BitScanForward(unsigned long* idx, unsigned long v)
{ if (!(unsigned)v) return 0; *idx = (unsigned long)__builtin_ctz((unsigned)v); return 1; }
#define __popcnt64(x) ((unsigned long long)__builtin_popcountll((unsigned long long)(x)))
#define __popcnt(x)   ((unsigned)__builtin_popcount((unsigned)(x)))
#ifndef __max
#define __max(a,b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef __min
#define __min(a,b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef _ASSERT
#define _ASSERT(x) assert(x)
#endif
