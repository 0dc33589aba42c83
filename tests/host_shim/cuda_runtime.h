// Host stand-in for <cuda_runtime.h> — ONLY for tests/filter_soundness.cpp, which compiles the device headers
// oxc_exact.cuh / oxc_filtered.cuh with g++ (-ffp-contract=off: every float operation is one IEEE binary32 rounding, which
// is what the __f*_rn intrinsics guarantee on the device).  Test infrastructure; nothing in the product includes it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __grid_constant__

struct __align__(16) float4 { float x, y, z, w; };
struct __align__(8) float2 { float x, y; };
struct __align__(8) uint2 { unsigned int x, y; };
struct __align__(16) uint4 { unsigned int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

template <typename T>
static inline T __ldg(const T* p) { return *p; }

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }
// cvt.rzi.u32.f32 / cvt.rzi.s32.f32: truncate, saturate, NaN -> 0
static inline unsigned int __float2uint_rz(float f) {
  if (!(f > 0.0f)) return 0u;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (unsigned int)f;
}
static inline int __float2int_rz(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return (-2147483647 - 1);
  return (int)f;
}
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned int)x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
