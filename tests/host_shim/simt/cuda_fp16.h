// Host stand-in for <cuda_fp16.h> (see cuda_runtime.h in this directory): half -> float only.
#pragma once
#include "../cuda_runtime.h"
struct __half { unsigned short x; };
static inline __half __ushort_as_half(unsigned short v) { return __half{v}; }
static inline float __half2float(__half h) {
  const unsigned int s = (unsigned int)(h.x & 0x8000u) << 16, e = (h.x >> 10) & 31u, m = h.x & 1023u;
  if (e == 0) return __uint_as_float(s) + (s ? -1.0f : 1.0f) * (float)m * 5.9604644775390625e-08f; // m * 2^-24 (exact)
  if (e == 31) return __uint_as_float(s | 0x7F800000u | (m ? (0x00400000u | (m << 13)) : 0u));      // Inf / quieted NaN
  return __uint_as_float(s | ((e + 112u) << 23) | (m << 13));
}
