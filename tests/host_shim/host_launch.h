// Host stand-ins for the CUDA execution model, for kernels whose threads are independent of one another (one thread per mesh
// instance / per pixel): tests/host_kernels.cpp runs the KERNEL FUNCTIONS THEMSELVES, one simulated thread at a time, by setting
// blockIdx / threadIdx and calling them.  Cross-thread primitives are stubbed so that such kernels compile (a shuffle returns
// the caller's own value, a barrier does nothing): results that depend on them — the block sums k_cull_meshes reduces for the
// scan — are NOT meaningful here and are not compared.  Test infrastructure only; see cuda_runtime.h in this directory.
#pragma once
#include <climits>
#include <cstddef>

#include "cuda_runtime.h"

struct HostDim3 { unsigned int x = 0, y = 0, z = 0; };
static thread_local HostDim3 blockIdx, threadIdx;
static thread_local HostDim3 blockDim, gridDim;

#undef __global__
#define __global__ static __attribute__((unused))
#define __shared__ /* a per-thread local here: only kernels whose threads do not communicate are run */
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline unsigned __activemask() { return 1u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __match_any_sync(unsigned, unsigned long long) { return 1u; }
static inline unsigned __match_any_sync(unsigned, unsigned) { return 1u; }
static inline void __nanosleep(unsigned) {}
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return v; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
template <typename T>
static inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
template <typename T>
static inline T atomicOr(T* p, T v) { const T o = *p; *p = o | v; return o; }
template <typename T>
static inline T atomicXor(T* p, T v) { const T o = *p; *p = o ^ v; return o; }
template <typename T>
static inline T atomicAnd(T* p, T v) { const T o = *p; *p = o & v; return o; }
template <typename T>
static inline T atomicMax(T* p, T v) { const T o = *p; if (v > o) *p = v; return o; }
template <typename T>
static inline T atomicMin(T* p, T v) { const T o = *p; if (v < o) *p = v; return o; }
template <typename T>
static inline T atomicExch(T* p, T v) { const T o = *p; *p = v; return o; }
template <typename T>
static inline T atomicCAS(T* p, T c, T v) { const T o = *p; if (o == c) *p = v; return o; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
