// kernels_mgpu.cuh — the multi-GPU exchange steps of the visibility pipeline (the reference is single-GPU).
//
// Mesh instances are sharded over one process per GPU (oxc_set_shard*).  Two things are global per frame:
//   (1) the Hi-Z pyramid between the two passes.  mip 0 is a POINT SAMPLE of the depth image and max over ranks commutes
//       with sampling, so only mip-0 texels travel — and only those a rank actually drew into: every rank's image starts from
//       the same external depth, so a texel whose sample pixel still carries the clear id holds the same value everywhere.
//       k_mgpu_hiz_push samples the rank's packed vis buffer and max-reduces the texels it owns a fragment of STRAIGHT INTO
//       EVERY PEER'S exchange buffer over NVLink (red.relaxed.sys.max on CUDA-IPC mapped peer memory): no collective launch,
//       no staging copy, ~1/world of the image on the wire.  A flag per (parity, rank) written after the push is the cross-GPU
//       barrier; k_mgpu_hiz_collect spins on the local flags, moves the reduced texels into the pyramid and zeroes the exchange
//       buffer for its next use (two frames later: double-buffered by frame parity).
//   (2) the frame's outputs: per-pixel max of the packed vis buffer (ncclAllReduce, u64 max == reverse-Z depth test, NVLS
//       in-switch reduction) and the survivor lists (count + fixed-capacity id segments, ncclAllGather).
#pragma once
#include "kernels_hiz.cuh"

namespace oxc {

constexpr int MGPU_MAX_RANKS = 16;
constexpr uint32_t OXC_STATUS_PEER_TIMEOUT_BIT = 1u << 4; // == OXC_STATUS_PEER_TIMEOUT (include/oxcull.h)

struct MgpuPeers {
  uint32_t* xbuf[MGPU_MAX_RANKS];  // peer r's exchange buffer: [2 parities][hw * hh] u32 depth bits (this rank's own at [rank])
  uint32_t* flags[MGPU_MAX_RANKS]; // peer r's flag array: [2 parities][MGPU_MAX_RANKS]
  uint32_t rank, world;
};

#ifdef OXC_HOST_SOUNDNESS_HARNESS
// host builds of the headers (tests/): the emulated multi-rank test runs the ranks as threads of one process, "peer memory" is
// plain shared memory
OXC_DI void red_max_sys(uint32_t* addr, uint32_t v) { atomicMax(addr, v); }
OXC_DI void st_release_sys(uint32_t* addr, uint32_t v) { __atomic_store_n(addr, v, __ATOMIC_RELEASE); }
OXC_DI uint32_t ld_acquire_sys(const uint32_t* addr) { return __atomic_load_n(addr, __ATOMIC_ACQUIRE); }
OXC_DI unsigned long long global_timer_ns() { return 0ull; }
OXC_DI uint4 ld_cg_v4(const uint4* p) { return *p; }
OXC_DI uint32_t ld_cg_u32(const uint32_t* p) { return *p; }
#else
OXC_DI void red_max_sys(uint32_t* addr, uint32_t v) { asm volatile("red.relaxed.sys.global.max.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory"); }
OXC_DI void st_release_sys(uint32_t* addr, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory"); }
OXC_DI uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
OXC_DI unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// peers wrote with system-scope reductions that completed before their flags: a weak load after the acquire is enough, but the
// lines may sit stale in this SM's L1 from two frames ago -> bypass it
OXC_DI uint4 ld_cg_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
OXC_DI uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
#endif

// One thread per Hi-Z mip-0 texel.  seq = *d_seq + 1 is the exchange this launch belongs to (device-side counter: the same
// launch sequence replays correctly from a CUDA graph).
__global__ void __launch_bounds__(256) k_mgpu_hiz_push(const unsigned long long* __restrict__ vis, uint32_t width, uint32_t height, uint32_t hw,
                                                       uint32_t hh, uint32_t hw_shift, uint32_t hh_shift, const __grid_constant__ MgpuPeers peers,
                                                       const uint32_t* __restrict__ d_seq) {
  const uint32_t parity = (*d_seq + 1u) & 1u;
  const size_t n = (size_t)hw * hh;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(i % hw), y = (uint32_t)(i / hw);
    uint32_t sx = (uint32_t)(((uint64_t)(x + 1) * width) >> hw_shift), sy = (uint32_t)(((uint64_t)(y + 1) * height) >> hh_shift); // hiz.slang:92-95
    sx = sx > width - 1 ? width - 1 : sx;
    sy = sy > height - 1 ? height - 1 : sy;
    const unsigned long long v = __ldg(&vis[(size_t)sy * width + sx]);
    const uint32_t d = (uint32_t)(v >> 32);
    red_max_sys(peers.xbuf[peers.rank] + parity * n + i, d); // own contribution (a peer's may already be there: same scope as theirs)
    if ((uint32_t)v != OXC_VIS_CLEAR) {                     // this rank drew the sample pixel: tell everybody
      for (uint32_t r = 0; r < peers.world; r++)
        if (r != peers.rank) red_max_sys(peers.xbuf[r] + parity * n + i, d);
    }
  }
}

// After the push kernel has completed (stream order => its peer writes are performed): bump the exchange counter and raise
// this rank's flag on every rank.
__global__ void k_mgpu_signal(const __grid_constant__ MgpuPeers peers, uint32_t* d_seq) {
  const uint32_t seq = *d_seq + 1u;
  __threadfence_system();
  if (threadIdx.x < peers.world) st_release_sys(peers.flags[threadIdx.x] + (seq & 1u) * MGPU_MAX_RANKS + peers.rank, seq);
  __syncthreads();
  if (threadIdx.x == 0) *d_seq = seq;
}

// Wait until every rank has raised its flag for this exchange, then move the reduced texels into pyramid level 0 and zero the
// exchange buffer.  Bounded spin (%globaltimer; default 30 s, OXC_MGPU_TIMEOUT_MS): a missing peer raises OXC_STATUS_PEER_TIMEOUT
// instead of hanging the GPU for good.
__global__ void __launch_bounds__(256) k_mgpu_hiz_collect(const __grid_constant__ MgpuPeers peers, const uint32_t* __restrict__ d_seq, float* mip0,
                                                          size_t n, uint32_t* status, unsigned long long timeout_ns) {
  __shared__ uint32_t ok_s;
  const uint32_t seq = *d_seq; // k_mgpu_signal of this exchange has run (stream order)
  if (threadIdx.x == 0) ok_s = 1;
  __syncthreads();
  if (threadIdx.x < peers.world) {
    const uint32_t* f = peers.flags[peers.rank] + (seq & 1u) * MGPU_MAX_RANKS + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
      if (global_timer_ns() - t0 > timeout_ns) { ok_s = 0; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (!ok_s && threadIdx.x == 0 && blockIdx.x == 0) atomicOr(status, OXC_STATUS_PEER_TIMEOUT_BIT);
  uint4* src = reinterpret_cast<uint4*>(peers.xbuf[peers.rank] + (seq & 1u) * n);
  uint4* dst = reinterpret_cast<uint4*>(mip0);
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = ld_cg_v4(src + i);
    dst[i] = v;
    src[i] = make_uint4(0, 0, 0, 0);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t* s1 = peers.xbuf[peers.rank] + (seq & 1u) * n + i;
    const uint32_t v = ld_cg_u32(s1);
    reinterpret_cast<uint32_t*>(mip0)[i] = v;
    *s1 = 0;
  }
}

// Survivor list + counters of this frame -> staging slot (the context's own list is rewritten by the next frame while the
// exchange of this one may still be running on a side stream).  Overflow of the gather capacity is an ERROR, not a truncation
// the host has to notice by itself: the sticky status bit is raised.
__global__ void __launch_bounds__(256) k_mgpu_stage_survivors(const OxcMeshletInstanceVisibility* __restrict__ vis, const uint32_t* __restrict__ ids,
                                                              uint32_t capacity, uint32_t* cnt_stage, uint32_t* ids_stage, uint32_t* status) {
  const uint32_t total = vis->total_visible_meshlet_instances, early = vis->early_visible_meshlet_instances, late = vis->late_visible_meshlet_instances;
  uint32_t n = early + late;
  if (n > capacity) {
    n = capacity;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, (uint32_t)OXC_STATUS_SURVIVOR_OVERFLOW);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { cnt_stage[0] = total; cnt_stage[1] = early; cnt_stage[2] = late; cnt_stage[3] = n; }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) ids_stage[i] = ids[i];
}

} // namespace oxc
