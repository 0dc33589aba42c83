// kernels_hiz.cuh — Hi-Z depth-pyramid build.  Reference: passes/hiz.slang:171-267 (AMD-SPD style single
// pass: each 256-thread workgroup reduces a 64x64 tile of mip 0 down to mip 6; the last workgroup does
// mips 7..12) recorded by generate_hiz (Passes/CullGeometry.cpp:10-59).
//
// mip 0 is a POINT SAMPLE of the depth image (hiz.slang:92-95: nearest sampler at uv=(texel+1)/hiz_extent),
// source texel = min(W-1, ((x+1)*W) >> log2(hizW)); mip k = 2x2 min of mip k-1 (hiz.slang:77-83).
// `min` is order independent for finite inputs, so the result is bit-exact whatever the reduction tree.
#pragma once
#include "oxc_exact.cuh"

namespace oxc {

struct HizBuildParams {
  const float* depth;       // D32F image, or the packed 64-bit vis buffer viewed as floats
  uint32_t elem_stride;     // 1: plain float image; 2: packed u64 image (depth = high word)
  uint32_t elem_offset;     // 0 / 1
  uint32_t width, height;   // depth image
  float* hiz;
  uint32_t hw, hh, levels;  // pyramid
  uint32_t hw_shift, hh_shift;
  uint32_t mode;            // 0: sample depth -> all mips; 1: sample depth -> mip 0 only; 2: mips 1.. from the mip 0 already in the pyramid
  uint32_t level_offset[OXC_HIZ_MAX_LEVELS];
};

OXC_DI float hiz_sample(const HizBuildParams& p, uint32_t x, uint32_t y) {
  uint32_t sx = (uint32_t)(((uint64_t)(x + 1) * p.width) >> p.hw_shift);
  uint32_t sy = (uint32_t)(((uint64_t)(y + 1) * p.height) >> p.hh_shift);
  sx = sx > p.width - 1 ? p.width - 1 : sx;
  sy = sy > p.height - 1 ? p.height - 1 : sy;
  return __ldg(p.depth + ((size_t)sy * p.width + sx) * p.elem_stride + p.elem_offset);
}

// One CTA per 64x64 tile of mip 0 (requires hw, hh multiples of 64).  Thread (tx,ty) of a 16x16 layout owns
// a 4x4 block: 16 point samples -> mip0 (4 x 128-bit stores) -> mip1 (2 x 64-bit) -> mip2 (1) in registers;
// mips 3..6 through 1 KB of shared memory.
__global__ void __launch_bounds__(256) k_hiz_tiles(const __grid_constant__ HizBuildParams p) {
  __shared__ float s2[16][16];
  __shared__ float s3[8][8];
  __shared__ float s4[4][4];
  __shared__ float s5[2][2];
  const uint32_t tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const uint32_t x0 = blockIdx.x * 64 + tx * 4, y0 = blockIdx.y * 64 + ty * 4;
  float v[4][4];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) v[j][i] = p.mode == 2 ? 0.0f : hiz_sample(p, x0 + i, y0 + j);
  float* m0 = p.hiz + p.level_offset[0];
  if (p.mode == 2) { // mip 0 was produced earlier (and max-reduced across GPUs): read it back instead of sampling
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float4 r = *reinterpret_cast<const float4*>(m0 + (size_t)(y0 + j) * p.hw + x0);
      v[j][0] = r.x; v[j][1] = r.y; v[j][2] = r.z; v[j][3] = r.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++)
      *reinterpret_cast<float4*>(m0 + (size_t)(y0 + j) * p.hw + x0) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
    if (p.mode == 1) return;
  }
  // mip 1
  float q[2][2];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int i = 0; i < 2; i++)
      q[j][i] = omin(omin(v[2 * j][2 * i], v[2 * j][2 * i + 1]), omin(v[2 * j + 1][2 * i], v[2 * j + 1][2 * i + 1]));
  {
    float* m1 = p.hiz + p.level_offset[1];
    const uint32_t w1 = p.hw >> 1;
#pragma unroll
    for (int j = 0; j < 2; j++)
      *reinterpret_cast<float2*>(m1 + (size_t)((y0 >> 1) + j) * w1 + (x0 >> 1)) = make_float2(q[j][0], q[j][1]);
  }
  // mip 2
  const float d2 = omin(omin(q[0][0], q[0][1]), omin(q[1][0], q[1][1]));
  p.hiz[p.level_offset[2] + (size_t)(y0 >> 2) * (p.hw >> 2) + (x0 >> 2)] = d2;
  s2[ty][tx] = d2;
  __syncthreads();
  const uint32_t t = threadIdx.x;
  if (t < 64) { // mip 3: 8x8 per tile
    const uint32_t x = t & 7, y = t >> 3;
    const float d = omin(omin(s2[2 * y][2 * x], s2[2 * y][2 * x + 1]), omin(s2[2 * y + 1][2 * x], s2[2 * y + 1][2 * x + 1]));
    s3[y][x] = d;
    p.hiz[p.level_offset[3] + (size_t)(blockIdx.y * 8 + y) * (p.hw >> 3) + blockIdx.x * 8 + x] = d;
  }
  __syncthreads();
  if (t < 16) { // mip 4
    const uint32_t x = t & 3, y = t >> 2;
    const float d = omin(omin(s3[2 * y][2 * x], s3[2 * y][2 * x + 1]), omin(s3[2 * y + 1][2 * x], s3[2 * y + 1][2 * x + 1]));
    s4[y][x] = d;
    p.hiz[p.level_offset[4] + (size_t)(blockIdx.y * 4 + y) * (p.hw >> 4) + blockIdx.x * 4 + x] = d;
  }
  __syncthreads();
  if (t < 4) { // mip 5
    const uint32_t x = t & 1, y = t >> 1;
    const float d = omin(omin(s4[2 * y][2 * x], s4[2 * y][2 * x + 1]), omin(s4[2 * y + 1][2 * x], s4[2 * y + 1][2 * x + 1]));
    s5[y][x] = d;
    p.hiz[p.level_offset[5] + (size_t)(blockIdx.y * 2 + y) * (p.hw >> 5) + blockIdx.x * 2 + x] = d;
  }
  __syncthreads();
  if (t == 0) // mip 6
    p.hiz[p.level_offset[6] + (size_t)blockIdx.y * (p.hw >> 6) + blockIdx.x] =
        omin(omin(s5[0][0], s5[0][1]), omin(s5[1][0], s5[1][1]));
}

// Tail: one CTA reduces level first_level-1 -> ... -> levels-1 (the reference's last-workgroup path,
// hiz.slang:236-266).  Coordinates clamp so non-square pyramids are well defined.
__global__ void __launch_bounds__(1024) k_hiz_tail(const __grid_constant__ HizBuildParams p, uint32_t first_level) {
  for (uint32_t l = first_level; l < p.levels; l++) {
    uint32_t pw = p.hw >> (l - 1), ph = p.hh >> (l - 1), mw = p.hw >> l, mh = p.hh >> l;
    pw = pw < 1 ? 1 : pw; ph = ph < 1 ? 1 : ph; mw = mw < 1 ? 1 : mw; mh = mh < 1 ? 1 : mh;
    const float* src = p.hiz + p.level_offset[l - 1];
    float* dst = p.hiz + p.level_offset[l];
    for (uint32_t i = threadIdx.x; i < mw * mh; i += blockDim.x) {
      const uint32_t x = i % mw, y = i / mw;
      const uint32_t xa = min(2 * x, pw - 1), xb = min(2 * x + 1, pw - 1), ya = min(2 * y, ph - 1), yb = min(2 * y + 1, ph - 1);
      const float a = src[(size_t)ya * pw + xa], b = src[(size_t)ya * pw + xb];
      const float c = src[(size_t)yb * pw + xa], d = src[(size_t)yb * pw + xb];
      dst[i] = omin(omin(a, b), omin(c, d));
    }
    __syncthreads(); // block-scope visibility of dst before it becomes src
  }
}

// Generic mip 0 for pyramids smaller than one tile (tiny test resolutions).
__global__ void k_hiz_mip0_generic(const __grid_constant__ HizBuildParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.hw * p.hh) return;
  p.hiz[p.level_offset[0] + i] = hiz_sample(p, i % p.hw, i / p.hw);
}

__global__ void k_fill_u32(uint32_t* dst, uint32_t value, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = value;
}

} // namespace oxc
