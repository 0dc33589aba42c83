// oxc_exact.cuh — the canonical binary32 arithmetic of the pipeline on the device.
//
// Every function here reproduces, operation for operation, the evaluation order the CPU oracle commits
// to (oracle/oxc_oracle.h header): IEEE RN, no fma contraction, denormals kept, IEEE divide / sqrt.
// The __f*_rn intrinsics are never contracted by nvcc, whatever -fmad says.
// Reference lines: Oxylus/src/Render/Shaders/cull.slang, scene.slang, common/math.slang.
#pragma once
#include <cuda_fp16.h>

#include "oxc_types.cuh"

namespace oxc {

#define OXC_DI __device__ __forceinline__

OXC_DI float fm(float a, float b) { return __fmul_rn(a, b); }
OXC_DI float fa(float a, float b) { return __fadd_rn(a, b); }
OXC_DI float fs(float a, float b) { return __fsub_rn(a, b); }
OXC_DI float fd(float a, float b) { return __fdiv_rn(a, b); }
OXC_DI float fsq(float a) { return __fsqrt_rn(a); }
OXC_DI float omin(float a, float b) { return a < b ? a : b; }
OXC_DI float omax(float a, float b) { return a > b ? a : b; }

// dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z
OXC_DI float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return fa(fa(fm(ax, bx), fm(ay, by)), fm(az, bz));
}
OXC_DI float length3(float x, float y, float z) { return fsq(dot3(x, y, z, x, y, z)); }

// common/math.slang:193-201 dequantize_half.  Identical for every one of the 65536 inputs to
// "hardware half->float, except denormals flush to (signed) zero" (tests/test_gpu_units.py checks all).
OXC_DI float dequantize_half(uint32_t h) {
  uint32_t s = (h & 0x8000u) << 16;
  uint32_t em = h & 0x7fffu;
  uint32_t r = (em + (112u << 10)) << 13;
  r = (em < (1u << 10)) ? 0u : r;
  r += (em >= (31u << 10)) ? (112u << 23) : 0u;
  return __uint_as_float(s | r);
}

// Same function through the hardware half->float conversion: identical for all 65536 inputs up to NaN
// payloads (hardware quiets signalling NaNs; every consumer only compares).  ~4 instructions instead of ~10.
OXC_DI float dequantize_half_hw(uint32_t h) {
  const float f = __half2float(__ushort_as_half((unsigned short)h));
  return (h & 0x7C00u) == 0u ? __uint_as_float((h & 0x8000u) << 16) : f; // denormals flush to signed zero
}

// scene.slang:408-418: i8 / 127.0
OXC_DI float s8_over_127(int v) { return fd((float)v, 127.0f); }

// row . (x,y,z,w): ((r.x*x + r.y*y) + r.z*z) + r.w*w
OXC_DI float row_dot4(float4 r, float x, float y, float z, float w) {
  return fa(fa(fa(fm(r.x, x), fm(r.y, y)), fm(r.z, z)), fm(r.w, w));
}
// w == 1.0f: r.w * 1.0f == r.w exactly
OXC_DI float row_dot_p1(float4 r, float x, float y, float z) { return fa(fa(fa(fm(r.x, x), fm(r.y, y)), fm(r.z, z)), r.w); }

struct ScreenAabb {
  float minx, miny, minz, maxx, maxy, maxz;
};

// cull.slang:12-47.  mvp rows r0..r3.  e = FULL extent.  Returns false for `none`.
// SX/SY/SZ: mul(mvp, (e.x,0,0,0))[i] = ((M[i][0]*e.x + M[i][1]*0) + M[i][2]*0) + M[i][3]*0 == M[i][0]*e.x
// up to the sign of a zero, which no consumer below can observe.
OXC_DI bool project_aabb(const float4 r0, const float4 r1, const float4 r2, const float4 r3, float near_clip, float cx,
                         float cy, float cz, float ex, float ey, float ez, ScreenAabb& out) {
  const float SXx = fm(r0.x, ex), SXy = fm(r1.x, ex), SXz = fm(r2.x, ex), SXw = fm(r3.x, ex);
  const float SYx = fm(r0.y, ey), SYy = fm(r1.y, ey), SYz = fm(r2.y, ey), SYw = fm(r3.y, ey);
  const float SZx = fm(r0.z, ez), SZy = fm(r1.z, ez), SZz = fm(r2.z, ez), SZw = fm(r3.z, ez);
  const float px = fs(cx, fm(ex, 0.5f)), py = fs(cy, fm(ey, 0.5f)), pz = fs(cz, fm(ez, 0.5f));
  float X[8], Y[8], Z[8], W[8];
  X[0] = row_dot_p1(r0, px, py, pz); Y[0] = row_dot_p1(r1, px, py, pz);
  Z[0] = row_dot_p1(r2, px, py, pz); W[0] = row_dot_p1(r3, px, py, pz);
#define OXC_ADDV(d, s, V) X[d] = fa(X[s], V##x); Y[d] = fa(Y[s], V##y); Z[d] = fa(Z[s], V##z); W[d] = fa(W[s], V##w);
  OXC_ADDV(1, 0, SZ) OXC_ADDV(2, 0, SY) OXC_ADDV(3, 2, SZ) OXC_ADDV(4, 0, SX)
  OXC_ADDV(5, 4, SZ) OXC_ADDV(6, 4, SY) OXC_ADDV(7, 6, SZ)
#undef OXC_ADDV
  float depth = W[7];
#pragma unroll
  for (int i = 6; i >= 0; i--) depth = omin(W[i], depth);
  if (depth < near_clip) return false;
  float mnx, mny, mnz, mxx, mxy, mxz;
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    const float dx = fd(X[i], W[i]), dy = fd(Y[i], W[i]), dz = fd(Z[i], W[i]);
    if (i == 7) { mnx = mxx = dx; mny = mxy = dy; mnz = mxz = dz; }
    else {
      mnx = omin(dx, mnx); mny = omin(dy, mny); mnz = omin(dz, mnz);
      mxx = omax(dx, mxx); mxy = omax(dy, mxy); mxz = omax(dz, mxz);
    }
  }
  out.minx = fa(fm(mnx, 0.5f), 0.5f); out.miny = fa(fm(mny, 0.5f), 0.5f); out.minz = mnz;
  out.maxx = fa(fm(mxx, 0.5f), 0.5f); out.maxy = fa(fm(mxy, 0.5f), 0.5f); out.maxz = mxz;
  return true;
}

// cull.slang:73-81 with the six normalised planes hoisted per instance (InstCull::plane)
OXC_DI bool test_frustum_planes(const float4* __restrict__ planes, float cx, float cy, float cz, float ex, float ey,
                                float ez) {
  const float hx = fm(ex, 0.5f), hy = fm(ey, 0.5f), hz = fm(ez, 0.5f);
  float4 pl[6];
#pragma unroll
  for (int i = 0; i < 6; i++) pl[i] = __ldg(&planes[i]); // all six loads in flight before the first test
  bool inside = true;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float4 p = pl[i];
    const float sx = __uint_as_float(__float_as_uint(hx) ^ (__float_as_uint(p.x) & 0x80000000u));
    const float sy = __uint_as_float(__float_as_uint(hy) ^ (__float_as_uint(p.y) & 0x80000000u));
    const float sz = __uint_as_float(__float_as_uint(hz) ^ (__float_as_uint(p.z) & 0x80000000u));
    inside = inside && !(dot3(fa(cx, sx), fa(cy, sy), fa(cz, sz), p.x, p.y, p.z) <= -p.w);
  }
  return inside;
}

// ceil(log2(float(n))) in integers (oracle: orc_ceil_log2_u32)
OXC_DI uint32_t ceil_log2_u32(uint32_t n) { return n <= 1u ? 0u : 32u - (uint32_t)__clz((int)(n - 1u)); }

// cull.slang:86-135.  hiz_off = per-level float offsets (shared memory copy).
OXC_DI bool test_occlusion(const ScreenAabb& a, const float* __restrict__ hiz, uint32_t hw_u, uint32_t hh_u,
                           uint32_t levels, const uint32_t* hiz_off) {
  const float hw = (float)hw_u, hh = (float)hh_u;
  const uint32_t min_tx = __float2uint_rz(omax(fm(a.minx, hw), 0.0f));
  const uint32_t min_ty = __float2uint_rz(omax(fm(a.miny, hh), 0.0f));
  const uint32_t max_tx = __float2uint_rz(omin(fm(a.maxx, hw), fs(hw, 1.0f)));
  const uint32_t max_ty = __float2uint_rz(omin(fm(a.maxy, hh), fs(hh, 1.0f)));
  const uint32_t sx = max_tx - min_tx, sy = max_ty - min_ty;
  const uint32_t max_size = sx > sy ? sx : sy;
  uint32_t mip = ceil_log2_u32(max_size);
  mip = mip > levels - 1 ? levels - 1 : mip;
  const float u = fd(fm(fa((float)min_tx, (float)max_tx), 0.5f), hw);
  const float v = fd(fm(fa((float)min_ty, (float)max_ty), 0.5f), hh);
  uint32_t mw = hw_u >> mip, mh = hh_u >> mip;
  mw = mw < 1 ? 1 : mw;
  mh = mh < 1 ? 1 : mh;
  const int bx = __float2int_rz(floorf(fs(fm(u, (float)mw), 0.5f)));
  const int by = __float2int_rz(floorf(fs(fm(v, (float)mh), 0.5f)));
  const int mx = (int)mw - 1, my = (int)mh - 1;
  const int x0 = min(max(bx, 0), mx), x1 = min(max(bx + 1, 0), mx);
  const int y0 = min(max(by, 0), my), y1 = min(max(by + 1, 0), my);
  const float* lvl = hiz + hiz_off[mip];
  const float p00 = __ldg(lvl + (size_t)y0 * mw + x0), p10 = __ldg(lvl + (size_t)y0 * mw + x1);
  const float p01 = __ldg(lvl + (size_t)y1 * mw + x0), p11 = __ldg(lvl + (size_t)y1 * mw + x1);
  const float d = omin(omin(p00, p10), omin(p01, p11));
  return a.maxz <= fs(d, 1e-7f);
}

// cull_meshlets_hiz.slang:53-58 cone part with the per-instance terms hoisted.  Returns cone_visible.
OXC_DI bool cone_visible_positional(const InstCull* __restrict__ ic, float cx, float cy, float cz, float ex, float ey,
                                    float ez, float ax, float ay, float az, float cutoff, float camx, float camy,
                                    float camz) {
  if (cutoff >= 1.0f) return true;
  const float4 n0 = __ldg(&ic->nrm[0]), n1 = __ldg(&ic->nrm[1]), n2 = __ldg(&ic->nrm[2]);
  // mul(N, axis)[i] = (r0[i]*a.x + r1[i]*a.y) + r2[i]*a.z
  const float nx = fa(fa(fm(n0.x, ax), fm(n1.x, ay)), fm(n2.x, az));
  const float ny = fa(fa(fm(n0.y, ax), fm(n1.y, ay)), fm(n2.y, az));
  const float nz = fa(fa(fm(n0.z, ax), fm(n1.z, ay)), fm(n2.z, az));
  const float len = length3(nx, ny, nz);
  const float wax = fd(nx, len), way = fd(ny, len), waz = fd(nz, len);
  const float4 w0 = __ldg(&ic->world_row[0]), w1 = __ldg(&ic->world_row[1]), w2 = __ldg(&ic->world_row[2]);
  const float wcx = row_dot_p1(w0, cx, cy, cz), wcy = row_dot_p1(w1, cx, cy, cz), wcz = row_dot_p1(w2, cx, cy, cz);
  const float hx = fm(ex, 0.5f), hy = fm(ey, 0.5f), hz = fm(ez, 0.5f);
  const float wr = fm(length3(hx, hy, hz), n0.w); // to_world_radius: radius * max row length
  const float dx = fs(wcx, camx), dy = fs(wcy, camy), dz = fs(wcz, camz);
  // test_cone: dot(d, axis) >= cutoff * length(d) + radius
  const bool culled = dot3(dx, dy, dz, wax, way, waz) >= fa(fm(cutoff, length3(dx, dy, dz)), wr);
  return !culled;
}

// world-space cone axis only (directional cone test, cull_meshlets_hpb.slang:53-54)
OXC_DI void world_cone_axis(const InstCull* __restrict__ ic, float ax, float ay, float az, float& wax, float& way,
                            float& waz) {
  const float4 n0 = __ldg(&ic->nrm[0]), n1 = __ldg(&ic->nrm[1]), n2 = __ldg(&ic->nrm[2]);
  const float nx = fa(fa(fm(n0.x, ax), fm(n1.x, ay)), fm(n2.x, az));
  const float ny = fa(fa(fm(n0.y, ax), fm(n1.y, ay)), fm(n2.y, az));
  const float nz = fa(fa(fm(n0.z, ax), fm(n1.z, ay)), fm(n2.z, az));
  const float len = length3(nx, ny, nz);
  wax = fd(nx, len); way = fd(ny, len); waz = fd(nz, len);
}

// mul(A, B) with column-major storage (m[col*4+row]); out row-major rows as float4
OXC_DI void mul_mm_rows(const float* __restrict__ a, const float* __restrict__ b, float4 rows[4]) {
#define A(i, j) a[(j) * 4 + (i)]
#define B(i, j) b[(j) * 4 + (i)]
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
      r[j] = fa(fa(fa(fm(A(i, 0), B(0, j)), fm(A(i, 1), B(1, j))), fm(A(i, 2), B(2, j))), fm(A(i, 3), B(3, j)));
    rows[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
#undef A
#undef B
}

// cull.slang:58-71 + normalize_plane :49-51
OXC_DI void frustum_planes(const float4 rows[4], float4 planes[6]) {
  float4 p[6];
  p[0] = make_float4(fa(rows[3].x, rows[0].x), fa(rows[3].y, rows[0].y), fa(rows[3].z, rows[0].z), fa(rows[3].w, rows[0].w));
  p[1] = make_float4(fs(rows[3].x, rows[0].x), fs(rows[3].y, rows[0].y), fs(rows[3].z, rows[0].z), fs(rows[3].w, rows[0].w));
  p[2] = make_float4(fa(rows[3].x, rows[1].x), fa(rows[3].y, rows[1].y), fa(rows[3].z, rows[1].z), fa(rows[3].w, rows[1].w));
  p[3] = make_float4(fs(rows[3].x, rows[1].x), fs(rows[3].y, rows[1].y), fs(rows[3].z, rows[1].z), fs(rows[3].w, rows[1].w));
  p[4] = rows[2];
  p[5] = make_float4(fs(rows[3].x, rows[2].x), fs(rows[3].y, rows[2].y), fs(rows[3].z, rows[2].z), fs(rows[3].w, rows[2].w));
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float len = length3(p[i].x, p[i].y, p[i].z);
    planes[i] = make_float4(fd(p[i].x, len), fd(p[i].y, len), fd(p[i].z, len), fd(p[i].w, len));
  }
}

// full test_frustum from rows (mesh-level test in cull_meshes.slang:34)
OXC_DI bool test_frustum_rows(const float4 planes[6], float cx, float cy, float cz, float ex, float ey, float ez) {
  const float hx = fm(ex, 0.5f), hy = fm(ey, 0.5f), hz = fm(ez, 0.5f);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float4 p = planes[i];
    const float sx = __uint_as_float(__float_as_uint(hx) ^ (__float_as_uint(p.x) & 0x80000000u));
    const float sy = __uint_as_float(__float_as_uint(hy) ^ (__float_as_uint(p.y) & 0x80000000u));
    const float sz = __uint_as_float(__float_as_uint(hz) ^ (__float_as_uint(p.z) & 0x80000000u));
    if (dot3(fa(cx, sx), fa(cy, sy), fa(cz, sz), p.x, p.y, p.z) <= -p.w) return false;
  }
  return true;
}

// log2 for decision paths (VSM clipmap selection): the same operation sequence as oracle/oxc_oracle.c orc_log2_canonical, so the
// result is bit-identical on both sides.  |error| < 1e-7 absolute on [2^-126, 2^127].
OXC_DI float canonical_log2(float x) {
  if (!(x > 0.0f)) return -3.0e38f;
  if (!(x <= 3.0e38f)) return 3.0e38f;
  int e = 0;
  if (x < 1.17549435e-38f) { x = fm(x, 16777216.0f); e = -24; }
  const uint32_t bits = __float_as_uint(x);
  e += (int)(bits >> 23) - 127;
  float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421354f) { m = fm(m, 0.5f); e += 1; }
  const float t = fd(fs(m, 1.0f), fa(m, 1.0f));
  const float t2 = fm(t, t);
  float p = 0.412198562f;
  p = fa(0.577078044f, fm(t2, p));
  p = fa(0.961796701f, fm(t2, p));
  p = fa(2.88539004f, fm(t2, p));
  return fa((float)e, fm(t, p));
}

// cull.slang:169-171: determinant(f32x3x3(c0.xyw, c1.xyw, c2.xyw)) >= 0.0001
OXC_DI bool triangle_backface(float4 c0, float4 c1, float4 c2) {
  const float m00 = c0.x, m01 = c0.y, m02 = c0.w;
  const float m10 = c1.x, m11 = c1.y, m12 = c1.w;
  const float m20 = c2.x, m21 = c2.y, m22 = c2.w;
  const float t0 = fm(m00, fs(fm(m11, m22), fm(m12, m21)));
  const float t1 = fm(m01, fs(fm(m10, m22), fm(m12, m20)));
  const float t2 = fm(m02, fs(fm(m10, m21), fm(m11, m20)));
  return fa(fs(t0, t1), t2) >= 0.0001f;
}

} // namespace oxc
