// oxc_filtered.cuh — filtered predicates for the meshlet cull.
//
// Every OUTPUT of the cull is an integer decision (visible / not, texel indices), so a decision may be taken
// with cheaper arithmetic (fma, MUFU rcp/rsqrt) whenever a rigorous error bound proves the canonical
// evaluation (oxc_exact.cuh == the CPU oracle) must agree; only margin-ambiguous items run the canonical
// path.  The result is bit-identical to the canonical path for EVERY input; the bounds are derived below
// with u = 2^-24 (binary32 unit roundoff) and carry >= 1.9x slack.
#pragma once
#include "oxc_exact.cuh"

namespace oxc {

enum Tri : int { TRI_FALSE = 0, TRI_TRUE = 1, TRI_AMBIGUOUS = 2 };

#ifdef OXC_HOST_SOUNDNESS_HARNESS
// tests/filter_soundness.cpp compiles this header for the HOST and supplies the two approximate units itself: the exact
// value perturbed adversarially within the error the PTX ISA allows, so the bounds below are tested against the worst case
// the hardware may produce, not only against what one GPU happens to return.
float rcp_approx(float x);
float rsqrt_approx(float x);
#else
OXC_DI float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); // max relative error 2^-23 (PTX ISA)
  return r;
}
OXC_DI float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); // max relative error 2^-22.4 (PTX ISA)
  return r;
}
#endif

// Whole-instance frustum shortcut (k_cull_meshes).  U = union AABB of the decoded meshlet boxes of a LOD (ua = min xyz, max
// xyz), inflated for the rounding of c +- h.  The canonical test rejects a box inside U only if fl(dot(p,n)) <= -w; its rounding
// error is <= 3.1u * sum|p_i| <= 3.1u * B with B = sum_i max(|Umin_i|, |Umax_i|), and the real p-vertex value of any such box
// is >= the n-vertex value of U.  Require n-vertex(U) + w > 2^-18 (B + |w|)  (64u: > 8x slack incl. the rounding of this very
// evaluation) on all six planes: then no meshlet of the instance can fail the canonical frustum test.
OXC_DI bool union_box_inside_frustum(const float4* planes, const float* __restrict__ ua, bool have_aabb) {
  bool inside = have_aabb;
  float B = 0.0f;
  float lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float mn = ua[a], mx = ua[3 + a];
    const float pad = fmaxf(fabsf(mn), fabsf(mx)) * 4.76837158203125e-07f; // 2^-21 relative inflation
    lo[a] = mn - pad; hi[a] = mx + pad;
    B += fmaxf(fabsf(lo[a]), fabsf(hi[a]));
    inside = inside && (mn <= mx); // NaN / empty => false
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const float4 pl = planes[k];
    const float vx = pl.x >= 0.0f ? lo[0] : hi[0], vy = pl.y >= 0.0f ? lo[1] : hi[1], vz = pl.z >= 0.0f ? lo[2] : hi[2];
    const float sv = fmaf(vx, pl.x, fmaf(vy, pl.y, fmaf(vz, pl.z, pl.w)));
    inside = inside && (sv > (B + fabsf(pl.w)) * 3.814697265625e-06f);
  }
  return inside;
}

// Centre-inside frustum filter.  The canonical test (test_frustum_planes) rejects on plane i iff
// fl(dot(c (+) s*h, n_i)) <= -w_i with s = sign(n_i) and h >= 0, i.e. it evaluates Σ n_k c_k + Σ |n_k| h_k (each
// term rounded).  With M = Σ|c_k| + Σ h_k and |n_k| <= 1 + 4u the canonical value differs from the real one by
// <= 4.3u M, and the real one is >= Σ n_k c_k.  D = fma-chain(Σ n_k c_k + w_i) carries <= 3u (M + |w_i|).  Hence
//     D > 2^-19 (M + |w_i|)   (= 32u: > 4x slack)   ==>  the canonical test does NOT reject on plane i.
// True for all six planes => visible, exactly as the canonical test decides.  NaN / Inf anywhere makes a comparison
// false => "unknown" => the canonical path runs.
OXC_DI bool frustum_centre_inside(const float4* __restrict__ planes, float cx, float cy, float cz, float ex, float ey, float ez) {
  // h >= 0 is a precondition of the bound: a negative (or NaN) extent is not something a mesh builder produces, but the canonical
  // test is defined for it (its p-vertex then lies on the other side of the centre) => leave such boxes to the canonical path
  if (!(ex >= 0.0f && ey >= 0.0f && ez >= 0.0f)) return false;
  const float M = (fabsf(cx) + fabsf(cy)) + (fabsf(cz) + 0.5f * (fabsf(ex) + fabsf(ey) + fabsf(ez)));
  const float Mk = M * 1.9073486328125e-06f; // 2^-19
  bool inside = true;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float4 pl = __ldg(&planes[i]);
    const float D = fmaf(cx, pl.x, fmaf(cy, pl.y, fmaf(cz, pl.z, pl.w)));
    inside = inside && (D > fmaf(fabsf(pl.w), 1.9073486328125e-06f, Mk));
  }
  return inside;
}

// ------------------------------------------------------------------------------------------------
// Cone test (cull.slang:173-175 via cull_meshlets_hiz.slang:53-58).  Returns cone_VISIBLE as a Tri.
//
// Canonical: axis = n/|n| (3 IEEE divides), lhs = dot(d, axis), rhs = cutoff*|d| + wr, culled = lhs >= rhs,
//            with n = N*a, d = world*c - cam, wr = |h|*maxscale.
// n, d, h are computed HERE with the canonical operation order, so both evaluations share them bit for bit.
// Fast:  L = dot(d,n), RHS = (cutoff*|d| + wr)*|n|   (the same inequality multiplied by |n| > 0).
// Error budget relative to S = (|d| + wr)*|n|:   fast side <= 19u*S  (3u for L; |d|,|n|,|h| via x*rsqrt(x):
//   3u + 2^-22.4 + u each; two fmas), canonical side <= 12u*S (3.5u per axis component, 3u for the dot,
//   2.5u |d|, 4u wr, 2u for the final mul/add).  Threshold 2^-17*S = 128u*S  => > 4x slack.
// Degenerate inputs (|n| = 0, NaN, Inf) make every comparison false => TRI_AMBIGUOUS => canonical path.
// ------------------------------------------------------------------------------------------------
struct ConeInputs {
  float nx, ny, nz, dx, dy, dz, hx, hy, hz, maxscale;
};

OXC_DI ConeInputs cone_inputs(const InstCull* __restrict__ ic, float cx, float cy, float cz, float ex, float ey, float ez,
                              float ax, float ay, float az, float camx, float camy, float camz) {
  ConeInputs c;
  const float4 n0 = __ldg(&ic->nrm[0]), n1 = __ldg(&ic->nrm[1]), n2 = __ldg(&ic->nrm[2]);
  c.nx = fa(fa(fm(n0.x, ax), fm(n1.x, ay)), fm(n2.x, az));
  c.ny = fa(fa(fm(n0.y, ax), fm(n1.y, ay)), fm(n2.y, az));
  c.nz = fa(fa(fm(n0.z, ax), fm(n1.z, ay)), fm(n2.z, az));
  const float4 w0 = __ldg(&ic->world_row[0]), w1 = __ldg(&ic->world_row[1]), w2 = __ldg(&ic->world_row[2]);
  c.dx = fs(row_dot_p1(w0, cx, cy, cz), camx);
  c.dy = fs(row_dot_p1(w1, cx, cy, cz), camy);
  c.dz = fs(row_dot_p1(w2, cx, cy, cz), camz);
  c.hx = fm(ex, 0.5f); c.hy = fm(ey, 0.5f); c.hz = fm(ez, 0.5f);
  c.maxscale = n0.w;
  return c;
}

// canonical tail on the shared inputs (== cone_visible_positional)
OXC_DI bool cone_visible_exact(const ConeInputs& c, float cutoff) {
  const float len = length3(c.nx, c.ny, c.nz);
  const float wax = fd(c.nx, len), way = fd(c.ny, len), waz = fd(c.nz, len);
  const float wr = fm(length3(c.hx, c.hy, c.hz), c.maxscale);
  const bool culled = dot3(c.dx, c.dy, c.dz, wax, way, waz) >= fa(fm(cutoff, length3(c.dx, c.dy, c.dz)), wr);
  return !culled;
}

OXC_DI Tri cone_visible_fast(const ConeInputs& c, float cutoff) {
  const float nn = fmaf(c.nz, c.nz, fmaf(c.ny, c.ny, c.nx * c.nx));
  const float dd = fmaf(c.dz, c.dz, fmaf(c.dy, c.dy, c.dx * c.dx));
  const float hh = fmaf(c.hz, c.hz, fmaf(c.hy, c.hy, c.hx * c.hx));
  const float L = fmaf(c.dz, c.nz, fmaf(c.dy, c.ny, c.dx * c.nx));
  const float len_n = nn * rsqrt_approx(nn), len_d = dd * rsqrt_approx(dd);
  const float wr = (hh > 0.0f ? hh * rsqrt_approx(hh) : hh) * c.maxscale; // |h| = 0 is legitimate (degenerate box); a NaN stays a NaN (=> ambiguous)
  const float rhs = fmaf(cutoff, len_d, wr) * len_n;
  const float T = (len_d + wr) * len_n * 7.62939453125e-06f; // 2^-17
  const float diff = L - rhs;
  if (diff > T) return TRI_FALSE;   // surely culled => not visible
  if (diff < -T) return TRI_TRUE;   // surely not culled
  return TRI_AMBIGUOUS;
}

// ------------------------------------------------------------------------------------------------
// Occlusion (cull.slang:12-47 project_aabb + :86-135 test_occlusion) with the 24 IEEE divides replaced by
// 8 MUFU reciprocals.  Returns VISIBLE as a Tri (visible = project failed || !occluded).
//
// The clip-space corners X/Y/Z/W are built with the canonical operation order (identical bits), so the
// `depth < near` early-out is exact.  Only q = X * rcp(W) differs from the canonical X / W:
//   |q_fast - q_canon| <= (2^-23 + 2^-24 + 2^-24) |q| < 2^-22 * 1.6 |q|
// uv = q*0.5 + 0.5 (one rounding each side), t = uv * hiz_size (exact power-of-two scaling):
//   |t_fast - t_canon| <= size * 2^-22 * 1.05 (|q| + 1)        =>  delta = size * 2^-21 * (|q|max + 1)
// A float->texel conversion can only differ across an integer boundary, so it is safe when t is farther than
// delta from every integer it could cross (clamped regions are safe by construction).  With identical texel
// integers, mip / uv centre / the four Hi-Z loads / d are identical; the final compare max.z <= d - 1e-7 is
// safe when |max.z_fast - thr| > 2^-21 |max.z_fast|.  Anything else => TRI_AMBIGUOUS => canonical path.
// ------------------------------------------------------------------------------------------------
OXC_DI bool texel_safe_lo(float t_raw, float delta) { // u32(max(t, 0))
  if (t_raw < 1.0f - delta) return true;              // -> 0 on both sides
  return fabsf(t_raw - rintf(t_raw)) > delta;
}
OXC_DI bool texel_safe_hi(float t_raw, float size_m1, float delta) { // u32(min(t, size-1))
  if (t_raw > size_m1 + delta) return true;           // clamped to size-1 on both sides
  if (t_raw < 1.0f - delta) return true;              // -> 0 on both sides
  if (t_raw >= size_m1 - delta) return false;         // clamp boundary itself
  return fabsf(t_raw - rintf(t_raw)) > delta;
}

OXC_DI Tri occlusion_visible_fast(const float4 r0, const float4 r1, const float4 r2, const float4 r3, float near_clip,
                                  float cx, float cy, float cz, float ex, float ey, float ez, const float* __restrict__ hiz,
                                  uint32_t hw_u, uint32_t hh_u, uint32_t levels, const uint32_t* hiz_off, bool mvp_ok) {
  // Preconditions that rule out NaN / Inf in the corner arithmetic (then fminf/fmaxf == the canonical
  // `a < b ? a : b` up to the sign of zero, which no consumer observes): |mvp| <= 2^60 (InstCull flag), bounds
  // finite (|half| <= 65504 => |products| < 2^77), W >= 1e-18 so the reciprocal stays finite.
  if (!mvp_ok || !(fabsf(cx) <= 65504.0f && fabsf(cy) <= 65504.0f && fabsf(cz) <= 65504.0f && fabsf(ex) <= 65504.0f &&
                   fabsf(ey) <= 65504.0f && fabsf(ez) <= 65504.0f))
    return TRI_AMBIGUOUS;
  // --- canonical corner construction (same ops as project_aabb) ---
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
  const float depth = fminf(fminf(fminf(W[0], W[1]), fminf(W[2], W[3])), fminf(fminf(W[4], W[5]), fminf(W[6], W[7])));
  if (depth < near_clip) return TRI_TRUE; // project_aabb == none => visible (cull_meshlets_hiz.slang:62-64)
  if (!(depth >= 1e-18f)) return TRI_AMBIGUOUS;
  // --- fast perspective divide ---
  float mnx, mny, mxx, mxy, mxz;
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    const float r = rcp_approx(W[i]);
    const float dx = X[i] * r, dy = Y[i] * r, dz = Z[i] * r;
    if (i == 7) { mnx = mxx = dx; mny = mxy = dy; mxz = dz; }
    else {
      mnx = fminf(dx, mnx); mny = fminf(dy, mny);
      mxx = fmaxf(dx, mxx); mxy = fmaxf(dy, mxy); mxz = fmaxf(dz, mxz);
    }
  }
  const float hw = (float)hw_u, hh = (float)hh_u;
  const float uminx = fa(fm(mnx, 0.5f), 0.5f), uminy = fa(fm(mny, 0.5f), 0.5f);
  const float umaxx = fa(fm(mxx, 0.5f), 0.5f), umaxy = fa(fm(mxy, 0.5f), 0.5f);
  const float tminx = fm(uminx, hw), tminy = fm(uminy, hh), tmaxx = fm(umaxx, hw), tmaxy = fm(umaxy, hh);
  const float qx = fmaxf(fabsf(mnx), fabsf(mxx)) + 1.0f, qy = fmaxf(fabsf(mny), fabsf(mxy)) + 1.0f;
  const float dlx = hw * 4.76837158203125e-07f * qx, dly = hh * 4.76837158203125e-07f * qy; // size * 2^-21 * (|q|+1)
  const bool safe = texel_safe_lo(tminx, dlx) && texel_safe_lo(tminy, dly) && texel_safe_hi(tmaxx, hw - 1.0f, dlx) &&
                    texel_safe_hi(tmaxy, hh - 1.0f, dly);
  if (!safe) return TRI_AMBIGUOUS; // also catches NaN / Inf (every comparison false)
  // --- test_occlusion on (provably identical) integers ---
  const uint32_t min_tx = __float2uint_rz(omax(tminx, 0.0f)), min_ty = __float2uint_rz(omax(tminy, 0.0f));
  const uint32_t max_tx = __float2uint_rz(omin(tmaxx, hw - 1.0f)), max_ty = __float2uint_rz(omin(tmaxy, hh - 1.0f));
  const uint32_t sx = max_tx - min_tx, sy = max_ty - min_ty;
  const uint32_t max_size = sx > sy ? sx : sy;
  uint32_t mip = ceil_log2_u32(max_size);
  mip = mip > levels - 1 ? levels - 1 : mip;
  // x / 2^k == x * 2^-k exactly: hiz extents are powers of two
  const float u = fm(fm(fa((float)min_tx, (float)max_tx), 0.5f), 1.0f / hw);
  const float v = fm(fm(fa((float)min_ty, (float)max_ty), 0.5f), 1.0f / hh);
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
  const float thr = fs(d, 1e-7f);
  const float gap = mxz - thr;
  if (!(fabsf(gap) > fabsf(mxz) * 4.76837158203125e-07f)) return TRI_AMBIGUOUS; // 2^-21 |max.z|
  return gap <= 0.0f ? TRI_FALSE : TRI_TRUE; // occluded => not visible
}

// Early pass against the per-frame CLEARED pyramid (d = 0 everywhere, so
// occluded == max.z <= -1e-7).  Take the far-plane p-vertex corner and build its clip z / w with
// project_aabb's own operation order (X, then Y, then Z increments).  If Zc > 0 and Wc > 0 then either the
// projection fails (=> visible) or max.z >= Zc/Wc >= +0 or max.z is NaN; in every case `max.z <= -1e-7` is
// false.  Exact reasoning, no error bound needed.  Otherwise (box grazing the far plane): full evaluation.
OXC_DI bool cleared_hiz_surely_visible(const float4 r2, const float4 r3, float cx, float cy, float cz, float ex, float ey,
                                       float ez) {
  const float px = fs(cx, fm(ex, 0.5f)), py = fs(cy, fm(ey, 0.5f)), pz = fs(cz, fm(ez, 0.5f));
  float Zc = row_dot_p1(r2, px, py, pz), Wc = row_dot_p1(r3, px, py, pz);
  if (!(__float_as_uint(r2.x) & 0x80000000u)) { Zc = fa(Zc, fm(r2.x, ex)); Wc = fa(Wc, fm(r3.x, ex)); }
  if (!(__float_as_uint(r2.y) & 0x80000000u)) { Zc = fa(Zc, fm(r2.y, ey)); Wc = fa(Wc, fm(r3.y, ey)); }
  if (!(__float_as_uint(r2.z) & 0x80000000u)) { Zc = fa(Zc, fm(r2.z, ez)); Wc = fa(Wc, fm(r3.z, ez)); }
  return Zc > 0.0f && Wc > 0.0f;
}

} // namespace oxc
