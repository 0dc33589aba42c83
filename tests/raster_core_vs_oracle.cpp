// raster_core_vs_oracle.cpp — HOST cross-check of the software raster's per-triangle core (oxylus_b200/csrc/oxc_raster_core.cuh:
// snapping, set-up, 32- and 64-bit stepped edge functions, tie-break, depth interpolation, packed max, Sutherland-Hodgman
// clipping) against the oracle's specification (oracle/oxc_oracle.c raster_triangle / raster_triangle_clipped), pixel for
// pixel, without a GPU.
//
// The device header is compiled for the host through tests/host_shim/ (each __f*_rn intrinsic = one IEEE binary32 operation
// under -ffp-contract=off; atomicMax = max).  The few lines of warp scheduling around the core (which lane takes which
// triangle, the chunk queue for screen-filling triangles) are not part of this check — the GPU parity tests cover them; what
// is compared here is the arithmetic that decides WHICH pixels a triangle covers and WHAT value each gets.
// Test infrastructure only (built and run by tests/test_device_source_cpu.py).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  const unsigned long long old = *p;
  if (v > old) *p = v;
  return old;
}

#include "oxc_raster_core.cuh"

#include "../oracle/oxc_oracle.h"

namespace {
using namespace oxc;

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double range(double a, double b) { return a + (b - a) * uniform(); }
  uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
};

// what k_raster_visbuffer / k_raster_clip_queue do with one triangle that passed the near / backface test
// pixelwise = true: every pixel of the bounding box through raster_pixel (direct 64-bit edge functions) — the form the
// warp-cooperative and chunk-queue paths (k_raster_big) evaluate — instead of the incrementally stepped walk of raster_small
bool device_draw(const float clip[3][4], uint32_t data, uint32_t W, uint32_t H, unsigned long long* vis, bool pixelwise) {
  auto draw = [&](const TriSetup& s) {
    if (!pixelwise) { raster_small(s, data, vis, W); return; }
    for (int py = s.py0; py <= s.py1; py++)
      for (int px = s.px0; px <= s.px1; px++) raster_pixel(s, px, py, data, vis, W);
  };
  const float fW = (float)W, fH = (float)H;
  const float4 c0 = make_float4(clip[0][0], clip[0][1], clip[0][2], clip[0][3]), c1 = make_float4(clip[1][0], clip[1][1], clip[1][2], clip[1][3]),
               c2 = make_float4(clip[2][0], clip[2][1], clip[2][2], clip[2][3]);
  const ScreenVert v0 = to_screen(c0, fW, fH), v1 = to_screen(c1, fW, fH), v2 = to_screen(c2, fW, fH);
  TriSetup s;
  const int rc = tri_setup(v0, v1, v2, W, H, s);
  if (rc == TRI_DRAW) { draw(s); return false; }
  if (rc != TRI_INVALID_VERTEX) return false;
  // clip queue
  float4 poly[2][12];
  int cur;
  const int n = clip_polygon(c0, c1, c2, poly, cur);
  for (int i = 1; i + 1 < n; i++) {
    if (tri_setup(to_screen(poly[cur][0], fW, fH), to_screen(poly[cur][i], fW, fH), to_screen(poly[cur][i + 1], fW, fH), W, H, s) != TRI_DRAW) continue;
    s.narrow = false; // pieces may be large: 64-bit edge functions (kernels_tri.cuh clip_and_draw)
    draw(s);
  }
  return true;
}

void random_triangle(Rng& r, float clip[3][4], uint32_t W, uint32_t H, int kind) {
  // screen-space construction, then un-projected with a random w so the divide matters
  double cx = r.range(-0.1, 1.1) * W, cy = r.range(-0.1, 1.1) * H;
  double size;
  switch (kind) {
    case 0: size = r.range(0.05, 1.5); break;      // sub-pixel .. pixel
    case 1: size = r.range(1.0, 12.0); break;      // small
    case 2: size = r.range(10.0, 63.0); break;     // up to the 32-bit edge-function limit
    case 3: size = r.range(64.0, 400.0); break;    // 64-bit edge functions
    default: size = r.range(0.5, 30.0); break;
  }
  for (int k = 0; k < 3; k++) {
    double sx = cx + r.range(-size, size), sy = cy + r.range(-size, size);
    if (kind == 5) { sx = floor(sx) + (r.below(3) == 0 ? 0.5 : (double)r.below(256) / 256.0); sy = floor(sy) + (r.below(3) == 0 ? 0.5 : (double)r.below(256) / 256.0); } // on the grid / on sample centres
    const double w = kind == 6 ? r.range(-2.0, 4.0) : r.range(0.2, 50.0); // kind 6: around the camera (w <= 0 possible)
    const double z = r.range(0.0, 1.0) * (r.below(20) == 0 ? 1.2 : 1.0);   // some depths outside [0, 1]
    clip[k][0] = (float)((sx / W * 2.0 - 1.0) * w);
    clip[k][1] = (float)((sy / H * 2.0 - 1.0) * w);
    clip[k][2] = (float)(z * w);
    clip[k][3] = (float)w;
  }
  if (kind == 7) { // far outside the snap range on one vertex
    clip[r.below(3)][r.below(2)] *= 1e6f;
  }
}

} // namespace

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 200000ull;
  Rng r(0x0C115EEDull + 99);
  uint64_t mismatched_images = 0, compared = 0, clipped = 0, clip_disagree = 0, small_disagree = 0, drawn_pixels = 0;
  static const uint32_t sizes[][2] = {{160, 90}, {64, 48}, {257, 131}, {1, 1}, {640, 8}};
  std::vector<unsigned long long> img_dev, img_dev2;
  std::vector<uint64_t> img_orc;
  for (uint64_t batch = 0; batch * 256 < n; batch++) {
    const uint32_t* sz = sizes[batch % 5];
    const uint32_t W = sz[0], H = sz[1];
    img_dev.assign((size_t)W * H, 0);
    img_orc.assign((size_t)W * H, 0);
    orc_clear_visbuffer(img_orc.data(), W, H);
    for (size_t i = 0; i < img_dev.size(); i++) img_dev[i] = img_orc[i];
    img_dev2 = img_dev;
    float prev[3][4] = {};
    for (uint32_t t = 0; t < 256; t++) {
      float clip[3][4];
      const int kind = (int)r.below(8);
      random_triangle(r, clip, W, H, kind);
      if (t % 4 == 1) { // shares the edge (v0, v1) of the previous triangle with opposite winding: watertightness / no double hits
        for (int k = 0; k < 4; k++) { clip[0][k] = prev[1][k]; clip[1][k] = prev[0][k]; }
      }
      for (int a = 0; a < 3; a++) for (int k = 0; k < 4; k++) prev[a][k] = clip[a][k];
      const uint32_t data = (uint32_t)((batch * 256 + t) << 8) | (uint32_t)r.below(64);
      const bool dev_clip = device_draw(clip, data, W, H, img_dev.data(), false);
      device_draw(clip, data, W, H, img_dev2.data(), true);
      const int orc_clip = orc_raster_triangle(clip, data, W, H, img_orc.data());
      clipped += (uint64_t)orc_clip;
      if ((int)dev_clip != orc_clip) clip_disagree++;
      // the small-primitive predicate
      const float4 c0 = make_float4(clip[0][0], clip[0][1], clip[0][2], clip[0][3]), c1 = make_float4(clip[1][0], clip[1][1], clip[1][2], clip[1][3]),
                   c2 = make_float4(clip[2][0], clip[2][1], clip[2][2], clip[2][3]);
      if ((int)tri_covers_no_sample(c0, c1, c2, (float)W, (float)H, W, H) != orc_triangle_covers_no_sample(clip, W, H)) small_disagree++;
    }
    compared++;
    bool same = true;
    for (size_t i = 0; i < img_dev.size(); i++) {
      same = same && img_dev[i] == img_orc[i] && img_dev2[i] == img_orc[i];
      drawn_pixels += (uint32_t)img_orc[i] != 0xFFFFFFFFu;
    }
    if (!same) {
      if (mismatched_images < 3) std::fprintf(stderr, "MISMATCH: image of batch %llu (%u x %u)\n", (unsigned long long)batch, W, H);
      mismatched_images++;
    }
  }
  std::printf("%llu images of 256 triangles compared, %llu differ; %llu triangles took the clip path (%llu path disagreements); "
              "%llu small-primitive disagreements; %llu covered pixels in the final images\n",
              (unsigned long long)compared, (unsigned long long)mismatched_images, (unsigned long long)clipped, (unsigned long long)clip_disagree,
              (unsigned long long)small_disagree, (unsigned long long)drawn_pixels);
  const bool fail = mismatched_images || clip_disagree || small_disagree || drawn_pixels == 0 || clipped == 0;
  std::printf("%s\n", fail ? "FAILED" : "ok");
  return fail ? 1 : 0;
}
