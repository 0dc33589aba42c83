// device_vs_oracle.cpp — HOST cross-check of the kernels' canonical arithmetic (oxylus_b200/csrc/oxc_exact.cuh) against the CPU
// oracle (oracle/liboxc_oracle.so), bit for bit, without a GPU.
//
// The device headers are compiled for the host through tests/host_shim/ (each __f*_rn intrinsic = one IEEE binary32 operation
// under -ffp-contract=off — exactly what the intrinsics guarantee on the device), so this compares the PRODUCT'S SOURCE for
// dequantize_half, mat4 products, frustum plane extraction + test, project_aabb, test_occlusion, the canonical log2, the
// backface determinant and ceil(log2) with the oracle's independent C restatement of the Slang shaders on millions of random
// and special inputs.  The GPU parity tests compare the same functions as compiled by nvcc; this one runs in the CPU tier.
// Test infrastructure only (built and run by tests/test_device_source_cpu.py).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define OXC_HOST_SOUNDNESS_HARNESS
#include "oxc_filtered.cuh"

#include "../oracle/oxc_oracle.h"

namespace oxc { // unused here, but declared by the header
float rcp_approx(float x) { return 1.0f / x; }
float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
} // namespace oxc

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

uint64_t g_fail = 0;
void expect(bool ok, const char* what) {
  if (!ok) {
    if (g_fail < 10) std::fprintf(stderr, "MISMATCH: %s\n", what);
    g_fail++;
  }
}
bool same_bits(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

void random_matrix(Rng& r, float* m, int kind) {
  if (kind == 0) { // perspective x view-ish: the shapes the cull sees
    for (int i = 0; i < 16; i++) m[i] = 0.0f;
    const double f = 1.0 / tan(r.range(0.5, 1.4) * 0.5), zn = 0.1, zf = r.range(50, 3000);
    m[0] = (float)(f / r.range(1.0, 2.4)); m[5] = (float)-f; m[10] = (float)(zn / (zf - zn)); m[11] = -1.0f; m[14] = (float)(zf * zn / (zf - zn));
  } else if (kind == 1) { // rotation x scale + translation
    double q[4], n = 0;
    for (double& c : q) { c = r.range(-1, 1); n += c * c; }
    n = sqrt(n);
    for (double& c : q) c /= n;
    const double s = r.range(0.3, 3.0), x = q[0], y = q[1], z = q[2], w = q[3];
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                         2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    for (int c = 0; c < 3; c++)
      for (int rr = 0; rr < 3; rr++) m[c * 4 + rr] = (float)(R[rr * 3 + c] * s);
    m[3] = m[7] = m[11] = 0.0f;
    m[12] = (float)r.range(-50, 50); m[13] = (float)r.range(-20, 20); m[14] = (float)r.range(-100, 5); m[15] = 1.0f;
  } else { // anything
    for (int i = 0; i < 16; i++) m[i] = (float)r.range(-3, 3);
  }
}

void rows_to_colmajor(const float4 rows[4], float* m) {
  for (int i = 0; i < 4; i++) { m[0 * 4 + i] = rows[i].x; m[1 * 4 + i] = rows[i].y; m[2 * 4 + i] = rows[i].z; m[3 * 4 + i] = rows[i].w; }
}

float half_value(Rng& r) { return dequantize_half((uint32_t)r.below(65536)); }

} // namespace

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 300000ull;
  Rng r(0x0C115EEDull + 77);

  // ---- a1: dequantize_half, both device decoders, all 65536 inputs ----
  for (uint32_t h = 0; h < 65536; h++) {
    const float want = orc_dequantize_half((uint16_t)h);
    expect(same_bits(dequantize_half(h), want), "dequantize_half");
    expect(same_bits(dequantize_half_hw(h), want), "dequantize_half_hw");
  }
  // ---- ceil(log2) in integers, canonical log2 ----
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t v = i < 70000 ? (uint32_t)i : (uint32_t)r.next();
    expect(ceil_log2_u32(v) == orc_ceil_log2_u32(v), "ceil_log2_u32");
    float x = __uint_as_float((uint32_t)r.next());
    if (i % 4 == 0) x = (float)r.range(1e-3, 1e3);
    expect(same_bits(canonical_log2(x), orc_log2_canonical(x)), "canonical_log2");
  }
  const float specials[] = {0.0f, -0.0f, 1.0f, 2.0f, 0.5f, 1.41421354f, 1.41421366f, 1.17549435e-38f, 1e-45f, 3.4e38f, __uint_as_float(0x7F800000u),
                            __uint_as_float(0x7FC00000u), -1.0f};
  for (float x : specials) expect(same_bits(canonical_log2(x), orc_log2_canonical(x)), "canonical_log2 (special)");

  // ---- matrices, planes, projection, occlusion ----
  std::vector<float> hiz_data;
  OrcHiz hiz{};
  for (uint64_t i = 0; i < n; i++) {
    float a[16], b[16], want[16], got[16];
    random_matrix(r, a, (int)(i % 3 == 2 ? 2 : 0));
    random_matrix(r, b, (int)(i % 3 == 2 ? 2 : 1));
    float4 rows[4];
    mul_mm_rows(a, b, rows);
    rows_to_colmajor(rows, got);
    orc_mat4_mul(a, b, want);
    bool eq = true;
    for (int k = 0; k < 16; k++) eq = eq && same_bits(got[k], want[k]);
    expect(eq, "mul(projection_view, world)");

    float c[3], e[3];
    for (int k = 0; k < 3; k++) { c[k] = half_value(r); e[k] = fabsf(half_value(r)); }
    if (i % 3 == 0) for (int k = 0; k < 3; k++) { c[k] = (float)r.range(-4, 4); e[k] = (float)r.range(0.0, 2.0); }
    if (c[0] != c[0] || c[1] != c[1] || c[2] != c[2] || e[0] != e[0] || e[1] != e[1] || e[2] != e[2]) continue; // NaN halves: not bounds

    // frustum: per-instance hoisted planes (test_frustum_planes, what the meshlet cull runs) and the row form (cull_meshes)
    float4 planes[6];
    frustum_planes(rows, planes);
    const int want_fr = orc_test_frustum(want, c, e);
    expect((int)test_frustum_planes(planes, c[0], c[1], c[2], e[0], e[1], e[2]) == want_fr, "test_frustum (hoisted planes)");
    expect((int)test_frustum_rows(planes, c[0], c[1], c[2], e[0], e[1], e[2]) == want_fr, "test_frustum (rows)");

    // project_aabb
    const float near_clip = 0.1f;
    ScreenAabb sa{};
    OrcScreenAabb oa{};
    const bool got_ok = project_aabb(rows[0], rows[1], rows[2], rows[3], near_clip, c[0], c[1], c[2], e[0], e[1], e[2], sa);
    const int want_ok = orc_project_aabb(want, near_clip, c, e, &oa);
    expect((int)got_ok == want_ok, "project_aabb (none / some)");
    if (got_ok && want_ok) {
      expect(same_bits(sa.minx, oa.min[0]) && same_bits(sa.miny, oa.min[1]) && same_bits(sa.minz, oa.min[2]) && same_bits(sa.maxx, oa.max[0]) &&
                 same_bits(sa.maxy, oa.max[1]) && same_bits(sa.maxz, oa.max[2]),
             "project_aabb (screen box)");
      // test_occlusion against a random pyramid
      if (i % 2048 == 0 || hiz_data.empty()) {
        static const uint32_t sizes[][2] = {{1024, 1024}, {512, 256}, {64, 64}, {2048, 2048}, {2, 1}};
        const uint32_t* s = sizes[r.below(5)];
        orc_hiz_layout(s[0], s[1], &hiz);
        hiz_data.resize(orc_hiz_total_texels(s[0], s[1]));
        for (float& d : hiz_data) d = r.below(4) ? (float)r.range(0.0, 0.3) : 0.0f;
        hiz.data = hiz_data.data();
      }
      expect((int)test_occlusion(sa, hiz.data, hiz.width, hiz.height, hiz.levels, hiz.level_offset) == orc_test_occlusion(&oa, &hiz), "test_occlusion");
      // the same with a hand-made screen box: huge / inverted / off-screen rectangles (u32 wrap of max - min, clamps)
      OrcScreenAabb ob;
      ScreenAabb sb;
      for (int k = 0; k < 2; k++) { ob.min[k] = (float)r.range(-0.5, 1.5); ob.max[k] = (float)r.range(-0.5, 1.5); }
      ob.min[2] = 0.0f; ob.max[2] = (float)r.range(0.0, 0.3);
      sb.minx = ob.min[0]; sb.miny = ob.min[1]; sb.minz = ob.min[2]; sb.maxx = ob.max[0]; sb.maxy = ob.max[1]; sb.maxz = ob.max[2];
      expect((int)test_occlusion(sb, hiz.data, hiz.width, hiz.height, hiz.levels, hiz.level_offset) == orc_test_occlusion(&ob, &hiz), "test_occlusion (hand-made box)");
    }

    // backface determinant (cull.slang:169-171)
    float clip[3][4];
    for (auto& v : clip) for (float& x : v) x = (float)r.range(-2, 2);
    if (i % 4 == 0) for (int k = 0; k < 4; k++) clip[2][k] = clip[0][k] + (clip[1][k] - clip[0][k]) * 0.5f; // nearly degenerate
    const float4 c0 = make_float4(clip[0][0], clip[0][1], clip[0][2], clip[0][3]), c1 = make_float4(clip[1][0], clip[1][1], clip[1][2], clip[1][3]),
                 c2 = make_float4(clip[2][0], clip[2][1], clip[2][2], clip[2][3]);
    expect((int)triangle_backface(c0, c1, c2) == orc_test_triangle_backface(clip), "triangle_backface");
  }
  std::printf("%llu mismatches\n%s\n", (unsigned long long)g_fail, g_fail ? "FAILED" : "ok");
  return g_fail ? 1 : 0;
}
