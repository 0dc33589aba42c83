/*
 * oxc_oracle.c — CPU ORACLE (test infrastructure; see oxc_oracle.h for the rules and the
 * "parity unpinned" statement).  Plain C11, compiled with
 *   gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -shared
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/Oxylus/src/Render/Shaders unless noted).
 */
#define _GNU_SOURCE
#include "oxc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * scalar helpers
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* saturating, truncating conversions (PTX cvt.rzi.{u32,s32}.f32; NaN -> 0) */
static inline uint32_t to_u32_f32(float x) {
  if (!(x > 0.0f)) return 0u;
  if (x >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)x;
}
static inline int32_t to_i32_f32(float x) {
  if (x != x) return 0;
  if (x <= -2147483648.0f) return INT32_MIN;
  if (x >= 2147483648.0f) return INT32_MAX;
  return (int32_t)x;
}
static inline uint32_t to_u32_f64(double x) {
  if (!(x > 0.0)) return 0u;
  if (x >= 4294967296.0) return 0xFFFFFFFFu;
  return (uint32_t)x;
}
static inline int32_t to_i32_f64(double x) {
  if (x != x) return 0;
  if (x <= -2147483648.0) return INT32_MIN;
  if (x >= 2147483648.0) return INT32_MAX;
  return (int32_t)x;
}
static inline float rsqrt__f32(float x) { return sqrtf(x); } /* "real sqrt" */
static inline double rsqrt__f64(double x) { return sqrt(x); }
static inline float rfloor_f32(float x) { return floorf(x); }
static inline double rfloor_f64(double x) { return floor(x); }
/* asfloat(asuint(v) ^ (asuint(n) & 0x80000000)) — cull.slang:76-77 */
static inline float xor_sign_f32(float v, float n) { return bits2f(f2bits(v) ^ (f2bits(n) & 0x80000000u)); }
static inline double xor_sign_f64(double v, double n) { return signbit(n) ? -v : v; }

uint32_t orc_ceil_log2_u32(uint32_t n) {
  if (n <= 1u) return 0u; /* log2(0) = -inf -> clamp 0; log2(1) = 0 */
  return 32u - (uint32_t)__builtin_clz(n - 1u);
}

#define REAL float
#define SUF(x) x##_f32
#include "oxc_oracle_math.inc"
#undef REAL
#undef SUF

#define REAL double
#define SUF(x) x##_f64
#include "oxc_oracle_math.inc"
#undef REAL
#undef SUF

/* ------------------------------------------------------------------------------------------------
 * common/math.slang:193-201 dequantize_half
 * ---------------------------------------------------------------------------------------------- */
float orc_dequantize_half(uint16_t h) {
  uint32_t s = (uint32_t)(h & 0x8000) << 16;
  int32_t em = h & 0x7fff;
  int32_t r = (int32_t)((uint32_t)(em + (112 << 10)) << 13);
  r = (em < (1 << 10)) ? 0 : r;                /* denormals flush to zero */
  r += (em >= (31 << 10)) ? (112 << 23) : 0;   /* inf / nan exponent fix-up */
  return bits2f(s | (uint32_t)r);
}

/* scene.slang:401-435 */
void orc_bounds_decode(const OxcMeshletBounds* b, float center[3], float extent[3], float cone_axis[3],
                       float* cone_cutoff) {
  for (int i = 0; i < 3; i++) {
    center[i] = orc_dequantize_half(b->aabb_center[i]);
    extent[i] = orc_dequantize_half(b->aabb_extent[i]);
  }
  cone_axis[0] = (float)(int32_t)b->cone_axis_xy[0] / 127.0f;
  cone_axis[1] = (float)(int32_t)b->cone_axis_xy[1] / 127.0f;
  cone_axis[2] = (float)(int32_t)b->cone_axis_z / 127.0f;
  *cone_cutoff = (float)b->cone_cutoff / 127.0f;
}

void orc_mat4_mul(const float a[16], const float b[16], float out[16]) { mul_mm_f32(a, b, out); }

int orc_project_aabb(const float mvp[16], float near_clip, const float c[3], const float e[3], OrcScreenAabb* out) {
  Vec3_f32 cc = {c[0], c[1], c[2]}, ee = {e[0], e[1], e[2]};
  ScreenAabb_f32 sa;
  if (!project_aabb_f32(mvp, near_clip, cc, ee, &sa)) return 0;
  memcpy(out->min, sa.min, sizeof sa.min);
  memcpy(out->max, sa.max, sizeof sa.max);
  return 1;
}

int orc_test_frustum(const float mvp[16], const float c[3], const float e[3]) {
  Vec3_f32 cc = {c[0], c[1], c[2]}, ee = {e[0], e[1], e[2]};
  return test_frustum_f32(mvp, cc, ee);
}

int orc_test_occlusion(const OrcScreenAabb* aabb, const OrcHiz* hiz) {
  ScreenAabb_f32 sa;
  memcpy(sa.min, aabb->min, sizeof sa.min);
  memcpy(sa.max, aabb->max, sizeof sa.max);
  return test_occlusion_f32(&sa, hiz);
}

int orc_test_cone(const float center[3], float radius, const float axis[3], float cutoff, const float cam[3]) {
  Vec3_f32 c = {center[0], center[1], center[2]}, a = {axis[0], axis[1], axis[2]}, p = {cam[0], cam[1], cam[2]};
  return test_cone_f32(c, radius, a, cutoff, p);
}

/* cull.slang:177-179 */
int orc_test_cone_directional(const float axis[3], float cutoff, const float view_dir[3]) {
  Vec3_f32 a = {axis[0], axis[1], axis[2]}, v = {view_dir[0], view_dir[1], view_dir[2]};
  return dot3_f32(a, v) >= cutoff;
}

/* cull.slang:169-171: determinant(f32x3x3(c0.xyw, c1.xyw, c2.xyw)) >= 0.0001 */
static inline float det3_f32(const float m[3][3]) {
  return (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}
int orc_test_triangle_backface(const float clip[3][4]) {
  float m[3][3];
  for (int i = 0; i < 3; i++) { m[i][0] = clip[i][0]; m[i][1] = clip[i][1]; m[i][2] = clip[i][3]; }
  return det3_f32(m) >= 0.0001f;
}

/* ------------------------------------------------------------------------------------------------
 * Hi-Z layout — Oxylus/include/Asset/Texture.hpp:144-146 (calculate_mip_count) and
 * RendererInstance.cpp:583-586 (min(., 13))
 * ---------------------------------------------------------------------------------------------- */
uint32_t orc_hiz_level_count(uint32_t w, uint32_t h) {
  uint32_t m = w > h ? w : h;
  uint32_t levels = 0;
  while (m) { levels++; m >>= 1; } /* floor(log2(m)) + 1 */
  return levels < OXC_HIZ_MAX_LEVELS ? levels : OXC_HIZ_MAX_LEVELS;
}
void orc_hiz_layout(uint32_t w, uint32_t h, OrcHiz* hiz) {
  hiz->width = w;
  hiz->height = h;
  hiz->levels = orc_hiz_level_count(w, h);
  uint32_t off = 0;
  for (uint32_t l = 0; l < OXC_HIZ_MAX_LEVELS; l++) {
    hiz->level_offset[l] = off;
    if (l < hiz->levels) {
      uint32_t mw = w >> l, mh = h >> l;
      if (mw < 1) mw = 1;
      if (mh < 1) mh = 1;
      off += mw * mh;
    }
  }
}
uint32_t orc_hiz_total_texels(uint32_t w, uint32_t h) {
  OrcHiz t;
  orc_hiz_layout(w, h, &t);
  uint32_t l = t.levels - 1;
  uint32_t mw = w >> l, mh = h >> l;
  if (mw < 1) mw = 1;
  if (mh < 1) mh = 1;
  return t.level_offset[l] + mw * mh;
}

/* passes/hiz.slang:92-95 load(): uv = (texel + 1) / hiz_extent sampled with NearestSamplerClamped from a
 * depth image of a different size  =>  source texel = min(W-1, floor((x+1) * W / hizW)); exact in
 * integers because hiz extents are powers of two.  mip k: hiz.slang:77-83,137-166. */
void orc_build_hiz(const float* depth, uint32_t width, uint32_t height, OrcHiz* hiz) {
  const uint32_t hw = hiz->width, hh = hiz->height;
  float* m0 = hiz->data + hiz->level_offset[0];
  for (uint32_t y = 0; y < hh; y++) {
    uint64_t sy = ((uint64_t)(y + 1) * height) / hh;
    if (sy > height - 1) sy = height - 1;
    for (uint32_t x = 0; x < hw; x++) {
      uint64_t sx = ((uint64_t)(x + 1) * width) / hw;
      if (sx > width - 1) sx = width - 1;
      m0[(size_t)y * hw + x] = depth[(size_t)sy * width + sx]; /* transform_z is identity, CullGeometry.cpp:44 */
    }
  }
  for (uint32_t l = 1; l < hiz->levels; l++) {
    uint32_t pw = hw >> (l - 1), ph = hh >> (l - 1);
    if (pw < 1) pw = 1;
    if (ph < 1) ph = 1;
    uint32_t mw = hw >> l, mh = hh >> l;
    if (mw < 1) mw = 1;
    if (mh < 1) mh = 1;
    const float* src = hiz->data + hiz->level_offset[l - 1];
    float* dst = hiz->data + hiz->level_offset[l];
    for (uint32_t y = 0; y < mh; y++)
      for (uint32_t x = 0; x < mw; x++) {
        uint32_t x0 = 2 * x, x1 = 2 * x + 1, y0 = 2 * y, y1 = 2 * y + 1;
        if (x1 > pw - 1) x1 = pw - 1; /* only reachable for non-square pyramids */
        if (y1 > ph - 1) y1 = ph - 1;
        if (x0 > pw - 1) x0 = pw - 1;
        if (y0 > ph - 1) y0 = ph - 1;
        float a = src[(size_t)y0 * pw + x0], b = src[(size_t)y0 * pw + x1];
        float c = src[(size_t)y1 * pw + x0], d = src[(size_t)y1 * pw + x1];
        float ab = a < b ? a : b, cd = c < d ? c : d; /* reduce(): hiz.slang:77-83 */
        dst[(size_t)y * mw + x] = ab < cd ? ab : cd;
      }
  }
}

/* ------------------------------------------------------------------------------------------------
 * scene access (blob offsets instead of device addresses)
 * ---------------------------------------------------------------------------------------------- */
static inline const OxcMeshLOD* mesh_lod(const OrcScene* s, const OxcMesh* m, uint32_t lod) {
  return (const OxcMeshLOD*)(s->blob + m->lods) + lod;
}
static inline const OxcMeshletBounds* lod_bounds(const OrcScene* s, const OxcMeshLOD* l) {
  return (const OxcMeshletBounds*)(s->blob + l->meshlet_bounds);
}

/* ------------------------------------------------------------------------------------------------
 * passes/cull_meshes.slang:17-85
 * ---------------------------------------------------------------------------------------------- */
void orc_cull_meshes(const OrcScene* scene, const OxcCullCamera* cam, uint32_t flags, uint32_t first, uint32_t count,
                     OxcMeshletInstance* meshlet_instances, OxcMeshletInstanceVisibility* vis,
                     OxcDispatchIndirectCommand* cull_meshlets_cmd) {
  /* scratch_buffer init: CullGeometry.cpp:97-100 */
  vis->total_visible_meshlet_instances = 0;
  vis->early_visible_meshlet_instances = 0;
  vis->late_visible_meshlet_instances = 0;
  cull_meshlets_cmd->x = 0; cull_meshlets_cmd->y = 1; cull_meshlets_cmd->z = 1;

  OxcMeshInstance* mesh_instances = (OxcMeshInstance*)scene->mesh_instances; /* RWStructuredBuffer :13 */
  uint32_t n = cam->mesh_instance_count;                                      /* :28 */
  uint32_t lo = first, hi = n;
  if (count != 0xFFFFFFFFu && first + count < hi) hi = first + count;
  for (uint32_t mi = lo; mi < hi; mi++) {
    uint32_t meshlet_count = 0, lod_index = 0;
    const OxcMeshInstance inst = mesh_instances[mi];
    const OxcMesh* mesh = &scene->meshes[inst.mesh_index];
    const float* world = scene->transforms[inst.transform_index].world;
    float mvp[16];
    mul_mm_f32(cam->projection_view, world, mvp); /* :32 */
    Vec3_f32 bc = {mesh->bounds.aabb_center[0], mesh->bounds.aabb_center[1], mesh->bounds.aabb_center[2]};
    Vec3_f32 be = {mesh->bounds.aabb_extent[0], mesh->bounds.aabb_extent[1], mesh->bounds.aabb_extent[2]};
    if ((flags & OXC_CULL_TEST_FRUSTUM) && test_frustum_f32(mvp, bc, be)) { /* :34 */
      if (flags & OXC_CULL_SELECT_LOD) {                                     /* :35-57 */
        Vec4_f32 c4 = {bc.x, bc.y, bc.z, 1.0f}, e4 = {be.x, be.y, be.z, 0.0f};
        Vec4_f32 wc = mul_mv_f32(world, c4), we = mul_mv_f32(world, e4);
        Vec3_f32 aabb_extent = {fabsf(we.x), fabsf(we.y), fabsf(we.z)};
        float rough_extent = rmax_f32(aabb_extent.x, rmax_f32(aabb_extent.y, aabb_extent.z));
        Vec3_f32 d = {wc.x - cam->position[0], wc.y - cam->position[1], wc.z - cam->position[2]};
        float rough_dist = rmax_f32(length3_f32(d) - 0.5f * rough_extent, 0.0f);
        const float fov90_distance_to_screen_ratio = 2.0f;
        float pixel_size_at_1m = fov90_distance_to_screen_ratio / rmax_f32(cam->resolution[0], cam->resolution[1]);
        float aabb_size_at_1m = rough_extent / rough_dist;
        float rough_aabb_pixel_size = aabb_size_at_1m / pixel_size_at_1m;
        for (uint32_t i = 1; i < mesh->lod_count; i++) {
          float rough_pixel_error = rough_aabb_pixel_size * mesh_lod(scene, mesh, i)->error;
          if (rough_pixel_error < cam->acceptable_lod_error) lod_index = i;
          else break;
        }
      }
      meshlet_count = mesh_lod(scene, mesh, lod_index)->meshlet_count; /* :59 */
    }
    /* :63-84 — serial order == wave order */
    uint32_t base = vis->total_visible_meshlet_instances;
    vis->total_visible_meshlet_instances += meshlet_count;
    uint32_t needed = (vis->total_visible_meshlet_instances + 64u - 1u) / 64u; /* CULLING_MESHLET_COUNT */
    if (needed > cull_meshlets_cmd->x) cull_meshlets_cmd->x = needed;
    if (meshlet_count > 0) {
      mesh_instances[mi].lod_index = lod_index; /* :76 */
      for (uint32_t i = 0; i < meshlet_count; i++) {
        meshlet_instances[base + i].mesh_instance_index = mi;
        meshlet_instances[base + i].meshlet_index = i;
      }
    }
  }
}

/* gather everything one meshlet-instance needs (cull_meshlets_hiz.slang:30-40) */
typedef struct MeshletCtx {
  const float* world;
  OxcMeshInstance inst;
  Vec3_f32 c, e, axis;
  float cutoff;
} MeshletCtx;

static inline void fetch_meshlet(const OrcScene* s, OxcMeshletInstance mi, MeshletCtx* o) {
  o->inst = s->mesh_instances[mi.mesh_instance_index];
  o->world = s->transforms[o->inst.transform_index].world;
  const OxcMesh* mesh = &s->meshes[o->inst.mesh_index];
  const OxcMeshLOD* lod = mesh_lod(s, mesh, o->inst.lod_index);
  const OxcMeshletBounds* b = lod_bounds(s, lod) + mi.meshlet_index;
  float c[3], e[3], a[3];
  orc_bounds_decode(b, c, e, a, &o->cutoff);
  o->c = (Vec3_f32){c[0], c[1], c[2]};
  o->e = (Vec3_f32){e[0], e[1], e[2]};
  o->axis = (Vec3_f32){a[0], a[1], a[2]};
}

/* ------------------------------------------------------------------------------------------------
 * passes/cull_meshlets_hiz.slang:19-88
 * ---------------------------------------------------------------------------------------------- */
void orc_cull_meshlets_hiz(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, const OxcCullCamera* cam,
                           uint32_t flags, const OrcHiz* hiz, OxcMeshletInstanceVisibility* vis,
                           uint32_t* visible_indices, uint32_t* mask, OxcDispatchIndirectCommand* cull_triangles_cmd) {
  cull_triangles_cmd->x = 0; cull_triangles_cmd->y = 1; cull_triangles_cmd->z = 1; /* CullGeometry.cpp:125-127 */
  const Vec3_f32 campos = {cam->position[0], cam->position[1], cam->position[2]};
  const uint32_t total = vis->total_visible_meshlet_instances;
  for (uint32_t i = 0; i < total; i++) { /* :25-28 */
    MeshletCtx m;
    fetch_meshlet(scene, meshlet_instances[i], &m);
    uint32_t mask_index = 0, bit = 0;
    int was_visible = 1;
    if (flags & OXC_CULL_TEST_OCCLUSION) { /* :45-51 */
      uint32_t vi = m.inst.meshlet_instance_visibility_offset + meshlet_instances[i].meshlet_index;
      mask_index = vi / 32;
      bit = 1u << (vi - mask_index * 32);
      was_visible = (mask[mask_index] & bit) != 0;
    }
    int visible = meshlet_visible_hiz_f32(cam->projection_view, m.world, cam->near_clip, campos, m.c, m.e, m.axis,
                                          m.cutoff, flags, was_visible, hiz);
    if (visible && (!(flags & OXC_CULL_LATE_PASS) || !was_visible)) { /* :67-79 */
      uint32_t index;
      if (!(flags & OXC_CULL_LATE_PASS)) index = vis->early_visible_meshlet_instances++;
      else index = (vis->late_visible_meshlet_instances++) + vis->early_visible_meshlet_instances;
      visible_indices[index] = i;
      cull_triangles_cmd->x++;
    }
    if (flags & (OXC_CULL_TEST_OCCLUSION | OXC_CULL_LATE_PASS)) { /* :81-87 */
      if (visible) mask[mask_index] |= bit;
      else mask[mask_index] &= ~bit;
    }
  }
}

static void meshlet_flags_generic(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                  const OxcCullCamera* cam, uint32_t flags, const OrcHiz* hiz,
                                  const OxcMeshletInstanceVisibility* vis, const uint32_t* mask_in, uint8_t* out,
                                  int use_f64) {
  const uint32_t total = vis->total_visible_meshlet_instances;
  for (uint32_t i = 0; i < total; i++) {
    MeshletCtx m;
    fetch_meshlet(scene, meshlet_instances[i], &m);
    int was_visible = 1;
    if (flags & OXC_CULL_TEST_OCCLUSION) {
      uint32_t vi = m.inst.meshlet_instance_visibility_offset + meshlet_instances[i].meshlet_index;
      was_visible = (mask_in[vi / 32] >> (vi & 31)) & 1;
    }
    if (use_f64) {
      double pv[16], w[16];
      for (int k = 0; k < 16; k++) { pv[k] = cam->projection_view[k]; w[k] = m.world[k]; }
      Vec3_f64 campos = {cam->position[0], cam->position[1], cam->position[2]};
      Vec3_f64 c = {m.c.x, m.c.y, m.c.z}, e = {m.e.x, m.e.y, m.e.z};
      /* cone axis / cutoff re-derived in f64 from the s8 values would differ from the f32 decode only in the
       * division rounding; keep the f32-decoded inputs so only the predicate arithmetic is re-evaluated */
      Vec3_f64 ax = {m.axis.x, m.axis.y, m.axis.z};
      out[i] = (uint8_t)meshlet_visible_hiz_f64(pv, w, (double)cam->near_clip, campos, c, e, ax, (double)m.cutoff,
                                                flags, was_visible, hiz);
    } else {
      Vec3_f32 campos = {cam->position[0], cam->position[1], cam->position[2]};
      out[i] = (uint8_t)meshlet_visible_hiz_f32(cam->projection_view, m.world, cam->near_clip, campos, m.c, m.e, m.axis,
                                                m.cutoff, flags, was_visible, hiz);
    }
  }
}

void orc_cull_meshlets_hiz_f64(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                               const OxcCullCamera* cam, uint32_t flags, const OrcHiz* hiz,
                               const OxcMeshletInstanceVisibility* vis, const uint32_t* mask_in, uint8_t* out_visible) {
  meshlet_flags_generic(scene, meshlet_instances, cam, flags, hiz, vis, mask_in, out_visible, 1);
}
void orc_cull_meshlets_hiz_f32_flags(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                     const OxcCullCamera* cam, uint32_t flags, const OrcHiz* hiz,
                                     const OxcMeshletInstanceVisibility* vis, const uint32_t* mask_in,
                                     uint8_t* out_visible) {
  meshlet_flags_generic(scene, meshlet_instances, cam, flags, hiz, vis, mask_in, out_visible, 0);
}

/* ------------------------------------------------------------------------------------------------
 * passes/cull_meshlets.slang:21-73
 * ---------------------------------------------------------------------------------------------- */
void orc_cull_meshlets(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, const OxcCullCamera* cam,
                       OxcMeshletInstanceVisibility* vis, uint32_t* visible_indices,
                       OxcDispatchIndirectCommand* cull_triangles_cmd) {
  cull_triangles_cmd->x = 0; cull_triangles_cmd->y = 1; cull_triangles_cmd->z = 1;
  const Vec3_f32 campos = {cam->position[0], cam->position[1], cam->position[2]};
  const uint32_t total = vis->total_visible_meshlet_instances;
  for (uint32_t i = 0; i < total; i++) {
    MeshletCtx m;
    fetch_meshlet(scene, meshlet_instances[i], &m);
    float mvp[16];
    mul_mm_f32(cam->projection_view, m.world, mvp);                                     /* :40 */
    int cone_vis = cone_visible_f32(m.world, m.c, m.e, m.axis, m.cutoff, campos);      /* :49-52 */
    if (cone_vis && test_frustum_f32(mvp, m.c, m.e))                                    /* :54 */
      visible_indices[cull_triangles_cmd->x++] = i;                                     /* :55-70 */
  }
}

/* ------------------------------------------------------------------------------------------------
 * multi-view batched cull: per view the cone + frustum part of cull_meshlets_hpb.slang:39-60
 * (directional: test_cone_directional with view_dir = camera.position, :54) or of
 * cull_meshlets.slang:49-54 (positional)
 * ---------------------------------------------------------------------------------------------- */
void orc_cull_meshlets_multiview(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, uint32_t total,
                                 const OxcCullCamera* views, uint32_t n_views, int directional, uint32_t* view_bits,
                                 uint32_t* view_counts) {
  for (uint32_t v = 0; v < OXC_MAX_VIEWS; v++) view_counts[v] = 0;
  for (uint32_t i = 0; i < total; i++) {
    MeshletCtx m;
    fetch_meshlet(scene, meshlet_instances[i], &m);
    uint32_t bits = 0;
    for (uint32_t v = 0; v < n_views; v++) {
      const OxcCullCamera* cam = &views[v];
      const Vec3_f32 campos = {cam->position[0], cam->position[1], cam->position[2]};
      int cone_vis;
      if (directional) {
        if (m.cutoff >= 1.0f) cone_vis = 1;
        else {
          Vec3_f32 axis = world_cone_axis_f32(m.world, m.axis);
          cone_vis = !(dot3_f32(axis, campos) >= m.cutoff);
        }
      } else {
        cone_vis = cone_visible_f32(m.world, m.c, m.e, m.axis, m.cutoff, campos);
      }
      if (!cone_vis) continue;
      float mvp[16];
      mul_mm_f32(cam->projection_view, m.world, mvp);
      if (test_frustum_f32(mvp, m.c, m.e)) { bits |= 1u << v; view_counts[v]++; }
    }
    view_bits[i] = bits;
  }
}

/* ------------------------------------------------------------------------------------------------
 * triangles: scene.slang:336-382 (Meshlet::indices / positions), :478-484 (decode_position),
 * passes/cull_triangles.slang:27-90
 * ---------------------------------------------------------------------------------------------- */
typedef struct TriMeshlet {
  const OxcMesh* mesh;
  const OxcMeshLOD* lod;
  OxcMeshlet meshlet;
  float mvp[16];
} TriMeshlet;

static void fetch_tri_meshlet(const OrcScene* s, const OxcMeshletInstance* mis, uint32_t meshlet_instance_index,
                              const OxcCullCamera* cam, TriMeshlet* t) {
  OxcMeshletInstance mi = mis[meshlet_instance_index];               /* :45 */
  OxcMeshInstance inst = s->mesh_instances[mi.mesh_instance_index];  /* :46 */
  t->mesh = &s->meshes[inst.mesh_index];                             /* :47 */
  t->lod = mesh_lod(s, t->mesh, inst.lod_index);                     /* :48 */
  t->meshlet = ((const OxcMeshlet*)(s->blob + t->lod->meshlets))[mi.meshlet_index]; /* :49 */
  mul_mm_f32(cam->projection_view, s->transforms[inst.transform_index].world, t->mvp); /* :51-52 */
}

/* scene.slang:336-342 get_micro_index */
static inline uint32_t micro_index(const uint32_t* buf, uint32_t byte_offset) {
  uint32_t pack = buf[byte_offset >> 2];
  return (pack >> ((byte_offset & 3) * 8)) & 0xFF;
}

/* clip positions of triangle `tri` (cull_triangles.slang:60-66); returns vertex clip coords */
static void tri_clip(const OrcScene* s, const TriMeshlet* t, uint32_t tri, float clip[3][4]) {
  const uint32_t* micro = (const uint32_t*)(s->blob + t->lod->local_triangle_indices);
  const uint32_t* vidx = (const uint32_t*)(s->blob + t->lod->indirect_vertex_indices);
  const uint16_t* pos = (const uint16_t*)(s->blob + t->mesh->vertex_positions);
  uint32_t base = t->meshlet.local_triangle_index_offset + tri * 3; /* scene.slang:366 */
  for (int c = 0; c < 3; c++) {
    uint32_t local = micro_index(micro, base + (uint32_t)c);
    uint32_t v = vidx[t->meshlet.indirect_vertex_index_offset + local];
    Vec4_f32 p = {orc_dequantize_half(pos[v * 4 + 0]), orc_dequantize_half(pos[v * 4 + 1]),
                  orc_dequantize_half(pos[v * 4 + 2]), 1.0f};
    Vec4_f32 cp = mul_mv_f32(t->mvp, p);
    clip[c][0] = cp.x; clip[c][1] = cp.y; clip[c][2] = cp.z; clip[c][3] = cp.w;
  }
}

static inline int tri_passes(const float clip[3][4]) {
  int passed = clip[0][2] >= 0.0f && clip[1][2] >= 0.0f && clip[2][2] >= 0.0f; /* :68 */
  return passed && !orc_test_triangle_backface(clip);                            /* :69 */
}

void orc_cull_triangles(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                        const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                        const OxcCullCamera* cam, uint32_t id_base, uint32_t* reordered_indices,
                        OxcDrawIndexedIndirectCommand* draw_cmd) {
  /* CullGeometry.cpp:380-382 */
  draw_cmd->index_count = 0; draw_cmd->instance_count = 1; draw_cmd->first_index = 0;
  draw_cmd->vertex_offset = 0; draw_cmd->first_instance = 0;
  for (uint32_t g = 0; g < pass_count; g++) { /* one workgroup per surviving meshlet, :34-37 */
    uint32_t mii = visible_indices[pass_first + g];
    TriMeshlet t;
    fetch_tri_meshlet(scene, meshlet_instances, mii, cam, &t);
    for (uint32_t tri = 0; tri < t.meshlet.triangle_count && tri < 64; tri++) { /* :59 */
      float clip[3][4];
      tri_clip(scene, &t, tri, clip);
      if (!tri_passes(clip)) continue;
      uint32_t off = draw_cmd->index_count; /* :78,84 */
      uint32_t masked = (mii + id_base) << OXC_VIS_PRIMITIVE_BITS; /* :85 */
      uint32_t ti = tri * 3;
      reordered_indices[off + 0] = masked | ((ti + 0) & OXC_VIS_PRIMITIVE_MASK);
      reordered_indices[off + 1] = masked | ((ti + 1) & OXC_VIS_PRIMITIVE_MASK);
      reordered_indices[off + 2] = masked | ((ti + 2) & OXC_VIS_PRIMITIVE_MASK);
      draw_cmd->index_count += 3;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * SW raster — SURVEY §8a row R.  No reference implementation exists (the reference uses the HW
 * rasteriser, DrawGeometry.cpp:104-190); this is the SPECIFICATION the CUDA raster follows:
 *   1. clip = mvp * (pos,1) per vertex (visbuffer_encode_ms.slang:135-137); triangle kept iff
 *      cull_triangles' test passes (near z>=0 on all three, det(xyw) < 1e-4).
 *   2. triangles with any w <= 0 are dropped (no homogeneous clipping; documented limitation).
 *   3. rw = 1/w; ndc = xyz*rw; screen = (ndc.xy*0.5+0.5)*(W,H); fixed point fx = (int)floor(s*256+0.5),
 *      dropped if any |fx| > 2^22.
 *   4. area2 (int64, orient2d) must be < 0 (front face == negative clip-space determinant, the sign
 *      cull.slang:169-171 keeps); vertices 1,2 are swapped so the edge functions are >= 0 inside.
 *   5. sample at pixel centres (px*256+128); edge tie-break: a pixel exactly on edge a->b is inside
 *      iff (dy > 0) || (dy == 0 && dx < 0)  (watertight, no double hits).
 *   6. depth z = (za + (float)E_b * dzb) + (float)E_c * dzc with the per-triangle gradients
 *      dzb = (zb - za) / (float)area2, dzc = (zc - za) / (float)area2 (a,b,c = the re-oriented vertices,
 *      E_b / E_c = edge functions opposite b / c, int64 -> float round-to-nearest); fragments with
 *      z outside [0,1] are clipped (near/far); -0.0 is stored as +0.0.  value = asuint(z)<<32 | (id<<8 | tri); max wins
 *      (reverse-Z GreaterOrEqual; equal depth -> larger id, deterministic).
 * ---------------------------------------------------------------------------------------------- */
void orc_clear_visbuffer(uint64_t* vis, uint32_t width, uint32_t height) {
  for (size_t i = 0; i < (size_t)width * height; i++) vis[i] = (uint64_t)OXC_VIS_CLEAR; /* depth 0 | ~0u */
}

static inline int64_t orient2d(int64_t ax, int64_t ay, int64_t bx, int64_t by, int64_t cx, int64_t cy) {
  return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}
static inline int edge_bias(int64_t ax, int64_t ay, int64_t bx, int64_t by) {
  int64_t dx = bx - ax, dy = by - ay;
  return ((dy > 0) || (dy == 0 && dx < 0)) ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------------
 * Alpha-tested discard of the vis-buffer encode — visbuffer_encode.slang:54-66:
 *     if (material.flags & HasAlbedoImage) {
 *       alpha = material.sample_albedo_color(grad).a        // scene.slang:115-124: albedo_color * image.SampleGrad(sampler, uv)
 *       if (alpha < clamp(material.get_alpha_cutoff(), 0.001, 1.0)) discard;
 *     }
 * uv is the rasteriser's perspective-correct interpolation of the vertex shader's tex_coord (:36,:44; (0,0) for a mesh
 * without texture coordinates, scene.slang:355-361).  Interpolation and sampling happen in fixed-function hardware whose
 * arithmetic is not specified bit for bit, so THIS is the specification the CUDA raster follows (include/oxcull.h repeats it):
 *   1. per vertex of the drawn triangle: rw = 1 / w (the value of raster step 3) and (u, v).
 *   2. at a covered sample the raster's own integer edge functions E_a, E_b, E_c (step 5; exact, >= 0, sum = 2*area > 0) are
 *      the screen-space barycentric weights; perspective correction the way a hardware rasteriser does it:
 *        p_i = (float)E_i * rw_i,  l_i = p_i * (1 / ((p_a + p_b) + p_c)),  u = (l_a*u_a + l_b*u_b) + l_c*u_c, v likewise
 *      — every term is non-negative: no cancellation, however small or thin the triangle.
 *   3. triangles that take the clip path: Sutherland-Hodgman carries (u, v) with the position (clip space is linear in the
 *      attributes): a cut vertex gets uv = uv_I + t * (uv_O - uv_I) with the same t, same operation order; every piece of the
 *      fan is then an ordinary triangle for steps 1-2.
 *   4. the texel fetch: filter / address / mipmap modes of the material's sampler (default linear, linear, repeat — Texture.hpp:
 *      38-45; no anisotropy, no LOD bias or clamp: the reference sets none).  One level — linear: x = u*width - 0.5, x0 = floor(x),
 *      wx = x - x0, texels x0 and x0 + 1 (wrapped), top = a00 + wx*(a10 - a00), bot likewise, a = top + wy*(bot - top); nearest:
 *      texel floor(u*width); alpha of a texel = byte / 255; float -> int saturates, NaN -> 0.
 *      Level selection (images with a mip chain, or samplers whose mag and min filters differ) — SampleGrad with
 *      ddx / ddy(tex_coord) (visbuffer_encode.slang:57-60) and the isotropic LOD rule of the Vulkan specification:
 *        derivatives are the "fine" quad differences: with qx = px & ~1, qy = py & ~1,
 *          ddx(uv) = uv(qx + 1, py) - uv(qx, py),  ddy(uv) = uv(px, qy + 1) - uv(px, qy),
 *        uv(.) = steps 1-2 evaluated with the edge functions stepped to that pixel (exact integers; outside the triangle the same
 *        formula extrapolates, as helper invocations do);
 *        mx = ddx(u)*width0, my = ddx(v)*height0, rho_x^2 = mx*mx + my*my, rho_y^2 likewise, rho^2 = (rho_x^2 > rho_y^2 ? rho_x^2 :
 *        rho_y^2), lambda = 0.5 * log2_canonical(rho^2)  (orc_log2_canonical; NaN / 0 -> very negative);
 *        lambda > 0 selects the min filter, else the mag filter; lc = !(lambda > 0) ? 0 : min(lambda, levels - 1);
 *        mipmap linear: d = floor(lc), e = min(d + 1, levels - 1), f = lc - d, a = (1 - f)*a_d + f*a_e;
 *        mipmap nearest: level (lc <= 0.5 ? 0 : ceil(lc + 0.5) - 1).
 *      Level l is max(1, width0 >> l) x max(1, height0 >> l) texels and follows level l - 1 in memory, tightly packed.
 *   5. keep iff !(albedo_color.a * a < cutoff), cutoff = clamp(dequantize_half(alpha_cutoff), 0.001, 1.0) (NaN stays NaN).
 * ---------------------------------------------------------------------------------------------- */
typedef struct OrcAlphaMaterial {
  const uint8_t* texels;
  uint32_t width, height, format, levels, mag_filter, min_filter, mipmap_mode, address_u, address_v;
  float albedo_a, cutoff;
} OrcAlphaMaterial;

static inline int alpha_f2i(float f) {
  return !(f == f) ? 0 : (f >= 2147483648.0f ? 2147483647 : (f <= -2147483648.0f ? (-2147483647 - 1) : (int)f));
}
static inline uint32_t alpha_wrap(int64_t i, uint32_t n, uint32_t mode) {
  const int64_t N = (int64_t)n;
  if (mode == OXC_ADDRESS_CLAMP_TO_EDGE) return (uint32_t)(i < 0 ? 0 : (i > N - 1 ? N - 1 : i));
  if (mode == OXC_ADDRESS_MIRRORED_REPEAT) {
    int64_t r = i % (2 * N);
    if (r < 0) r += 2 * N;
    return (uint32_t)(r < N ? r : 2 * N - 1 - r);
  }
  int64_t r = i % N;
  if (r < 0) r += N;
  return (uint32_t)r;
}
static inline float alpha_texel(const OrcAlphaMaterial* a, const uint8_t* base, uint32_t w, uint32_t x, uint32_t y) {
  const size_t i = (size_t)y * w + x;
  const uint32_t t = a->format == OXC_IMAGE_R8_UNORM ? base[i] : base[i * 4 + 3];
  return (float)t / 255.0f;
}
/* one level of the image with one filter */
static float alpha_sample_level(const OrcAlphaMaterial* a, uint32_t level, uint32_t filter, float u, float v) {
  const uint8_t* base = a->texels;
  uint32_t w = a->width, h = a->height;
  for (uint32_t l = 0; l < level; l++) { /* levels are tightly packed one after the other */
    base += (size_t)w * h * (a->format == OXC_IMAGE_R8_UNORM ? 1u : 4u);
    w = w > 1 ? w >> 1 : 1;
    h = h > 1 ? h >> 1 : 1;
  }
  const float fw = (float)w, fh = (float)h;
  if (filter == OXC_FILTER_NEAREST) {
    const int ix = alpha_f2i(floorf(u * fw)), iy = alpha_f2i(floorf(v * fh));
    return alpha_texel(a, base, w, alpha_wrap(ix, w, a->address_u), alpha_wrap(iy, h, a->address_v));
  }
  const float x = u * fw - 0.5f, y = v * fh - 0.5f;
  const float x0 = floorf(x), y0 = floorf(y);
  const float wx = x - x0, wy = y - y0;
  const int ix = alpha_f2i(x0), iy = alpha_f2i(y0);
  const uint32_t xa = alpha_wrap(ix, w, a->address_u), xb = alpha_wrap((int64_t)ix + 1, w, a->address_u);
  const uint32_t ya = alpha_wrap(iy, h, a->address_v), yb = alpha_wrap((int64_t)iy + 1, h, a->address_v);
  const float a00 = alpha_texel(a, base, w, xa, ya), a10 = alpha_texel(a, base, w, xb, ya), a01 = alpha_texel(a, base, w, xa, yb),
              a11 = alpha_texel(a, base, w, xb, yb);
  const float top = a00 + wx * (a10 - a00), bot = a01 + wx * (a11 - a01);
  return top + wy * (bot - top);
}
/* steps 1-2: uv from three edge-function values; rw / uv in the raster's (a, b, c) order */
static void alpha_interpolate(const int64_t e[3], const float rw[3], const float uv[3][2], float* u, float* v) {
  const float p0 = (float)e[0] * rw[0], p1 = (float)e[1] * rw[1], p2 = (float)e[2] * rw[2];
  const float inv = 1.0f / ((p0 + p1) + p2);
  const float l0 = p0 * inv, l1 = p1 * inv, l2 = p2 * inv;
  *u = (l0 * uv[0][0] + l1 * uv[1][0]) + l2 * uv[2][0];
  *v = (l0 * uv[0][1] + l1 * uv[1][1]) + l2 * uv[2][1];
}
/* steps 2, 4, 5 at pixel (px, py): e = its three edge-function values, ex / ey = their increments per pixel in x / y */
static int alpha_keep_fragment(const OrcAlphaMaterial* a, int64_t px, int64_t py, const int64_t e[3], const int64_t ex[3],
                               const int64_t ey[3], const float rw[3], const float uv[3][2]) {
  float u, v;
  alpha_interpolate(e, rw, uv, &u, &v);
  float alpha;
  if (a->levels <= 1 && a->mag_filter == a->min_filter) {
    alpha = alpha_sample_level(a, 0, a->min_filter, u, v); /* nothing to select */
  } else {
    const int64_t ox = -(px & 1), oy = -(py & 1); /* to the quad's first column / row */
    int64_t eA[3], eB[3], eC[3], eD[3];
    for (int i = 0; i < 3; i++) {
      eA[i] = e[i] + ox * ex[i]; eB[i] = eA[i] + ex[i];
      eC[i] = e[i] + oy * ey[i]; eD[i] = eC[i] + ey[i];
    }
    float uA, vA, uB, vB, uC, vC, uD, vD;
    alpha_interpolate(eA, rw, uv, &uA, &vA); alpha_interpolate(eB, rw, uv, &uB, &vB);
    alpha_interpolate(eC, rw, uv, &uC, &vC); alpha_interpolate(eD, rw, uv, &uD, &vD);
    const float w0 = (float)a->width, h0 = (float)a->height;
    const float mx = (uB - uA) * w0, my = (vB - vA) * h0, nx = (uD - uC) * w0, ny = (vD - vC) * h0;
    const float rx2 = mx * mx + my * my, ry2 = nx * nx + ny * ny;
    const float r2 = rx2 > ry2 ? rx2 : ry2;
    const float lambda = 0.5f * orc_log2_canonical(r2);
    const uint32_t filter = lambda > 0.0f ? a->min_filter : a->mag_filter;
    const float q = (float)(a->levels > 1 ? a->levels - 1 : 0);
    const float lc = !(lambda > 0.0f) ? 0.0f : (lambda > q ? q : lambda);
    if (a->mipmap_mode == OXC_MIPMAP_NEAREST) {
      uint32_t d = lc <= 0.5f ? 0u : (uint32_t)ceilf(lc + 0.5f) - 1u;
      if (d > (uint32_t)q) d = (uint32_t)q;
      alpha = alpha_sample_level(a, d, filter, u, v);
    } else {
      const uint32_t d = (uint32_t)floorf(lc), dn = d + 1u > (uint32_t)q ? (uint32_t)q : d + 1u;
      const float f = lc - (float)d;
      const float lo = alpha_sample_level(a, d, filter, u, v), hi = alpha_sample_level(a, dn, filter, u, v);
      alpha = (1.0f - f) * lo + f * hi;
    }
  }
  return !(a->albedo_a * alpha < a->cutoff);
}
/* returns 0 when the material is not alpha tested (no albedo image) */
static int alpha_material_setup(const OxcMaterialTable* tab, uint32_t material_index, OrcAlphaMaterial* a) {
  if (!tab || material_index >= tab->material_count) return 0;
  const OxcMaterial* m = &tab->materials[material_index];
  if (!(m->flags & OXC_MATERIAL_HAS_ALBEDO_IMAGE) || m->albedo_image_index >= tab->image_count) return 0;
  const OxcAlphaImage* im = &tab->images[m->albedo_image_index];
  a->texels = (const uint8_t*)im->texels_dev; /* the oracle's images live in host memory */
  a->width = im->width; a->height = im->height; a->format = im->format; a->levels = im->level_count ? im->level_count : 1;
  a->mag_filter = OXC_FILTER_LINEAR; a->min_filter = OXC_FILTER_LINEAR; a->mipmap_mode = OXC_MIPMAP_LINEAR;
  a->address_u = OXC_ADDRESS_REPEAT; a->address_v = OXC_ADDRESS_REPEAT;
  if (tab->samplers && m->sampler_index < tab->sampler_count) {
    const OxcSamplerDesc* sd = &tab->samplers[m->sampler_index];
    a->mag_filter = sd->mag_filter; a->min_filter = sd->min_filter; a->mipmap_mode = sd->mipmap_mode;
    a->address_u = sd->address_u; a->address_v = sd->address_v;
  }
  a->albedo_a = orc_dequantize_half(m->albedo_color[3]);
  const float c = orc_dequantize_half(m->alpha_cutoff);
  a->cutoff = !(c == c) ? c : (c < 0.001f ? 0.001f : (c > 1.0f ? 1.0f : c));
  return 1;
}

/* non-NULL while orc_raster_overdraw runs: fragments are counted into this W x H image instead of being written to the vis buffer */
static __thread uint32_t* tl_overdraw = NULL;

static void raster_triangle_uv(const float clip[3][4], const float (*uv)[2], const OrcAlphaMaterial* am, uint32_t data, uint32_t W,
                               uint32_t H, uint64_t* vis);
static void raster_triangle(const float clip[3][4], uint32_t data, uint32_t W, uint32_t H, uint64_t* vis) {
  raster_triangle_uv(clip, NULL, NULL, data, W, H, vis);
}

/* the raster specification above; am != NULL adds the alpha test (uv = the three vertices' texture coordinates) */
static void raster_triangle_uv(const float clip[3][4], const float (*uv)[2], const OrcAlphaMaterial* am, uint32_t data, uint32_t W,
                               uint32_t H, uint64_t* vis) {
  if (!(clip[0][3] > 0.0f && clip[1][3] > 0.0f && clip[2][3] > 0.0f)) return; /* step 2 */
  int64_t fx[3], fy[3];
  float z[3], rws[3];
  for (int i = 0; i < 3; i++) {
    float rw = 1.0f / clip[i][3];
    rws[i] = rw;
    float nx = clip[i][0] * rw, ny = clip[i][1] * rw;
    z[i] = clip[i][2] * rw;
    float sx = (nx * 0.5f + 0.5f) * (float)W, sy = (ny * 0.5f + 0.5f) * (float)H;
    float qx = floorf(sx * 256.0f + 0.5f), qy = floorf(sy * 256.0f + 0.5f);
    if (!(fabsf(qx) <= 4194304.0f && fabsf(qy) <= 4194304.0f)) return; /* step 3 (also rejects NaN) */
    fx[i] = (int64_t)qx;
    fy[i] = (int64_t)qy;
  }
  int64_t area2 = orient2d(fx[0], fy[0], fx[1], fy[1], fx[2], fy[2]);
  if (area2 >= 0) return; /* step 4 */
  /* swap 1 <-> 2 so that orientation is positive */
  int64_t ax = fx[0], ay = fy[0], bx = fx[2], by = fy[2], cx = fx[1], cy = fy[1];
  float za = z[0], zb = z[2], zc = z[1];
  area2 = -area2;
  int64_t minx = ax < bx ? (ax < cx ? ax : cx) : (bx < cx ? bx : cx);
  int64_t maxx = ax > bx ? (ax > cx ? ax : cx) : (bx > cx ? bx : cx);
  int64_t miny = ay < by ? (ay < cy ? ay : cy) : (by < cy ? by : cy);
  int64_t maxy = ay > by ? (ay > cy ? ay : cy) : (by > cy ? by : cy);
  int64_t px0 = (minx - 128 + 255) >> 8, px1 = (maxx - 128) >> 8;
  int64_t py0 = (miny - 128 + 255) >> 8, py1 = (maxy - 128) >> 8;
  if (px0 < 0) px0 = 0;
  if (py0 < 0) py0 = 0;
  if (px1 > (int64_t)W - 1) px1 = (int64_t)W - 1;
  if (py1 > (int64_t)H - 1) py1 = (int64_t)H - 1;
  const int b0 = edge_bias(bx, by, cx, cy), b1 = edge_bias(cx, cy, ax, ay), b2 = edge_bias(ax, ay, bx, by);
  const float fa = (float)area2;
  const float dzb = (zb - za) / fa, dzc = (zc - za) / fa; /* per-triangle depth gradients w.r.t. the edge functions */
  for (int64_t py = py0; py <= py1; py++)
    for (int64_t px = px0; px <= px1; px++) {
      int64_t sxp = px * 256 + 128, syp = py * 256 + 128;
      int64_t e0 = orient2d(bx, by, cx, cy, sxp, syp); /* weight of a */
      int64_t e1 = orient2d(cx, cy, ax, ay, sxp, syp); /* weight of b */
      int64_t e2 = orient2d(ax, ay, bx, by, sxp, syp); /* weight of c */
      if ((e0 + b0) < 0 || (e1 + b1) < 0 || (e2 + b2) < 0) continue;
      float zz = (za + (float)e1 * dzb) + (float)e2 * dzc;
      if (!(zz >= 0.0f && zz <= 1.0f)) continue;
      if (am) { /* discard, visbuffer_encode.slang:62-64; (a, b, c) = vertices (0, 2, 1) */
        const int64_t e[3] = {e0, e1, e2};
        /* orient2d(a, b, p) = (bx-ax)*(py-ay) - (by-ay)*(px-ax): per pixel (256 sub-pixels) d/dpx = -(by-ay)*256, d/dpy = (bx-ax)*256 */
        const int64_t ex[3] = {-(cy - by) * 256, -(ay - cy) * 256, -(by - ay) * 256};
        const int64_t ey[3] = {(cx - bx) * 256, (ax - cx) * 256, (bx - ax) * 256};
        const float rwo[3] = {rws[0], rws[2], rws[1]};
        const float uvo[3][2] = {{uv[0][0], uv[0][1]}, {uv[2][0], uv[2][1]}, {uv[1][0], uv[1][1]}};
        if (!alpha_keep_fragment(am, px, py, e, ex, ey, rwo, uvo)) continue;
      }
      if (tl_overdraw) { tl_overdraw[(size_t)py * W + (size_t)px] += 1u; continue; } /* RENDER_OVERDRAW, visbuffer_encode.slang:68-70 */
      uint32_t zbits = f2bits(zz);
      if (zbits == 0x80000000u) zbits = 0u; /* -0.0 -> +0.0 so unsigned order == depth order */
      uint64_t v = ((uint64_t)zbits << 32) | (uint64_t)data;
      uint64_t* p = &vis[(size_t)py * W + (size_t)px];
      if (v > *p) *p = v;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Small-primitive cull — north_star item; the reference has none (cull_triangles.slang:59-90), so it is OPT-IN and
 * specified here: a triangle that passed cull_triangles' test is additionally culled iff all three vertices snap
 * (raster spec steps 2-3: w > 0, |fx|, |fy| <= 2^22) and the snapped bounding box holds no sample centre
 * (px1 < px0 or py1 < py0 after clamping to the image, exactly the bounds raster_triangle loops over).  Such a triangle
 * produces no fragment, so the vis buffer is unchanged; only the index buffer / the triangle count shrink.
 * ---------------------------------------------------------------------------------------------- */
int orc_triangle_covers_no_sample(const float clip[3][4], uint32_t W, uint32_t H) {
  int64_t fx[3], fy[3];
  for (int i = 0; i < 3; i++) {
    if (!(clip[i][3] > 0.0f)) return 0;
    float rw = 1.0f / clip[i][3];
    float nx = clip[i][0] * rw, ny = clip[i][1] * rw;
    float sx = (nx * 0.5f + 0.5f) * (float)W, sy = (ny * 0.5f + 0.5f) * (float)H;
    float qx = floorf(sx * 256.0f + 0.5f), qy = floorf(sy * 256.0f + 0.5f);
    if (!(fabsf(qx) <= 4194304.0f && fabsf(qy) <= 4194304.0f)) return 0;
    fx[i] = (int64_t)qx;
    fy[i] = (int64_t)qy;
  }
  int64_t minx = fx[0] < fx[1] ? (fx[0] < fx[2] ? fx[0] : fx[2]) : (fx[1] < fx[2] ? fx[1] : fx[2]);
  int64_t maxx = fx[0] > fx[1] ? (fx[0] > fx[2] ? fx[0] : fx[2]) : (fx[1] > fx[2] ? fx[1] : fx[2]);
  int64_t miny = fy[0] < fy[1] ? (fy[0] < fy[2] ? fy[0] : fy[2]) : (fy[1] < fy[2] ? fy[1] : fy[2]);
  int64_t maxy = fy[0] > fy[1] ? (fy[0] > fy[2] ? fy[0] : fy[2]) : (fy[1] > fy[2] ? fy[1] : fy[2]);
  int64_t px0 = (minx - 128 + 255) >> 8, px1 = (maxx - 128) >> 8;
  int64_t py0 = (miny - 128 + 255) >> 8, py1 = (maxy - 128) >> 8;
  if (px0 < 0) px0 = 0;
  if (py0 < 0) py0 = 0;
  if (px1 > (int64_t)W - 1) px1 = (int64_t)W - 1;
  if (py1 > (int64_t)H - 1) py1 = (int64_t)H - 1;
  return px1 < px0 || py1 < py0;
}

/* cull_triangles with the opt-in small-primitive cull: same outputs as orc_cull_triangles minus the culled triangles;
 * returns how many were culled */
uint64_t orc_cull_triangles_small_primitive(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                            const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                                            const OxcCullCamera* cam, uint32_t id_base, uint32_t W, uint32_t H,
                                            uint32_t* reordered_indices, OxcDrawIndexedIndirectCommand* draw_cmd) {
  uint64_t culled = 0;
  draw_cmd->index_count = 0; draw_cmd->instance_count = 1; draw_cmd->first_index = 0;
  draw_cmd->vertex_offset = 0; draw_cmd->first_instance = 0;
  for (uint32_t g = 0; g < pass_count; g++) {
    uint32_t mii = visible_indices[pass_first + g];
    TriMeshlet t;
    fetch_tri_meshlet(scene, meshlet_instances, mii, cam, &t);
    for (uint32_t tri = 0; tri < t.meshlet.triangle_count && tri < 64; tri++) {
      float clip[3][4];
      tri_clip(scene, &t, tri, clip);
      if (!tri_passes(clip)) continue;
      if (orc_triangle_covers_no_sample(clip, W, H)) { culled++; continue; }
      if (reordered_indices) {
        uint32_t off = draw_cmd->index_count;
        uint32_t masked = (mii + id_base) << OXC_VIS_PRIMITIVE_BITS;
        uint32_t ti = tri * 3;
        reordered_indices[off + 0] = masked | ((ti + 0) & OXC_VIS_PRIMITIVE_MASK);
        reordered_indices[off + 1] = masked | ((ti + 1) & OXC_VIS_PRIMITIVE_MASK);
        reordered_indices[off + 2] = masked | ((ti + 2) & OXC_VIS_PRIMITIVE_MASK);
      }
      draw_cmd->index_count += 3;
    }
  }
  return culled;
}

/* ------------------------------------------------------------------------------------------------
 * Clipped raster (round 2: implemented by the CUDA raster — oxc_raster_visbuffer queues exactly these triangles for
 * k_raster_clip_queue — and used by the frame functions below).
 * The plain spec drops a triangle when a vertex has w <= 0 or a snapped coordinate exceeds 2^22 (steps 2-3); a hardware
 * rasteriser clips such triangles instead (DrawGeometry.cpp:104-190 relies on it: geometry around the camera).  Only those
 * triangles take this path, so every triangle the plain spec draws is drawn identically:
 *   Sutherland-Hodgman in clip space against, in this order, near (w - z >= 0), left (w + x >= 0), right (w - x >= 0),
 *   bottom (w + y >= 0), top (w - y >= 0); a vertex with distance exactly 0 is inside.  An edge that crosses a plane is
 *   cut at  P = I + t * (O - I),  t = dI / (dI - dO)  evaluated FROM THE INSIDE VERTEX I TO THE OUTSIDE VERTEX O (so the two
 *   triangles sharing the edge compute the same point whatever their winding), per component, f32, this operation order.
 *   The polygon (<= 8 vertices) is drawn as the fan (P0, Pi, Pi+1) with the plain rules (snapping, tie-break, depth).
 * ---------------------------------------------------------------------------------------------- */
static int tri_dropped_by_range(const float clip[3][4], uint32_t W, uint32_t H) {
  for (int i = 0; i < 3; i++) {
    if (!(clip[i][3] > 0.0f)) return 1;
    float rw = 1.0f / clip[i][3];
    float nx = clip[i][0] * rw, ny = clip[i][1] * rw;
    float sx = (nx * 0.5f + 0.5f) * (float)W, sy = (ny * 0.5f + 0.5f) * (float)H;
    float qx = floorf(sx * 256.0f + 0.5f), qy = floorf(sy * 256.0f + 0.5f);
    if (!(fabsf(qx) <= 4194304.0f && fabsf(qy) <= 4194304.0f)) return 1;
  }
  return 0;
}

static inline float clip_plane_distance(const float v[4], int plane) {
  switch (plane) {
    case 0: return v[3] - v[2]; /* near (reverse-Z: z <= w) */
    case 1: return v[3] + v[0];
    case 2: return v[3] - v[0];
    case 3: return v[3] + v[1];
    default: return v[3] - v[1];
  }
}

typedef void (*RasterFn)(const float clip[3][4], uint32_t data, uint32_t W, uint32_t H, uint64_t* vis);
static void raster_triangle_clipped_with(const float clip[3][4], uint32_t data, uint32_t W, uint32_t H, uint64_t* vis, RasterFn draw) {
  float poly[2][12][4];
  int n = 3, cur = 0;
  memcpy(poly[0], clip, sizeof(float) * 12);
  for (int plane = 0; plane < 5 && n >= 3; plane++) {
    float(*in)[4] = poly[cur];
    float(*out)[4] = poly[cur ^ 1];
    int m = 0;
    for (int i = 0; i < n; i++) {
      const float* A = in[i];
      const float* B = in[(i + 1) % n];
      const float dA = clip_plane_distance(A, plane), dB = clip_plane_distance(B, plane);
      const int inA = dA >= 0.0f, inB = dB >= 0.0f;
      if (inA) { memcpy(out[m], A, 16); m++; }
      if (inA != inB) {
        const float* I = inA ? A : B;
        const float* O = inA ? B : A;
        const float dI = inA ? dA : dB, dO = inA ? dB : dA;
        const float t = dI / (dI - dO);
        for (int k = 0; k < 4; k++) out[m][k] = I[k] + t * (O[k] - I[k]);
        m++;
      }
    }
    n = m;
    cur ^= 1;
  }
  if (n < 3) return;
  for (int i = 1; i + 1 < n; i++) {
    float tri[3][4];
    memcpy(tri[0], poly[cur][0], 16); memcpy(tri[1], poly[cur][i], 16); memcpy(tri[2], poly[cur][i + 1], 16);
    draw(tri, data, W, H, vis);
  }
}
static void raster_triangle_clipped(const float clip[3][4], uint32_t data, uint32_t W, uint32_t H, uint64_t* vis) {
  raster_triangle_clipped_with(clip, data, W, H, vis, raster_triangle);
}

/* one clip-space triangle through the rules above (plain raster, or the clipped one when the plain rules drop it): the unit the
 * host cross-check of the CUDA raster core compares against (tests/raster_core_vs_oracle.cpp).  Returns 1 if it took the clip path. */
int orc_raster_triangle(const float clip[3][4], uint32_t data, uint32_t width, uint32_t height, uint64_t* vis) {
  if (tri_dropped_by_range(clip, width, height)) { raster_triangle_clipped(clip, data, width, height, vis); return 1; }
  raster_triangle(clip, data, width, height, vis);
  return 0;
}

void orc_raster_visbuffer_clip(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                               const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                               const OxcCullCamera* cam, uint32_t id_base, uint32_t width, uint32_t height, uint64_t* vis,
                               uint64_t* triangles_rasterised, uint64_t* triangles_clipped) {
  uint64_t ntri = 0, nclip = 0;
  for (uint32_t g = 0; g < pass_count; g++) {
    uint32_t mii = visible_indices[pass_first + g];
    TriMeshlet t;
    fetch_tri_meshlet(scene, meshlet_instances, mii, cam, &t);
    for (uint32_t tri = 0; tri < t.meshlet.triangle_count && tri < 64; tri++) {
      float clip[3][4];
      tri_clip(scene, &t, tri, clip);
      if (!tri_passes(clip)) continue;
      ntri++;
      uint32_t data = ((mii + id_base) << OXC_VIS_PRIMITIVE_BITS) | (tri & OXC_VIS_PRIMITIVE_MASK);
      if (tri_dropped_by_range(clip, width, height)) { nclip++; raster_triangle_clipped(clip, data, width, height, vis); }
      else raster_triangle(clip, data, width, height, vis);
    }
  }
  if (triangles_rasterised) *triangles_rasterised += ntri;
  if (triangles_clipped) *triangles_clipped += nclip;
}

void orc_raster_visbuffer(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                          const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                          const OxcCullCamera* cam, uint32_t id_base, uint32_t width, uint32_t height, uint64_t* vis,
                          uint64_t* triangles_rasterised) {
  uint64_t ntri = 0;
  for (uint32_t g = 0; g < pass_count; g++) {
    uint32_t mii = visible_indices[pass_first + g];
    TriMeshlet t;
    fetch_tri_meshlet(scene, meshlet_instances, mii, cam, &t);
    for (uint32_t tri = 0; tri < t.meshlet.triangle_count && tri < 64; tri++) {
      float clip[3][4];
      tri_clip(scene, &t, tri, clip);
      if (!tri_passes(clip)) continue;
      ntri++;
      uint32_t data = ((mii + id_base) << OXC_VIS_PRIMITIVE_BITS) | (tri & OXC_VIS_PRIMITIVE_MASK);
      raster_triangle(clip, data, width, height, vis);
    }
  }
  if (triangles_rasterised) *triangles_rasterised += ntri;
}

/* uv of the three corners of triangle `tri` (visbuffer_encode.slang:36; scene.slang:355-361, 491-497) */
static void tri_uv(const OrcScene* s, const TriMeshlet* t, uint32_t tri, float uv[3][2]) {
  const uint32_t* micro = (const uint32_t*)(s->blob + t->lod->local_triangle_indices);
  const uint32_t* vidx = (const uint32_t*)(s->blob + t->lod->indirect_vertex_indices);
  const uint16_t* tc = t->mesh->texture_coords ? (const uint16_t*)(s->blob + t->mesh->texture_coords) : NULL;
  uint32_t base = t->meshlet.local_triangle_index_offset + tri * 3;
  for (int c = 0; c < 3; c++) {
    uint32_t local = micro_index(micro, base + (uint32_t)c);
    uint32_t v = vidx[t->meshlet.indirect_vertex_index_offset + local];
    uv[c][0] = tc ? orc_dequantize_half(tc[v * 2 + 0]) : 0.0f;
    uv[c][1] = tc ? orc_dequantize_half(tc[v * 2 + 1]) : 0.0f;
  }
}

/* Sutherland-Hodgman of raster_triangle_clipped_with carrying (u, v) (alpha spec step 3): vertex = x y z w u v */
static void raster_triangle_clipped_uv(const float clip[3][4], const float uv[3][2], const OrcAlphaMaterial* am, uint32_t data,
                                       uint32_t W, uint32_t H, uint64_t* vis) {
  float poly[2][12][6];
  int n = 3, cur = 0;
  for (int i = 0; i < 3; i++) { memcpy(poly[0][i], clip[i], 16); poly[0][i][4] = uv[i][0]; poly[0][i][5] = uv[i][1]; }
  for (int plane = 0; plane < 5 && n >= 3; plane++) {
    float(*in)[6] = poly[cur];
    float(*out)[6] = poly[cur ^ 1];
    int m = 0;
    for (int i = 0; i < n; i++) {
      const float* A = in[i];
      const float* B = in[(i + 1) % n];
      const float dA = clip_plane_distance(A, plane), dB = clip_plane_distance(B, plane);
      const int inA = dA >= 0.0f, inB = dB >= 0.0f;
      if (inA) { memcpy(out[m], A, 24); m++; }
      if (inA != inB) {
        const float* I = inA ? A : B;
        const float* O = inA ? B : A;
        const float dI = inA ? dA : dB, dO = inA ? dB : dA;
        const float t = dI / (dI - dO);
        for (int k = 0; k < 6; k++) out[m][k] = I[k] + t * (O[k] - I[k]);
        m++;
      }
    }
    n = m;
    cur ^= 1;
  }
  if (n < 3) return;
  for (int i = 1; i + 1 < n; i++) {
    float tri[3][4], tuv[3][2];
    const int idx[3] = {0, i, i + 1};
    for (int k = 0; k < 3; k++) { memcpy(tri[k], poly[cur][idx[k]], 16); tuv[k][0] = poly[cur][idx[k]][4]; tuv[k][1] = poly[cur][idx[k]][5]; }
    raster_triangle_uv(tri, tuv, am, data, W, H, vis);
  }
}

/* one clip-space triangle of an alpha-tested material through the whole specification (plain or clip path) */
int orc_raster_triangle_alpha(const OxcMaterialTable* table, uint32_t material_index, const float clip[3][4], const float uv[3][2],
                              uint32_t data, uint32_t width, uint32_t height, uint64_t* vis) {
  OrcAlphaMaterial am;
  const OrcAlphaMaterial* a = alpha_material_setup(table, material_index, &am) ? &am : NULL;
  if (tri_dropped_by_range(clip, width, height)) {
    if (a) raster_triangle_clipped_uv(clip, uv, a, data, width, height, vis);
    else raster_triangle_clipped(clip, data, width, height, vis);
    return 1;
  }
  raster_triangle_uv(clip, uv, a, data, width, height, vis);
  return 0;
}
float orc_alpha_sample(const OxcAlphaImage* image, const OxcSamplerDesc* sampler /* NULL: linear + repeat */, uint32_t level, float u, float v) {
  OrcAlphaMaterial a;
  memset(&a, 0, sizeof a);
  a.texels = (const uint8_t*)image->texels_dev; a.width = image->width; a.height = image->height; a.format = image->format;
  uint32_t filter = OXC_FILTER_LINEAR;
  if (sampler) { filter = sampler->min_filter; a.address_u = sampler->address_u; a.address_v = sampler->address_v; }
  return alpha_sample_level(&a, level, filter, u, v);
}

/* orc_raster_visbuffer_clip with the alpha-tested discard of visbuffer_encode.slang:54-66 (specification above raster_triangle).
 * table->images[].texels_dev are HOST pointers here.  triangles_rasterised counts the triangles that pass the cull, as before. */
void orc_raster_visbuffer_alpha(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                                const OxcCullCamera* cam, uint32_t id_base, uint32_t width, uint32_t height, uint64_t* vis,
                                const OxcMaterialTable* table, uint64_t* triangles_rasterised, uint64_t* alpha_tested_triangles) {
  uint64_t ntri = 0, nalpha = 0;
  for (uint32_t g = 0; g < pass_count; g++) {
    uint32_t mii = visible_indices[pass_first + g];
    TriMeshlet t;
    fetch_tri_meshlet(scene, meshlet_instances, mii, cam, &t);
    const uint32_t material_index = scene->mesh_instances[meshlet_instances[mii].mesh_instance_index].material_index; /* :46 */
    OrcAlphaMaterial am;
    const OrcAlphaMaterial* a = alpha_material_setup(table, material_index, &am) ? &am : NULL;
    for (uint32_t tri = 0; tri < t.meshlet.triangle_count && tri < 64; tri++) {
      float clip[3][4], uv[3][2];
      tri_clip(scene, &t, tri, clip);
      if (!tri_passes(clip)) continue;
      ntri++;
      uint32_t data = ((mii + id_base) << OXC_VIS_PRIMITIVE_BITS) | (tri & OXC_VIS_PRIMITIVE_MASK);
      if (!a) {
        if (tri_dropped_by_range(clip, width, height)) raster_triangle_clipped(clip, data, width, height, vis);
        else raster_triangle(clip, data, width, height, vis);
        continue;
      }
      nalpha++;
      tri_uv(scene, &t, tri, uv);
      if (tri_dropped_by_range(clip, width, height)) raster_triangle_clipped_uv(clip, uv, a, data, width, height, vis);
      else raster_triangle_uv(clip, uv, a, data, width, height, vis);
    }
  }
  if (triangles_rasterised) *triangles_rasterised += ntri;
  if (alpha_tested_triangles) *alpha_tested_triangles += nalpha;
}

/* The encode pass's overdraw counter (visbuffer_encode.slang:15,68-70 / visbuffer_encode_ms.slang:189-191, RENDER_OVERDRAW;
 * MainGeometryContext::draw_overdraw, RendererInstance.cpp:771-776): overdraw[pixel] += 1 for every fragment the fragment shader
 * reaches its atomic with — covered sample (raster spec steps 1-5), depth inside [0, 1] (near / far clipping), not discarded by the
 * alpha test.  The depth COMPARISON plays no part: the shader has a side effect and a discard and declares no early fragment
 * tests, so the test runs after it; the counter is how many triangles were shaded at the pixel, not how many won.
 * table may be NULL (no material is alpha tested).  overdraw is accumulated into (the reference clears it with the vis buffer,
 * RendererInstance.cpp:649-679). */
void orc_raster_overdraw(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, const uint32_t* visible_indices,
                         uint32_t pass_first, uint32_t pass_count, const OxcCullCamera* cam, uint32_t width, uint32_t height,
                         uint32_t* overdraw, const OxcMaterialTable* table) {
  tl_overdraw = overdraw;
  orc_raster_visbuffer_alpha(scene, meshlet_instances, visible_indices, pass_first, pass_count, cam, 0, width, height, NULL, table, NULL, NULL);
  tl_overdraw = NULL;
}

void orc_resolve_visbuffer(const uint64_t* vis, uint32_t width, uint32_t height, uint32_t* vis32, float* depth) {
  for (size_t i = 0; i < (size_t)width * height; i++) {
    if (vis32) vis32[i] = (uint32_t)(vis[i] & 0xFFFFFFFFu); /* visbuffer.slang:67-70 */
    if (depth) depth[i] = bits2f((uint32_t)(vis[i] >> 32));
  }
}

/* ------------------------------------------------------------------------------------------------
 * cull.slang:137-166 test_vsm_page + passes/cull_meshlets_hpb.slang:27-99
 * ---------------------------------------------------------------------------------------------- */
/* max(0, ceil(log2(x))) on the float's bits: exact, no libm (x <= 0 or NaN -> 0; log2(+inf) -> 255) */
uint32_t orc_ceil_log2_f32(float x) {
  if (!(x > 1.0f)) return 0u;
  uint32_t b = f2bits(x);
  if ((b >> 23) == 255u) return 255u; /* inf */
  int32_t e = (int32_t)(b >> 23) - 127;
  return (uint32_t)(e + ((b & 0x7FFFFFu) ? 1 : 0));
}

static inline float fract_f32(float x) { return x - floorf(x); } /* Slang fract */

/* SampleLevel(NearestSamplerClamped, uv, level) != 0 on the R8UI pyramid: texel = clamp(floor(uv * size), 0, size-1) */
static int hpb_tap(const uint8_t* hpb, uint32_t hpb_size, uint32_t layers, uint32_t layer, uint32_t level, float u, float v) {
  size_t off = 0;
  for (uint32_t l = 0; l < level; l++) { uint32_t sl = hpb_size >> l; if (sl < 1) sl = 1; off += (size_t)layers * sl * sl; }
  uint32_t sz = hpb_size >> level; if (sz < 1) sz = 1;
  int32_t x = to_i32_f32(floorf(u * (float)sz)), y = to_i32_f32(floorf(v * (float)sz));
  if (x < 0) x = 0;
  if (x > (int32_t)sz - 1) x = (int32_t)sz - 1;
  if (y < 0) y = 0;
  if (y > (int32_t)sz - 1) y = (int32_t)sz - 1;
  return hpb[off + ((size_t)layer * sz + (size_t)y) * sz + (size_t)x] != 0;
}

static int test_vsm_page(const ScreenAabb_f32* a, const uint8_t* hpb, uint32_t hpb_size, uint32_t levels, uint32_t layers,
                         uint32_t layer, const int32_t page_offset[2]) {
  const float hs = (float)hpb_size;                                     /* :148-149 (square pyramid) */
  const float pox = (float)page_offset[0] / hs, poy = (float)page_offset[1] / hs; /* :151 */
  const float minu = a->min[0], minv = a->min[1], maxu = a->max[0], maxv = a->max[1];
  const float box_w = (maxu - minu) * hs, box_h = (maxv - minv) * hs;   /* :156-157 */
  uint32_t mip = orc_ceil_log2_f32(rmax_f32(box_w, box_h));             /* :158 */
  if (mip > levels - 1) mip = levels - 1;
  const int tl = hpb_tap(hpb, hpb_size, layers, layer, mip, fract_f32(minu + pox), fract_f32(minv + poy)); /* :160 */
  const int tr = hpb_tap(hpb, hpb_size, layers, layer, mip, fract_f32(maxu + pox), fract_f32(minv + poy)); /* :161 */
  const int bl = hpb_tap(hpb, hpb_size, layers, layer, mip, fract_f32(minu + pox), fract_f32(maxv + poy)); /* :162 */
  const int br = hpb_tap(hpb, hpb_size, layers, layer, mip, fract_f32(maxu + pox), fract_f32(maxv + poy)); /* :163 */
  return tl | tr | bl | br;
}

void orc_cull_meshlets_hpb(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, const OxcCullCamera* cam,
                           const OxcVirtualClipmap* clipmaps, const uint32_t* dirty_flags, uint32_t clipmap_count,
                           const uint8_t* hpb, uint32_t hpb_size, uint32_t hpb_levels, const OxcMeshletInstanceVisibility* vis,
                           uint32_t* visible_indices, OxcDispatchIndirectCommand* cmd) {
  cmd->x = 0; cmd->y = 1; cmd->z = 1;
  const Vec3_f32 view_dir = {cam->position[0], cam->position[1], cam->position[2]};
  const uint32_t total = vis->total_visible_meshlet_instances;
  for (uint32_t i = 0; i < total; i++) { /* :39-40 */
    MeshletCtx m;
    fetch_meshlet(scene, meshlet_instances[i], &m);
    float mvp[16];
    mul_mm_f32(cam->projection_view, m.world, mvp); /* :44 */
    int cone_vis = 1;                               /* :53-54 */
    if (!(m.cutoff >= 1.0f)) {
      Vec3_f32 axis = world_cone_axis_f32(m.world, m.axis);
      cone_vis = !(dot3_f32(axis, view_dir) >= m.cutoff);
    }
    if (!(cone_vis && test_frustum_f32(mvp, m.c, m.e))) continue; /* :56 */
    int visible = 0;
    for (uint32_t ci = 0; ci < clipmap_count; ci++) { /* :59-79 */
      if (dirty_flags[ci] == 0u) continue;
      float cmvp[16];
      mul_mm_f32(clipmaps[ci].projection_view_mat, m.world, cmvp);
      if (!test_frustum_f32(cmvp, m.c, m.e)) continue;
      ScreenAabb_f32 sa;
      if (project_aabb_f32(cmvp, clipmaps[ci].z_near, m.c, m.e, &sa))
        visible = test_vsm_page(&sa, hpb, hpb_size, hpb_levels, clipmap_count, ci, clipmaps[ci].page_offset);
      else visible = 1;
      if (visible) break;
    }
    if (visible) visible_indices[cmd->x++] = i; /* :81-98 */
  }
}

/* ------------------------------------------------------------------------------------------------
 * passes/rmvsm_downsample_hpb.slang:15-33 via Shadowmaps.cpp:331-366: level 0 from the page table
 * (cached = visible && backed && dirty, rmvsm.slang:16-28 [Flags] Visible=1 Dirty=2 Backed=4), level k
 * from level k-1 with the shader's `(tl | tr | bl | br) == 1` (values are 0/1 bytes).
 * hpb layout == oxc_cull_meshlets_hpb's: level l is layers x s_l x s_l bytes, s_l = max(1, size >> l).
 * ---------------------------------------------------------------------------------------------- */
void orc_build_hpb(const uint32_t* page_table, uint32_t size, uint32_t layers, uint8_t* hpb, uint32_t levels) {
  size_t src_off = 0, dst_off = 0;
  for (uint32_t l = 0; l < levels; l++) {
    uint32_t s = size >> l; if (s < 1) s = 1;                /* Shadowmaps.cpp:342-346 */
    uint32_t ps = l ? (size >> (l - 1)) : size; if (ps < 1) ps = 1;
    for (uint32_t z = 0; z < layers; z++)
      for (uint32_t y = 0; y < s; y++)
        for (uint32_t x = 0; x < s; x++) {
          uint8_t cached;
          if (l == 0) {                                     /* :24-26 */
            uint32_t page = page_table[((size_t)z * size + y) * size + x];
            cached = (page & 1u) != 0 && (page & 4u) != 0 && (page & 2u) != 0;
          } else {                                          /* :27-32; out-of-range loads return 0 (robust image access) */
            const uint8_t* src = hpb + src_off + (size_t)z * ps * ps;
            uint32_t x0 = x * 2, y0 = y * 2;
            uint8_t tl = (x0 < ps && y0 < ps) ? src[(size_t)y0 * ps + x0] : 0;
            uint8_t tr = (x0 < ps && y0 + 1 < ps) ? src[(size_t)(y0 + 1) * ps + x0] : 0;
            uint8_t bl = (x0 + 1 < ps && y0 < ps) ? src[(size_t)y0 * ps + x0 + 1] : 0;
            uint8_t br = (x0 + 1 < ps && y0 + 1 < ps) ? src[(size_t)(y0 + 1) * ps + x0 + 1] : 0;
            cached = (uint8_t)((tl | tr | bl | br) == 1);
          }
          hpb[dst_off + ((size_t)z * s + y) * s + x] = cached;
        }
    src_off = dst_off;
    dst_off += (size_t)layers * s * s;
  }
}

/* ------------------------------------------------------------------------------------------------
 * VSM page marking — passes/rmvsm_mark_visible_pages.slang:19-84 with rmvsm.slang:116-221.
 * log2 is the canonical polynomial below (strict f32 operations: the CUDA side evaluates the very same sequence).
 * ---------------------------------------------------------------------------------------------- */
float orc_log2_canonical(float x) {
  if (!(x > 0.0f)) return -3.0e38f;              /* zero, negative, NaN */
  if (!(x <= 3.0e38f)) return 3.0e38f;           /* +inf */
  int32_t e = 0;
  if (x < 1.17549435e-38f) { x = x * 16777216.0f; e = -24; } /* denormal: scale into the normal range (exact) */
  uint32_t bits = f2bits(x);
  e += (int32_t)(bits >> 23) - 127;
  float m = bits2f((bits & 0x007FFFFFu) | 0x3F800000u); /* [1, 2) */
  if (m > 1.41421354f) { m = m * 0.5f; e += 1; }        /* [0.7071, 1.4142] */
  const float t = (m - 1.0f) / (m + 1.0f);
  const float t2 = t * t;
  /* log2(m) = 2/ln2 * (t + t^3/3 + t^5/5 + t^7/7 + ...) */
  float p = 0.412198562f;                        /* 2/(7 ln 2) */
  p = 0.577078044f + t2 * p;                     /* 2/(5 ln 2) */
  p = 0.961796701f + t2 * p;                     /* 2/(3 ln 2) */
  p = 2.88539004f + t2 * p;                      /* 2/ln 2 */
  return (float)e + t * p;
}

static Vec4_f32 unproject_uv_h(const float inv_pv[16], float u, float v, float depth) { /* scene.slang:189-193 */
  Vec4_f32 ndc = {u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f};
  Vec4_f32 h = mul_mv_f32(inv_pv, ndc);
  Vec4_f32 w = {h.x / h.w, h.y / h.w, h.z / h.w, 1.0f};
  return w;
}

void orc_mark_visible_pages(const float inv_pv[16], const float resolution[2], const OxcVirtualClipmap* clipmaps,
                            const OxcVsmContext* vsm, const float* depth, uint32_t* page_tables, uint32_t* page_occupancy,
                            uint32_t* request_count, int32_t* requests, uint32_t request_capacity) {
  const int32_t W = vsm->depth_extent[0], H = vsm->depth_extent[1], size = vsm->page_table_size;
  /* rmvsm.slang:147-154 get_first_clipmap_texel_length */
  const float scale_ratio = (float)(size - 1) / (float)size;
  const float effective_width = vsm->first_clipmap_width * scale_ratio;
  const float texel_length = (effective_width * 2.0f) / vsm->virtual_extent;
  for (int32_t y = 0; y < H; y++)
    for (int32_t x = 0; x < W; x++) {
      const float d = depth[(size_t)y * W + x];
      if (d == 0.0f) continue;                                                        /* :43-45 */
      const float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;  /* :47 */
      const Vec4_f32 wp = unproject_uv_h(inv_pv, u, v, d);
      /* :156-186 (VMS_USE_DIAG_PIXEL_FOOTPRINT): distance between the unprojected left / right ends of the pixel */
      const float ox = (1.0f / resolution[0]) * 0.5f, oy = (1.0f / resolution[1]) * 0.5f;
      const Vec4_f32 a = unproject_uv_h(inv_pv, u - ox, v + oy, d), b = unproject_uv_h(inv_pv, u + ox, v + oy, d);
      const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
      const float dist = sqrtf((dx * dx + dy * dy) + dz * dz);
      float level_f = orc_log2_canonical(dist / texel_length);
      level_f = level_f > 0.0f ? level_f : 0.0f;                                      /* max(level_f, 0) */
      const float lf = ceilf(vsm->clipmap_selection_bias + level_f);
      uint32_t ci = lf >= 4294967296.0f ? 0xFFFFFFFFu : (lf > 0.0f ? (uint32_t)lf : 0u); /* u32() saturates */
      if (ci > (uint32_t)(vsm->clipmap_count - 1)) ci = (uint32_t)(vsm->clipmap_count - 1);
      const OxcVirtualClipmap* cm = &clipmaps[ci];
      Vec4_f32 wp1 = {wp.x, wp.y, wp.z, 1.0f};
      const Vec4_f32 lh = mul_mv_f32(cm->projection_view_mat, wp1);                   /* :214-221 */
      const float cu = (lh.x / lh.w + 1.0f) * 0.5f, cv = (lh.y / lh.w + 1.0f) * 0.5f;
      if (cu < 0.0f || cv < 0.0f || cu > 1.0f || cv > 1.0f) continue;                  /* :200-203 (NaN passes, as in the shader) */
      const float fxv = floorf(cu * (float)size), fyv = floorf(cv * (float)size);
      const int32_t vx = to_i32_f32(fxv), vy = to_i32_f32(fyv);
      if (vx < 0 || vy < 0 || vx > size - 1 || vy > size - 1) continue;               /* :129-133 */
      /* com::mod(offset, size) on floats: x - y * floor(x / y)  (common/math.slang:99-101) */
      const float fox = (float)(vx + cm->page_offset[0]), foy = (float)(vy + cm->page_offset[1]), fs = (float)size;
      const int32_t wx = to_i32_f32(fox - fs * floorf(fox / fs)), wy = to_i32_f32(foy - fs * floorf(foy / fs));
      if (wx < 0 || wy < 0 || wx > size - 1 || wy > size - 1) continue;
      uint32_t* page = &page_tables[((size_t)ci * size + (size_t)wy) * size + (size_t)wx];
      const uint32_t prev = *page;
      *page = prev | 1u;                                                              /* VSMPageState.Visible */
      if (!(prev & 1u)) {
        if (prev & 4u) page_occupancy[prev >> 16] = 1u;                               /* backed: :76-80 */
        else {
          const uint32_t k = (*request_count)++;                                      /* :81-82 */
          if (k < request_capacity) { requests[k * 3 + 0] = wx; requests[k * 3 + 1] = wy; requests[k * 3 + 2] = (int32_t)ci; }
        }
      }
    }
}

/* ------------------------------------------------------------------------------------------------
 * passes/visbuffer_decode.slang:42-183, geometry part: vis texel -> triangle re-fetch ->
 * analytic barycentrics + screen-space derivatives (compute_partial_derivatives :42-92), interpolated
 * texture coordinate with its gradients (:33-40,118-125 without the material's uv transform), geometric
 * world normal (:146-147) oct-encoded (common/encoding.slang:17-21).  Material / texture sampling and the
 * tangent frame built from it (:127-183) need the engine's material + image tables and stay out of scope.
 * Canonical arithmetic as in oxc_oracle.h; `1.0 / x` and `a / b` are IEEE divisions.
 * ---------------------------------------------------------------------------------------------- */
static inline Vec3_f32 decode_normal_f32(uint32_t packed) { /* scene.slang:486-489 */
  int32_t p = (int32_t)packed;
  Vec3_f32 n = {(float)((p >> 20) & 1023) / 511.0f - 1.0f, (float)((p >> 10) & 1023) / 511.0f - 1.0f,
                (float)(p & 1023) / 511.0f - 1.0f};
  return n;
}

static inline void vec3_to_oct_f32(Vec3_f32 v, float out[2]) { /* common/encoding.slang:17-21 */
  const float inv = 1.0f / ((fabsf(v.x) + fabsf(v.y)) + fabsf(v.z));
  const float px = v.x * inv, py = v.y * inv;
  if (v.z <= 0.0f) {
    out[0] = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
    out[1] = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
  } else {
    out[0] = px; out[1] = py;
  }
}

void orc_decode_visbuffer(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, uint32_t total,
                          const OxcCullCamera* cam, const uint32_t* vis32, uint32_t width, uint32_t height,
                          float* lambda_out, float* ddx_out, float* ddy_out, float* uv_normal_out, float* uv_grad_out) {
  const float* pv = cam->projection_view;
  for (uint32_t y = 0; y < height; y++)
    for (uint32_t x = 0; x < width; x++) {
      const size_t pix = (size_t)y * width + x;
      float L[4] = {0, 0, 0, 0}, DX[4] = {0, 0, 0, 0}, DY[4] = {0, 0, 0, 0}, UN[4] = {0, 0, 0, 0}, UG[4] = {0, 0, 0, 0};
      const uint32_t texel = vis32[pix];                                                  /* :96 */
      const uint32_t mii = (texel >> OXC_VIS_PRIMITIVE_BITS) & 0xFFFFFFu;                 /* visbuffer.slang:34 */
      const uint32_t tri = texel & OXC_VIS_PRIMITIVE_MASK;
      if (texel == 0xFFFFFFFFu || mii == 0xFFFFFEu || mii >= total) goto store;           /* :97-99 discard (+ range guard) */
      {
        const OxcMeshletInstance mi = meshlet_instances[mii];                             /* :103 */
        const OxcMeshInstance inst = scene->mesh_instances[mi.mesh_instance_index];       /* :104 */
        const OxcMesh* mesh = &scene->meshes[inst.mesh_index];                            /* :105 */
        const float* world = scene->transforms[inst.transform_index].world;               /* :107 */
        const OxcMeshLOD* lod = mesh_lod(scene, mesh, inst.lod_index);                    /* :108 */
        const OxcMeshlet meshlet = ((const OxcMeshlet*)(scene->blob + lod->meshlets))[mi.meshlet_index]; /* :109 */
        const uint32_t* micro = (const uint32_t*)(scene->blob + lod->local_triangle_indices);
        const uint32_t* vidx = (const uint32_t*)(scene->blob + lod->indirect_vertex_indices);
        const uint16_t* pos = (const uint16_t*)(scene->blob + mesh->vertex_positions);
        uint32_t idx[3];
        const uint32_t base = meshlet.local_triangle_index_offset + tri * 3;              /* scene.slang:366 */
        for (int c = 0; c < 3; c++) idx[c] = vidx[meshlet.indirect_vertex_index_offset + micro_index(micro, base + (uint32_t)c)];
        L[3] = 2.0f;
        if (idx[0] > mesh->vertex_count - 1 || idx[1] > mesh->vertex_count - 1 || idx[2] > mesh->vertex_count - 1) goto store; /* :115-117 */
        L[3] = 1.0f;
        Vec4_f32 wp[3];
        Vec3_f32 nrm[3];
        float tc[3][2];
        for (int c = 0; c < 3; c++) {
          Vec4_f32 p = {orc_dequantize_half(pos[idx[c] * 4 + 0]), orc_dequantize_half(pos[idx[c] * 4 + 1]),
                        orc_dequantize_half(pos[idx[c] * 4 + 2]), 1.0f};
          wp[c] = mul_mv_f32(world, p);                                                   /* :121 to_world_positions */
          if (mesh->vertex_normals) nrm[c] = decode_normal_f32(((const uint32_t*)(scene->blob + mesh->vertex_normals))[idx[c]]);
          else { nrm[c].x = 0; nrm[c].y = 0; nrm[c].z = 0; }
          if (mesh->texture_coords) {                                                     /* scene.slang:390-399 */
            const uint16_t* t = (const uint16_t*)(scene->blob + mesh->texture_coords) + (size_t)idx[c] * 2;
            tc[c][0] = orc_dequantize_half(t[0]); tc[c][1] = orc_dequantize_half(t[1]);
          } else { tc[c][0] = 0; tc[c][1] = 0; }
        }
        /* fullscreen.slang:11-17: tex_coord at the pixel centre; NDC = tex_coord * 2 - 1 (:122) */
        const float u = ((float)x + 0.5f) / (float)width, v = ((float)y + 0.5f) / (float)height;
        const float ndcx = u * 2.0f - 1.0f, ndcy = v * 2.0f - 1.0f;
        /* compute_partial_derivatives :42-92 */
        float inv_w[3], nx[3], ny[3];
        for (int c = 0; c < 3; c++) {
          Vec4_f32 w1 = {wp[c].x, wp[c].y, wp[c].z, 1.0f};
          Vec4_f32 cp = mul_mv_f32(pv, w1);                                               /* :45-47 */
          inv_w[c] = 1.0f / cp.w;                                                         /* :50 */
          nx[c] = cp.x * inv_w[c]; ny[c] = cp.y * inv_w[c];                               /* :51-53 */
        }
        /* :58 determinant(f32x2x2(ndc_2 - ndc_1, ndc_0 - ndc_1)) = a.x*b.y - a.y*b.x */
        const float ax = nx[2] - nx[1], ay = ny[2] - ny[1], bx = nx[0] - nx[1], by = ny[0] - ny[1];
        const float inv_det = 1.0f / (ax * by - ay * bx);
        float ddx[3], ddy[3];
        ddx[0] = ((ny[1] - ny[2]) * inv_det) * inv_w[0];                                  /* :60 */
        ddx[1] = ((ny[2] - ny[0]) * inv_det) * inv_w[1];
        ddx[2] = ((ny[0] - ny[1]) * inv_det) * inv_w[2];
        ddy[0] = ((nx[2] - nx[1]) * inv_det) * inv_w[0];                                  /* :62 */
        ddy[1] = ((nx[0] - nx[2]) * inv_det) * inv_w[1];
        ddy[2] = ((nx[1] - nx[0]) * inv_det) * inv_w[2];
        float ddx_sum = (ddx[0] * 1.0f + ddx[1] * 1.0f) + ddx[2] * 1.0f;                  /* :63 dot(v, 1.0) */
        float ddy_sum = (ddy[0] * 1.0f + ddy[1] * 1.0f) + ddy[2] * 1.0f;                  /* :64 */
        const float dvx = ndcx - nx[0], dvy = ndcy - ny[0];                               /* :66 */
        const float interp_inv_w = (inv_w[0] + dvx * ddx_sum) + dvy * ddy_sum;            /* :67 */
        const float interp_w = 1.0f / interp_inv_w;                                       /* :68 */
        float lam[3];
        lam[0] = interp_w * ((inv_w[0] + dvx * ddx[0]) + dvy * ddy[0]);                   /* :69-73 */
        lam[1] = interp_w * (dvx * ddx[1] + dvy * ddy[1]);
        lam[2] = interp_w * (dvx * ddx[2] + dvy * ddy[2]);
        const float torx = 2.0f / cam->resolution[0], tory = 2.0f / cam->resolution[1];   /* :74 */
        for (int c = 0; c < 3; c++) { ddx[c] = ddx[c] * torx; ddy[c] = ddy[c] * -tory; }  /* :75-76 */
        ddx_sum = ddx_sum * torx; ddy_sum = ddy_sum * -tory;                              /* :77-78 */
        const float interp_ddx_w = 1.0f / (interp_inv_w + ddx_sum);                       /* :80 */
        const float interp_ddy_w = 1.0f / (interp_inv_w + ddy_sum);                       /* :81 */
        for (int c = 0; c < 3; c++) {                                                     /* :82-83 */
          ddx[c] = interp_ddx_w * (lam[c] * interp_inv_w + ddx[c]) - lam[c];
          ddy[c] = interp_ddy_w * (lam[c] * interp_inv_w + ddy[c]) - lam[c];
        }
        for (int c = 0; c < 3; c++) { L[c] = lam[c]; DX[c] = ddx[c]; DY[c] = ddy[c]; }
        /* gradient_of :33-40: mul(row vector, 3x2 matrix) */
        for (int j = 0; j < 2; j++) {
          UN[j] = (lam[0] * tc[0][j] + lam[1] * tc[1][j]) + lam[2] * tc[2][j];
          UG[j] = (ddx[0] * tc[0][j] + ddx[1] * tc[1][j]) + ddx[2] * tc[2][j];
          UG[2 + j] = (ddy[0] * tc[0][j] + ddy[1] * tc[1][j]) + ddy[2] * tc[2][j];
        }
        /* :146-147 world_normal = normalize(mul(lambda, to_world_normals(normals))); scene.slang:291-302,319-322 */
        {
          Vec3_f32 b0 = {world[0], world[1], world[2]}, b1 = {world[4], world[5], world[6]}, b2 = {world[8], world[9], world[10]};
          Vec3_f32 r0 = cross3_f32(b1, b2), r1 = cross3_f32(b2, b0), r2 = cross3_f32(b0, b1);
          Vec3_f32 wn[3];
          for (int c = 0; c < 3; c++) {
            wn[c].x = (r0.x * nrm[c].x + r1.x * nrm[c].y) + r2.x * nrm[c].z;
            wn[c].y = (r0.y * nrm[c].x + r1.y * nrm[c].y) + r2.y * nrm[c].z;
            wn[c].z = (r0.z * nrm[c].x + r1.z * nrm[c].y) + r2.z * nrm[c].z;
          }
          Vec3_f32 n = {(lam[0] * wn[0].x + lam[1] * wn[1].x) + lam[2] * wn[2].x,
                        (lam[0] * wn[0].y + lam[1] * wn[1].y) + lam[2] * wn[2].y,
                        (lam[0] * wn[0].z + lam[1] * wn[1].z) + lam[2] * wn[2].z};
          const float len = length3_f32(n);
          Vec3_f32 nn = {n.x / len, n.y / len, n.z / len};
          vec3_to_oct_f32(nn, &UN[2]);                                                    /* :170 */
        }
      }
    store:
      for (int c = 0; c < 4; c++) {
        if (lambda_out) lambda_out[pix * 4 + c] = L[c];
        if (ddx_out) ddx_out[pix * 4 + c] = DX[c];
        if (ddy_out) ddy_out[pix * 4 + c] = DY[c];
        if (uv_normal_out) uv_normal_out[pix * 4 + c] = UN[c];
        if (uv_grad_out) uv_grad_out[pix * 4 + c] = UG[c];
      }
    }
}

/* ------------------------------------------------------------------------------------------------
 * passes/terrain_cull.slang:19-83; TerrainData::patch_corner / decode_height scene.slang:648-660
 * ---------------------------------------------------------------------------------------------- */
void orc_cull_terrain(const OxcTerrainData* t, const float* patch_minmax, const OxcCullCamera* cam, uint32_t flags,
                      const OrcHiz* hiz, uint32_t* visible_patches, uint32_t* mask, OxcDrawIndirectCommand* draw_cmd) {
  draw_cmd->vertex_count = 4; draw_cmd->instance_count = 0; draw_cmd->first_vertex = 0; draw_cmd->first_instance = 0; /* Terrain.cpp:168-170 */
  const uint32_t patch_total = t->patch_count[0] * t->patch_count[1]; /* :21 */
  for (uint32_t patch_index = 0; patch_index < patch_total; patch_index++) {
    const uint32_t px = patch_index % t->patch_count[0], py = patch_index / t->patch_count[0]; /* :26 */
    /* patch_corner: world_min + (f32x2(patch + corner) / f32x2(patch_count)) * world_size */
    const float g0x = (float)(px + 0) / (float)t->patch_count[0], g0y = (float)(py + 0) / (float)t->patch_count[1];
    const float g1x = (float)(px + 1) / (float)t->patch_count[0], g1y = (float)(py + 1) / (float)t->patch_count[1];
    const float cminx = t->world_min[0] + g0x * t->world_size[0], cminy = t->world_min[1] + g0y * t->world_size[1];
    const float cmaxx = t->world_min[0] + g1x * t->world_size[0], cmaxy = t->world_min[1] + g1y * t->world_size[1];
    const float bx = patch_minmax[2 * patch_index + 0], by = patch_minmax[2 * patch_index + 1]; /* :30 */
    Vec3_f32 c, e;
    c.x = (cminx + cmaxx) * 0.5f;
    c.y = t->base_height + ((bx + by) * 0.5f) * t->height_scale; /* decode_height */
    c.z = (cminy + cmaxy) * 0.5f;
    e.x = cmaxx - cminx;
    e.y = rmax_f32(t->height_scale * (by - bx), 1e-3f);
    e.z = cmaxy - cminy;
    const uint32_t mask_word = patch_index / 32u, mask_bit = 1u << (patch_index % 32u);
    const int was_visible = (mask[mask_word] & mask_bit) != 0u; /* :45 */
    int visible = (flags & OXC_CULL_LATE_PASS) ? 1 : was_visible; /* :47 */
    if (flags & OXC_CULL_TEST_FRUSTUM) visible = visible && test_frustum_f32(cam->projection_view, c, e); /* :49-51 */
    if ((flags & (OXC_CULL_TEST_OCCLUSION | OXC_CULL_LATE_PASS)) != 0 && visible) { /* :53-57 */
      ScreenAabb_f32 sa;
      if (project_aabb_f32(cam->projection_view, cam->near_clip, c, e, &sa)) visible = !test_occlusion_f32(&sa, hiz);
    }
    const int emit = visible && (!(flags & OXC_CULL_LATE_PASS) || !was_visible); /* :59 */
    if (flags & (OXC_CULL_TEST_OCCLUSION | OXC_CULL_LATE_PASS)) { /* :61-67 */
      if (visible) mask[mask_word] |= mask_bit; else mask[mask_word] &= ~mask_bit;
    }
    if (emit) visible_patches[draw_cmd->instance_count++] = patch_index; /* :70-82 */
  }
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline
 * ---------------------------------------------------------------------------------------------- */
typedef struct BaselineJob {
  const OrcScene* scene;
  const OxcMeshletInstance* mis;
  const OxcCullCamera* cam;
  uint32_t lo, hi;
  int mode;
  uint32_t* out; /* chunk-local draw list, written at out[lo..] */
  uint64_t n_out;
  float planes[6][4]; /* mode 0 */
} BaselineJob;

/* Oxylus/include/Utils/OxMath.hpp:54-80 calc_frustum_planes (glm m[col][row]); plane.w negated */
static void calc_frustum_planes(const float m[16], float planes[6][4]) {
#define M(c, r) m[(c) * 4 + (r)]
  for (int i = 0; i < 4; i++) planes[0][i] = M(i, 3) + M(i, 0);
  for (int i = 0; i < 4; i++) planes[1][i] = M(i, 3) - M(i, 0);
  for (int i = 0; i < 4; i++) planes[2][i] = M(i, 3) + M(i, 1);
  for (int i = 0; i < 4; i++) planes[3][i] = M(i, 3) - M(i, 1);
  for (int i = 0; i < 4; i++) planes[4][i] = M(i, 3) + M(i, 2);
  for (int i = 0; i < 4; i++) planes[5][i] = M(i, 3) - M(i, 2);
#undef M
  for (int p = 0; p < 6; p++) {
    float len = sqrtf((planes[p][0] * planes[p][0] + planes[p][1] * planes[p][1]) + planes[p][2] * planes[p][2]);
    for (int i = 0; i < 4; i++) planes[p][i] /= len;
    planes[p][3] = -planes[p][3];
  }
}

static void* baseline_worker(void* arg) {
  BaselineJob* j = (BaselineJob*)arg;
  const Vec3_f32 campos = {j->cam->position[0], j->cam->position[1], j->cam->position[2]};
  uint64_t n = 0;
  for (uint32_t i = j->lo; i < j->hi; i++) {
    MeshletCtx m;
    fetch_meshlet(j->scene, j->mis[i], &m);
    int vis;
    if (j->mode == 0) {
      /* world-space AABB of the meshlet box (8 corners through `world`), then
       * AABB::is_on_frustum / is_on_or_forward_plane — Oxylus/src/Render/BoundingVolume.cpp:72-88 */
      float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
      for (int k = 0; k < 8; k++) {
        Vec4_f32 p = {m.c.x + ((k & 1) ? 0.5f : -0.5f) * m.e.x, m.c.y + ((k & 2) ? 0.5f : -0.5f) * m.e.y,
                      m.c.z + ((k & 4) ? 0.5f : -0.5f) * m.e.z, 1.0f};
        Vec4_f32 w = mul_mv_f32(m.world, p);
        mn[0] = fminf(mn[0], w.x); mn[1] = fminf(mn[1], w.y); mn[2] = fminf(mn[2], w.z);
        mx[0] = fmaxf(mx[0], w.x); mx[1] = fmaxf(mx[1], w.y); mx[2] = fmaxf(mx[2], w.z);
      }
      float cx = (mx[0] + mn[0]) * 0.5f, cy = (mx[1] + mn[1]) * 0.5f, cz = (mx[2] + mn[2]) * 0.5f;
      float ex = (mx[0] - mn[0]) * 0.5f, ey = (mx[1] - mn[1]) * 0.5f, ez = (mx[2] - mn[2]) * 0.5f;
      vis = 1;
      for (int p = 0; p < 6 && vis; p++) {
        const float* pl = j->planes[p];
        float r = ex * fabsf(pl[0]) + ey * fabsf(pl[1]) + ez * fabsf(pl[2]);
        float dist = (pl[0] * cx + pl[1] * cy + pl[2] * cz) - pl[3]; /* Plane::get_distance */
        vis = -r <= dist;
      }
    } else {
      float mvp[16];
      mul_mm_f32(j->cam->projection_view, m.world, mvp);
      vis = cone_visible_f32(m.world, m.c, m.e, m.axis, m.cutoff, campos) && test_frustum_f32(mvp, m.c, m.e);
    }
    if (vis) j->out[j->lo + n++] = i;
  }
  j->n_out = n;
  return NULL;
}

uint64_t orc_cpu_baseline_cull(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, uint32_t total,
                               const OxcCullCamera* cam, int mode, int n_threads, uint32_t* out_indices) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  BaselineJob jobs[256];
  pthread_t th[256];
  float pv[16];
  memcpy(pv, cam->projection_view, sizeof pv);
  for (int t = 0; t < n_threads; t++) {
    BaselineJob* j = &jobs[t];
    j->scene = scene; j->mis = meshlet_instances; j->cam = cam; j->mode = mode; j->out = out_indices; j->n_out = 0;
    j->lo = (uint32_t)(((uint64_t)total * (uint64_t)t) / (uint64_t)n_threads);
    j->hi = (uint32_t)(((uint64_t)total * (uint64_t)(t + 1)) / (uint64_t)n_threads);
    calc_frustum_planes(pv, j->planes);
  }
  if (n_threads == 1) baseline_worker(&jobs[0]);
  else {
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, baseline_worker, &jobs[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  }
  /* compact the per-chunk draw lists (the std::vector append of the reference-shaped loop) */
  uint64_t n = jobs[0].n_out;
  for (int t = 1; t < n_threads; t++) {
    memmove(out_indices + n, out_indices + jobs[t].lo, jobs[t].n_out * sizeof(uint32_t));
    n += jobs[t].n_out;
  }
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * Threaded two-pass frame (bench.py --impl reference): same decisions as the serial passes above;
 * the meshlet culls and the raster are chunked over n_threads (per-thread vis buffers are avoided by
 * a CAS max on the shared image).
 * ---------------------------------------------------------------------------------------------- */
typedef struct FrameJob {
  const OrcScene* scene;
  const OxcMeshletInstance* mis;
  const OxcCullCamera* cam;
  const OrcHiz* hiz;
  uint32_t flags, lo, hi;
  uint32_t* mask;       /* read: mask_in snapshot; written bits go to set/clr lists */
  uint32_t* out;        /* survivors at out[lo..] */
  uint8_t* new_visible; /* per meshlet instance */
  uint64_t n_out;
  /* raster part */
  const uint32_t* visible;
  uint32_t width, height;
  uint64_t* vis;
  uint64_t ntri;
  uint32_t* next;       /* shared cursor: raster workers take chunks of FRAME_RASTER_CHUNK survivors (near meshlets cost 10x) */
  uint32_t n_items;
} FrameJob;
#define FRAME_RASTER_CHUNK 32u

static void* frame_cull_worker(void* arg) {
  FrameJob* j = (FrameJob*)arg;
  const Vec3_f32 campos = {j->cam->position[0], j->cam->position[1], j->cam->position[2]};
  uint64_t n = 0;
  /* consecutive meshlet instances share their mesh instance: mvp and the six planes are computed once per run */
  uint32_t cached_inst = 0xFFFFFFFFu;
  float mvp[16], planes[6][4];
  for (uint32_t i = j->lo; i < j->hi; i++) {
    MeshletCtx m;
    fetch_meshlet(j->scene, j->mis[i], &m);
    if (j->mis[i].mesh_instance_index != cached_inst) {
      cached_inst = j->mis[i].mesh_instance_index;
      mul_mm_f32(j->cam->projection_view, m.world, mvp);
      frustum_planes_f32(mvp, planes);
    }
    int was_visible = 1;
    if (j->flags & OXC_CULL_TEST_OCCLUSION) {
      uint32_t vi = m.inst.meshlet_instance_visibility_offset + j->mis[i].meshlet_index;
      was_visible = (j->mask[vi / 32] >> (vi & 31)) & 1;
    }
    int visible = meshlet_visible_hiz_hoisted_f32(mvp, (const float(*)[4])planes, m.world, j->cam->near_clip, campos, m.c, m.e,
                                                  m.axis, m.cutoff, j->flags, was_visible, j->hiz);
    j->new_visible[i] = (uint8_t)visible;
    if (visible && (!(j->flags & OXC_CULL_LATE_PASS) || !was_visible)) j->out[j->lo + n++] = i;
  }
  j->n_out = n;
  return NULL;
}

static void raster_triangle_atomic(const float clip[3][4], uint32_t data, uint32_t W, uint32_t H, uint64_t* vis);

static void* frame_raster_worker(void* arg) {
  FrameJob* j = (FrameJob*)arg;
  uint64_t ntri = 0;
  for (;;) {
    const uint32_t lo = __atomic_fetch_add(j->next, FRAME_RASTER_CHUNK, __ATOMIC_RELAXED);
    if (lo >= j->n_items) break;
    const uint32_t hi = lo + FRAME_RASTER_CHUNK < j->n_items ? lo + FRAME_RASTER_CHUNK : j->n_items;
    for (uint32_t g = lo; g < hi; g++) {
      uint32_t mii = j->visible[g];
      TriMeshlet t;
      fetch_tri_meshlet(j->scene, j->mis, mii, j->cam, &t);
      /* transform the meshlet's vertices once (as visbuffer_encode_ms.slang:135-137 does per vertex); the values are the
       * ones tri_clip computes per corner, so the image is unchanged */
      const uint32_t* micro = (const uint32_t*)(j->scene->blob + t.lod->local_triangle_indices);
      const uint32_t* vidx = (const uint32_t*)(j->scene->blob + t.lod->indirect_vertex_indices);
      const uint16_t* pos = (const uint16_t*)(j->scene->blob + t.mesh->vertex_positions);
      float vclip[OXC_MESHLET_MAX_VERTICES][4];
      const uint32_t nv = t.meshlet.vertex_count < OXC_MESHLET_MAX_VERTICES ? t.meshlet.vertex_count : OXC_MESHLET_MAX_VERTICES;
      for (uint32_t v = 0; v < nv; v++) {
        const uint32_t vi = vidx[t.meshlet.indirect_vertex_index_offset + v];
        Vec4_f32 pp = {orc_dequantize_half(pos[vi * 4 + 0]), orc_dequantize_half(pos[vi * 4 + 1]),
                       orc_dequantize_half(pos[vi * 4 + 2]), 1.0f};
        Vec4_f32 cp = mul_mv_f32(t.mvp, pp);
        vclip[v][0] = cp.x; vclip[v][1] = cp.y; vclip[v][2] = cp.z; vclip[v][3] = cp.w;
      }
      for (uint32_t tri = 0; tri < t.meshlet.triangle_count && tri < 64; tri++) {
        float clip[3][4];
        const uint32_t base = t.meshlet.local_triangle_index_offset + tri * 3;
        const uint32_t l0 = micro_index(micro, base), l1 = micro_index(micro, base + 1), l2 = micro_index(micro, base + 2);
        if (l0 < nv && l1 < nv && l2 < nv) {
          memcpy(clip[0], vclip[l0], 16); memcpy(clip[1], vclip[l1], 16); memcpy(clip[2], vclip[l2], 16);
        } else {
          tri_clip(j->scene, &t, tri, clip);
        }
        if (!tri_passes(clip)) continue;
        ntri++;
        const uint32_t data = (mii << OXC_VIS_PRIMITIVE_BITS) | (tri & OXC_VIS_PRIMITIVE_MASK);
        /* the product's raster clips what the plain spec drops (round 2): the frame follows */
        if (tri_dropped_by_range(clip, j->width, j->height)) raster_triangle_clipped_with(clip, data, j->width, j->height, j->vis, raster_triangle_atomic);
        else raster_triangle_atomic(clip, data, j->width, j->height, j->vis);
      }
    }
  }
  j->ntri = ntri;
  return NULL;
}

/* mask rewrite (cull_meshlets_hiz.slang:81-87) for a range of meshlet instances; neighbouring ranges share words */
static void* frame_mask_worker(void* arg) {
  FrameJob* j = (FrameJob*)arg;
  for (uint32_t i = j->lo; i < j->hi; i++) {
    OxcMeshletInstance mi = j->mis[i];
    uint32_t vi = j->scene->mesh_instances[mi.mesh_instance_index].meshlet_instance_visibility_offset + mi.meshlet_index;
    if (j->new_visible[i]) __atomic_fetch_or(&j->mask[vi / 32], 1u << (vi & 31), __ATOMIC_RELAXED);
    else __atomic_fetch_and(&j->mask[vi / 32], ~(1u << (vi & 31)), __ATOMIC_RELAXED);
  }
  return NULL;
}

/* same as raster_triangle but with an atomic max so threads can share the image */
static void raster_triangle_atomic(const float clip[3][4], uint32_t data, uint32_t W, uint32_t H, uint64_t* vis) {
  /* rasterise into a tiny private list is overkill: reuse raster_triangle's maths through a CAS loop */
  if (!(clip[0][3] > 0.0f && clip[1][3] > 0.0f && clip[2][3] > 0.0f)) return;
  int64_t fx[3], fy[3];
  float z[3];
  for (int i = 0; i < 3; i++) {
    float rw = 1.0f / clip[i][3];
    float nx = clip[i][0] * rw, ny = clip[i][1] * rw;
    z[i] = clip[i][2] * rw;
    float sx = (nx * 0.5f + 0.5f) * (float)W, sy = (ny * 0.5f + 0.5f) * (float)H;
    float qx = floorf(sx * 256.0f + 0.5f), qy = floorf(sy * 256.0f + 0.5f);
    if (!(fabsf(qx) <= 4194304.0f && fabsf(qy) <= 4194304.0f)) return;
    fx[i] = (int64_t)qx;
    fy[i] = (int64_t)qy;
  }
  int64_t area2 = orient2d(fx[0], fy[0], fx[1], fy[1], fx[2], fy[2]);
  if (area2 >= 0) return;
  int64_t ax = fx[0], ay = fy[0], bx = fx[2], by = fy[2], cx = fx[1], cy = fy[1];
  float za = z[0], zb = z[2], zc = z[1];
  area2 = -area2;
  int64_t minx = ax < bx ? (ax < cx ? ax : cx) : (bx < cx ? bx : cx);
  int64_t maxx = ax > bx ? (ax > cx ? ax : cx) : (bx > cx ? bx : cx);
  int64_t miny = ay < by ? (ay < cy ? ay : cy) : (by < cy ? by : cy);
  int64_t maxy = ay > by ? (ay > cy ? ay : cy) : (by > cy ? by : cy);
  int64_t px0 = (minx - 128 + 255) >> 8, px1 = (maxx - 128) >> 8;
  int64_t py0 = (miny - 128 + 255) >> 8, py1 = (maxy - 128) >> 8;
  if (px0 < 0) px0 = 0;
  if (py0 < 0) py0 = 0;
  if (px1 > (int64_t)W - 1) px1 = (int64_t)W - 1;
  if (py1 > (int64_t)H - 1) py1 = (int64_t)H - 1;
  const int b0 = edge_bias(bx, by, cx, cy), b1 = edge_bias(cx, cy, ax, ay), b2 = edge_bias(ax, ay, bx, by);
  const float fa = (float)area2;
  const float dzb = (zb - za) / fa, dzc = (zc - za) / fa; /* per-triangle depth gradients w.r.t. the edge functions */
  for (int64_t py = py0; py <= py1; py++)
    for (int64_t px = px0; px <= px1; px++) {
      int64_t sxp = px * 256 + 128, syp = py * 256 + 128;
      int64_t e0 = orient2d(bx, by, cx, cy, sxp, syp);
      int64_t e1 = orient2d(cx, cy, ax, ay, sxp, syp);
      int64_t e2 = orient2d(ax, ay, bx, by, sxp, syp);
      if ((e0 + b0) < 0 || (e1 + b1) < 0 || (e2 + b2) < 0) continue;
      float zz = (za + (float)e1 * dzb) + (float)e2 * dzc;
      if (!(zz >= 0.0f && zz <= 1.0f)) continue;
      uint32_t zbits = f2bits(zz);
      if (zbits == 0x80000000u) zbits = 0u; /* -0.0 -> +0.0 so unsigned order == depth order */
      uint64_t v = ((uint64_t)zbits << 32) | (uint64_t)data;
      uint64_t* p = &vis[(size_t)py * W + (size_t)px];
      uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
      while (v > old && !__atomic_compare_exchange_n(p, &old, v, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    }
}

static void run_jobs(FrameJob* jobs, int n, void* (*fn)(void*)) {
  pthread_t th[256];
  if (n == 1) { fn(&jobs[0]); return; }
  for (int t = 0; t < n; t++) pthread_create(&th[t], NULL, fn, &jobs[t]);
  for (int t = 0; t < n; t++) pthread_join(th[t], NULL);
}

uint64_t orc_cpu_frame(const OrcScene* scene, const OxcCullCamera* cam, uint32_t width, uint32_t height, uint32_t hiz_w,
                       uint32_t hiz_h, uint32_t* mask, const float* occluder_depth, int n_threads,
                       OxcMeshletInstance* meshlet_instances, uint32_t* visible_indices, uint64_t* vis,
                       OxcMeshletInstanceVisibility* vc, uint64_t* triangles) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  OxcDispatchIndirectCommand cmd;
  orc_cull_meshes(scene, cam, OXC_CULL_TEST_ALL, 0, 0xFFFFFFFFu, meshlet_instances, vc, &cmd);
  const uint32_t total = vc->total_visible_meshlet_instances;
  OrcHiz hiz;
  orc_hiz_layout(hiz_w, hiz_h, &hiz);
  hiz.data = (float*)calloc(orc_hiz_total_texels(hiz_w, hiz_h), sizeof(float)); /* cleared to 0 each frame */
  uint8_t* newvis = (uint8_t*)malloc(total ? total : 1);
  float* depth = (float*)malloc((size_t)width * height * sizeof(float));
  FrameJob jobs[256];
  uint64_t ntri = 0;
  /* vis = clear, then occluder depth (depth written by passes outside the path) */
  for (size_t i = 0; i < (size_t)width * height; i++)
    vis[i] = ((uint64_t)(occluder_depth ? f2bits(occluder_depth[i]) : 0u) << 32) | OXC_VIS_CLEAR;

  for (int pass = 0; pass < 2; pass++) {
    uint32_t flags = OXC_CULL_TEST_ALL | (pass ? OXC_CULL_LATE_PASS : 0);
    for (int t = 0; t < n_threads; t++) {
      FrameJob* j = &jobs[t];
      memset(j, 0, sizeof *j);
      j->scene = scene; j->mis = meshlet_instances; j->cam = cam; j->hiz = &hiz; j->flags = flags; j->mask = mask;
      j->out = visible_indices + (pass ? vc->early_visible_meshlet_instances : 0) ; j->new_visible = newvis;
      j->lo = (uint32_t)(((uint64_t)total * (uint64_t)t) / (uint64_t)n_threads);
      j->hi = (uint32_t)(((uint64_t)total * (uint64_t)(t + 1)) / (uint64_t)n_threads);
    }
    /* survivors are staged in a scratch list (chunk-local positions), then compacted */
    uint32_t* scratch = (uint32_t*)malloc((size_t)(total ? total : 1) * sizeof(uint32_t));
    for (int t = 0; t < n_threads; t++) jobs[t].out = scratch;
    run_jobs(jobs, n_threads, frame_cull_worker);
    uint32_t base = pass ? vc->early_visible_meshlet_instances : 0, n = 0;
    for (int t = 0; t < n_threads; t++) {
      memcpy(visible_indices + base + n, scratch + jobs[t].lo, jobs[t].n_out * sizeof(uint32_t));
      n += (uint32_t)jobs[t].n_out;
    }
    free(scratch);
    if (pass) vc->late_visible_meshlet_instances = n; else vc->early_visible_meshlet_instances = n;
    /* mask rewrite (cull_meshlets_hiz.slang:81-87): same [lo, hi) ranges of meshlet instances as the cull */
    for (int t = 0; t < n_threads; t++) {
      jobs[t].lo = (uint32_t)(((uint64_t)total * (uint64_t)t) / (uint64_t)n_threads);
      jobs[t].hi = (uint32_t)(((uint64_t)total * (uint64_t)(t + 1)) / (uint64_t)n_threads);
    }
    run_jobs(jobs, n_threads, frame_mask_worker);
    /* raster this pass's survivors: chunks of FRAME_RASTER_CHUNK from a shared cursor */
    uint32_t cursor = 0;
    for (int t = 0; t < n_threads; t++) {
      FrameJob* j = &jobs[t];
      j->visible = visible_indices + base; j->width = width; j->height = height; j->vis = vis;
      j->next = &cursor; j->n_items = n;
    }
    run_jobs(jobs, n_threads, frame_raster_worker);
    for (int t = 0; t < n_threads; t++) ntri += jobs[t].ntri;
    if (!pass) { /* generate_hiz from the early depth */
      orc_resolve_visbuffer(vis, width, height, NULL, depth);
      orc_build_hiz(depth, width, height, &hiz);
    }
  }
  if (triangles) *triangles = ntri;
  free(hiz.data); free(newvis); free(depth);
  return (uint64_t)vc->early_visible_meshlet_instances + vc->late_visible_meshlet_instances;
}
