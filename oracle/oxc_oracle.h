/*
 * oxc_oracle.h — CPU ORACLE for the meshlet visibility pipeline.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (liboxcull.so) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (oxylusengine/Oxylus @ 30560c65) has no test, golden image or
 * fixture for its render path and its Vulkan/Slang path cannot be built or run
 * here (needs C++23 deducing-this, xmake, vuk, slangc, a Vulkan ICD).  This file is a plain-C
 * restatement of the Slang shaders, one function per shader function, each citing the lines it
 * follows.  It is pinned by hand-computed known-answer tests (tests/test_oracle_*.py), by an f64
 * re-evaluation of every predicate (the *_f64 entry points) and by code review against the cited
 * lines — not by reference-generated vectors.
 *
 * Canonical arithmetic (the reference compiles with SLANG_FLOATING_POINT_MODE_FAST,
 * ResourceCompiler/private/Session.cpp:53, so its own bits are driver-dependent; this is the
 * evaluation order both the oracle and the CUDA kernels commit to):
 *   - IEEE-754 binary32, round-to-nearest-even, NO fma contraction, denormals kept
 *   - dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z            (vec4: ... + a.w*b.w)
 *   - mul(M,v)[i]   = ((M[i][0]*v.x + M[i][1]*v.y) + M[i][2]*v.z) + M[i][3]*v.w    (M[i] = row i)
 *   - mul(A,B)[i][j]= ((A[i][0]*B[0][j] + A[i][1]*B[1][j]) + A[i][2]*B[2][j]) + A[i][3]*B[3][j]
 *   - length(v) = sqrt(dot(v,v)); normalize(v) = v / length(v); vector / scalar = per-component divide
 *   - cross, determinant: textbook cofactor order, left to right
 *   - float -> u32/i32 conversions saturate and truncate toward zero (PTX cvt.rzi semantics)
 *   - ceil(log2(float(n))) for integer n is evaluated in integers: 0 for n <= 1, else 32 - clz(n-1)
 *     (identical to libm for every n the Hi-Z sizes allow; tests/test_oracle_units.py checks it)
 */
#ifndef OXC_ORACLE_H_
#define OXC_ORACLE_H_

#include "../include/oxcull.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Host scene: tables + blob; Mesh/MeshLOD u64 members are byte offsets into blob (OxcSceneDesc). */
typedef OxcSceneDesc OrcScene;

/* Hi-Z pyramid on the host: level l is w>>l x h>>l (min 1) floats at data + level_offset[l]. */
typedef struct OrcHiz {
  float* data;
  uint32_t width, height, levels;
  uint32_t level_offset[OXC_HIZ_MAX_LEVELS];
} OrcHiz;

typedef struct OrcScreenAabb { float min[3]; float max[3]; } OrcScreenAabb;

/* ---- unit functions (cull.slang / scene.slang / common/math.slang) ---- */
float orc_dequantize_half(uint16_t h);                                   /* common/math.slang:193-201 */
void orc_bounds_decode(const OxcMeshletBounds* b, float center[3], float extent[3], float cone_axis[3],
                       float* cone_cutoff);                               /* scene.slang:401-435 */
void orc_mat4_mul(const float a[16], const float b[16], float out[16]);  /* mul(A,B), column-major storage */
int orc_project_aabb(const float mvp[16], float near_clip, const float c[3], const float e[3],
                     OrcScreenAabb* out);                                 /* cull.slang:12-47; 0 = none */
int orc_test_frustum(const float mvp[16], const float c[3], const float e[3]); /* cull.slang:57-84 */
int orc_test_occlusion(const OrcScreenAabb* aabb, const OrcHiz* hiz);    /* cull.slang:86-135 */
int orc_test_cone(const float center[3], float radius, const float axis[3], float cutoff,
                  const float cam[3]);                                    /* cull.slang:173-175 */
int orc_test_cone_directional(const float axis[3], float cutoff, const float view_dir[3]); /* :177-179 */
int orc_test_triangle_backface(const float clip[3][4]);                  /* cull.slang:169-171 */
uint32_t orc_ceil_log2_u32(uint32_t n);
uint32_t orc_hiz_level_count(uint32_t w, uint32_t h);                    /* Texture.hpp:144-146, min(.,13) */
void orc_hiz_layout(uint32_t w, uint32_t h, OrcHiz* hiz);                /* fills levels / offsets (data untouched) */
uint32_t orc_hiz_total_texels(uint32_t w, uint32_t h);

/* ---- passes ---- */
/* cull_meshes.slang:17-85.  first/count restrict to a mesh-instance shard (count==0xFFFFFFFF: all).
 * Writes lod_index back into scene->mesh_instances (cast away const, like the RW buffer). */
void orc_cull_meshes(const OrcScene* scene, const OxcCullCamera* cam, uint32_t flags, uint32_t first,
                     uint32_t count, OxcMeshletInstance* meshlet_instances,
                     OxcMeshletInstanceVisibility* vis, OxcDispatchIndirectCommand* cull_meshlets_cmd);

/* cull_meshlets_hiz.slang:19-88 */
void orc_cull_meshlets_hiz(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                           const OxcCullCamera* cam, uint32_t flags, const OrcHiz* hiz,
                           OxcMeshletInstanceVisibility* vis, uint32_t* visible_indices, uint32_t* mask,
                           OxcDispatchIndirectCommand* cull_triangles_cmd);
/* same decisions in binary64 (margin classification): out_visible[i] in {0,1} for every i < total */
void orc_cull_meshlets_hiz_f64(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                               const OxcCullCamera* cam, uint32_t flags, const OrcHiz* hiz,
                               const OxcMeshletInstanceVisibility* vis, const uint32_t* mask_in,
                               uint8_t* out_visible);
void orc_cull_meshlets_hiz_f32_flags(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                     const OxcCullCamera* cam, uint32_t flags, const OrcHiz* hiz,
                                     const OxcMeshletInstanceVisibility* vis, const uint32_t* mask_in,
                                     uint8_t* out_visible);

/* cull_meshlets.slang:21-73 */
void orc_cull_meshlets(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                       const OxcCullCamera* cam, OxcMeshletInstanceVisibility* vis, uint32_t* visible_indices,
                       OxcDispatchIndirectCommand* cull_triangles_cmd);

/* multi-view: per view cone + frustum (cull_meshlets_hpb.slang:27-60 without the page test) */
void orc_cull_meshlets_multiview(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                 uint32_t total, const OxcCullCamera* views, uint32_t n_views, int directional,
                                 uint32_t* view_bits, uint32_t* view_counts);

/* hiz.slang:171-267 (+ CullGeometry.cpp:10-59): mip0 = nearest sample of depth at uv=(texel+1)/hiz_extent,
 * mip k = 2x2 min of mip k-1.  hiz->data must hold orc_hiz_total_texels floats. */
void orc_build_hiz(const float* depth, uint32_t width, uint32_t height, OrcHiz* hiz);

/* cull_triangles.slang:27-90: pass_first/pass_count select the survivors of this pass. */
void orc_cull_triangles(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                        const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                        const OxcCullCamera* cam, uint32_t id_base, uint32_t* reordered_indices,
                        OxcDrawIndexedIndirectCommand* draw_cmd);

/* SW raster (spec in DESIGN.md §raster): same triangle cull, then rasterise with
 * max on asuint(depth)<<32 | (id<<8 | tri).  vis must be pre-cleared (orc_clear_visbuffer). */
void orc_clear_visbuffer(uint64_t* vis, uint32_t width, uint32_t height);
/* one clip-space triangle (after the near / backface test) through the raster specification, clipped when the plain rules drop
 * it; returns 1 if the clip path was taken */
int orc_raster_triangle(const float clip[3][4], uint32_t data, uint32_t width, uint32_t height, uint64_t* vis);
/* VSM page marking (rmvsm_mark_visible_pages.slang) and the canonical log2 it uses */
float orc_log2_canonical(float x);
void orc_mark_visible_pages(const float inv_projection_view[16], const float resolution[2], const OxcVirtualClipmap* clipmaps,
                            const OxcVsmContext* vsm, const float* depth, uint32_t* page_tables, uint32_t* page_occupancy,
                            uint32_t* request_count, int32_t* requests, uint32_t request_capacity);

/* north_star's small-primitive cull (opt-in; no reference equivalent): specification in oxc_oracle.c */
int orc_triangle_covers_no_sample(const float clip[3][4], uint32_t width, uint32_t height);
uint64_t orc_cull_triangles_small_primitive(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                            const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                                            const OxcCullCamera* cam, uint32_t id_base, uint32_t width, uint32_t height,
                                            uint32_t* reordered_indices /* may be NULL: count only */,
                                            OxcDrawIndexedIndirectCommand* draw_cmd);

void orc_raster_visbuffer(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                          const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                          const OxcCullCamera* cam, uint32_t id_base, uint32_t width, uint32_t height,
                          uint64_t* vis, uint64_t* triangles_rasterised);
void orc_resolve_visbuffer(const uint64_t* vis, uint32_t width, uint32_t height, uint32_t* vis32, float* depth);
/* SPECIFICATION ONLY (not implemented by the CUDA raster yet): as orc_raster_visbuffer, but triangles the plain spec drops
 * for w <= 0 / coordinate overflow are clipped in clip space against near + the four side planes and drawn as a fan. */
void orc_raster_visbuffer_clip(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                               const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                               const OxcCullCamera* cam, uint32_t id_base, uint32_t width, uint32_t height, uint64_t* vis,
                               uint64_t* triangles_rasterised, uint64_t* triangles_clipped);

/* Alpha-tested discard of the vis-buffer encode (visbuffer_encode.slang:54-66; specification in oxc_oracle.c above
 * raster_triangle, repeated in include/oxcull.h).  Image texels are HOST pointers for the oracle. */
/* one clip-space triangle of material `material_index` (uv = its three texture coordinates) through the raster + alpha
 * specification, plain or clip path; returns 1 if it took the clip path */
int orc_raster_triangle_alpha(const OxcMaterialTable* table, uint32_t material_index, const float clip[3][4], const float uv[3][2],
                              uint32_t data, uint32_t width, uint32_t height, uint64_t* vis);
float orc_alpha_sample(const OxcAlphaImage* image, const OxcSamplerDesc* sampler, uint32_t level, float u, float v); /* one level, min_filter */
void orc_raster_visbuffer_alpha(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                                const uint32_t* visible_indices, uint32_t pass_first, uint32_t pass_count,
                                const OxcCullCamera* cam, uint32_t id_base, uint32_t width, uint32_t height, uint64_t* vis,
                                const OxcMaterialTable* table, uint64_t* triangles_rasterised, uint64_t* alpha_tested_triangles);

/* RENDER_OVERDRAW of the encode pass (visbuffer_encode.slang:68-70): += 1 per shaded fragment (specification in oxc_oracle.c) */
void orc_raster_overdraw(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, const uint32_t* visible_indices,
                         uint32_t pass_first, uint32_t pass_count, const OxcCullCamera* cam, uint32_t width, uint32_t height,
                         uint32_t* overdraw, const OxcMaterialTable* table /* may be NULL */);

/* passes/cull_meshlets_hpb.slang:27-99 + cull.slang:137-166 test_vsm_page.  hpb: levels of (layers x s x s) bytes. */
void orc_cull_meshlets_hpb(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, const OxcCullCamera* cam,
                           const OxcVirtualClipmap* clipmaps, const uint32_t* dirty_flags, uint32_t clipmap_count,
                           const uint8_t* hpb, uint32_t hpb_size, uint32_t hpb_levels, const OxcMeshletInstanceVisibility* vis,
                           uint32_t* visible_indices, OxcDispatchIndirectCommand* cull_triangles_cmd);
uint32_t orc_ceil_log2_f32(float x); /* ceil(log2(x)) clamped at 0 from below, evaluated on the float's bits */

/* passes/rmvsm_downsample_hpb.slang:15-33 via Shadowmaps.cpp:331-366: page table (layers x size x size
 * u32, rmvsm.slang VSMPageState bits) -> the R8UI pyramid oxc_cull_meshlets_hpb consumes. */
void orc_build_hpb(const uint32_t* page_table, uint32_t size, uint32_t layers, uint8_t* hpb, uint32_t levels);

/* passes/visbuffer_decode.slang:42-183, geometry part.  Five float4 planes (any may be NULL):
 * lambda.xyz + status (0 discarded, 1 decoded, 2 vertex index out of range), ddx.xyz, ddy.xyz,
 * (uv.xy, oct(world_normal).xy), (uv_ddx.xy, uv_ddy.xy). */
void orc_decode_visbuffer(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances, uint32_t total,
                          const OxcCullCamera* cam, const uint32_t* vis32, uint32_t width, uint32_t height,
                          float* lambda_out, float* ddx_out, float* ddy_out, float* uv_normal_out, float* uv_grad_out);

/* passes/terrain_cull.slang:19-83 */
void orc_cull_terrain(const OxcTerrainData* terrain, const float* patch_minmax, const OxcCullCamera* cam, uint32_t flags,
                      const OrcHiz* hiz, uint32_t* visible_patches, uint32_t* mask, OxcDrawIndirectCommand* draw_cmd);

/* ---- CPU baseline: the reference's CPU primitives over meshlet bounds ----
 * mode 0: AABB::is_on_frustum (BoundingVolume.cpp:72-88) with planes from math::calc_frustum_planes
 *         (OxMath.hpp:54-80) on world-space AABBs of dequantised bounds
 * mode 1: shader-equivalent cone + test_frustum (cull.slang:57-84,173-175) — same decisions as
 *         cull_meshlets.slang
 * Appends survivors to out_indices per thread chunk (draw-list build shaped like Scene.cpp:1230-1261);
 * returns survivor count.  n_threads >= 1 (pthreads, static chunking). */
uint64_t orc_cpu_baseline_cull(const OrcScene* scene, const OxcMeshletInstance* meshlet_instances,
                               uint32_t total, const OxcCullCamera* cam, int mode, int n_threads,
                               uint32_t* out_indices);
/* full two-pass frame on the CPU with n_threads (the "reference arm" of bench.py --impl reference):
 * cull_meshes -> early cull -> raster -> hiz -> late cull -> raster.  Returns late+early survivor count. */
uint64_t orc_cpu_frame(const OrcScene* scene, const OxcCullCamera* cam, uint32_t width, uint32_t height,
                       uint32_t hiz_w, uint32_t hiz_h, uint32_t* mask, const float* occluder_depth,
                       int n_threads, OxcMeshletInstance* meshlet_instances, uint32_t* visible_indices,
                       uint64_t* vis, OxcMeshletInstanceVisibility* vis_counts, uint64_t* triangles);

#ifdef __cplusplus
}
#endif
#endif
