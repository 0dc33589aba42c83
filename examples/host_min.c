/* host_min.c — the smallest plain-C host of liboxcull.so: one hand-built meshlet (a quad), one frame of the visibility
 * path (cull_meshes -> cull_meshlets -> raster -> resolve), results back on the host; `host_min alpha` then runs the same frame
 * with a material table whose 2x2 checker image makes the encode pass discard two quadrants of the quad
 * (visbuffer_encode.slang:54-66).
 *
 *   gcc -std=c11 -Iinclude examples/host_min.c -Loxylus_b200 -loxcull -Wl,-rpath,$PWD/oxylus_b200 -o host_min
 *
 * Mirrors what RendererInstance::update / render do with the engine's tables (Scene.cpp:1226-1290,
 * RendererInstance.cpp:842-884).  Without a CUDA device it stops at oxc_create with OXC_E_NO_DEVICE — there is no CPU
 * fallback (tests/test_abi_cpu.py builds and runs it to check exactly that). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oxcull.h"

static uint16_t half_bits(float v) { /* round-to-nearest half for the few constants below */
  uint32_t u;
  memcpy(&u, &v, 4);
  int s = (u >> 16) & 0x8000, em = (int)(u & 0x7fffffff);
  int h = (em - (112 << 23) + (1 << 12)) >> 13;
  if (em < (113 << 23)) h = 0;
  return (uint16_t)(s | h);
}

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    int rc_ = (call);                                                                                 \
    if (rc_ != OXC_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, oxc_last_error()); return rc_ == OXC_E_NO_DEVICE ? 3 : 1; } \
  } while (0)

int main(int argc, char** argv) {
  const int with_alpha = argc > 1 && strcmp(argv[1], "alpha") == 0; /* host_min alpha: second frame with a material table */
  enum { W = 64, H = 48 };
  /* ---- blob: positions | meshlet | bounds | micro indices | vertex indices | lod table (all 16-byte aligned) ---- */
  _Alignas(16) uint8_t blob[512];
  memset(blob, 0, sizeof blob);
  const float pos[4][3] = {{-0.5f, -0.5f, 0.5f}, {0.5f, -0.5f, 0.5f}, {0.5f, 0.5f, 0.5f}, {-0.5f, 0.5f, 0.5f}};
  uint16_t* vp = (uint16_t*)(blob + 0);
  for (int v = 0; v < 4; v++)
    for (int a = 0; a < 3; a++) vp[v * 4 + a] = half_bits(pos[v][a]);
  OxcMeshlet ml = {0, 0, 4, 2};
  memcpy(blob + 32, &ml, sizeof ml);
  OxcMeshletBounds mb;
  memset(&mb, 0, sizeof mb);
  mb.aabb_center[2] = half_bits(0.5f);
  mb.aabb_extent[0] = half_bits(1.0f); mb.aabb_extent[1] = half_bits(1.0f); mb.aabb_extent[2] = half_bits(0.01f);
  mb.cone_cutoff = 127; /* cone test disabled */
  memcpy(blob + 48, &mb, sizeof mb);
  const uint8_t micro[8] = {0, 2, 1, 0, 3, 2, 0, 0}; /* front faces: negative xyw determinant (cull.slang:169-171) */
  memcpy(blob + 64, micro, 8);
  const uint32_t vidx[4] = {0, 1, 2, 3};
  memcpy(blob + 80, vidx, 16);
  OxcMeshLOD lod;
  memset(&lod, 0, sizeof lod);
  lod.meshlets = 32; lod.meshlet_bounds = 48; lod.local_triangle_indices = 64; lod.indirect_vertex_indices = 80;
  lod.meshlet_count = 1; lod.meshlet_bounds_count = 1; lod.local_triangle_indices_count = 8; lod.indirect_vertex_indices_count = 4;
  memcpy(blob + 96, &lod, sizeof lod);
  uint16_t* tc = (uint16_t*)(blob + 160); /* texture coordinates: uv = pos.xy + 0.5, half2 per vertex (scene.slang:491-497) */
  for (int v = 0; v < 4; v++) { tc[v * 2 + 0] = half_bits(pos[v][0] + 0.5f); tc[v * 2 + 1] = half_bits(pos[v][1] + 0.5f); }

  OxcMesh mesh;
  memset(&mesh, 0, sizeof mesh);
  mesh.vertex_positions = 0; mesh.vertex_count = 4; mesh.lod_count = 1; mesh.lods = 96; mesh.texture_coords = 160;
  mesh.bounds.aabb_center[2] = 0.5f;
  mesh.bounds.aabb_extent[0] = 1.0f; mesh.bounds.aabb_extent[1] = 1.0f; mesh.bounds.aabb_extent[2] = 0.01f;
  OxcMeshInstance inst;
  memset(&inst, 0, sizeof inst);
  OxcTransformWorld xf;
  memset(&xf, 0, sizeof xf);
  xf.world[0] = xf.world[5] = xf.world[10] = xf.world[15] = 1.0f;
  OxcCullCamera cam;
  memset(&cam, 0, sizeof cam);
  cam.projection_view[0] = cam.projection_view[5] = cam.projection_view[10] = cam.projection_view[15] = 1.0f; /* clip == local */
  cam.position[2] = 10.0f; cam.acceptable_lod_error = 2.0f; cam.resolution[0] = W; cam.resolution[1] = H;
  cam.near_clip = 0.01f; cam.mesh_instance_count = 1;

  OxcCreateInfo info;
  memset(&info, 0, sizeof info);
  info.max_mesh_instances = 1; info.max_meshlet_instances = 1; info.hiz_width = 32; info.hiz_height = 32;
  OxcContext* ctx = NULL;
  CHECK(oxc_create(0, &info, &ctx));
  OxcSceneDesc desc = {&mesh, 1, &inst, 1, &xf, 1, blob, sizeof blob};
  CHECK(oxc_set_scene(ctx, &desc, NULL));

  void *vis64 = NULL, *vis32_dev = NULL;
  CHECK(oxc_device_alloc(ctx, (uint64_t)W * H * 8, &vis64));
  CHECK(oxc_device_alloc(ctx, (uint64_t)W * H * 4, &vis32_dev));
  CHECK(oxc_clear_visbuffer(ctx, (uint64_t*)vis64, W, H, NULL));
  CHECK(oxc_cull_meshes(ctx, &cam, OXC_CULL_TEST_ALL, NULL));
  CHECK(oxc_cull_meshlets(ctx, &cam, OXC_CULL_TEST_FRUSTUM, /*use_hiz=*/0, NULL)); /* plain variant, CullGeometry.cpp:275 */
  CHECK(oxc_raster_visbuffer(ctx, &cam, OXC_CULL_TEST_ALL, W, H, (uint64_t*)vis64, 0, NULL));
  CHECK(oxc_resolve_visbuffer(ctx, (const uint64_t*)vis64, W, H, (uint32_t*)vis32_dev, NULL, NULL));
  static uint32_t vis32[W * H];
  CHECK(oxc_copy(ctx, vis32, vis32_dev, sizeof vis32, /*device->host*/ 1, NULL));
  CHECK(oxc_sync(ctx, NULL));
  unsigned covered = 0;
  for (int i = 0; i < W * H; i++) covered += vis32[i] != OXC_VIS_CLEAR;
  printf("%s: %u of %d pixels covered by the quad (expected %d)\n", oxc_version(), covered, W * H, (W / 2) * (H / 2));
  if (!with_alpha) {
    oxc_device_free(ctx, vis64);
    oxc_device_free(ctx, vis32_dev);
    oxc_destroy(ctx);
    return covered == (W / 2) * (H / 2) ? 0 : 2;
  }
  /* ---- the same frame with an alpha-tested material: 2x2 R8 checker, nearest + clamp, cutoff 0.5 ---- */
  const uint8_t texels[4] = {255, 0, 0, 255};
  void* tex_dev = NULL;
  CHECK(oxc_device_alloc(ctx, sizeof texels, &tex_dev));
  CHECK(oxc_copy(ctx, tex_dev, texels, sizeof texels, /*host->device*/ 0, NULL));
  OxcMaterial mat;
  memset(&mat, 0, sizeof mat);
  mat.albedo_color[3] = half_bits(1.0f);
  mat.alpha_cutoff = half_bits(0.5f);
  mat.flags = OXC_MATERIAL_HAS_ALBEDO_IMAGE | OXC_MATERIAL_ALPHA_MASK;
  OxcAlphaImage image = {tex_dev, 2, 2, OXC_IMAGE_R8_UNORM, 0};
  OxcSamplerDesc sampler = {OXC_FILTER_NEAREST, OXC_FILTER_NEAREST, OXC_MIPMAP_NEAREST, OXC_ADDRESS_CLAMP_TO_EDGE, OXC_ADDRESS_CLAMP_TO_EDGE};
  OxcMaterialTable table = {&mat, 1, &image, 1, &sampler, 1};
  CHECK(oxc_set_materials(ctx, &table, NULL)); /* MeshInstance::material_index is 0 */
  CHECK(oxc_clear_visbuffer(ctx, (uint64_t*)vis64, W, H, NULL));
  CHECK(oxc_raster_visbuffer(ctx, &cam, OXC_CULL_TEST_ALL, W, H, (uint64_t*)vis64, 0, NULL));
  CHECK(oxc_resolve_visbuffer(ctx, (const uint64_t*)vis64, W, H, (uint32_t*)vis32_dev, NULL, NULL));
  CHECK(oxc_copy(ctx, vis32, vis32_dev, sizeof vis32, /*device->host*/ 1, NULL));
  CHECK(oxc_check_status(ctx, NULL, NULL));
  unsigned kept = 0, wrong = 0;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const int in_quad = x >= W / 4 && x < 3 * W / 4 && y >= H / 4 && y < 3 * H / 4;
      const int opaque_texel = (x < W / 2) == (y < H / 2); /* texels (0,0) and (1,1) are 255 */
      const int drawn = vis32[y * W + x] != OXC_VIS_CLEAR;
      kept += drawn;
      wrong += drawn != (in_quad && opaque_texel);
    }
  printf("alpha-tested: %u of %u quad pixels kept, %u pixels differ from the checker (expected %d kept, 0 differ)\n", kept, covered, wrong,
         (W / 2) * (H / 2) / 2);
  oxc_device_free(ctx, tex_dev);
  oxc_device_free(ctx, vis64);
  oxc_device_free(ctx, vis32_dev);
  oxc_destroy(ctx);
  return covered == (W / 2) * (H / 2) && kept == (W / 2) * (H / 2) / 2 && wrong == 0 ? 0 : 2;
}
