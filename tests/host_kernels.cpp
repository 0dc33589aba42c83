// host_kernels.cpp — runs thread-independent KERNELS of the product on the HOST, one simulated thread at a time, so the CPU tier
// can compare them with the oracle on the same synthetic scenes the GPU parity tests use (tests/test_host_kernels_cpu.py):
//   k_cull_meshes            cull_meshes.slang:17-61: mesh-level frustum test, LOD selection, meshlet counts, lod_index write-back,
//                            InstCull / InstGeom (one thread per mesh instance)
//   k_decode_visbuffer       visbuffer_decode.slang:42-183, geometry part (one thread per pixel)
// The kernel source is compiled unchanged through tests/host_shim/ (each __f*_rn intrinsic = one IEEE binary32 operation under
// -ffp-contract=off; blockIdx / threadIdx are variables; cross-thread primitives are stubs whose results are not used here).
// TEST INFRASTRUCTURE ONLY: nothing in oxylus_b200/ builds, links or loads this; the product has no CPU path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "host_launch.h"

#define OXC_HOST_SOUNDNESS_HARNESS
#include "kernels_cull.cuh"
#include "kernels_decode.cuh"

namespace oxc {
float rcp_approx(float x) { return 1.0f / x; }
float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
} // namespace oxc

using namespace oxc;

namespace {
// what oxc_set_scene does on the device, on the host: tables copied, every blob offset rebased to an address
struct HostScene {
  std::vector<uint8_t> blob;
  std::vector<OxcMesh> meshes;
  std::vector<OxcMeshInstance> mesh_instances;
  std::vector<OxcTransformWorld> transforms;
  std::vector<float> lod_aabb;
  std::vector<InstCull> inst;
  std::vector<InstGeom> geom;
  std::vector<uint32_t> counts, block_sums;
};
} // namespace

extern "C" {

void* hk_scene_create(const OxcSceneDesc* sc) {
  HostScene* h = new HostScene();
  h->blob.assign(sc->blob, sc->blob + sc->blob_size);
  h->meshes.assign(sc->meshes, sc->meshes + sc->mesh_count);
  h->mesh_instances.assign(sc->mesh_instances, sc->mesh_instances + sc->mesh_instance_count);
  h->transforms.assign(sc->transforms, sc->transforms + sc->transform_count);
  const uint64_t base = reinterpret_cast<uint64_t>(h->blob.data());
  h->lod_aabb.assign((size_t)sc->mesh_count * OXC_MESH_MAX_LODS * 6, 0.0f);
  for (uint32_t m = 0; m < sc->mesh_count; m++) {
    OxcMesh& me = h->meshes[m];
    OxcMeshLOD* lods = reinterpret_cast<OxcMeshLOD*>(h->blob.data() + me.lods);
    for (uint32_t l = 0; l < OXC_MESH_MAX_LODS; l++) {
      float* o = &h->lod_aabb[((size_t)m * OXC_MESH_MAX_LODS + l) * 6];
      o[0] = o[1] = o[2] = 1.0f; o[3] = o[4] = o[5] = -1.0f; // empty => shortcut disabled
      if (l >= me.lod_count) continue;
      OxcMeshLOD& d = lods[l];
      // union AABB of the decoded meshlet boxes (k_lod_union_aabb, a warp kernel: restated here, it only feeds the shortcut flag)
      float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
      bool bad = d.meshlet_bounds_count == 0;
      const uint4* b = reinterpret_cast<const uint4*>(h->blob.data() + d.meshlet_bounds);
      for (uint32_t i = 0; i < d.meshlet_bounds_count; i++) {
        const uint4 v = b[i];
        const float c[3] = {dequantize_half(v.x & 0xFFFFu), dequantize_half(v.x >> 16), dequantize_half(v.y & 0xFFFFu)};
        const float e[3] = {dequantize_half(v.z & 0xFFFFu), dequantize_half(v.z >> 16), dequantize_half(v.w & 0xFFFFu)};
        for (int a = 0; a < 3; a++) {
          const float hh = fabsf(e[a]) * 0.5f, a0 = c[a] - hh, a1 = c[a] + hh;
          bad = bad || !(fabsf(a0) <= 3.0e38f) || !(fabsf(a1) <= 3.0e38f) || !(e[a] >= 0.0f);
          mn[a] = fminf(mn[a], a0); mx[a] = fmaxf(mx[a], a1);
        }
      }
      if (!bad) { o[0] = mn[0]; o[1] = mn[1]; o[2] = mn[2]; o[3] = mx[0]; o[4] = mx[1]; o[5] = mx[2]; }
      d.indices += base; d.meshlets += base; d.meshlet_bounds += base; d.local_triangle_indices += base; d.indirect_vertex_indices += base;
    }
    me.lods += base;
    me.vertex_positions += base;
    if (me.vertex_normals) me.vertex_normals += base;
    if (me.texture_coords) me.texture_coords += base;
  }
  const size_t n = h->mesh_instances.size();
  h->inst.resize(n); h->geom.resize(n); h->counts.assign(n, 0);
  h->block_sums.assign(n / CULL_MESHES_THREADS + 2, 0);
  return h;
}
void hk_scene_destroy(void* s) { delete static_cast<HostScene*>(s); }

static MeshesParams meshes_params(HostScene* h, const OxcCullCamera* cam, uint32_t flags, uint32_t first, uint32_t count, int select) {
  MeshesParams p{};
  p.meshes = h->meshes.data(); p.mesh_instances = h->mesh_instances.data(); p.transforms = h->transforms.data();
  p.inst = h->inst.data(); p.geom = h->geom.data(); p.lod_aabb = h->lod_aabb.data(); p.counts = h->counts.data(); p.block_sums = h->block_sums.data();
  p.first = first; p.count = count; p.flags = flags; p.select = select; p.cam = *cam; p.cam_dev = nullptr;
  return p;
}

// k_cull_meshes over the shard [first, first + count): counts_out[i] = meshlets emitted by mesh instance first + i,
// lod_index_out[j] = MeshInstance::lod_index of EVERY mesh instance after the pass (write-back, cull_meshes.slang:76)
int hk_cull_meshes(void* s, const OxcCullCamera* cam, uint32_t flags, uint32_t first, uint32_t count, uint32_t* counts_out, uint32_t* lod_index_out) {
  HostScene* h = static_cast<HostScene*>(s);
  if (first + (uint64_t)count > h->mesh_instances.size()) return -1;
  const MeshesParams p = meshes_params(h, cam, flags, first, count, 1);
  const uint32_t blocks = (count + CULL_MESHES_THREADS - 1) / CULL_MESHES_THREADS;
  blockDim.x = CULL_MESHES_THREADS; gridDim.x = blocks;
  for (uint32_t b = 0; b < blocks; b++)
    for (uint32_t t = 0; t < (uint32_t)CULL_MESHES_THREADS; t++) {
      blockIdx.x = b; threadIdx.x = t;
      k_cull_meshes(p);
    }
  for (uint32_t i = 0; i < count; i++) counts_out[i] = h->counts[i];
  for (size_t j = 0; j < h->mesh_instances.size(); j++) lod_index_out[j] = h->mesh_instances[j].lod_index;
  return 0;
}

// k_decode_visbuffer after hk_cull_meshes of the same camera (it resolves LODs and pointers, like the frame's oxc_cull_meshes)
int hk_decode(void* s, const OxcCullCamera* cam, const uint32_t* vis32, const uint64_t* vis64, uint32_t width, uint32_t height,
              const OxcMeshletInstance* meshlet_instances, uint32_t total, uint32_t prim_bits, float* lambda, float* ddx, float* ddy, float* uv_normal,
              float* uv_grad) {
  HostScene* h = static_cast<HostScene*>(s);
  OxcMeshletInstanceVisibility vis{};
  vis.total_visible_meshlet_instances = total;
  DecodeParams p{};
  p.vis64 = reinterpret_cast<const unsigned long long*>(vis64); p.vis32 = vis32; p.meshlet_instances = meshlet_instances; p.vis = &vis;
  p.inst = h->inst.data(); p.geom = h->geom.data(); p.id_base = nullptr;
  p.lambda = reinterpret_cast<float4*>(lambda); p.ddx = reinterpret_cast<float4*>(ddx); p.ddy = reinterpret_cast<float4*>(ddy);
  p.uv_normal = reinterpret_cast<float4*>(uv_normal); p.uv_grad = reinterpret_cast<float4*>(uv_grad);
  const float* m = cam->projection_view;
  for (int i = 0; i < 4; i++) p.pv_row[i] = make_float4(m[i], m[4 + i], m[8 + i], m[12 + i]); // oxc_decode_visbuffer
  p.res_x = cam->resolution[0]; p.res_y = cam->resolution[1];
  p.width = width; p.height = height; p.prim_bits = prim_bits;
  blockDim.x = DECODE_TX; blockDim.y = DECODE_TY;
  gridDim.x = (width + DECODE_TX - 1) / DECODE_TX; gridDim.y = (height + DECODE_TY - 1) / DECODE_TY;
  for (uint32_t by = 0; by < gridDim.y; by++)
    for (uint32_t bx = 0; bx < gridDim.x; bx++)
      for (uint32_t ty = 0; ty < (uint32_t)DECODE_TY; ty++)
        for (uint32_t tx = 0; tx < (uint32_t)DECODE_TX; tx++) {
          blockIdx.x = bx; blockIdx.y = by; threadIdx.x = tx; threadIdx.y = ty;
          k_decode_visbuffer(p);
        }
  return 0;
}

} // extern "C"
