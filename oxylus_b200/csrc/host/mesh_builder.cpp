// mesh_builder.cpp — host-side producer of the Mesh / MeshLOD / Meshlet / MeshletBounds blob the visibility path
// consumes.  Mirrors build_gltf_mesh (Oxylus/src/Asset/AssetManager_GLTF.cpp:481-771) from the point
// where the glTF accessors have been read: vertex fetch remap, half / snorm quantisation, per-LOD meshlet build, meshlet
// AABB + normal cone, blob layout and offset bookkeeping (upload_gltf_mesh :773-800 is oxc_set_scene's rebase).
//
// The reference delegates four steps to meshoptimizer v1.2 (xmake/packages.lua:9), which is not part of
// /root/reference and not installed here.  They are restated from the library's published definitions:
//   meshopt_optimizeVertexFetchRemap  new index = order of first use in the index buffer, unused vertices dropped
//   meshopt_quantizeHalf              round-to-nearest by +0x1000 then truncate, exponents below -14 flush to zero,
//                                     overflow -> inf, NaN -> 0x7e00
//   meshopt_quantizeSnorm(v, N)       int(clamp(v, -1, 1) * (2^(N-1) - 1) + (v >= 0 ? 0.5 : -0.5))
//   meshopt_computeMeshletBounds      cone axis = centre of Ritter's bounding sphere of the unit triangle normals,
//                                     cutoff = sqrt(1 - mindp^2) widened by the s8 quantisation error (+1 ulp of s8);
//                                     mindp <= 0.1 -> cutoff 127 (cone test disabled, cull.slang:173 `cutoff >= 1.0`)
// meshopt_buildMeshlets: cluster_mode 1 runs a greedy spatial clusteriser of the same published scheme (adjacency-first
// growth, nearest-centroid reseeding — spatial_triangle_order below), cluster_mode 0 keeps the caller's triangle order; either
// way meshlets are then formed by a linear scan that closes a meshlet when the next triangle would exceed 64 vertices or 64
// triangles (Model::MAX_MESHLET_INDICES / MAX_MESHLET_PRIMITIVES, Model.hpp:27-28).  meshopt_simplifyWithAttributes: coarser
// LODs are either index buffers the caller supplies or, with auto_lods, the chain of :596-641 produced by the edge-collapse
// simplifier in mesh_simplifier.cpp.  PARITY UNPINNED against meshoptimizer.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "../../../include/oxcull.h"
#include "mesh_simplifier.hpp"

namespace {

thread_local std::string g_builder_error;

inline uint16_t quantize_half(float v) {
  uint32_t ui;
  std::memcpy(&ui, &v, 4);
  const int s = (int)((ui >> 16) & 0x8000u);
  const int em = (int)(ui & 0x7fffffffu);
  int h = (em - (112 << 23) + (1 << 12)) >> 13; // bias exponent, round to nearest
  h = (em < (113 << 23)) ? 0 : h;               // underflow: flush to zero
  h = (em >= (143 << 23)) ? 0x7c00 : h;         // overflow: infinity
  h = (em > (255 << 23)) ? 0x7e00 : h;          // NaN
  return (uint16_t)(s | h);
}

inline int quantize_snorm(float v, int bits) {
  const float scale = (float)((1 << (bits - 1)) - 1);
  const float round = v >= 0.0f ? 0.5f : -0.5f;
  v = v >= -1.0f ? v : -1.0f;
  v = v <= 1.0f ? v : 1.0f;
  return (int)(v * scale + round);
}

struct Blob {
  std::vector<uint8_t> bytes;
  uint64_t append(const void* data, size_t size, size_t align) { // blob_append, AssetManager_GLTF.cpp:466-479
    const size_t off = (bytes.size() + align - 1) / align * align;
    bytes.resize(off + size);
    if (size) std::memcpy(bytes.data() + off, data, size);
    return off;
  }
};

// Ritter's bounding sphere, the variant meshoptimizer uses for cluster bounds.
void bounding_sphere(float out[4], const float (*pts)[3], size_t count) {
  size_t pmin[3] = {0, 0, 0}, pmax[3] = {0, 0, 0};
  for (size_t i = 0; i < count; i++)
    for (int a = 0; a < 3; a++) {
      pmin[a] = pts[i][a] < pts[pmin[a]][a] ? i : pmin[a];
      pmax[a] = pts[i][a] > pts[pmax[a]][a] ? i : pmax[a];
    }
  float best = 0.0f;
  int axis = 0;
  for (int a = 0; a < 3; a++) {
    const float* p1 = pts[pmin[a]];
    const float* p2 = pts[pmax[a]];
    const float d2 = (p2[0] - p1[0]) * (p2[0] - p1[0]) + (p2[1] - p1[1]) * (p2[1] - p1[1]) + (p2[2] - p1[2]) * (p2[2] - p1[2]);
    if (d2 > best) { best = d2; axis = a; }
  }
  const float* p1 = pts[pmin[axis]];
  const float* p2 = pts[pmax[axis]];
  float c[3] = {(p1[0] + p2[0]) / 2, (p1[1] + p2[1]) / 2, (p1[2] + p2[2]) / 2};
  float r = std::sqrt(best) / 2;
  for (size_t i = 0; i < count; i++) {
    const float* p = pts[i];
    const float d2 = (p[0] - c[0]) * (p[0] - c[0]) + (p[1] - c[1]) * (p[1] - c[1]) + (p[2] - c[2]) * (p[2] - c[2]);
    if (d2 > r * r) {
      const float d = std::sqrt(d2);
      const float k = 0.5f + (r / d) / 2;
      c[0] = c[0] * k + p[0] * (1 - k);
      c[1] = c[1] * k + p[1] * (1 - k);
      c[2] = c[2] * k + p[2] * (1 - k);
      r = (r + d) / 2;
    }
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = r;
}

struct ScanMeshlet {
  uint32_t vertex_offset, triangle_offset, vertex_count, triangle_count;
};

// Spatial clusteriser (cluster_mode 1) — the role meshopt_buildMeshlets' kd-tree guided growth plays in the reference
// (AssetManager_GLTF.cpp:630-676).  It only REORDERS the triangles; the linear scan below then closes meshlets at the same
// 64-vertex / 64-triangle limits, so "order + scan" is the whole definition.  Greedy growth, meshoptimizer's published scheme:
//   * the next triangle is the unused one adjacent to the current meshlet that adds the fewest new vertices (one less when
//     it is the last unused triangle of one of its vertices); ties: nearest centroid to the meshlet's centre, then lowest index;
//   * when no unused triangle touches the meshlet, the unused triangle whose centroid is nearest to the meshlet's centre
//     (uniform grid over the centroids instead of a kd-tree: same query, simpler structure);
//   * a meshlet is closed when the chosen triangle does not fit.
// Deterministic (no hashing, ties by index).  Not bit-compatible with meshoptimizer's clusters (parity unpinned).
std::vector<uint32_t> spatial_triangle_order(const std::vector<uint32_t>& indices, const std::vector<float>& positions, uint32_t vertex_count) {
  const uint32_t T = (uint32_t)(indices.size() / 3);
  const uint32_t NONE = 0xFFFFFFFFu;
  // vertex -> triangles (CSR)
  std::vector<uint32_t> adj_off(vertex_count + 1, 0);
  for (uint32_t i = 0; i < T * 3; i++) adj_off[indices[i] + 1]++;
  for (uint32_t v = 0; v < vertex_count; v++) adj_off[v + 1] += adj_off[v];
  std::vector<uint32_t> adj(T * 3), fill(adj_off.begin(), adj_off.end() - 1);
  for (uint32_t t = 0; t < T; t++)
    for (int k = 0; k < 3; k++) adj[fill[indices[t * 3 + k]]++] = t;
  std::vector<uint32_t> live(vertex_count, 0);
  for (uint32_t v = 0; v < vertex_count; v++) live[v] = adj_off[v + 1] - adj_off[v];
  // centroids + uniform grid
  std::vector<float> cen((size_t)T * 3);
  float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
  for (uint32_t t = 0; t < T; t++)
    for (int a = 0; a < 3; a++) {
      const float c = (positions[(size_t)indices[t * 3] * 3 + a] + positions[(size_t)indices[t * 3 + 1] * 3 + a] + positions[(size_t)indices[t * 3 + 2] * 3 + a]) / 3.0f;
      cen[(size_t)t * 3 + a] = c;
      lo[a] = c < lo[a] ? c : lo[a]; hi[a] = c > hi[a] ? c : hi[a];
    }
  int G = 1;
  while ((size_t)G * G * G * 8 < T && G < 128) G++;
  float inv_cell[3];
  for (int a = 0; a < 3; a++) inv_cell[a] = hi[a] > lo[a] ? (float)G / (hi[a] - lo[a]) : 0.0f;
  auto cell_of = [&](const float* c, int out[3]) {
    for (int a = 0; a < 3; a++) {
      int g = (int)((c[a] - lo[a]) * inv_cell[a]);
      out[a] = g < 0 ? 0 : (g >= G ? G - 1 : g);
    }
  };
  std::vector<uint32_t> cell_off((size_t)G * G * G + 1, 0), cell_tri(T);
  std::vector<uint32_t> tri_cell(T);
  for (uint32_t t = 0; t < T; t++) {
    int g[3];
    cell_of(&cen[(size_t)t * 3], g);
    tri_cell[t] = (uint32_t)((g[2] * G + g[1]) * G + g[0]);
    cell_off[tri_cell[t] + 1]++;
  }
  for (size_t c = 0; c < cell_off.size() - 1; c++) cell_off[c + 1] += cell_off[c];
  {
    std::vector<uint32_t> f(cell_off.begin(), cell_off.end() - 1);
    for (uint32_t t = 0; t < T; t++) cell_tri[f[tri_cell[t]]++] = t;
  }
  std::vector<uint32_t> cell_live(cell_off.size() - 1);
  for (size_t c = 0; c + 1 < cell_off.size(); c++) cell_live[c] = cell_off[c + 1] - cell_off[c];
  std::vector<uint8_t> used(T, 0);
  auto nearest_unused = [&](const float* q) -> uint32_t {
    int g[3];
    cell_of(q, g);
    uint32_t best = NONE;
    float best_d = 3e38f;
    for (int ring = 0; ring < G; ring++) {
      // once a candidate exists, one more ring is enough when the ring's inner distance exceeds it; keep it simple: search until a
      // ring beyond the first hit has been scanned
      bool any_cell = false;
      for (int z = g[2] - ring; z <= g[2] + ring; z++)
        for (int y = g[1] - ring; y <= g[1] + ring; y++)
          for (int x = g[0] - ring; x <= g[0] + ring; x++) {
            if (x < 0 || y < 0 || z < 0 || x >= G || y >= G || z >= G) continue;
            if (ring && x != g[0] - ring && x != g[0] + ring && y != g[1] - ring && y != g[1] + ring && z != g[2] - ring && z != g[2] + ring) continue; // shell only
            const size_t c = (size_t)(z * G + y) * G + x;
            if (!cell_live[c]) continue;
            any_cell = true;
            for (uint32_t k = cell_off[c]; k < cell_off[c + 1]; k++) {
              const uint32_t t = cell_tri[k];
              if (used[t]) continue;
              const float dx = cen[(size_t)t * 3] - q[0], dy = cen[(size_t)t * 3 + 1] - q[1], dz = cen[(size_t)t * 3 + 2] - q[2];
              const float d = dx * dx + dy * dy + dz * dz;
              if (d < best_d || (d == best_d && t < best)) { best_d = d; best = t; }
            }
          }
      (void)any_cell;
      if (best != NONE) {
        // cells of the next shells are at least `ring` cells away along some axis: stop once that bound exceeds the best distance
        float min_cell = 3e38f;
        for (int a = 0; a < 3; a++)
          if (inv_cell[a] > 0.0f) { const float cs = 1.0f / inv_cell[a]; min_cell = cs < min_cell ? cs : min_cell; }
        if (min_cell == 3e38f || (float)ring * min_cell * (float)ring * min_cell >= best_d) break;
      }
    }
    return best;
  };
  std::vector<uint32_t> order;
  order.reserve(T);
  std::vector<uint32_t> slot(vertex_count, NONE); // meshlet-local slot of a vertex
  std::vector<uint32_t> mverts;
  uint32_t mtris = 0;
  float centre[3] = {0, 0, 0};
  float last[3] = {cen.empty() ? 0.0f : cen[0], cen.empty() ? 0.0f : cen[1], cen.empty() ? 0.0f : cen[2]};
  auto close = [&]() {
    for (uint32_t v : mverts) slot[v] = NONE;
    mverts.clear();
    mtris = 0;
    centre[0] = centre[1] = centre[2] = 0.0f;
  };
  for (uint32_t done = 0; done < T; done++) {
    uint32_t best = NONE, best_extra = 4;
    float best_d = 3e38f;
    float mc[3] = {0, 0, 0};
    if (mtris) for (int a = 0; a < 3; a++) mc[a] = centre[a] / (float)mtris;
    for (uint32_t v : mverts)
      for (uint32_t k = adj_off[v]; k < adj_off[v + 1]; k++) {
        const uint32_t t = adj[k];
        if (used[t]) continue;
        const uint32_t a = indices[t * 3], b = indices[t * 3 + 1], c = indices[t * 3 + 2];
        uint32_t extra = (slot[a] == NONE) + (slot[b] == NONE && b != a) + (slot[c] == NONE && c != a && c != b);
        // a vertex whose last unused triangle this is gets "finished" by taking it: meshoptimizer's live-triangle bonus
        if (extra && (live[a] == 1 || live[b] == 1 || live[c] == 1)) extra--;
        const float dx = cen[(size_t)t * 3] - mc[0], dy = cen[(size_t)t * 3 + 1] - mc[1], dz = cen[(size_t)t * 3 + 2] - mc[2];
        const float d = dx * dx + dy * dy + dz * dz;
        if (extra < best_extra || (extra == best_extra && (d < best_d || (d == best_d && t < best)))) { best = t; best_extra = extra; best_d = d; }
      }
    if (best != NONE) { // the real number of new vertices (the bonus above is a preference, not a count)
      const uint32_t a = indices[best * 3], b = indices[best * 3 + 1], c = indices[best * 3 + 2];
      best_extra = (slot[a] == NONE) + (slot[b] == NONE && b != a) + (slot[c] == NONE && c != a && c != b);
    }
    if (best == NONE) {
      float q[3];
      if (mtris) for (int a = 0; a < 3; a++) q[a] = centre[a] / (float)mtris;
      else for (int a = 0; a < 3; a++) q[a] = last[a];
      best = nearest_unused(q);
      const uint32_t a = indices[best * 3], b = indices[best * 3 + 1], c = indices[best * 3 + 2];
      best_extra = (slot[a] == NONE) + (slot[b] == NONE && b != a) + (slot[c] == NONE && c != a && c != b);
    }
    if (mverts.size() + best_extra > OXC_MESHLET_MAX_VERTICES || mtris >= OXC_MESHLET_MAX_PRIMITIVES) {
      if (mtris) for (int a = 0; a < 3; a++) last[a] = centre[a] / (float)mtris;
      close();
    }
    for (int k = 0; k < 3; k++) {
      const uint32_t v = indices[best * 3 + k];
      if (slot[v] == NONE) { slot[v] = (uint32_t)mverts.size(); mverts.push_back(v); }
      live[v]--;
    }
    for (int a = 0; a < 3; a++) centre[a] += cen[(size_t)best * 3 + a];
    mtris++;
    used[best] = 1;
    cell_live[tri_cell[best]]--;
    order.push_back(best);
  }
  return order;
}

} // namespace

struct OxbMesh {
  std::vector<uint8_t> blob;
  OxcMesh mesh{};
  uint32_t meshlet_counts[OXC_MESH_MAX_LODS] = {};
  std::vector<uint32_t> vertex_remap; // input vertex -> blob vertex (0xFFFFFFFF = unused)
};

extern "C" {

const char* oxb_last_error(void) { return g_builder_error.c_str(); }

int oxb_build_mesh(const OxbMeshInput* in, OxbMesh** out) {
  if (!in || !out) { g_builder_error = "null argument"; return OXC_E_INVALID; }
  *out = nullptr;
  if (!in->positions || in->vertex_count == 0 || in->lod_count == 0 || in->lod_count > OXC_MESH_MAX_LODS || !in->lod_indices[0] ||
      in->lod_index_counts[0] < 3) {
    g_builder_error = "positions and at least one triangle of LOD 0 are required (build_gltf_mesh returns nullopt, :485-487,758-760)";
    return OXC_E_INVALID;
  }
  if (in->cluster_mode > 1) { g_builder_error = "cluster_mode must be 0 (caller order) or 1 (spatial)"; return OXC_E_INVALID; }
  if (in->auto_lods > 1) { g_builder_error = "auto_lods must be 0 or 1"; return OXC_E_INVALID; }
  if (in->auto_lods && in->lod_count != 1) { g_builder_error = "auto_lods generates LOD 1.. itself: pass LOD 0 only"; return OXC_E_INVALID; }
  for (uint32_t l = 0; l < in->lod_count; l++) {
    if (!in->lod_indices[l] || in->lod_index_counts[l] % 3u) { g_builder_error = "LOD index buffers must be triangle lists"; return OXC_E_INVALID; }
    for (uint32_t i = 0; i < in->lod_index_counts[l]; i++)
      if (in->lod_indices[l][i] >= in->vertex_count) { g_builder_error = "index out of range"; return OXC_E_INVALID; }
  }
  OxbMesh* m = new (std::nothrow) OxbMesh();
  if (!m) { g_builder_error = "out of memory"; return OXC_E_CUDA; }

  // ---- vertex fetch remap over LOD 0's index buffer (:515-528); coarser LODs only reference LOD-0 vertices ----
  const uint32_t NONE = 0xFFFFFFFFu;
  m->vertex_remap.assign(in->vertex_count, NONE);
  uint32_t vertex_count = 0;
  for (uint32_t i = 0; i < in->lod_index_counts[0]; i++) {
    uint32_t& r = m->vertex_remap[in->lod_indices[0][i]];
    if (r == NONE) r = vertex_count++;
  }
  for (uint32_t l = 1; l < in->lod_count; l++)
    for (uint32_t i = 0; i < in->lod_index_counts[l]; i++)
      if (m->vertex_remap[in->lod_indices[l][i]] == NONE) {
        g_builder_error = "coarser LODs may only use vertices LOD 0 uses (simplification without new vertices, :604-628)";
        delete m;
        return OXC_E_INVALID;
      }
  std::vector<float> positions((size_t)vertex_count * 3);
  for (uint32_t v = 0; v < in->vertex_count; v++)
    if (m->vertex_remap[v] != NONE) std::memcpy(&positions[(size_t)m->vertex_remap[v] * 3], &in->positions[(size_t)v * 3], 12);

  // ---- quantised vertex streams (:568-594) ----
  Blob blob;
  {
    std::vector<uint16_t> q((size_t)vertex_count * 4, 0);
    for (uint32_t v = 0; v < vertex_count; v++)
      for (int a = 0; a < 3; a++) q[(size_t)v * 4 + a] = quantize_half(positions[(size_t)v * 3 + a]);
    m->mesh.vertex_positions = blob.append(q.data(), q.size() * 2, 16);
  }
  if (in->normals) {
    std::vector<uint32_t> q(vertex_count, 0);
    for (uint32_t v = 0; v < in->vertex_count; v++) {
      const uint32_t r = m->vertex_remap[v];
      if (r == NONE) continue;
      const float* n = &in->normals[(size_t)v * 3];
      q[r] = ((uint32_t)(quantize_snorm(n[0], 10) + 511) << 20) | ((uint32_t)(quantize_snorm(n[1], 10) + 511) << 10) |
             (uint32_t)(quantize_snorm(n[2], 10) + 511); // :579-581
    }
    m->mesh.vertex_normals = blob.append(q.data(), q.size() * 4, 16);
  }
  if (in->texcoords) {
    std::vector<uint16_t> q((size_t)vertex_count * 2, 0);
    for (uint32_t v = 0; v < in->vertex_count; v++) {
      const uint32_t r = m->vertex_remap[v];
      if (r == NONE) continue;
      q[(size_t)r * 2 + 0] = quantize_half(in->texcoords[(size_t)v * 2 + 0]);
      q[(size_t)r * 2 + 1] = quantize_half(in->texcoords[(size_t)v * 2 + 1]);
    }
    m->mesh.texture_coords = blob.append(q.data(), q.size() * 2, 16);
  }
  m->mesh.vertex_count = vertex_count;

  // ---- per LOD: meshlets, bounds, index tables (:596-756) ----
  OxcMeshLOD lods[OXC_MESH_MAX_LODS];
  std::memset(lods, 0, sizeof lods);
  const float FMAX = std::numeric_limits<float>::max(), FLOW = std::numeric_limits<float>::lowest();
  float mesh_min[3] = {FMAX, FMAX, FMAX}, mesh_max[3] = {FLOW, FLOW, FLOW};
  // LOD index buffers in blob vertex numbering: the caller's, or the simplification chain of :596-641 (each LOD simplified
  // from the previous one to half its index count, error accumulated; the chain ends when the simplifier stalls more than
  // 50 % above the target, the step's relative error exceeds 0.5 or fewer than two triangles are left)
  std::vector<std::vector<uint32_t>> lod_indices(in->lod_count);
  std::vector<float> lod_errors(in->lod_count);
  for (uint32_t l = 0; l < in->lod_count; l++) {
    lod_indices[l].resize(in->lod_index_counts[l]);
    for (uint32_t i = 0; i < in->lod_index_counts[l]; i++) lod_indices[l][i] = m->vertex_remap[in->lod_indices[l][i]];
    lod_errors[l] = in->lod_errors[l];
  }
  if (in->auto_lods) {
    std::vector<float> normals;
    if (in->normals) {
      normals.resize((size_t)vertex_count * 3);
      for (uint32_t v = 0; v < in->vertex_count; v++)
        if (m->vertex_remap[v] != NONE) std::memcpy(&normals[(size_t)m->vertex_remap[v] * 3], &in->normals[(size_t)v * 3], 12);
    }
    for (uint32_t l = 1; l < OXC_MESH_MAX_LODS; l++) {
      const std::vector<uint32_t>& last = lod_indices.back();
      const size_t target = (last.size() + 5) / 6 * 3;
      float step_error = 0.0f;
      std::vector<uint32_t> simplified = oxb::simplify(last.data(), last.size(), positions.data(), in->normals ? normals.data() : nullptr, vertex_count,
                                                       target, std::numeric_limits<float>::max(), &step_error);
      const float error = lod_errors.back() + step_error;
      if (simplified.size() > target + target / 2 || step_error > 0.5f || simplified.size() < 6) break;
      lod_indices.push_back(std::move(simplified));
      lod_errors.push_back(error);
    }
  }
  const uint32_t lod_total = (uint32_t)lod_indices.size();
  for (uint32_t l = 0; l < lod_total; l++) {
    std::vector<uint32_t>& indices = lod_indices[l];
    const uint32_t index_count = (uint32_t)indices.size();
    if (in->cluster_mode == 1) { // spatial clustering: reorder the triangles, then scan
      const std::vector<uint32_t> order = spatial_triangle_order(indices, positions, vertex_count);
      std::vector<uint32_t> sorted(index_count);
      for (size_t k = 0; k < order.size(); k++)
        for (int c = 0; c < 3; c++) sorted[k * 3 + c] = indices[(size_t)order[k] * 3 + c];
      indices.swap(sorted);
    }

    // linear-scan clustering: a meshlet closes when the next triangle does not fit
    std::vector<ScanMeshlet> raw;
    std::vector<uint32_t> vertex_indices; // indirect_vertex_indices
    std::vector<uint8_t> micro;           // local_triangle_indices, (triangle_count*3 + 3) & ~3 bytes per meshlet (:689)
    std::vector<uint32_t> slot(vertex_count, NONE);
    ScanMeshlet cur{0, 0, 0, 0};
    auto close = [&]() {
      if (cur.triangle_count == 0) return;
      for (uint32_t k = 0; k < cur.vertex_count; k++) slot[vertex_indices[cur.vertex_offset + k]] = NONE;
      micro.resize(cur.triangle_offset + ((cur.triangle_count * 3u + 3u) & ~3u), 0);
      raw.push_back(cur);
      cur = ScanMeshlet{(uint32_t)vertex_indices.size(), (uint32_t)micro.size(), 0, 0};
    };
    for (uint32_t t = 0; t + 2 < index_count; t += 3) {
      const uint32_t a = indices[t], b = indices[t + 1], c = indices[t + 2];
      uint32_t extra = (slot[a] == NONE) + (slot[b] == NONE && b != a) + (slot[c] == NONE && c != a && c != b);
      if (cur.vertex_count + extra > OXC_MESHLET_MAX_VERTICES || cur.triangle_count >= OXC_MESHLET_MAX_PRIMITIVES) close();
      const uint32_t tri[3] = {a, b, c};
      for (int k = 0; k < 3; k++) {
        if (slot[tri[k]] == NONE) {
          slot[tri[k]] = cur.vertex_count++;
          vertex_indices.push_back(tri[k]);
        }
        micro.push_back((uint8_t)slot[tri[k]]);
      }
      cur.triangle_count++;
    }
    close();
    if (raw.empty()) break; // :680-682

    std::vector<OxcMeshlet> meshlets(raw.size());
    std::vector<OxcMeshletBounds> bounds(raw.size());
    for (size_t i = 0; i < raw.size(); i++) {
      const ScanMeshlet& r = raw[i];
      float bmin[3] = {FMAX, FMAX, FMAX}, bmax[3] = {FLOW, FLOW, FLOW};
      for (uint32_t k = 0; k < r.triangle_count * 3u; k++) { // :697-711
        const float* p = &positions[(size_t)vertex_indices[r.vertex_offset + micro[r.triangle_offset + k]] * 3];
        for (int a = 0; a < 3; a++) { bmin[a] = p[a] < bmin[a] ? p[a] : bmin[a]; bmax[a] = p[a] > bmax[a] ? p[a] : bmax[a]; }
      }
      // normal cone (meshopt_computeMeshletBounds)
      float normals[OXC_MESHLET_MAX_PRIMITIVES][3];
      size_t n_normals = 0;
      for (uint32_t t = 0; t < r.triangle_count; t++) {
        const float* p0 = &positions[(size_t)vertex_indices[r.vertex_offset + micro[r.triangle_offset + t * 3 + 0]] * 3];
        const float* p1 = &positions[(size_t)vertex_indices[r.vertex_offset + micro[r.triangle_offset + t * 3 + 1]] * 3];
        const float* p2 = &positions[(size_t)vertex_indices[r.vertex_offset + micro[r.triangle_offset + t * 3 + 2]] * 3];
        const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const float nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
        const float area = std::sqrt(nx * nx + ny * ny + nz * nz);
        if (area == 0.0f) continue; // degenerate triangles carry no normal
        normals[n_normals][0] = nx / area; normals[n_normals][1] = ny / area; normals[n_normals][2] = nz / area;
        n_normals++;
      }
      int axis_s8[3] = {0, 0, 0}, cutoff_s8 = 127;
      if (n_normals) {
        float ns[4];
        bounding_sphere(ns, normals, n_normals);
        float axis[3] = {ns[0], ns[1], ns[2]};
        const float len = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
        const float inv = len == 0.0f ? 0.0f : 1.0f / len;
        axis[0] *= inv; axis[1] *= inv; axis[2] *= inv;
        float mindp = 1.0f;
        for (size_t k = 0; k < n_normals; k++) {
          const float dp = normals[k][0] * axis[0] + normals[k][1] * axis[1] + normals[k][2] * axis[2];
          mindp = dp < mindp ? dp : mindp;
        }
        if (mindp > 0.1f) {
          const float cutoff = std::sqrt(1.0f - mindp * mindp);
          float err = 0.0f;
          for (int a = 0; a < 3; a++) {
            axis_s8[a] = quantize_snorm(axis[a], 8);
            err += std::fabs((float)axis_s8[a] / 127.0f - axis[a]);
          }
          cutoff_s8 = (int)(127.0f * (cutoff + err) + 1.0f); // round up: the 8-bit test must stay conservative
          cutoff_s8 = cutoff_s8 > 127 ? 127 : cutoff_s8;
        }
      }
      OxcMeshlet& ml = meshlets[i];
      ml.indirect_vertex_index_offset = r.vertex_offset; // :725-728
      ml.local_triangle_index_offset = r.triangle_offset;
      ml.vertex_count = r.vertex_count;
      ml.triangle_count = r.triangle_count;
      OxcMeshletBounds& b = bounds[i];
      for (int a = 0; a < 3; a++) { // :722-736
        b.aabb_center[a] = quantize_half((bmax[a] + bmin[a]) * 0.5f);
        b.aabb_extent[a] = quantize_half(bmax[a] - bmin[a]);
      }
      b.cone_axis_xy[0] = (int8_t)axis_s8[0]; b.cone_axis_xy[1] = (int8_t)axis_s8[1];
      b.cone_axis_z = (int8_t)axis_s8[2];
      b.cone_cutoff = (int8_t)cutoff_s8;
      if (l == 0)
        for (int a = 0; a < 3; a++) { mesh_min[a] = bmin[a] < mesh_min[a] ? bmin[a] : mesh_min[a]; mesh_max[a] = bmax[a] > mesh_max[a] ? bmax[a] : mesh_max[a]; }
    }
    OxcMeshLOD& d = lods[l];
    // :748-756, with 16-byte alignment where the kernels issue 128-bit loads (oxc_set_scene checks it)
    d.indices = blob.append(indices.data(), indices.size() * 4, 16);
    d.meshlets = blob.append(meshlets.data(), meshlets.size() * sizeof(OxcMeshlet), 16);
    d.meshlet_bounds = blob.append(bounds.data(), bounds.size() * sizeof(OxcMeshletBounds), 16);
    d.local_triangle_indices = blob.append(micro.data(), micro.size(), 16);
    d.indirect_vertex_indices = blob.append(vertex_indices.data(), vertex_indices.size() * 4, 16);
    d.indices_count = (uint32_t)indices.size();
    d.meshlet_count = (uint32_t)meshlets.size();
    d.meshlet_bounds_count = (uint32_t)bounds.size();
    d.local_triangle_indices_count = (uint32_t)micro.size();
    d.indirect_vertex_indices_count = (uint32_t)vertex_indices.size();
    d.error = lod_errors[l];
    m->meshlet_counts[l] = d.meshlet_count;
    m->mesh.lod_count++;
  }
  if (m->mesh.lod_count == 0) { // :758-760
    g_builder_error = "no LOD produced a meshlet";
    delete m;
    return OXC_E_INVALID;
  }
  for (int a = 0; a < 3; a++) { // :743-746
    m->mesh.bounds.aabb_center[a] = (mesh_max[a] + mesh_min[a]) * 0.5f;
    m->mesh.bounds.aabb_extent[a] = mesh_max[a] - mesh_min[a];
  }
  m->mesh.lods = blob.append(lods, (size_t)m->mesh.lod_count * sizeof(OxcMeshLOD), 16); // :762-763 lod_metadata_offset
  blob.bytes.resize((blob.bytes.size() + 15) / 16 * 16, 0);
  m->blob.swap(blob.bytes);
  *out = m;
  return OXC_OK;
}

uint64_t oxb_mesh_blob_size(const OxbMesh* m) { return m ? (uint64_t)m->blob.size() : 0; }
uint32_t oxb_mesh_lod0_meshlet_count(const OxbMesh* m) { return m ? m->meshlet_counts[0] : 0; }

int oxb_mesh_emit(const OxbMesh* m, uint64_t base_offset, uint8_t* dst, OxcMesh* mesh_out) {
  if (!m || !dst || !mesh_out) { g_builder_error = "null argument"; return OXC_E_INVALID; }
  if (base_offset & 15u) { g_builder_error = "base_offset must be 16-byte aligned"; return OXC_E_INVALID; }
  std::memcpy(dst, m->blob.data(), m->blob.size());
  OxcMesh me = m->mesh;
  me.vertex_positions += base_offset;
  me.vertex_normals = m->mesh.vertex_normals ? m->mesh.vertex_normals + base_offset : 0;
  me.texture_coords = m->mesh.texture_coords ? m->mesh.texture_coords + base_offset : 0;
  OxcMeshLOD* lods = reinterpret_cast<OxcMeshLOD*>(dst + m->mesh.lods);
  for (uint32_t l = 0; l < me.lod_count; l++) {
    OxcMeshLOD d;
    std::memcpy(&d, &lods[l], sizeof d);
    d.indices += base_offset; d.meshlets += base_offset; d.meshlet_bounds += base_offset;
    d.local_triangle_indices += base_offset; d.indirect_vertex_indices += base_offset;
    std::memcpy(&lods[l], &d, sizeof d);
  }
  me.lods += base_offset;
  *mesh_out = me;
  return OXC_OK;
}

void oxb_mesh_free(OxbMesh* m) { delete m; }

int64_t oxb_simplify(uint32_t* dst, const uint32_t* indices, uint64_t index_count, const float* positions, const float* normals, uint32_t vertex_count,
                     uint64_t target_index_count, float target_error, float* result_error) {
  if (!dst || !indices || !positions || index_count % 3u || target_index_count % 3u) {
    g_builder_error = "oxb_simplify: null argument or index counts that are not multiples of 3";
    return OXC_E_INVALID;
  }
  for (uint64_t i = 0; i < index_count; i++)
    if (indices[i] >= vertex_count) { g_builder_error = "index out of range"; return OXC_E_INVALID; }
  const std::vector<uint32_t> r = oxb::simplify(indices, (size_t)index_count, positions, normals, vertex_count, (size_t)target_index_count, target_error, result_error);
  if (!r.empty()) std::memcpy(dst, r.data(), r.size() * 4);
  return (int64_t)r.size();
}

} // extern "C"
