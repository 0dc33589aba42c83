// mesh_simplifier.cpp — LOD generation for the mesh builder: the step build_gltf_mesh delegates to
// meshopt_simplifyWithAttributes (Oxylus/src/Asset/AssetManager_GLTF.cpp:604-637 — normals as attributes with weight 1,
// meshopt_SimplifyLockBorder, target error FLT_MAX, target index count = half of the previous LOD, relative result error).
//
// meshoptimizer v1.2 (xmake/packages.lua:9) is neither part of /root/reference nor installed here.  This is an edge-collapse
// simplifier of the library's published scheme, written from that description:
//   * collapses move a vertex onto an EXISTING neighbour (the vertex buffer is shared by all LODs, :598-641);
//   * per position: sum of the incident triangles' plane quadrics, weight sqrt(area); seam edges add a plane through the edge,
//     perpendicular to the triangle, weight = edge length;
//   * per wedge (vertices that share a position but not their attributes): attribute quadrics built from the per-triangle
//     gradient of each attribute's linear interpolant;
//   * vertices are classified once — manifold (collapses anywhere), seam (exactly two wedges; collapses along its seam, both
//     wedges together), locked (mesh border = meshopt_SimplifyLockBorder, seam ends, non-manifold) — and a pass ranks every
//     allowed collapse by position + attribute error, applies them in order, locks both endpoints for the rest of the pass
//     and stops at half the remaining distance to the target, so later passes see refreshed quadrics;
//   * the reported error is the largest POSITION error of a performed collapse, relative to the mesh extent (the quantity
//     MeshLOD::error feeds into cull_meshes.slang:35-57).
// Three additions make the result checkable: the link condition (a 2-manifold stays a 2-manifold), a flip test against each
// triangle's ORIGINAL normal as well as its current one (rotations cannot accumulate into a fold over several passes), and
// exact, stable ordering (no hashing, ties by candidate order).  NOT bit-compatible with meshoptimizer's output: parity unpinned.
//
// Arithmetic: IEEE binary64, one rounding per operation in the order written (the library is built with -ffp-contract=off);
// oracle/pysimplify.py restates this file operation for operation and tests/test_builder_cpu.py compares them bit for bit.
#include "mesh_simplifier.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>

namespace oxb {
namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint8_t MANIFOLD = 0, SEAM = 1, LOCKED = 2;
constexpr double SEAM_EDGE_WEIGHT = 1.0;

struct V3 { double x, y, z; };
inline V3 sub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// a00 a11 a22 a10 a20 a21 b0 b1 b2 c w
struct Quadric { double v[11]; };
inline Quadric q_zero() { Quadric q; for (double& x : q.v) x = 0.0; return q; }
inline Quadric q_plane(double a, double b, double c, double d, double w) {
  const double aw = a * w, bw = b * w, cw = c * w, dw = d * w;
  return {{a * aw, b * bw, c * cw, a * bw, a * cw, b * cw, a * dw, b * dw, c * dw, d * dw, w}};
}
inline void q_add(Quadric& q, const Quadric& r) { for (int i = 0; i < 11; i++) q.v[i] = q.v[i] + r.v[i]; }
inline double q_eval(const Quadric& q, const V3& p) {
  const double rx = q.v[0] * p.x + q.v[3] * p.y + q.v[4] * p.z;
  const double ry = q.v[3] * p.x + q.v[1] * p.y + q.v[5] * p.z;
  const double rz = q.v[4] * p.x + q.v[5] * p.y + q.v[2] * p.z;
  double r = rx * p.x + ry * p.y + rz * p.z;
  r = r + 2.0 * (q.v[6] * p.x + q.v[7] * p.y + q.v[8] * p.z);
  return r + q.v[9];
}
inline bool q_triangle(Quadric& out, const V3& p0, const V3& p1, const V3& p2) {
  V3 n = cross(sub(p1, p0), sub(p2, p0));
  const double ln = std::sqrt(dot(n, n));
  if (ln == 0.0) return false;
  n = {n.x / ln, n.y / ln, n.z / ln};
  out = q_plane(n.x, n.y, n.z, -dot(n, p0), std::sqrt(ln));
  return true;
}
inline bool q_seam_edge(Quadric& out, const V3& p0, const V3& p1, const V3& p2) {
  const V3 e = sub(p1, p0);
  const double ee = dot(e, e);
  if (ee == 0.0) return false;
  const V3 f = sub(p2, p0);
  const double t = dot(f, e) / ee;
  V3 n = {f.x - e.x * t, f.y - e.y * t, f.z - e.z * t};
  const double ln = std::sqrt(dot(n, n));
  if (ln == 0.0) return false;
  n = {n.x / ln, n.y / ln, n.z / ln};
  out = q_plane(n.x, n.y, n.z, -dot(n, p0), std::sqrt(ee) * SEAM_EDGE_WEIGHT);
  return true;
}
using Grad = std::array<double, 12>; // 3 attributes x (w*gx, w*gy, w*gz, w*d)
inline void q_attributes(Quadric& q, Grad& grads, const V3& p0, const V3& p1, const V3& p2, const float* a0, const float* a1, const float* a2) {
  const V3 e1 = sub(p1, p0), e2 = sub(p2, p0);
  const V3 n = cross(e1, e2);
  const double ln = std::sqrt(dot(n, n));
  const double w = std::sqrt(ln);
  const double d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2);
  const double den = d00 * d11 - d01 * d01;
  const double inv = den == 0.0 ? 0.0 : 1.0 / den;
  q = q_zero();
  for (int k = 0; k < 3; k++) {
    const double da1 = (double)a1[k] - (double)a0[k], da2 = (double)a2[k] - (double)a0[k];
    const double c1 = (da1 * d11 - da2 * d01) * inv;
    const double c2 = (da2 * d00 - da1 * d01) * inv;
    const V3 g = {c1 * e1.x + c2 * e2.x, c1 * e1.y + c2 * e2.y, c1 * e1.z + c2 * e2.z};
    const double d = (double)a0[k] - dot(g, p0);
    q_add(q, q_plane(g.x, g.y, g.z, d, w));
    grads[k * 4 + 0] = g.x * w; grads[k * 4 + 1] = g.y * w; grads[k * 4 + 2] = g.z * w; grads[k * 4 + 3] = d * w;
  }
  q.v[10] = w; // one weight per triangle, not one per attribute
}

inline uint64_t edge_key(uint32_t a, uint32_t b) { return ((uint64_t)a << 32) | b; }
struct EdgeSet {
  std::vector<uint64_t> keys;
  void finish() { std::sort(keys.begin(), keys.end()); }
  size_t count(uint32_t a, uint32_t b) const {
    const auto r = std::equal_range(keys.begin(), keys.end(), edge_key(a, b));
    return (size_t)(r.second - r.first);
  }
  bool has(uint32_t a, uint32_t b) const { return std::binary_search(keys.begin(), keys.end(), edge_key(a, b)); }
};

struct Candidate { double err; uint32_t v0, v1; double perr; };
struct Tri { uint32_t a, b, c, t; };

} // namespace

std::vector<uint32_t> simplify(const uint32_t* indices, size_t index_count, const float* positions, const float* normals, uint32_t vertex_count,
                               size_t target_index_count, float target_error, float* out_error) {
  std::vector<uint32_t> result(indices, indices + index_count);
  if (out_error) *out_error = 0.0f;
  if (index_count <= target_index_count || vertex_count == 0) return result;
  const uint32_t V = vertex_count;

  // ---- positions rescaled into the unit cube (extent of ALL vertices: one scale for the whole LOD chain) ----
  float lo32[3], hi32[3];
  for (int a = 0; a < 3; a++) lo32[a] = hi32[a] = positions[a];
  for (uint32_t v = 1; v < V; v++)
    for (int a = 0; a < 3; a++) {
      const float x = positions[(size_t)v * 3 + a];
      lo32[a] = x < lo32[a] ? x : lo32[a];
      hi32[a] = x > hi32[a] ? x : hi32[a];
    }
  const double lo[3] = {lo32[0], lo32[1], lo32[2]};
  double extent = (double)hi32[0] - lo[0];
  extent = std::max(extent, (double)hi32[1] - lo[1]);
  extent = std::max(extent, (double)hi32[2] - lo[2]);
  const double inv_extent = extent == 0.0 ? 0.0 : 1.0 / extent;
  std::vector<V3> P(V);
  for (uint32_t v = 0; v < V; v++)
    P[v] = {((double)positions[(size_t)v * 3] - lo[0]) * inv_extent, ((double)positions[(size_t)v * 3 + 1] - lo[1]) * inv_extent,
            ((double)positions[(size_t)v * 3 + 2] - lo[2]) * inv_extent};

  // ---- wedges: used vertices with bit-identical positions (canonical = lowest index, ring in index order) ----
  std::vector<uint8_t> used(V, 0);
  for (uint32_t i : result) used[i] = 1;
  std::vector<uint32_t> remap(V), wedge(V);
  for (uint32_t v = 0; v < V; v++) remap[v] = wedge[v] = v;
  {
    std::vector<uint32_t> ids;
    for (uint32_t v = 0; v < V; v++) if (used[v]) ids.push_back(v);
    auto bits = [&](uint32_t v, int a) { uint32_t u; std::memcpy(&u, &positions[(size_t)v * 3 + a], 4); return u; };
    std::sort(ids.begin(), ids.end(), [&](uint32_t x, uint32_t y) {
      for (int a = 0; a < 3; a++) { const uint32_t bx = bits(x, a), by = bits(y, a); if (bx != by) return bx < by; }
      return x < y;
    });
    for (size_t i = 0; i < ids.size();) {
      size_t j = i + 1;
      while (j < ids.size() && bits(ids[j], 0) == bits(ids[i], 0) && bits(ids[j], 1) == bits(ids[i], 1) && bits(ids[j], 2) == bits(ids[i], 2)) j++;
      for (size_t k = i; k < j; k++) { remap[ids[k]] = ids[i]; wedge[ids[k]] = ids[k + 1 < j ? k + 1 : i]; }
      i = j;
    }
  }

  // ---- classification ----
  size_t T = result.size() / 3;
  EdgeSet wedge_edges, pos_edges;
  for (size_t t = 0; t < T; t++)
    for (int e = 0; e < 3; e++) {
      const uint32_t a = result[3 * t + e], b = result[3 * t + (e + 1) % 3];
      wedge_edges.keys.push_back(edge_key(a, b));
      pos_edges.keys.push_back(edge_key(remap[a], remap[b]));
    }
  wedge_edges.finish();
  pos_edges.finish();
  std::vector<uint32_t> loop(V, NONE), loopback(V, NONE), outc(V, 0), inc(V, 0);
  std::vector<uint8_t> complex_(V, 0), pborder(V, 0);
  std::vector<Tri> seam_edges;
  for (size_t t = 0; t < T; t++)
    for (int e = 0; e < 3; e++) {
      const uint32_t a = result[3 * t + e], b = result[3 * t + (e + 1) % 3], c = result[3 * t + (e + 2) % 3];
      const uint32_t ra = remap[a], rb = remap[b];
      if (ra == rb || pos_edges.count(ra, rb) > 1) complex_[ra] = complex_[rb] = 1;
      const bool open_w = !wedge_edges.has(b, a);
      if (!pos_edges.has(rb, ra)) pborder[ra] = pborder[rb] = 1;
      else if (open_w) seam_edges.push_back({a, b, c, (uint32_t)t});
      if (open_w) {
        if (outc[a] == 0) loop[a] = b;
        if (inc[b] == 0) loopback[b] = a;
        outc[a]++;
        inc[b]++;
      }
    }
  std::vector<uint8_t> kind(V, LOCKED);
  for (uint32_t r = 0; r < V; r++) {
    if (!used[r] || remap[r] != r || complex_[r] || pborder[r]) continue;
    const uint32_t w = wedge[r];
    if (w == r) {
      if (outc[r] == 0 && inc[r] == 0) kind[r] = MANIFOLD;
    } else if (wedge[w] == r) {
      if (outc[r] == 1 && inc[r] == 1 && outc[w] == 1 && inc[w] == 1 && remap[loop[r]] == remap[loopback[w]] && remap[loopback[r]] == remap[loop[w]])
        kind[r] = SEAM;
    }
  }

  // ---- quadrics ----
  std::vector<Quadric> Q(V, q_zero());
  for (size_t t = 0; t < T; t++) {
    const uint32_t i0 = result[3 * t], i1 = result[3 * t + 1], i2 = result[3 * t + 2];
    Quadric q;
    if (q_triangle(q, P[i0], P[i1], P[i2])) { q_add(Q[remap[i0]], q); q_add(Q[remap[i1]], q); q_add(Q[remap[i2]], q); }
  }
  for (const Tri& s : seam_edges) {
    Quadric q;
    if (q_seam_edge(q, P[s.a], P[s.b], P[s.c])) { q_add(Q[remap[s.a]], q); q_add(Q[remap[s.b]], q); }
  }
  const bool attrs = normals != nullptr;
  std::vector<Quadric> QA;
  std::vector<Grad> G;
  if (attrs) {
    QA.assign(V, q_zero());
    Grad zero;
    zero.fill(0.0);
    G.assign(V, zero);
    for (size_t t = 0; t < T; t++) {
      const uint32_t i[3] = {result[3 * t], result[3 * t + 1], result[3 * t + 2]};
      Quadric q;
      Grad g;
      q_attributes(q, g, P[i[0]], P[i[1]], P[i[2]], &normals[(size_t)i[0] * 3], &normals[(size_t)i[1] * 3], &normals[(size_t)i[2] * 3]);
      for (int k = 0; k < 3; k++) {
        q_add(QA[i[k]], q);
        for (int j = 0; j < 12; j++) G[i[k]][j] = G[i[k]][j] + g[j];
      }
    }
  }

  auto pos_error = [&](uint32_t v0, uint32_t v1) {
    const Quadric& q = Q[remap[v0]];
    return std::fabs(q_eval(q, P[v1])) * (q.v[10] == 0.0 ? 0.0 : 1.0 / q.v[10]);
  };
  auto attr_error = [&](uint32_t v0, uint32_t v1) {
    const Quadric& q = QA[v0];
    const V3& p = P[v1];
    const float* a = &normals[(size_t)v1 * 3];
    double r = q_eval(q, p);
    for (int k = 0; k < 3; k++) {
      const double g = G[v0][k * 4] * p.x + G[v0][k * 4 + 1] * p.y + G[v0][k * 4 + 2] * p.z + G[v0][k * 4 + 3];
      const double ak = (double)a[k];
      r = r + ak * (ak * q.v[10] - 2.0 * g);
    }
    return std::fabs(r) * (q.v[10] == 0.0 ? 0.0 : 1.0 / q.v[10]);
  };
  // the other wedge of a seam collapse v0 -> v1: (s0, s1); false when the seam does not continue consistently
  auto seam_partner = [&](uint32_t v0, uint32_t v1, uint32_t& s0, uint32_t& s1) {
    s0 = wedge[v0];
    s1 = loop[v0] == v1 ? loopback[s0] : loop[s0];
    return s1 != NONE && remap[s1] == remap[v1];
  };
  auto collapse_error = [&](uint32_t v0, uint32_t v1) {
    double e = pos_error(v0, v1);
    if (attrs) {
      e = e + attr_error(v0, v1);
      if (kind[remap[v0]] == SEAM) {
        uint32_t s0, s1;
        seam_partner(v0, v1, s0, s1);
        e = e + attr_error(s0, s1);
      }
    }
    return e != e ? std::numeric_limits<double>::infinity() : e;
  };
  auto allowed = [&](uint32_t v0, uint32_t v1) {
    const uint8_t k0 = kind[remap[v0]];
    if (k0 == MANIFOLD) return true;
    if (k0 == SEAM && kind[remap[v1]] == SEAM && (loop[v0] == v1 || loopback[v0] == v1)) {
      uint32_t s0, s1;
      return seam_partner(v0, v1, s0, s1);
    }
    return false;
  };

  // each triangle remembers its ORIGINAL normal: a corner may move many times, the face may never turn away from where it started
  std::vector<V3> orig_n(T);
  for (size_t t = 0; t < T; t++) orig_n[t] = cross(sub(P[result[3 * t + 1]], P[result[3 * t]]), sub(P[result[3 * t + 2]], P[result[3 * t]]));
  double result_error = 0.0;
  const double error_limit = (double)target_error * (double)target_error;
  std::vector<uint32_t> tri_off(V + 1), tri_list, fill, collapse_remap(V), pos_collapse(V), order, mark_a(V, 0), mark_b(V, 0);
  std::vector<uint8_t> locked(V);
  std::vector<Candidate> cands;
  std::vector<Tri> ring0, ring1;
  uint32_t stamp = 0;
  while (result.size() > target_index_count) {
    T = result.size() / 3;
    // position -> triangles (in triangle order; a corner-degenerate triangle appears once per corner, the ring drops it)
    std::fill(tri_off.begin(), tri_off.end(), 0u);
    for (uint32_t i : result) tri_off[remap[i] + 1]++;
    for (uint32_t v = 0; v < V; v++) tri_off[v + 1] += tri_off[v];
    tri_list.resize(result.size());
    fill.assign(tri_off.begin(), tri_off.end() - 1);
    for (size_t t = 0; t < T; t++)
      for (int e = 0; e < 3; e++) tri_list[fill[remap[result[3 * t + e]]]++] = (uint32_t)t;

    cands.clear();
    for (size_t t = 0; t < T; t++)
      for (int e = 0; e < 3; e++) {
        const uint32_t i0 = result[3 * t + e], i1 = result[3 * t + (e + 1) % 3];
        const uint32_t r0 = remap[i0], r1 = remap[i1];
        if (r0 == r1 || r1 > r0) continue; // the opposite half-edge generates this pair
        const bool a01 = allowed(i0, i1), a10 = allowed(i1, i0);
        if (!a01 && !a10) continue;
        if (a01 && a10) {
          const double e01 = collapse_error(i0, i1), e10 = collapse_error(i1, i0);
          if (e10 < e01) cands.push_back({e10, i1, i0, pos_error(i1, i0)});
          else cands.push_back({e01, i0, i1, pos_error(i0, i1)});
        } else if (a01) {
          cands.push_back({collapse_error(i0, i1), i0, i1, pos_error(i0, i1)});
        } else {
          cands.push_back({collapse_error(i1, i0), i1, i0, pos_error(i1, i0)});
        }
      }
    if (cands.empty()) break;
    order.resize(cands.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cands[x].err < cands[y].err; });
    const size_t goal = (result.size() - target_index_count) / 3;
    const size_t edge_goal = goal / 2;
    const double error_goal = edge_goal < cands.size() ? 1.5 * cands[order[edge_goal]].err : std::numeric_limits<double>::infinity();
    for (uint32_t v = 0; v < V; v++) collapse_remap[v] = pos_collapse[v] = v;
    std::fill(locked.begin(), locked.end(), (uint8_t)0);
    size_t collapsed_tris = 0, collapses = 0;

    // current triangles of position r as position triples, r first, degenerate ones dropped
    auto ring = [&](uint32_t r, std::vector<Tri>& out) {
      out.clear();
      for (uint32_t k = tri_off[r]; k < tri_off[r + 1]; k++) {
        const size_t t = tri_list[k];
        uint32_t a = pos_collapse[remap[result[3 * t]]], b = pos_collapse[remap[result[3 * t + 1]]], c = pos_collapse[remap[result[3 * t + 2]]];
        if (a == b || b == c || c == a) continue;
        if (b == r) { const uint32_t x = a; a = b; b = c; c = x; }
        else if (c == r) { const uint32_t x = a; a = c; c = b; b = x; }
        out.push_back({a, b, c, (uint32_t)t});
      }
    };

    for (uint32_t ci : order) {
      const Candidate& cd = cands[ci];
      if (cd.err > error_limit || collapsed_tris >= goal) break;
      if (cd.err > error_goal && collapsed_tris > goal / 6) break;
      const uint32_t v0 = cd.v0, v1 = cd.v1, r0 = remap[v0], r1 = remap[v1];
      if (locked[r0] || locked[r1]) continue;
      ring(r0, ring0);
      uint32_t shared[2] = {NONE, NONE};
      size_t n_shared = 0;
      for (const Tri& tr : ring0)
        if (tr.b == r1 || tr.c == r1) {
          if (n_shared < 2) shared[n_shared] = tr.b == r1 ? tr.c : tr.b;
          n_shared++;
        }
      if (n_shared != 2 || shared[0] == shared[1]) continue;
      // link condition: the only common neighbours are the two vertices opposite the edge
      ring(r1, ring1);
      stamp++;
      for (const Tri& tr : ring0) { mark_a[tr.b] = stamp; mark_a[tr.c] = stamp; }
      bool link_ok = true;
      size_t common = 0;
      for (const Tri& tr : ring1)
        for (uint32_t x : {tr.b, tr.c}) {
          if (mark_b[x] == stamp) continue;
          mark_b[x] = stamp;
          if (mark_a[x] != stamp) continue;
          common++;
          if (x != shared[0] && x != shared[1]) link_ok = false;
        }
      if (!link_ok || common != 2) continue;
      bool flip = false;
      for (const Tri& tr : ring0) {
        if (tr.b == r1 || tr.c == r1) continue;
        const V3 n_old = cross(sub(P[tr.b], P[r0]), sub(P[tr.c], P[r0]));
        const V3 n_new = cross(sub(P[tr.b], P[r1]), sub(P[tr.c], P[r1]));
        const double nn = dot(n_new, n_new);
        const V3& n_ref = orig_n[tr.t];
        if (dot(n_old, n_new) <= 0.25 * std::sqrt(dot(n_old, n_old) * nn) || dot(n_ref, n_new) <= 0.25 * std::sqrt(dot(n_ref, n_ref) * nn)) { flip = true; break; }
      }
      if (flip) continue;
      uint32_t pa[2] = {v0, NONE}, pb[2] = {v1, NONE};
      int n_pairs = 1;
      if (kind[r0] == SEAM) { seam_partner(v0, v1, pa[1], pb[1]); n_pairs = 2; }
      for (int k = 0; k < n_pairs; k++) {
        collapse_remap[pa[k]] = pb[k];
        if (attrs) {
          q_add(QA[pb[k]], QA[pa[k]]);
          for (int j = 0; j < 12; j++) G[pb[k]][j] = G[pb[k]][j] + G[pa[k]][j];
        }
      }
      q_add(Q[r1], Q[r0]);
      pos_collapse[r0] = r1;
      locked[r0] = locked[r1] = 1;
      collapsed_tris += 2;
      collapses++;
      if (cd.perr > result_error) result_error = cd.perr;
    }
    if (collapses == 0) break;
    size_t w = 0;
    for (size_t t = 0; t < T; t++) {
      const uint32_t a = collapse_remap[result[3 * t]], b = collapse_remap[result[3 * t + 1]], c = collapse_remap[result[3 * t + 2]];
      if (remap[a] != remap[b] && remap[b] != remap[c] && remap[c] != remap[a]) {
        result[w] = a; result[w + 1] = b; result[w + 2] = c;
        orig_n[w / 3] = orig_n[t];
        w += 3;
      }
    }
    result.resize(w);
    // the open-edge loops follow the collapses
    for (std::vector<uint32_t>* tbl : {&loop, &loopback}) {
      std::vector<uint32_t> upd(*tbl);
      for (uint32_t i = 0; i < V; i++) {
        const uint32_t l = (*tbl)[i];
        if (l == NONE || collapse_remap[i] != i) continue;
        uint32_t r = collapse_remap[l];
        if (r == i) {
          const uint32_t l2 = (*tbl)[l];
          r = l2 == NONE ? NONE : collapse_remap[l2];
          r = r == i ? NONE : r;
        }
        upd[i] = r;
      }
      tbl->swap(upd);
    }
  }
  if (out_error) *out_error = (float)std::sqrt(result_error);
  return result;
}

} // namespace oxb
