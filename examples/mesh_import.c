/* examples/mesh_import.c — the asset-import side of the path in plain C11 (no GPU needed): a procedural sphere goes through
 * oxb_build_mesh with the spatial clusteriser and the generated LOD chain (the role of build_gltf_mesh,
 * Oxylus/src/Asset/AssetManager_GLTF.cpp:481-771), then its blob is emitted the way OxcSceneDesc expects it and the LOD table
 * is read back from the blob.
 *
 *   gcc -std=c11 -I include examples/mesh_import.c -L oxylus_b200 -loxcull -Wl,-rpath,$PWD/oxylus_b200 -lm -o mesh_import
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oxcull.h"

#define RINGS 48
#define SEGMENTS 96

int main(void) {
  /* UV sphere: (RINGS - 1) interior rings x SEGMENTS vertices + 2 poles */
  const uint32_t vertex_count = (RINGS - 1) * SEGMENTS + 2;
  const uint32_t triangle_count = 2 * SEGMENTS + (RINGS - 2) * SEGMENTS * 2;
  float* positions = malloc(sizeof(float) * 3 * vertex_count);
  float* normals = malloc(sizeof(float) * 3 * vertex_count);
  uint32_t* indices = malloc(sizeof(uint32_t) * 3 * triangle_count);
  if (!positions || !normals || !indices) return 1;
  const double pi = 3.14159265358979323846;
  uint32_t v = 0;
  for (int r = 1; r < RINGS; r++)
    for (int s = 0; s < SEGMENTS; s++, v++) {
      const double th = pi * r / RINGS, ph = 2 * pi * s / SEGMENTS;
      normals[3 * v + 0] = (float)(sin(th) * cos(ph));
      normals[3 * v + 1] = (float)cos(th);
      normals[3 * v + 2] = (float)(sin(th) * sin(ph));
      for (int a = 0; a < 3; a++) positions[3 * v + a] = 1.5f * normals[3 * v + a];
    }
  const uint32_t north = v, south = v + 1;
  for (int a = 0; a < 3; a++) { normals[3 * north + a] = a == 1 ? 1.0f : 0.0f; normals[3 * south + a] = a == 1 ? -1.0f : 0.0f; }
  for (int a = 0; a < 3; a++) { positions[3 * north + a] = 1.5f * normals[3 * north + a]; positions[3 * south + a] = 1.5f * normals[3 * south + a]; }
  uint32_t n = 0;
  for (int s = 0; s < SEGMENTS; s++) { /* caps */
    const uint32_t s1 = (uint32_t)((s + 1) % SEGMENTS);
    indices[n++] = north; indices[n++] = s1; indices[n++] = (uint32_t)s;
    const uint32_t base = (RINGS - 2) * SEGMENTS;
    indices[n++] = south; indices[n++] = base + (uint32_t)s; indices[n++] = base + s1;
  }
  for (int r = 0; r < RINGS - 2; r++)
    for (int s = 0; s < SEGMENTS; s++) {
      const uint32_t s1 = (uint32_t)((s + 1) % SEGMENTS);
      const uint32_t a = (uint32_t)r * SEGMENTS + (uint32_t)s, b = (uint32_t)r * SEGMENTS + s1;
      const uint32_t c = (uint32_t)(r + 1) * SEGMENTS + (uint32_t)s, d = (uint32_t)(r + 1) * SEGMENTS + s1;
      indices[n++] = a; indices[n++] = b; indices[n++] = d;
      indices[n++] = a; indices[n++] = d; indices[n++] = c;
    }
  if (n != 3 * triangle_count) return 1;

  OxbMeshInput in;
  memset(&in, 0, sizeof in);
  in.positions = positions;
  in.normals = normals;
  in.vertex_count = vertex_count;
  in.lod_count = 1;              /* LOD 0 only ... */
  in.lod_indices[0] = indices;
  in.lod_index_counts[0] = 3 * triangle_count;
  in.lod_errors[0] = 0.0f;
  in.cluster_mode = 1;           /* spatial clusteriser (meshopt_buildMeshlets' role) */
  in.auto_lods = 1;              /* ... the chain is generated (meshopt_simplifyWithAttributes' role, :596-641) */
  OxbMesh* mesh = NULL;
  if (oxb_build_mesh(&in, &mesh) != OXC_OK) {
    fprintf(stderr, "oxb_build_mesh: %s\n", oxb_last_error());
    return 2;
  }
  /* one mesh at offset 0 of the scene's geometry blob */
  const uint64_t blob_size = oxb_mesh_blob_size(mesh);
  uint8_t* blob = malloc((size_t)blob_size);
  OxcMesh record;
  if (!blob || oxb_mesh_emit(mesh, 0, blob, &record) != OXC_OK) return 2;
  printf("vertices %u  blob %llu bytes  lods %u  lod0 meshlets %u\n", record.vertex_count, (unsigned long long)blob_size, record.lod_count,
         oxb_mesh_lod0_meshlet_count(mesh));
  int ok = record.lod_count >= 4 && record.vertex_count == vertex_count;
  uint32_t prev = 0;
  float prev_error = -1.0f;
  for (uint32_t l = 0; l < record.lod_count; l++) {
    OxcMeshLOD lod;
    memcpy(&lod, blob + record.lods + (uint64_t)l * sizeof lod, sizeof lod);
    printf("  lod %u: %7u indices  %5u meshlets  error %.5f\n", l, lod.indices_count, lod.meshlet_count, lod.error);
    if (l && lod.indices_count > prev * 3 / 4) ok = 0;   /* about half of the previous LOD */
    if (lod.error <= prev_error) ok = 0;                 /* accumulated error grows */
    prev = lod.indices_count;
    prev_error = lod.error;
  }
  oxb_mesh_free(mesh);
  free(blob); free(indices); free(normals); free(positions);
  if (!ok) { fprintf(stderr, "unexpected LOD chain\n"); return 4; }
  printf("ok\n");
  return 0;
}
