// mesh_simplifier.hpp — internal interface of the builder's LOD simplifier (mesh_simplifier.cpp); the public entry points are
// oxb_simplify and OxbMeshInput::auto_lods (include/oxcull.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace oxb {

// Edge-collapse simplification of a triangle list towards target_index_count indices; vertices are never moved or created.
// normals may be NULL (positions only).  *result_error: largest position error of a performed collapse, relative to the
// extent of the vertex buffer.  The result can stay above the target (locked borders, topology) — the caller decides.
std::vector<uint32_t> simplify(const uint32_t* indices, size_t index_count, const float* positions, const float* normals, uint32_t vertex_count,
                               size_t target_index_count, float target_error, float* result_error);

} // namespace oxb
