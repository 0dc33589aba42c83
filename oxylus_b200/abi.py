"""numpy / ctypes mirrors of include/oxcull.h (which mirrors Oxylus/include/Scene/SceneGPU.hpp).

Sizes are asserted against the reference's scalar-layout sizes: MeshletBounds 16 B
(SceneGPU.hpp:84-90), MeshInstance 20 B (:111-117), Meshlet 16 B (:119-124), MeshLOD 64 B (:126-140),
Mesh 64 B (:142-152), CullCamera 96 B (:222-229).
"""
import ctypes as C

import numpy as np

TRANSFORM_DT = np.dtype([("world", "<f4", (16,))])
MESHLET_BOUNDS_DT = np.dtype(
    [
        ("aabb_center", "<u2", (3,)),
        ("cone_axis_xy", "i1", (2,)),
        ("aabb_extent", "<u2", (3,)),
        ("cone_axis_z", "i1"),
        ("cone_cutoff", "i1"),
    ]
)
MESH_BOUNDS_DT = np.dtype([("aabb_center", "<f4", (3,)), ("aabb_extent", "<f4", (3,))])
VISIBILITY_DT = np.dtype([("total", "<u4"), ("early", "<u4"), ("late", "<u4")])
MESHLET_INSTANCE_DT = np.dtype([("mesh_instance_index", "<u4"), ("meshlet_index", "<u4")])
MESH_INSTANCE_DT = np.dtype(
    [
        ("mesh_index", "<u4"),
        ("lod_index", "<u4"),
        ("material_index", "<u4"),
        ("transform_index", "<u4"),
        ("meshlet_instance_visibility_offset", "<u4"),
    ]
)
MESHLET_DT = np.dtype(
    [
        ("indirect_vertex_index_offset", "<u4"),
        ("local_triangle_index_offset", "<u4"),
        ("vertex_count", "<u4"),
        ("triangle_count", "<u4"),
    ]
)
MESH_LOD_DT = np.dtype(
    [
        ("indices", "<u8"),
        ("meshlets", "<u8"),
        ("meshlet_bounds", "<u8"),
        ("local_triangle_indices", "<u8"),
        ("indirect_vertex_indices", "<u8"),
        ("indices_count", "<u4"),
        ("meshlet_count", "<u4"),
        ("meshlet_bounds_count", "<u4"),
        ("local_triangle_indices_count", "<u4"),
        ("indirect_vertex_indices_count", "<u4"),
        ("error", "<f4"),
    ]
)
MESH_DT = np.dtype(
    [
        ("vertex_positions", "<u8"),
        ("vertex_normals", "<u8"),
        ("texture_coords", "<u8"),
        ("vertex_count", "<u4"),
        ("lod_count", "<u4"),
        ("lods", "<u8"),
        ("bounds", MESH_BOUNDS_DT),
    ]
)
CULL_CAMERA_DT = np.dtype(
    [
        ("projection_view", "<f4", (16,)),
        ("position", "<f4", (3,)),
        ("acceptable_lod_error", "<f4"),
        ("resolution", "<f4", (2,)),
        ("near_clip", "<f4"),
        ("mesh_instance_count", "<u4"),
    ]
)
DISPATCH_CMD_DT = np.dtype([("x", "<u4"), ("y", "<u4"), ("z", "<u4")])
DRAW_CMD_DT = np.dtype(
    [
        ("index_count", "<u4"),
        ("instance_count", "<u4"),
        ("first_index", "<u4"),
        ("vertex_offset", "<i4"),
        ("first_instance", "<u4"),
    ]
)

TERRAIN_DT = np.dtype([("world_min", "<f4", (2,)), ("world_size", "<f4", (2,)), ("patch_count", "<u4", (2,)),
                       ("base_height", "<f4"), ("height_scale", "<f4")])
DRAW_INDIRECT_DT = np.dtype([("vertex_count", "<u4"), ("instance_count", "<u4"), ("first_vertex", "<u4"), ("first_instance", "<u4")])
CLIPMAP_DT = np.dtype([("projection_view_mat", "<f4", (16,)), ("page_offset", "<i4", (2,)), ("z_near", "<f4")])
assert CLIPMAP_DT.itemsize == 76
assert TERRAIN_DT.itemsize == 32 and DRAW_INDIRECT_DT.itemsize == 16
assert TRANSFORM_DT.itemsize == 64
assert MESHLET_BOUNDS_DT.itemsize == 16
assert VISIBILITY_DT.itemsize == 12
assert MESHLET_INSTANCE_DT.itemsize == 8
assert MESH_INSTANCE_DT.itemsize == 20
assert MESHLET_DT.itemsize == 16
assert MESH_LOD_DT.itemsize == 64
assert MESH_DT.itemsize == 64
assert CULL_CAMERA_DT.itemsize == 96
assert DISPATCH_CMD_DT.itemsize == 12
assert DRAW_CMD_DT.itemsize == 20

# CullFlag — SceneGPU.hpp:345-353
CULL_NONE = 0
CULL_TEST_FRUSTUM = 1 << 0
CULL_SELECT_LOD = 1 << 1
CULL_TEST_OCCLUSION = 1 << 2
CULL_LATE_PASS = 1 << 3
CULL_TEST_ALL = CULL_TEST_FRUSTUM | CULL_SELECT_LOD | CULL_TEST_OCCLUSION

HIZ_MAX_LEVELS = 13
MAX_VIEWS = 16
VIS_PRIMITIVE_BITS = 8
VIS_WIDE_PRIMITIVE_BITS = 6
VIS_CLEAR = 0xFFFFFFFF
STATUS_MESHLET_OVERFLOW, STATUS_BAD_GEOMETRY, STATUS_SURVIVOR_OVERFLOW, STATUS_ID_OVERFLOW, STATUS_PEER_TIMEOUT = 1, 2, 4, 8, 16
STATUS_CLIP_OVERFLOW = 32
STATUS_BAD_MATERIAL = 64

# Material — SceneGPU.hpp:67-82 / scene.slang:51-66 (56 B); MaterialFlag scene.slang:34-49
MATERIAL_DT = np.dtype(
    [
        ("albedo_color", "<u2", (4,)),
        ("emissive_color", "<u2", (3,)),
        ("roughness_factor", "<u2"),
        ("metallic_factor", "<u2"),
        ("alpha_cutoff", "<u2"),
        ("flags", "<u4"),
        ("sampler_index", "<u4"),
        ("albedo_image_index", "<u4"),
        ("normal_image_index", "<u4"),
        ("emissive_image_index", "<u4"),
        ("metallic_roughness_image_index", "<u4"),
        ("occlusion_image_index", "<u4"),
        ("uv_size", "<u2", (2,)),
        ("uv_offset", "<u2", (2,)),
    ]
)
assert MATERIAL_DT.itemsize == 56
MATERIAL_HAS_ALBEDO_IMAGE, MATERIAL_ALPHA_MASK = 1, 1 << 8
IMAGE_RGBA8_UNORM, IMAGE_R8_UNORM = 0, 1
FILTER_LINEAR, FILTER_NEAREST = 0, 1
ADDRESS_REPEAT, ADDRESS_CLAMP_TO_EDGE, ADDRESS_MIRRORED_REPEAT = 0, 1, 2
MIPMAP_LINEAR, MIPMAP_NEAREST = 0, 1
ALPHA_IMAGE_DT = np.dtype([("texels", "<u8"), ("width", "<u4"), ("height", "<u4"), ("format", "<u4"), ("level_count", "<u4")])
SAMPLER_DT = np.dtype([("mag_filter", "<u4"), ("min_filter", "<u4"), ("mipmap_mode", "<u4"), ("address_u", "<u4"), ("address_v", "<u4")])


def sampler(mag=FILTER_LINEAR, min=FILTER_LINEAR, mip=MIPMAP_LINEAR, u=ADDRESS_REPEAT, v=ADDRESS_REPEAT):  # noqa: A002
    """one SAMPLER_DT record (defaults = the reference's default sampler, Texture.hpp:38-45)"""
    return (mag, min, mip, u, v)


class MaterialTable(C.Structure):
    """OxcMaterialTable"""

    _fields_ = [
        ("materials", C.c_void_p),
        ("material_count", C.c_uint32),
        ("images", C.c_void_p),
        ("image_count", C.c_uint32),
        ("samplers", C.c_void_p),
        ("sampler_count", C.c_uint32),
    ]


# VSMPageState — Shaders/rmvsm.slang:16-28 ([Flags] enum)
VSM_PAGE_VISIBLE = 1
VSM_PAGE_DIRTY = 2
VSM_PAGE_BACKED = 4


class SceneDesc(C.Structure):
    """OxcSceneDesc"""

    _fields_ = [
        ("meshes", C.c_void_p),
        ("mesh_count", C.c_uint32),
        ("mesh_instances", C.c_void_p),
        ("mesh_instance_count", C.c_uint32),
        ("transforms", C.c_void_p),
        ("transform_count", C.c_uint32),
        ("blob", C.c_void_p),
        ("blob_size", C.c_uint64),
    ]


class CreateInfo(C.Structure):
    """OxcCreateInfo"""

    _fields_ = [
        ("max_mesh_instances", C.c_uint32),
        ("max_meshlet_instances", C.c_uint32),
        ("hiz_width", C.c_uint32),
        ("hiz_height", C.c_uint32),
        ("alloc_reordered_indices", C.c_uint32),
        ("max_views", C.c_uint32),
        ("max_mask_bits", C.c_uint32),
        ("wide_ids", C.c_uint32),
    ]


class Outputs(C.Structure):
    """OxcOutputs"""

    _fields_ = [
        ("visibility", C.c_void_p),
        ("cull_meshlets_cmd", C.c_void_p),
        ("cull_triangles_cmd", C.c_void_p),
        ("draw_cmd", C.c_void_p),
        ("meshlet_instances", C.c_void_p),
        ("visible_meshlet_instances_indices", C.c_void_p),
        ("meshlet_instance_visibility_mask", C.c_void_p),
        ("reordered_indices", C.c_void_p),
        ("mesh_instances", C.c_void_p),
        ("hiz", C.c_void_p),
        ("hiz_level_offset", C.c_uint32 * HIZ_MAX_LEVELS),
        ("hiz_levels", C.c_uint32),
        ("hiz_width", C.c_uint32),
        ("hiz_height", C.c_uint32),
        ("visibility_mask_words", C.c_uint32),
        ("view_visibility_bits", C.c_void_p),
        ("view_visible_counts", C.c_void_p),
        ("raster_triangle_count", C.c_void_p),
        ("status_flags", C.c_void_p),
        ("vis_primitive_bits", C.c_uint32),
    ]


VSM_CONTEXT_DT = np.dtype([("page_size", "<i4"), ("page_table_size", "<i4"), ("physical_page_table_size", "<i4"), ("curr_clipmap_index", "<i4"),
                           ("clipmap_count", "<i4"), ("depth_extent", "<i4", (2,)), ("first_clipmap_width", "<f4"),
                           ("clipmap_selection_bias", "<f4"), ("virtual_extent", "<f4"), ("z_length", "<f4"),
                           ("directional_light_dir", "<f4", (3,))])
assert VSM_CONTEXT_DT.itemsize == 56


class MgpuInfo(C.Structure):
    """OxcMgpuInfo"""

    _fields_ = [
        ("active", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32), ("survivor_capacity", C.c_uint32),
        ("hiz_over_peer_memory", C.c_uint32),
        ("gathered_counts", C.c_void_p * 2),
        ("gathered_ids", C.c_void_p * 2),
    ]


MGPU_ID_BYTES = 128


class DecodeTargets(C.Structure):
    """OxcDecodeTargets: device pointers of the five float4 planes (any may be NULL)"""

    _fields_ = [("lambda_", C.c_void_p), ("ddx", C.c_void_p), ("ddy", C.c_void_p), ("uv_normal", C.c_void_p), ("uv_grad", C.c_void_p)]


MESH_MAX_LODS = 8


class MeshInput(C.Structure):
    """OxbMeshInput"""

    _fields_ = [
        ("positions", C.c_void_p),
        ("normals", C.c_void_p),
        ("texcoords", C.c_void_p),
        ("vertex_count", C.c_uint32),
        ("lod_count", C.c_uint32),
        ("lod_indices", C.c_void_p * MESH_MAX_LODS),
        ("lod_index_counts", C.c_uint32 * MESH_MAX_LODS),
        ("lod_errors", C.c_float * MESH_MAX_LODS),
        ("cluster_mode", C.c_uint32),
        ("auto_lods", C.c_uint32),
    ]


class OrcHiz(C.Structure):
    """OrcHiz (oracle/oxc_oracle.h) — also used as a plain layout helper"""

    _fields_ = [
        ("data", C.c_void_p),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("levels", C.c_uint32),
        ("level_offset", C.c_uint32 * HIZ_MAX_LEVELS),
    ]


class FrameResult(C.Structure):
    """OxrFrameResult"""

    _fields_ = [
        ("total", C.c_uint32),
        ("early", C.c_uint32),
        ("late", C.c_uint32),
        ("draw_index_count_early", C.c_uint32),
        ("draw_index_count_late", C.c_uint32),
        ("raster_triangles", C.c_uint64),
    ]


def hiz_extent(width: int, height: int):
    """RendererInstance.cpp:573-577: bit_ceil((W+1)>>1) per axis."""

    def bit_ceil(v):
        return 1 if v <= 1 else 1 << (int(v) - 1).bit_length()

    return bit_ceil((width + 1) >> 1), bit_ceil((height + 1) >> 1)


def hiz_layout(w: int, h: int):
    """(levels, offsets[levels], total_texels) — Texture.hpp:144-146 + min(.,13)."""
    levels = min(int(max(w, h)).bit_length(), HIZ_MAX_LEVELS)
    offs, off = [], 0
    for l in range(levels):
        offs.append(off)
        off += max(1, w >> l) * max(1, h >> l)
    return levels, offs, off
