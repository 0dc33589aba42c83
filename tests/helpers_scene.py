"""Hand-built minimal scenes for known-answer tests."""
import numpy as np

from oxylus_b200 import abi, synth


def quad_scene(width, height, depth_a=0.5, depth_b=0.5, attributes=False):
    """One mesh, one meshlet: a quad (2 triangles, 4 vertices) spanning x,y in [-0.5,0.5] in a clip space where
    projection_view = diag(1,1,1,1) (so clip == local, w = 1).  Winding is front-facing (negative xyw determinant)."""
    pos = np.array([[-0.5, -0.5, depth_a], [0.5, -0.5, depth_a], [0.5, 0.5, depth_b], [-0.5, 0.5, depth_b]], dtype=np.float32)
    # front-facing == negative determinant of [x y w] rows == clockwise in (x right, y up)
    tris = np.array([[0, 2, 1], [0, 3, 2]], dtype=np.uint8)
    chunks, off = [], 0

    def put(a, align=16):
        nonlocal off
        pad = (-off) % align
        if pad:
            chunks.append(np.zeros(pad, dtype=np.uint8))
            off += pad
        start = off
        raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        chunks.append(raw)
        off += raw.size
        return start

    pq = np.zeros((4, 4), dtype=np.uint16)
    pq[:, :3] = synth.quantize_half(pos)
    mesh = np.zeros(1, dtype=abi.MESH_DT)
    mesh["vertex_positions"] = put(pq)
    mesh["vertex_count"] = 4
    mesh["lod_count"] = 1
    if attributes:  # normals (0,0,1) packed 10:10:10 (scene.slang:486-489), uv = xy + 0.5 as half2
        mesh["vertex_normals"] = put(np.full(4, (511 << 20) | (511 << 10) | 1022, dtype=np.uint32))
        mesh["texture_coords"] = put(synth.quantize_half(pos[:, :2] + np.float32(0.5)))
    meshlet = np.zeros(1, dtype=abi.MESHLET_DT)
    meshlet["vertex_count"] = 4
    meshlet["triangle_count"] = 2
    bounds = np.zeros(1, dtype=abi.MESHLET_BOUNDS_DT)
    bounds["aabb_center"][0] = synth.quantize_half(np.array([0.0, 0.0, (depth_a + depth_b) / 2], dtype=np.float32))
    bounds["aabb_extent"][0] = synth.quantize_half(np.array([1.0, 1.0, abs(depth_a - depth_b) + 0.01], dtype=np.float32))
    bounds["cone_cutoff"] = 127
    micro = np.zeros(8, dtype=np.uint8)
    micro[:6] = tris.reshape(-1)
    lod = np.zeros(1, dtype=abi.MESH_LOD_DT)
    lod["meshlets"] = put(meshlet)
    lod["meshlet_bounds"] = put(bounds)
    lod["local_triangle_indices"] = put(micro)
    lod["indirect_vertex_indices"] = put(np.arange(4, dtype=np.uint32))
    lod["meshlet_count"] = 1
    lod["meshlet_bounds_count"] = 1
    lod["local_triangle_indices_count"] = 8
    lod["indirect_vertex_indices_count"] = 4
    mesh["lods"] = put(lod)
    mesh["bounds"]["aabb_center"][0] = (0.0, 0.0, (depth_a + depth_b) / 2)
    mesh["bounds"]["aabb_extent"][0] = (1.0, 1.0, abs(depth_a - depth_b) + 0.01)
    blob = np.concatenate(chunks)
    blob = np.concatenate([blob, np.zeros((-blob.size) % 16, dtype=np.uint8)])
    inst = np.zeros(1, dtype=abi.MESH_INSTANCE_DT)
    xf = np.zeros(1, dtype=abi.TRANSFORM_DT)
    xf["world"][0] = np.eye(4, dtype=np.float32).reshape(16)
    sc = synth.Scene(mesh, inst, xf, blob, 1, width, height, 0)
    cam = np.zeros(1, dtype=abi.CULL_CAMERA_DT)
    cam["projection_view"][0] = np.eye(4, dtype=np.float32).reshape(16)
    cam["position"][0] = (0.0, 0.0, 10.0)
    cam["acceptable_lod_error"] = 2.0
    cam["resolution"][0] = (width, height)
    cam["near_clip"] = 0.01
    cam["mesh_instance_count"] = 1
    return sc, cam


def boxes_scene(centers, extents, width, height, cone_cutoff=127, translations=None):
    """One mesh per box: a single meshlet whose MeshletBounds are the given (half-quantised) centre / full extent and
    whose geometry is one tiny triangle at the centre.  Identity transforms.  Used for adversarial predicate tests."""
    centers = np.asarray(centers, dtype=np.float32)
    extents = np.asarray(extents, dtype=np.float32)
    n = len(centers)
    chunks, off = [], 0

    def put(a, align=16):
        nonlocal off
        pad = (-off) % align
        if pad:
            chunks.append(np.zeros(pad, dtype=np.uint8))
            off += pad
        start = off
        raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        chunks.append(raw)
        off += raw.size
        return start

    cq = synth.quantize_half(centers)
    eq = synth.quantize_half(extents)
    bounds = np.zeros(n, dtype=abi.MESHLET_BOUNDS_DT)
    bounds["aabb_center"] = cq
    bounds["aabb_extent"] = eq
    bounds["cone_cutoff"] = cone_cutoff
    cdec = cq.view(np.float16).astype(np.float32)
    edec = eq.view(np.float16).astype(np.float32)
    pos = np.zeros((n, 3, 4), dtype=np.uint16)
    tri = np.stack([cdec, cdec + np.float32([0.01, 0, 0]), cdec + np.float32([0, 0.01, 0])], axis=1)
    pos[:, :, :3] = synth.quantize_half(tri)
    pos_off = put(pos.reshape(-1, 4))
    meshlets = np.zeros(n, dtype=abi.MESHLET_DT)
    meshlets["indirect_vertex_index_offset"] = np.arange(n) * 3
    meshlets["local_triangle_index_offset"] = np.arange(n) * 4
    meshlets["vertex_count"] = 3
    meshlets["triangle_count"] = 1
    micro = np.tile(np.array([0, 1, 2, 0], dtype=np.uint8), n)
    vidx = np.arange(n * 3, dtype=np.uint32)
    m_off, b_off, mi_off, v_off = put(meshlets), put(bounds), put(micro), put(vidx)
    lods = np.zeros(n, dtype=abi.MESH_LOD_DT)
    lods["meshlets"] = m_off + np.arange(n) * 16
    lods["meshlet_bounds"] = b_off + np.arange(n) * 16
    lods["local_triangle_indices"] = mi_off
    lods["indirect_vertex_indices"] = v_off
    lods["meshlet_count"] = 1
    lods["meshlet_bounds_count"] = 1
    lods["local_triangle_indices_count"] = micro.size
    lods["indirect_vertex_indices_count"] = vidx.size
    l_off = put(lods, 16)
    meshes = np.zeros(n, dtype=abi.MESH_DT)
    meshes["vertex_positions"] = pos_off
    meshes["vertex_count"] = n * 3
    meshes["lod_count"] = 1
    meshes["lods"] = l_off + np.arange(n) * 64
    meshes["bounds"]["aabb_center"] = cdec
    meshes["bounds"]["aabb_extent"] = edec
    blob = np.concatenate(chunks)
    blob = np.concatenate([blob, np.zeros((-blob.size) % 16, dtype=np.uint8)])
    inst = np.zeros(n, dtype=abi.MESH_INSTANCE_DT)
    inst["mesh_index"] = np.arange(n)
    inst["meshlet_instance_visibility_offset"] = np.arange(n)
    if translations is None:
        xf = np.zeros(1, dtype=abi.TRANSFORM_DT)
        xf["world"][0] = np.eye(4, dtype=np.float32).reshape(16)
    else:
        xf = np.zeros(n, dtype=abi.TRANSFORM_DT)
        wm = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))  # [col][row]
        wm[:, 3, :3] = np.asarray(translations, dtype=np.float32)
        xf["world"] = wm.reshape(n, 16)
        inst["transform_index"] = np.arange(n)
    return synth.Scene(meshes, inst, xf, blob, n, width, height, 0), cdec, edec
