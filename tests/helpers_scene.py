"""Hand-built minimal scenes for known-answer tests."""
import numpy as np

from oxylus_b200 import abi, synth


def quad_scene(width, height, depth_a=0.5, depth_b=0.5):
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
