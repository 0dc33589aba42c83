import numpy as np

from oxylus_b200 import abi, synth


def test_scene_is_deterministic_and_exact(small_scene):
    a = small_scene
    b = synth.make_scene(6000, config_index=2, width=640, height=360, n_unique_meshes=16)
    assert a.max_meshlet_instance_count == 6000 == b.max_meshlet_instance_count
    np.testing.assert_array_equal(a.blob, b.blob)
    np.testing.assert_array_equal(a.transforms["world"], b.transforms["world"])
    np.testing.assert_array_equal(a.mesh_instances, b.mesh_instances)
    c = synth.make_scene(6000, config_index=3, width=640, height=360, n_unique_meshes=16)
    assert not np.array_equal(a.transforms["world"], c.transforms["world"])


def test_visibility_offsets_are_lod0_prefix_sums(small_scene):
    """Scene.cpp:1255-1260"""
    sc = small_scene
    lods = []
    for inst in sc.mesh_instances:
        m = sc.meshes[inst["mesh_index"]]
        lod0 = np.frombuffer(sc.blob, dtype=abi.MESH_LOD_DT, count=1, offset=int(m["lods"]))[0]
        lods.append(int(lod0["meshlet_count"]))
    want = np.concatenate([[0], np.cumsum(lods)[:-1]])
    np.testing.assert_array_equal(sc.mesh_instances["meshlet_instance_visibility_offset"], want)
    assert sum(lods) == sc.max_meshlet_instance_count


def test_blob_layout_alignment_and_bounds(small_scene):
    sc = small_scene
    for m in sc.meshes:
        assert m["vertex_positions"] % 16 == 0 and m["lods"] % 8 == 0
        lods = np.frombuffer(sc.blob, dtype=abi.MESH_LOD_DT, count=int(m["lod_count"]), offset=int(m["lods"]))
        for l in lods:
            assert l["meshlet_bounds"] % 16 == 0 and l["meshlets"] % 16 == 0
            ml = np.frombuffer(sc.blob, dtype=abi.MESHLET_DT, count=int(l["meshlet_count"]), offset=int(l["meshlets"]))
            assert np.all(ml["triangle_count"] <= 64) and np.all(ml["vertex_count"] <= 64) and np.all(ml["triangle_count"] >= 1)
            end = ml["local_triangle_index_offset"] + ml["triangle_count"] * 3
            assert end.max() <= l["local_triangle_indices_count"]


def test_camera_is_reverse_z():
    """Camera.cpp:36-54: near -> 1, far -> 0, y flipped."""
    cam = synth.make_camera(1920, 1080, 1)
    pv = cam["projection_view"][0].reshape(4, 4).T.astype(np.float64)
    near = pv @ np.array([0, 0, -0.1, 1.0])
    far = pv @ np.array([0, 0, -1000.0, 1.0])
    assert abs(near[2] / near[3] - 1.0) < 1e-5 and abs(far[2] / far[3]) < 1e-6
    up = pv @ np.array([0, 1.0, -10.0, 1.0])
    assert up[1] / up[3] < 0  # +Y world maps to -Y NDC (Vulkan y-down)


def test_splitmix_reference_values():
    # SplitMix64 with seed 0: first outputs of the canonical generator
    z = synth.splitmix64(0, 0, 3)
    assert [int(x) for x in z] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
