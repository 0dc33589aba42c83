"""Thread-independent KERNELS of the product run on the HOST — the kernel functions themselves, one simulated thread at a time
(tests/host_kernels.cpp through tests/host_shim/) — against the oracle, on the synthetic scenes the GPU parity tests use:
k_cull_meshes (mesh-level frustum + LOD selection + counts + lod_index write-back, whole scene and a shard),
k_decode_visbuffer (vis-buffer decode, every float).  The GPU tier checks the same source as compiled by nvcc.  Test
infrastructure only: the product has no CPU path (test_abi_cpu.py checks that)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oxylus_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = {
    "small": dict(n_meshlets=6000, width=320, height=180, n_unique_meshes=16),
    "box_ragged_lods": dict(n_meshlets=20000, width=320, height=180, n_unique_meshes=48, max_lods=4, ragged=True, placement="box"),
    "ragged_lods": dict(n_meshlets=9000, width=400, height=225, n_unique_meshes=24, max_lods=3, ragged=True),
}


@pytest.fixture(scope="module")
def hk(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostkernels") / "libhostkernels.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I", os.path.join(ROOT, "tests", "host_shim"),
                           "-I", os.path.join(ROOT, "oxylus_b200", "csrc"), os.path.join(ROOT, "tests", "host_kernels.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.hk_scene_create.restype = C.c_void_p
    lib.hk_scene_create.argtypes = [C.c_void_p]
    lib.hk_scene_destroy.argtypes = [C.c_void_p]
    vp, u32 = C.c_void_p, C.c_uint32
    lib.hk_cull_meshes.argtypes = [vp, vp, u32, u32, u32, vp, vp]
    lib.hk_decode.argtypes = [vp, vp, vp, vp, u32, u32, vp, u32, u32, vp, vp, vp, vp, vp]
    return lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _expand(counts, first):
    """the deterministic expansion order of k_expand_meshlet_instances: ascending mesh instance, ascending meshlet"""
    out = np.zeros(int(counts.sum()), dtype=abi.MESHLET_INSTANCE_DT)
    inst = np.repeat(np.arange(len(counts), dtype=np.uint32) + first, counts)
    start = np.repeat(np.cumsum(counts) - counts, counts)
    out["mesh_instance_index"] = inst
    out["meshlet_index"] = np.arange(len(inst), dtype=np.uint32) - start
    return out


@pytest.mark.parametrize("name", ["small", "box_ragged_lods"])
def test_cull_meshes_kernel_on_the_host(orc, hk, name):
    sc = synth.make_scene(config_index=3, **SCENES[name])
    hs = orc.HostScene(sc)
    h = hk.hk_scene_create(C.cast(hs.ref, C.c_void_p))
    n = sc.mesh_instance_count
    for yaw, flags in ((0.0, abi.CULL_TEST_ALL), (25.0, abi.CULL_TEST_ALL), (0.0, abi.CULL_TEST_FRUSTUM), (0.0, abi.CULL_SELECT_LOD)):
        cam = sc.camera(yaw)
        mi_ref, vis_ref, _ = orc.cull_meshes(hs, cam, flags)
        total = int(vis_ref["total"][0])
        counts = np.zeros(n, dtype=np.uint32)
        lod = np.zeros(n, dtype=np.uint32)
        assert hk.hk_cull_meshes(h, _p(cam), flags, 0, n, _p(counts), _p(lod)) == 0
        assert int(counts.sum()) == total
        np.testing.assert_array_equal(_expand(counts, 0), mi_ref[:total])
        np.testing.assert_array_equal(lod, hs.mesh_instances["lod_index"])
        if flags == abi.CULL_SELECT_LOD:
            assert total == 0                                  # cull_meshes.slang:34: nothing without TestFrustum
        # a shard (multi-GPU): same decisions, local numbering of the counts
        first, cnt = n // 3, n // 2
        mi_s, vis_s, _ = orc.cull_meshes(orc.HostScene(sc), cam, flags, first, cnt)
        cs = np.zeros(cnt, dtype=np.uint32)
        assert hk.hk_cull_meshes(h, _p(cam), flags, first, cnt, _p(cs), _p(lod)) == 0
        np.testing.assert_array_equal(_expand(cs, first), mi_s[: int(vis_s["total"][0])])
    if SCENES[name].get("max_lods", 1) > 1:
        assert len(np.unique(hs.mesh_instances["lod_index"])) > 1   # the LOD selection really selected
    hk.hk_scene_destroy(h)


@pytest.mark.parametrize("name", ["small", "ragged_lods"])
def test_decode_kernel_on_the_host(orc, hk, name):
    sc = synth.make_scene(config_index=2, **SCENES[name])
    hs = orc.HostScene(sc)
    w, hgt = sc.width, sc.height
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    for f in range(2):
        cam = sc.camera(2.0 * f)
        ref = orc.frame(hs, cam, w, hgt, mask, sc.occluder_depth)
    total = int(ref["visibility"]["total"][0])
    v32, _ = orc.resolve(ref["vis64"])
    v32 = v32.copy()
    v32[0, :7] = [(0xFFFFFE << 8) | 5, ((total + 3) << 8) | 1, (total << 8), 0xFFFFFF00, ((total - 1) << 8), 0, 0xFFFFFFFF]  # hostile texels
    want = orc.decode_visbuffer(hs, ref["meshlet_instances"], total, cam, v32)
    assert (want["lambda_"][:, :, 3] == 1.0).sum() > 100
    h = hk.hk_scene_create(C.cast(hs.ref, C.c_void_p))
    n = sc.mesh_instance_count
    counts, lod = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
    assert hk.hk_cull_meshes(h, _p(cam), abi.CULL_TEST_ALL, 0, n, _p(counts), _p(lod)) == 0   # resolves LODs + pointers, like the frame
    got = {k: np.full((hgt, w, 4), np.nan, dtype=np.float32) for k in want}
    mi = np.ascontiguousarray(ref["meshlet_instances"])
    assert hk.hk_decode(h, _p(cam), _p(v32), None, w, hgt, _p(mi), total, 8, _p(got["lambda_"]), _p(got["ddx"]), _p(got["ddy"]), _p(got["uv_normal"]),
                        _p(got["uv_grad"])) == 0
    for k in want:
        a, b = got[k].view(np.uint32), want[k].view(np.uint32)
        nan_both = np.isnan(got[k]) & np.isnan(want[k])
        assert np.array_equal(a[~nan_both], b[~nan_both]), k
    # the packed 64-bit image as input gives the same planes
    got2 = {k: np.zeros((hgt, w, 4), dtype=np.float32) for k in ("lambda_", "uv_normal")}
    v64 = (ref["vis64"] & np.uint64(0xFFFFFFFF00000000)) | v32.astype(np.uint64)
    assert hk.hk_decode(h, _p(cam), None, _p(v64), w, hgt, _p(mi), total, 8, _p(got2["lambda_"]), None, None, _p(got2["uv_normal"]), None) == 0
    for k in got2:
        nan_both = np.isnan(got2[k]) & np.isnan(want[k])
        assert np.array_equal(got2[k].view(np.uint32)[~nan_both], want[k].view(np.uint32)[~nan_both]), k
    hk.hk_scene_destroy(h)
