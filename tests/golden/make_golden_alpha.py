"""Generates tests/golden/alpha_small.json from the CPU oracle: the alpha-tested discard of the vis-buffer encode
(visbuffer_encode.slang:54-66) on the golden scene, with a single-level and with a mip-mapped material table.

Like frames_small.json this pins the ORACLE (and, in the GPU tier, the CUDA path) AGAINST REGRESSION only: the reference has no
vectors for this path, and its interpolation / sampling run in fixed-function hardware — the arithmetic is the written specification
above raster_triangle in oracle/oxc_oracle.c ("parity unpinned").

    python tests/golden/make_golden_alpha.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle as orc  # noqa: E402
from oxylus_b200 import abi, synth  # noqa: E402
from test_oracle_alpha import checker, material  # noqa: E402

SCENE = dict(n_meshlets=6000, config_index=2, width=640, height=360, n_unique_meshes=16)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tables():
    """name -> (materials, images, samplers): what both the oracle table and oxc_set_materials are built from"""
    rng = np.random.default_rng(2024)
    noise = rng.integers(0, 256, (16, 16), dtype=np.uint8)
    gradient = np.ascontiguousarray(np.tile(np.linspace(0, 255, 16).astype(np.uint8), (16, 1)))
    mats = np.array([material(), material(image=0, cutoff=0.5), material(image=1, cutoff=0.4, albedo_a=0.9, sampler=1),
                     material(image=2, cutoff=0.5, sampler=2)], dtype=abi.MATERIAL_DT)
    single = ([(checker(32, 4), abi.IMAGE_RGBA8_UNORM), (noise, abi.IMAGE_R8_UNORM), (gradient, abi.IMAGE_R8_UNORM)],
              np.array([abi.sampler(), abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST, u=abi.ADDRESS_CLAMP_TO_EDGE, v=abi.ADDRESS_CLAMP_TO_EDGE),
                        abi.sampler(u=abi.ADDRESS_MIRRORED_REPEAT)], dtype=abi.SAMPLER_DT))
    mipped = ([(orc.mip_chain(checker(32, 4)), abi.IMAGE_RGBA8_UNORM), (orc.mip_chain(noise), abi.IMAGE_R8_UNORM), (orc.mip_chain(gradient), abi.IMAGE_R8_UNORM)],
              np.array([abi.sampler(), abi.sampler(mip=abi.MIPMAP_NEAREST, u=abi.ADDRESS_CLAMP_TO_EDGE, v=abi.ADDRESS_MIRRORED_REPEAT),
                        abi.sampler(mag=abi.FILTER_NEAREST, min=abi.FILTER_LINEAR)], dtype=abi.SAMPLER_DT))
    return {"single_level": (mats, *single), "mip_mapped": (mats, *mipped)}


def scene():
    sc = synth.make_scene(**SCENE)
    sc.mesh_instances["material_index"] = np.arange(sc.mesh_instance_count) % 4
    return sc


def generate(frame_fn=None):
    """frame_fn(name, table parts, scene, camera, frame index) -> dict(vis64, early, late, ntri) replaces the oracle frame (the GPU tier
    passes the CUDA path in); None = the oracle"""
    out = {}
    for name, (mats, images, smp) in tables().items():
        sc = scene()
        hs = orc.HostScene(sc)
        tab = orc.MaterialTable(mats, images, smp)
        mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
        frames = []
        for f in range(2):
            cam = sc.camera(2.0 * f)
            if frame_fn is None:
                r = orc.frame(hs, cam, sc.width, sc.height, mask, sc.occluder_depth, materials=tab)
                r = dict(vis64=r["vis64"], early=r["early"], late=r["late"], ntri=r["ntri_early"] + r["ntri_late"])
            else:
                r = frame_fn(name, (mats, images, smp, tab), sc, cam, f)
            v32 = (r["vis64"] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            frames.append(dict(yaw=2.0 * f, early=int(r["early"]), late=int(r["late"]), ntri=int(r["ntri"]), vis64_sha=sha(r["vis64"]),
                               covered_pixels=int((v32 != 0xFFFFFFFF).sum())))
        out[name] = frames
    return dict(scene=SCENE, tables=out)


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "alpha_small.json")
    json.dump(generate(), open(path, "w"), indent=1)
    print("wrote", path)
