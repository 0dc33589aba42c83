#!/usr/bin/env python
"""Hostile-but-representable inputs through the whole two-pass frame: MeshletBounds fields set to NaN, +-Inf, negative extents,
+-0, the largest half, denormal-flushed values, garbage cone bytes; vertex positions with the same special values; transforms scaled by 1e-12 / 1e6, mirrored, with a zero
column, at the camera, far away.  No mesh builder produces such records, but the reference's shaders — and therefore the oracle —
are defined for them, and "bit-identical to the canonical evaluation for every input" has to hold for them too (this is the
scenario that found the cone filter turning a NaN radius into 0 and the frustum filter assuming h >= 0).
Run by tests/test_emulated_library_cpu.py against the SIMT-emulated library; works unchanged on a GPU.

    python tests/emulated_torture_check.py [seeds]"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc  # noqa: E402
from oxylus_b200 import abi, capi, synth  # noqa: E402

SPECIALS = np.array([0x0000, 0x8000, 0x0001, 0x03FF, 0x0400, 0x7BFF, 0xFBFF, 0x7C00, 0xFC00, 0x7E00, 0x3C00, 0xBC00], dtype=np.uint16)


def mutate(base, rng, mode):
    sc = copy.deepcopy(base)
    for m in sc.meshes:
        lods = np.frombuffer(sc.blob, dtype=abi.MESH_LOD_DT, count=int(m["lod_count"]), offset=int(m["lods"]))
        for d in lods:
            n = int(d["meshlet_bounds_count"])
            b = np.ndarray((n, 8), dtype=np.uint16, buffer=sc.blob.data, offset=int(d["meshlet_bounds"]))  # c.xyz | cone xy | e.xyz | cone z, cutoff
            hit = rng.random(n) < 0.15
            if mode in ("bounds", "all"):
                for c in (0, 1, 2, 4, 5, 6):
                    sel = hit & (rng.random(n) < 0.5)
                    b[sel, c] = SPECIALS[rng.integers(0, len(SPECIALS), int(sel.sum()))]
            if mode in ("cones", "all"):
                sel = rng.random(n) < 0.3
                b[sel, 3] = rng.integers(0, 65536, int(sel.sum())).astype(np.uint16)
                b[sel, 7] = rng.integers(0, 65536, int(sel.sum())).astype(np.uint16)
    if mode in ("vertices", "all"):  # vertex positions: u16 x 4 per vertex (scene.slang:478-484)
        for m in sc.meshes:
            n = int(m["vertex_count"])
            v = np.ndarray((n, 4), dtype=np.uint16, buffer=sc.blob.data, offset=int(m["vertex_positions"]))
            for c in range(3):
                sel = rng.random(n) < 0.02
                v[sel, c] = SPECIALS[rng.integers(0, len(SPECIALS), int(sel.sum()))]
    if mode in ("transforms", "all"):
        t = sc.transforms["world"]
        for i in np.nonzero(rng.random(len(t)) < 0.2)[0]:
            kind = int(rng.integers(0, 6))
            if kind == 0:
                t[i, :12] *= np.float32(1e-12)
            elif kind == 1:
                t[i, :12] *= np.float32(1e6)
            elif kind == 2:
                t[i, 0:3] *= np.float32(-1)
            elif kind == 3:
                t[i, 0:3] = 0
            elif kind == 4:
                t[i, 12:15] = (0, 0, 0)
            else:
                t[i, 12:15] *= np.float32(1e4)
    return sc


def frames_equal(sc, frames=2):
    hs = orc.HostScene(sc)
    w, h = sc.width, sc.height
    hw, hh = sc.hiz_extent()
    ctx = capi.Context(0, sc.mesh_instance_count, sc.max_meshlet_instance_count, hw, hh)
    ctx.set_scene(sc)
    vis, occ = ctx.alloc(w * h * 8), ctx.alloc(w * h * 4)
    ctx.upload(occ, sc.occluder_depth)
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    ok = True
    for f in range(frames):
        cam = sc.camera(3.0 * f)
        ref = orc.frame(hs, cam, w, h, mask, sc.occluder_depth)
        ctx.clear_visbuffer_with_depth(vis, occ, w, h)
        ctx.clear_hiz()
        ctx.cull_meshes(cam, abi.CULL_TEST_ALL)
        ctx.cull_meshlets(cam, abi.CULL_TEST_ALL, True)
        ctx.raster_visbuffer(cam, abi.CULL_TEST_ALL, w, h, vis)
        ctx.build_hiz_packed(vis, w, h)
        ctx.cull_meshlets(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, True)
        ctx.raster_visbuffer(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, w, h, vis)
        v = ctx.visibility()
        e, l = int(v["early"][0]), int(v["late"][0])
        same = (e, l) == (ref["early"], ref["late"])
        same = same and np.array_equal(np.sort(ctx.visible_indices(e + l)), np.sort(ref["visible"][: ref["early"] + ref["late"]]))
        same = same and np.array_equal(ctx.download(vis, np.uint64, w * h).reshape(h, w), ref["vis64"]) and np.array_equal(ctx.mask(), mask)
        ok = ok and bool(same)
    ctx.close()
    return ok


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    base = synth.make_scene(8000, config_index=2, width=320, height=180, n_unique_meshes=24, max_lods=2, ragged=True)
    bad = []
    for seed in range(seeds):
        for mode in ("bounds", "cones", "vertices", "transforms", "all"):
            if not frames_equal(mutate(base, np.random.default_rng(seed * 10 + 1), mode)):
                bad.append((seed, mode))
    print(f"{seeds * 5} hostile scenes x 2 frames: {'ok' if not bad else 'MISMATCH ' + repr(bad)}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
