#!/usr/bin/env python
"""CPU prototype + soundness check of an EXACT instance-level occlusion shortcut for the late meshlet cull
(not in the product yet).

Claim: if, for a mesh instance with union box U of its decoded meshlet boxes,
  (a) every clip-space corner of U has w >= near * (1 + 2^-10) + 2^-20 (|mvp row 3| . |corner| bound), and
  (b) zmax(U) + margin <= D - 1e-7 - margin, D = min over the Hi-Z texels T at level L that cover every texel any meshlet's
      canonical test could sample (instance texel rect at mip 0, expanded by 2.5 * 2^L texels, L = min(mip_I + 1, levels - 1)),
then the canonical per-meshlet test (project_aabb + test_occlusion) reports 'occluded' for EVERY meshlet of the instance.
The script evaluates the shortcut in float64 with the stated margins and compares with the oracle's per-meshlet f32
decisions on the late pass of several frames: any flagged instance with a non-occluded meshlet is a counter-example."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc  # noqa: E402
from oxylus_b200 import abi, synth  # noqa: E402


def lod_union_aabbs(sc):
    """(mesh, lod) -> (min[3], max[3]) of decoded meshlet boxes, as k_lod_union_aabb computes them"""
    out = {}
    for m, mesh in enumerate(sc.meshes):
        lods = np.frombuffer(sc.blob, dtype=abi.MESH_LOD_DT, count=int(mesh["lod_count"]), offset=int(mesh["lods"]))
        for l, lod in enumerate(lods):
            b = np.frombuffer(sc.blob, dtype=abi.MESHLET_BOUNDS_DT, count=int(lod["meshlet_count"]), offset=int(lod["meshlet_bounds"]))
            c = b["aabb_center"].copy().view(np.float16).astype(np.float64)
            e = np.abs(b["aabb_extent"].copy().view(np.float16).astype(np.float64)) * 0.5
            out[(m, l)] = ((c - e).min(axis=0), (c + e).max(axis=0))
    return out


def instance_flags(sc, hs, cam, hiz, unions):
    pv = cam["projection_view"][0].reshape(4, 4).T.astype(np.float64)
    near = float(cam["near_clip"][0])
    W, H, levels = hiz.w, hiz.h, hiz.levels
    flags = np.zeros(sc.mesh_instance_count, dtype=bool)
    for i, inst in enumerate(hs.mesh_instances):
        lo, hi = unions[(int(inst["mesh_index"]), int(inst["lod_index"]))]
        world = sc.transforms["world"][inst["transform_index"]].reshape(4, 4).T.astype(np.float64)
        mvp = pv @ world
        corners = np.array([[x, y, z, 1.0] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        clip = corners @ mvp.T
        wbound = np.abs(corners) @ np.abs(mvp[3])          # magnitude scale for the f32 rounding of w
        if not np.all(clip[:, 3] >= near * (1 + 2.0**-10) + wbound * 2.0**-20):
            continue                                          # (a)
        ndc = clip[:, :3] / clip[:, 3:4]
        zmax = ndc[:, 2].max()
        u = ndc[:, :2] * 0.5 + 0.5
        # instance texel rect at mip 0 with one texel of slack for the f32 rounding of the per-meshlet conversions
        tx0 = max(np.floor(u[:, 0].min() * W) - 1, 0); tx1 = min(np.floor(u[:, 0].max() * W) + 1, W - 1)
        ty0 = max(np.floor(u[:, 1].min() * H) - 1, 0); ty1 = min(np.floor(u[:, 1].max() * H) + 1, H - 1)
        if tx1 < tx0 or ty1 < ty0:
            continue
        size = max(tx1 - tx0, ty1 - ty0)
        mip_i = 0 if size <= 1 else int(np.ceil(np.log2(size)))
        L = min(mip_i + 1, levels - 1)
        s = 1 << L
        lw, lh = max(1, W >> L), max(1, H >> L)
        ax0 = int(max(np.floor((tx0 - 2.5 * s) / s), 0)); ax1 = int(min(np.floor((tx1 + 2.5 * s) / s), lw - 1))
        ay0 = int(max(np.floor((ty0 - 2.5 * s) / s), 0)); ay1 = int(min(np.floor((ty1 + 2.5 * s) / s), lh - 1))
        D = hiz.level(L)[ay0:ay1 + 1, ax0:ax1 + 1].min()
        margin = 1e-5 * (1 + abs(zmax))
        flags[i] = zmax + margin <= float(D) - 1e-7 - margin
    return flags


def check(n_meshlets, frames, **kw):
    sc = synth.make_scene(n_meshlets, config_index=2, **kw)
    hs = orc.HostScene(sc)
    unions = None
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    tot = dict(instances=0, flagged=0, meshlets=0, meshlets_in_flagged=0, occluded_or_culled=0, counter_examples=0)
    for f in range(frames):
        cam = sc.camera(2.0 * (f % 2) + 7.0 * (f // 2))
        r = orc.frame(hs, cam, sc.width, sc.height, mask.copy(), sc.occluder_depth)   # only to get this frame's Hi-Z and lods
        mask_before_late = r["mask_after_early"]
        if unions is None or True:
            unions = lod_union_aabbs(sc)
        flags = instance_flags(sc, hs, cam, r["hiz"], unions)
        vis_flags = orc.cull_meshlets_flags(hs, r["meshlet_instances"], r["visibility"], cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS,
                                            r["hiz"], mask_before_late)
        total = int(r["visibility"]["total"][0])
        inst_of = r["meshlet_instances"]["mesh_instance_index"][:total]
        in_flagged = flags[inst_of]
        bad = in_flagged & (vis_flags[:total] != 0)
        tot["instances"] += len(np.unique(inst_of)); tot["flagged"] += int(flags[np.unique(inst_of)].sum())
        tot["meshlets"] += total; tot["meshlets_in_flagged"] += int(in_flagged.sum())
        tot["occluded_or_culled"] += int((vis_flags[:total] == 0).sum()); tot["counter_examples"] += int(bad.sum())
        mask[:] = 0  # every frame judged from an empty mask: the late decision does not depend on it
    return tot


if __name__ == "__main__":
    for kw in (dict(n_meshlets=60000, width=1280, height=720, n_unique_meshes=32),
               dict(n_meshlets=60000, width=1280, height=720, n_unique_meshes=32, placement="box"),
               dict(n_meshlets=30000, width=800, height=450, n_unique_meshes=24, max_lods=3, ragged=True),
               dict(n_meshlets=300000, width=1920, height=1080, n_unique_meshes=64)):
        n = kw.pop("n_meshlets")
        t = check(n, 4, **kw)
        print(kw, t, "coverage of occluded/culled meshlets: %.1f %%" % (100.0 * t["meshlets_in_flagged"] / max(1, t["occluded_or_culled"])))
