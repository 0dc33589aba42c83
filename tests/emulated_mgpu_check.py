#!/usr/bin/env python
"""G emulated ranks == 1 context, bit for bit, through the product's multi-GPU API (oxc_mgpu_*) — the CPU-tier counterpart of
tools/check_multi_gpu.py.  Run by tests/test_emulated_library_cpu.py with OXC_LIB_PATH = the SIMT-emulated library and an
in-process NCCL stand-in on LD_LIBRARY_PATH: every rank is a THREAD with its own OxcContext (oxc_set_shard_auto), the Hi-Z
exchange writes into the other ranks' buffers ("peer memory" = shared memory here) and waits on their flags, the trailing
exchange goes through the communicator.  Compared with a single context over the whole scene: merged vis buffer, gathered
survivor ids (as a set), counters, every Hi-Z level on every rank, each rank's slice of the persistent mask.

    python tests/emulated_mgpu_check.py WORLD [MESHLETS] [alpha]

`alpha`: the mesh instances cycle through four materials and every context gets the material table (oxc_set_materials), so the
shards' rasters discard fragments (visbuffer_encode.slang:54-66): the holes travel through the Hi-Z exchange and the merge."""
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oxylus_b200 import abi, capi, synth  # noqa: E402  (not oxylus_b200.dist: it imports torch, which brings the real NCCL into the process)


def lod0_counts_of(scene):  # == oxylus_b200.dist.lod0_counts_of
    off = scene.mesh_instances["meshlet_instance_visibility_offset"].astype(np.int64)
    return np.diff(np.concatenate([off, [scene.max_meshlet_instance_count]]))


def partition_mesh_instances(counts, world):  # == oxylus_b200.dist.partition_mesh_instances
    counts = np.asarray(counts, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(counts)])
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(csum, int(csum[-1]) * r / world, side="left"))
        bounds.append(min(max(b, bounds[-1]), len(counts)))
    bounds.append(len(counts))
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def set_alpha_table(ctx):
    """four materials: opaque, RGBA8 checker (linear, repeat), R8 noise (nearest, clamp, albedo alpha 0.8), R8 gradient (mirror)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_alpha import checker, material

    rng = np.random.default_rng(5)
    images = [(checker(16, 2), abi.IMAGE_RGBA8_UNORM), (rng.integers(0, 256, (8, 8), dtype=np.uint8), abi.IMAGE_R8_UNORM),
              (np.ascontiguousarray(np.tile(np.linspace(0, 255, 16).astype(np.uint8), (16, 1))), abi.IMAGE_R8_UNORM)]
    mats = np.array([material(), material(image=0, cutoff=0.5), material(image=1, cutoff=0.3, albedo_a=0.8, sampler=1),
                     material(image=2, cutoff=0.5, sampler=2)], dtype=abi.MATERIAL_DT)
    smp = np.array([abi.sampler(), abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST, u=abi.ADDRESS_CLAMP_TO_EDGE, v=abi.ADDRESS_CLAMP_TO_EDGE),
                    abi.sampler(u=abi.ADDRESS_MIRRORED_REPEAT)], dtype=abi.SAMPLER_DT)
    dev = []
    for tex, fmt in images:
        d = ctx.alloc(tex.size)
        ctx.upload(d, tex)
        dev.append((d, tex.shape[1], tex.shape[0], fmt))
    ctx.set_materials(mats, dev, smp)


class Rank:
    def __init__(self, sc, shard=None, cap=None, alpha=False):
        hw, hh = sc.hiz_extent()
        self.sc, self.w, self.h = sc, sc.width, sc.height
        self.ctx = capi.Context(0, max(1, sc.mesh_instance_count), max(1, sc.max_meshlet_instance_count if cap is None else cap), hw, hh,
                                max_mask_bits=max(1, sc.max_meshlet_instance_count))
        if shard is not None:
            self.ctx.set_shard_auto(shard[0], shard[1])
        self.ctx.set_scene(sc)
        if alpha:
            set_alpha_table(self.ctx)
        self.vis = self.ctx.alloc(self.w * self.h * 8)
        self.occ = self.ctx.alloc(self.w * self.h * 4)
        self.ctx.upload(self.occ, sc.occluder_depth)
        self.mgpu = False

    def frame(self, cam):
        c, w, h, v = self.ctx, self.w, self.h, self.vis
        c.clear_visbuffer_with_depth(v, self.occ, w, h)
        c.clear_hiz()
        c.cull_meshes(cam, abi.CULL_TEST_ALL)
        c.cull_meshlets(cam, abi.CULL_TEST_ALL, True)
        c.raster_visbuffer(cam, abi.CULL_TEST_ALL, w, h, v)
        if self.mgpu:
            c.mgpu_exchange_hiz(v, w, h)
        else:
            c.build_hiz_packed(v, w, h)
        c.cull_meshlets(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, True)
        c.raster_visbuffer(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, w, h, v)

    def image(self):
        return self.ctx.download(self.vis, np.uint64, self.w * self.h)

    def hiz(self):
        return np.concatenate([l.ravel() for l in self.ctx.hiz_levels()]).view(np.uint32)


def main():
    world = int(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    alpha = len(sys.argv) > 3 and sys.argv[3] == "alpha"
    frames = 3
    sc = synth.make_scene(n, config_index=5, width=320, height=180, n_unique_meshes=24)
    if alpha:
        sc.mesh_instances["material_index"] = np.arange(sc.mesh_instance_count) % 4
    lod0 = lod0_counts_of(sc)
    parts = partition_mesh_instances(lod0, world)
    caps = [int(lod0[p[0]: p[0] + p[1]].sum()) for p in parts]
    uid = capi.Context.mgpu_unique_id()
    ref = Rank(sc, alpha=alpha)
    if alpha:  # the table must change the image, or the run would prove nothing about it
        plain = Rank(sc)
        plain.frame(sc.camera(0.0))
        ref.frame(sc.camera(0.0))
        assert not np.array_equal(plain.image(), ref.image())
        ref.ctx.reset_visibility_mask()
    want = []
    for f in range(frames):
        ref.frame(sc.camera(2.0 * (f % 2)))
        vis = ref.ctx.visibility()
        e, l = int(vis["early"][0]), int(vis["late"][0])
        want.append(dict(image=ref.image(), ids=np.sort(ref.ctx.visible_indices(e + l)), total=int(vis["total"][0]), early=e, late=l, hiz=ref.hiz(),
                         mask=np.unpackbits(ref.ctx.mask().view(np.uint8), bitorder="little")))
    results, errors = [None] * world, []
    barrier = threading.Barrier(world)

    def run(rank):
        try:
            r = Rank(sc, shard=parts[rank], cap=max(caps[rank], 1), alpha=alpha)
            info = r.ctx.mgpu_init(rank, world, uid, max(1024, caps[rank]))
            r.mgpu = True
            out = []
            for f in range(frames):
                r.frame(sc.camera(2.0 * (f % 2)))
                r.ctx.mgpu_exchange_frame(r.vis, r.w, r.h, slot=f & 1)
                r.ctx.check_status()
                cnt, ids = r.ctx.mgpu_gathered(f & 1)
                off = sc.mesh_instances["meshlet_instance_visibility_offset"].astype(np.int64)
                lo = int(off[parts[rank][0]]) if parts[rank][1] else 0
                mine = np.unpackbits(r.ctx.mask().view(np.uint8), bitorder="little")[lo: lo + caps[rank]]
                w_ = want[f]
                out.append(dict(image=bool(np.array_equal(r.image(), w_["image"])), ids=bool(np.array_equal(np.sort(np.concatenate(ids)), w_["ids"])),
                                counts=(int(cnt[:, 0].sum()), int(cnt[:, 1].sum()), int(cnt[:, 2].sum())) == (w_["total"], w_["early"], w_["late"]),
                                hiz=bool(np.array_equal(r.hiz(), w_["hiz"])), mask=bool(np.array_equal(mine, w_["mask"][lo: lo + caps[rank]]))))
                barrier.wait()  # nobody starts the next frame's Hi-Z exchange before everybody has read this frame's results
                if f == 0:  # collective resize of the survivor segments to what the frame needed (bench.py does this after warm-up)
                    seen = int(cnt[:, 3].max())
                    info2 = r.ctx.mgpu_set_survivor_capacity(max(64, 2 * seen + rank))  # different requests: the ranks agree on the largest
                    assert info2.survivor_capacity == max(64, 2 * seen + world - 1), (info2.survivor_capacity, seen)
            results[rank] = dict(frames=out, peer_memory=bool(info.hiz_over_peer_memory), local_total=int(r.ctx.visibility()["total"][0]))
            barrier.wait()
            r.ctx.mgpu_shutdown()
        except Exception as ex:  # noqa: BLE001
            errors.append(f"rank {rank}: {ex!r}")
            barrier.abort()

    threads = [threading.Thread(target=run, args=(k,)) for k in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    ok = not errors and all(r and all(all(f.values()) for f in r["frames"]) for r in results)
    ok = ok and all(r["peer_memory"] for r in results) and sum(r["local_total"] > 0 for r in results) >= min(world, 2)
    print(json.dumps({"check": "emulated_ranks_equal_single_context", "world": world, "meshlets": n, "alpha": alpha, "pass": bool(ok), "errors": errors, "ranks": results}))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
