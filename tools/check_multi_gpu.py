#!/usr/bin/env python
"""Multi-GPU parity check on real GPUs (run under torchrun, N >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multi_gpu.py

Every rank culls + rasterises its mesh-instance shard (CUDA kernels through the C ABI), with the real exchange
steps (id-base allgather, Hi-Z mip-0 max-reduce, vis-buffer max-reduce, survivor allgather).  Rank 0 then runs
the SAME scene on one GPU and asserts bit-identical results: packed vis buffer, sorted survivor ids, per-rank
mask slices, Hi-Z pyramid.  Correctness rule of SURVEY.md §8e: G GPUs == 1 GPU, bit for bit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oxylus_b200 import abi, capi, dist as oxdist, pipeline, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sc = synth.make_scene(300_000, config_index=5, width=1280, height=720, n_unique_meshes=64)
    lod0 = oxdist.lod0_counts_of(sc)
    parts = oxdist.partition_mesh_instances(lod0, world)
    cap = max(int(lod0[f:f + c].sum()) for f, c in parts)
    auto = os.environ.get("OXC_AUTO_ID_BASE", "1") == "1"
    pipe = pipeline.VisibilityPipeline(sc, device=local, shard=parts[rank], auto_id_base=auto)
    out = pipe.ctx.out
    vis_view = pipeline.wrap_device(out.visibility, 3, torch.int32, dev)
    ids_view = pipeline.wrap_device(out.visible_meshlet_instances_indices, cap, torch.int32, dev)
    hw, hh = sc.hiz_extent()
    mip0_view = pipeline.wrap_device(out.hiz, hw * hh, torch.int32, dev)
    vis_all = torch.zeros(world * 3, dtype=torch.int32, device=dev)
    ids_all = torch.zeros(world * cap, dtype=torch.int32, device=dev)

    def after_cull_meshes():
        dist.all_gather_into_tensor(vis_all, vis_view)
        pipe.id_base.copy_(vis_all.view(world, 3)[:rank, 0].sum())

    def between_passes():
        dist.all_reduce(mip0_view, op=dist.ReduceOp.MAX)

    def after_frame():
        oxdist.reduce_visbuffer(pipe.vis64)
        dist.all_gather_into_tensor(vis_all, vis_view)
        dist.all_gather_into_tensor(ids_all, ids_view)

    single = pipeline.VisibilityPipeline(sc, device=local) if rank == 0 else None
    ok = True
    for f in range(4):
        cam = sc.camera(2.0 * f)
        pipe.frame(cam, after_cull_meshes=None if auto else after_cull_meshes, between_passes=between_passes, after_frame=after_frame)
        torch.cuda.synchronize()
        counts = vis_all.view(world, 3).cpu().numpy()
        ids = ids_all.view(world, cap).cpu().numpy()
        merged = np.concatenate([ids[r, : counts[r, 1] + counts[r, 2]] for r in range(world)]).astype(np.uint32)
        img = pipe.vis64.cpu().numpy().view(np.uint64)
        masks = [None] * world
        dist.all_gather_object(masks, pipe.ctx.mask())
        hiz = np.concatenate([l.reshape(-1) for l in pipe.ctx.hiz_levels()])
        if rank == 0:
            single.frame(cam)
            torch.cuda.synchronize()
            c1 = single.counters()
            ref_ids = single.ctx.visible_indices(c1["early"] + c1["late"])
            ref_img = single.vis64.cpu().numpy().view(np.uint64)
            ref_hiz = np.concatenate([l.reshape(-1) for l in single.ctx.hiz_levels()])
            ref_mask = single.ctx.mask()
            # union of the ranks' owned mask bit ranges
            off = sc.mesh_instances["meshlet_instance_visibility_offset"].astype(np.int64)
            total_bits = sc.max_meshlet_instance_count
            merged_bits = np.zeros(len(ref_mask) * 32, dtype=np.uint8)
            for r, (first, count) in enumerate(parts):
                lo = int(off[first]) if count else total_bits
                hi = int(off[first + count]) if first + count < len(off) else total_bits
                bits = np.unpackbits(masks[r].view(np.uint8), bitorder="little")
                merged_bits[lo:hi] = bits[lo:hi]
            merged_mask = np.packbits(merged_bits, bitorder="little").view(np.uint32)
            checks = dict(
                totals=int(counts[:, 0].sum()) == c1["total"],
                early=int(counts[:, 1].sum()) == c1["early"], late=int(counts[:, 2].sum()) == c1["late"],
                survivors=np.array_equal(np.sort(merged), np.sort(ref_ids)),
                visbuffer=np.array_equal(img, ref_img), hiz=np.array_equal(hiz.view(np.uint32), ref_hiz.view(np.uint32)),
                mask=np.array_equal(merged_mask, ref_mask))
            print(f"frame {f}: world={world} survivors={len(ref_ids)} " + " ".join(f"{k}={'OK' if v else 'MISMATCH'}" for k, v in checks.items()),
                  flush=True)
            ok = ok and all(checks.values())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI-GPU PARITY: " + ("PASS" if ok else "FAIL"), flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
