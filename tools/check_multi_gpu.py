#!/usr/bin/env python
"""G GPUs == 1 GPU, bit for bit, through the product's multi-GPU API (oxc_mgpu_*):
    torchrun --standalone --nproc-per-node G tools/check_multi_gpu.py
Every rank runs the sharded pipeline (oxc_set_shard_auto + oxc_mgpu_exchange_hiz / _frame); rank 0 also runs the whole scene
on one GPU and compares: merged vis buffer, gathered survivor ids (as a set), the Hi-Z pyramid of every rank, the visibility
mask (each rank's slice).  torch.distributed only broadcasts the 128-byte communicator id and gathers the verdicts."""
import json
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
    dist.init_process_group("gloo")  # plumbing only: the data path is the product's own NCCL communicator + peer memory
    n = int(os.environ.get("OXC_CHECK_MESHLETS", 300_000))
    sc = synth.make_scene(n, config_index=5, width=1280, height=720, n_unique_meshes=64)
    lod0 = oxdist.lod0_counts_of(sc)
    parts = oxdist.partition_mesh_instances(lod0, world)
    cap = int(lod0[parts[rank][0]: parts[rank][0] + parts[rank][1]].sum())
    uid = [capi.Context.mgpu_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    pipe = pipeline.VisibilityPipeline(sc, device=local, shard=parts[rank], auto_id_base=True, shard_capacity=max(cap, 1),
                                       mgpu=dict(rank=rank, world=world, unique_id=uid[0], survivor_capacity=max(1024, cap)))
    ref = pipeline.VisibilityPipeline(sc, device=local) if rank == 0 else None
    ok, detail = True, {}
    torch.cuda.synchronize()
    dist.barrier()
    dbg = (lambda *a: print(f"[check r{rank}]", *a, file=sys.stderr, flush=True)) if os.environ.get("OXC_MGPU_DEBUG") else (lambda *a: None)
    for f in range(4):
        cam = sc.camera(2.0 * (f % 2))
        dbg("frame", f, "start")
        pipe.frame(cam)
        torch.cuda.synchronize()
        dbg("frame", f, "kernels + hiz exchange done; status", pipe.ctx.status_flags())
        pipe.exchange_frame(slot=f & 1)
        torch.cuda.synchronize()
        dbg("frame", f, "exchange_frame done")
        pipe.ctx.check_status()
        cnt, ids = pipe.ctx.mgpu_gathered(f & 1)
        hiz = np.concatenate([l.ravel() for l in pipe.ctx.hiz_levels()]).view(np.uint32)
        if rank == 0:
            ref.frame(cam)
            torch.cuda.synchronize()
            rc = ref.counters()
            r_ids = ref.ctx.visible_indices(rc["early"] + rc["late"])
            same_img = bool(torch.equal(pipe.vis64, ref.vis64))
            same_ids = bool(np.array_equal(np.sort(np.concatenate(ids)), np.sort(r_ids)))
            same_cnt = int(cnt[:, 0].sum()) == rc["total"] and int(cnt[:, 1].sum()) == rc["early"] and int(cnt[:, 2].sum()) == rc["late"]
            ref_hiz = np.concatenate([l.ravel() for l in ref.ctx.hiz_levels()]).view(np.uint32)
            ref_mask = ref.ctx.mask()
            detail[f] = dict(image=same_img, ids=same_ids, counts=same_cnt)
            ok = ok and same_img and same_ids and same_cnt
        else:
            ref_hiz, ref_mask = None, None
        box = [ref_hiz, ref_mask]
        dist.broadcast_object_list(box, src=0)
        same_hiz = bool(np.array_equal(hiz, box[0]))
        # this rank's slice of the persistent mask (bits are laid out in mesh-instance order)
        off = sc.mesh_instances["meshlet_instance_visibility_offset"].astype(np.int64)
        lo = int(off[parts[rank][0]]) if parts[rank][1] else 0
        hi = lo + cap
        mine = np.unpackbits(pipe.ctx.mask().view(np.uint8), bitorder="little")[lo:hi]
        want = np.unpackbits(box[1].view(np.uint8), bitorder="little")[lo:hi]
        verdict = [same_hiz and bool(np.array_equal(mine, want))]
        gathered = [None] * world
        dist.all_gather_object(gathered, verdict[0])
        if rank == 0:
            detail[f]["hiz_and_mask_per_rank"] = gathered
            ok = ok and all(gathered)
    info = pipe.mgpu
    if rank == 0:
        print(json.dumps({"check": "multi_gpu_equals_single_gpu", "world": world, "meshlets": n, "pass": bool(ok),
                          "hiz_over_peer_memory": bool(info.hiz_over_peer_memory), "frames": detail}), flush=True)
        ref.close()
    pipe.close()
    dist.destroy_process_group()
    sys.exit(0 if ok or rank != 0 else 1)


if __name__ == "__main__":
    main()
