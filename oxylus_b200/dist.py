"""Host-side sharding logic of the multi-GPU path (the reference is single-GPU).

The exchange itself is product code behind the C ABI (oxc_mgpu_*: Hi-Z over NVLink peer memory, vis-buffer max-reduce and
survivor allgather over NCCL — csrc/kernels_mgpu.cuh, oxcull.cu).  What lives here is what a host decides before it creates
its contexts — the contiguous mesh-instance ranges, balanced by LOD0 meshlet count, rank r owns (mask bits are laid out by
mesh-instance order, Scene.cpp:1255-1260, so a rank's slice of the persistent mask is private) — plus torch.distributed
restatements of the three exchange steps that run on CPU tensors under gloo, so tests/test_dist_cpu.py can check the
sharding algebra (global id bases, max-merge of the packed image, survivor segments) against the single-process oracle
frame without a GPU:
  1. all_gather of each rank's emitted meshlet-instance count -> exclusive prefix = this rank's id base
  2. all_reduce(MAX) of the packed depth|id image (depth bits of z in [0,1] are <= 0x3F800000, so the signed int64 order
     equals the unsigned order of the packing)
  3. all_gather of the survivor lists (count-prefixed, fixed-capacity segments)
"""
import numpy as np
import torch
import torch.distributed as dist


def partition_mesh_instances(lod0_meshlet_counts, world_size):
    """Contiguous mesh-instance ranges with ~equal LOD0 meshlet totals.  Returns [(first, count)] * world_size."""
    counts = np.asarray(lod0_meshlet_counts, dtype=np.int64)
    n = len(counts)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world_size)]


def lod0_counts_of(scene):
    """LOD0 meshlet count per mesh instance, recovered from the visibility offsets (Scene.cpp:1255-1260)."""
    off = scene.mesh_instances["meshlet_instance_visibility_offset"].astype(np.int64)
    return np.diff(np.concatenate([off, [scene.max_meshlet_instance_count]]))


def exchange_id_base(local_total: torch.Tensor, id_base_out: torch.Tensor, group=None):
    """local_total: int32[1] (this rank's visibility.total).  Writes this rank's global id base into id_base_out[0]
    and returns the gathered totals (int32[world])."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = torch.empty(world, dtype=local_total.dtype, device=local_total.device)
    dist.all_gather_into_tensor(gathered, local_total, group=group)
    prefix = torch.cumsum(gathered, 0) - gathered
    id_base_out.copy_(prefix[rank:rank + 1])
    return gathered


def reduce_visbuffer(vis64: torch.Tensor, group=None):
    """Per-pixel max of the packed depth|id image over ranks (reverse-Z: nearer = larger)."""
    dist.all_reduce(vis64, op=dist.ReduceOp.MAX, group=group)
    return vis64


def gather_survivors(local_ids: torch.Tensor, local_count: torch.Tensor, group=None):
    """local_ids: int32[capacity] (first local_count[0] valid, already global ids).  Returns (ids[world, capacity], counts[world])."""
    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=local_count.dtype, device=local_count.device)
    dist.all_gather_into_tensor(counts, local_count, group=group)
    ids = torch.empty(world * local_ids.numel(), dtype=local_ids.dtype, device=local_ids.device)
    dist.all_gather_into_tensor(ids, local_ids.reshape(-1), group=group)
    return ids.view(world, local_ids.numel()), counts


def merge_survivors(ids: torch.Tensor, counts: torch.Tensor):
    """Concatenate the valid prefix of every rank's segment (host side; used by tests / readback)."""
    ids = ids.cpu().numpy()
    counts = counts.cpu().numpy()
    return np.concatenate([ids[r, : int(counts[r])] for r in range(len(counts))]) if len(counts) else np.zeros(0, np.int32)
