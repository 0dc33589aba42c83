"""world_size-2 (and 3) gloo test of the multi-GPU host logic on CPU: mesh-instance sharding,
id-base exchange, vis-buffer max-reduce between the passes, survivor allgather.  Each rank runs the ORACLE on its
shard (the CUDA kernels need a GPU); the merged result must equal the single-process oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pyoracle as orc
    from oxylus_b200 import dist as oxdist, synth

    sc = synth.make_scene(12000, config_index=5, width=640, height=360, n_unique_meshes=16)
    parts = oxdist.partition_mesh_instances(oxdist.lod0_counts_of(sc), world)
    first, count = parts[rank]
    hs = orc.HostScene(sc)
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    cap = max(int(oxdist.lod0_counts_of(sc)[f:f + c].sum()) for f, c in parts)
    out = []
    for f in range(3):
        cam = sc.camera(2.0 * f)
        id_base_t = torch.zeros(1, dtype=torch.int32)

        def id_base_fn(total):
            oxdist.exchange_id_base(torch.tensor([total], dtype=torch.int32), id_base_t)
            return int(id_base_t.item())

        def reduce(img):
            t = torch.from_numpy(img.view(np.int64))
            oxdist.reduce_visbuffer(t)

        r = orc.frame(hs, cam, sc.width, sc.height, mask, sc.occluder_depth, first, count, id_base_fn, reduce, reduce)
        n = r["early"] + r["late"]
        ids = np.zeros(cap, dtype=np.int32)
        ids[:n] = r["visible"][:n].astype(np.int64) + r["id_base"]
        g_ids, g_counts = oxdist.gather_survivors(torch.from_numpy(ids), torch.tensor([n], dtype=torch.int32))
        merged = oxdist.merge_survivors(g_ids, g_counts)
        out.append(dict(vis64=r["vis64"].copy(), survivors=np.sort(merged), early=r["early"], late=r["late"], id_base=r["id_base"]))
    # this rank's mask slice: bits of its own mesh instances
    q.put((rank, out, mask, parts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_frames_equal_single_process(orc, world):
    from oxylus_b200 import dist as oxdist, synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, out, mask, parts = q.get(timeout=300)
        results[rank] = (out, mask, parts)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    sc = synth.make_scene(12000, config_index=5, width=640, height=360, n_unique_meshes=16)
    hs = orc.HostScene(sc)
    mask_ref = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    parts = results[0][2]
    offs = sc.mesh_instances["meshlet_instance_visibility_offset"]
    for f in range(3):
        ref = orc.frame(hs, sc.camera(2.0 * f), sc.width, sc.height, mask_ref, sc.occluder_depth)
        n = ref["early"] + ref["late"]
        assert sum(results[r][0][f]["early"] for r in range(world)) == ref["early"]
        assert sum(results[r][0][f]["late"] for r in range(world)) == ref["late"]
        for r in range(world):
            got = results[r][0][f]
            np.testing.assert_array_equal(got["vis64"], ref["vis64"])             # identical on every rank after the reduce
            np.testing.assert_array_equal(got["survivors"], np.sort(ref["visible"][:n]).astype(np.int32))
    # the union of the ranks' owned mask bit-ranges equals the single-process mask
    merged = np.zeros_like(mask_ref)
    bits_total = sc.max_meshlet_instance_count
    for r in range(world):
        first, count = parts[r]
        lo = int(offs[first]) if count else bits_total
        hi = int(offs[first + count]) if first + count < len(offs) else bits_total
        bits = np.unpackbits(results[r][1].view(np.uint8), bitorder="little")
        sel = np.zeros_like(bits)
        sel[lo:hi] = bits[lo:hi]
        merged |= np.packbits(sel, bitorder="little").view(np.uint32)
    np.testing.assert_array_equal(merged, mask_ref)


def test_partition_balances_and_covers():
    from oxylus_b200 import dist as oxdist

    rng = np.random.default_rng(0)
    counts = rng.integers(64, 257, size=1000)
    for world in (1, 2, 4, 8):
        parts = oxdist.partition_mesh_instances(counts, world)
        assert parts[0][0] == 0 and sum(c for _, c in parts) == len(counts)
        for (f0, c0), (f1, _) in zip(parts[:-1], parts[1:]):
            assert f0 + c0 == f1
        loads = [counts[f:f + c].sum() for f, c in parts]
        assert max(loads) - min(loads) <= 2 * 256
    assert oxdist.partition_mesh_instances([5], 4)[-1] == (1, 0) or sum(c for _, c in oxdist.partition_mesh_instances([5], 4)) == 1
