#!/usr/bin/env python
"""bench.py — meshlet visibility pipeline on B200 (BASELINE.json metric: meshlets culled/s + tris rasterised/s,
% of HBM roofline) with the CPU reference arm beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one frame of the hot path over the synthetic scene of BASELINE.json configs[1]
("1M meshlet instances, 1 camera, two-pass Hi-Z occlusion cull"): clear attachments -> cull_meshes ->
cull_meshlets(early) -> vis-buffer raster -> generate_hiz -> cull_meshlets(late) -> vis-buffer raster
(RendererInstance.cpp:842-884).  N > 1: weak scaling, 1M meshlet instances PER GPU, mesh-instance sharded,
with the real exchange steps (id-base allgather, vis-buffer max-reduce x2, survivor allgather) inside the step.

value   : whole-job meshlet instances culled / s with inputs resident in HBM (CUDA-graph replay of one frame,
          CUDA events per step on the launching stream, L2 flushed between steps, max over ranks)
e2e     : same metric through the reference-facing host API oxr_submit / oxr_wait (C++ RendererInstance mirror) with
          HOST buffers: camera + all transforms H2D from pinned memory, vis32 + survivor ids + counters D2H every frame
roofline: the late meshlet-cull kernel (cull_meshlets_hiz equivalent), algorithmic bytes of SURVEY.md §8d /
          its CUDA-event duration inside the timed loop, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference: the oracle port of the same frame (oracle/, pthreads on all host cores) on a
          bounded sample — the reference's own Vulkan path cannot be built or run here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "meshlet_instances_culled_per_s"
UNIT = "meshlet instances/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--meshlets", type=int, default=1_000_000, help="meshlet instances per GPU")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly instead of replaying a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: run the trailing exchange on the main stream (no overlap with the next frame)")
    ap.add_argument("--cpu-frames", type=int, default=2)
    return ap.parse_args()


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples, self.proc, self.thread = [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle

    pyoracle.lib()
    return pyoracle


def cpu_frames(scene, n_frames, n_threads):
    """The oracle port of the frame on the host cores (cpu_baseline / reference arm).  Returns (seconds per frame, info)."""
    orc = load_oracle()
    hs = orc.HostScene(scene)
    mask = np.zeros((scene.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    # untimed frames bring the persistent mask to the same steady state the GPU arm is timed in
    for f in range(4):
        orc.cpu_frame(hs, scene.camera(2.0 * (f % 2)), scene.width, scene.height, mask, scene.occluder_depth, n_threads)
    times, last = [], None
    for f in range(n_frames):
        cam = scene.camera(2.0 * (f % 2))
        t0 = time.perf_counter()
        last = orc.cpu_frame(hs, cam, scene.width, scene.height, mask, scene.occluder_depth, n_threads)
        times.append(time.perf_counter() - t0)
    return float(np.mean(times)), last


def cpu_frustum_loops(scene, cam, cores):
    """the reference's CPU primitives looped over the meshlet bounds + draw-list build
    (engine AABB::is_on_frustum test, and the shader-equivalent cone+frustum), 1 thread and all host threads;
    plus BASELINE.json configs[0] (10 k bounds, 1 camera, scalar)."""
    orc = load_oracle()
    hs = orc.HostScene(scene)
    mi, vis, _ = orc.cull_meshes(hs, cam, 7)
    total = int(vis["total"][0])
    out = {}
    for name, mode in (("engine_aabb_is_on_frustum", 0), ("shader_equivalent_cone_frustum", 1)):
        for threads in (1, cores):
            n = total if threads > 1 else min(total, 200_000)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                surv = orc.cpu_baseline_cull(hs, mi, n, cam, mode, threads)
                best = min(best, time.perf_counter() - t0)
            out[f"{name}_{threads}t"] = {"meshlets_per_s": n / best, "meshlets": n, "survivors": int(len(surv))}
    n = min(total, 10_000)
    best = 1e9
    for _ in range(20):
        t0 = time.perf_counter()
        orc.cpu_baseline_cull(hs, mi, n, cam, 1, 1)
        best = min(best, time.perf_counter() - t0)
    out["configs0_10k_bounds_scalar_1t"] = {"meshlets_per_s": n / best, "meshlets": n}
    # configs[0], first half: "meshlet build of one mesh" — the host-side builder (oxb_build_mesh: fetch remap, quantisation,
    # scan meshlets, AABBs + normal cones, blob) on a procedural 131k-triangle torus, one thread.  Host C++ of the
    # product, no GPU involved; meshoptimizer (the reference's builder) is not available here, so this is not a comparison.
    try:
        from oxylus_b200 import capi as _capi
        nu, nv = 512, 128
        uu, vv = np.meshgrid(np.arange(nu) / nu * 2 * np.pi, np.arange(nv) / nv * 2 * np.pi, indexing="ij")
        pos = np.stack([(2 + 0.7 * np.cos(vv)) * np.cos(uu), (2 + 0.7 * np.cos(vv)) * np.sin(uu), 0.7 * np.sin(vv)], axis=2).reshape(-1, 3)
        nrm = np.stack([np.cos(vv) * np.cos(uu), np.cos(vv) * np.sin(uu), np.sin(vv)], axis=2).reshape(-1, 3)
        i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
        a, b = i * nv + j, ((i + 1) % nu) * nv + j
        c, d = i * nv + (j + 1) % nv, ((i + 1) % nu) * nv + (j + 1) % nv
        idx = np.stack([a, b, d, a, d, c], axis=2).reshape(-1).astype(np.uint32)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            bm = _capi.BuiltMesh(pos, [(idx, 0.0)], normals=nrm)
            best = min(best, time.perf_counter() - t0)
        out["configs0_meshlet_build_1t"] = {"triangles_per_s": len(idx) / 3 / best, "triangles": int(len(idx) // 3),
                                            "meshlets": bm.lod0_meshlet_count, "ms": best * 1e3}
        bm.close()
    except Exception as e:  # the builder lives in liboxcull.so; never let a baseline extra break the bench line
        out["configs0_meshlet_build_1t"] = {"error": str(e)[:200]}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oxylus_b200 import synth

    cores = os.cpu_count() or 1
    scene = synth.make_scene(args.meshlets * max(1, args.gpus), config_index=2, width=args.width, height=args.height)
    # bounded sample: every step is one full frame of the same scene on all host threads
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_frames(scene, 1, cores)
    n = max(1, min(args.steps, 4))
    sec, last = cpu_frames(scene, n, cores)
    value = scene.max_meshlet_instance_count / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, 1),
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, scene),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} full frames of the {scene.max_meshlet_instance_count}-meshlet scene, oracle port, pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "triangles_per_s": (last["triangles"] / sec) if last else None,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, scene):
    return {"workload": "BASELINE.json configs[1]: 1M meshlet instances per GPU, 1 camera, two-pass Hi-Z occlusion cull + vis-buffer raster",
            "meshlet_instances_per_gpu": args.meshlets, "resolution": [args.width, args.height],
            "hiz": list(scene.hiz_extent()), "mesh_instances": scene.mesh_instance_count, "unique_meshes": len(scene.meshes),
            "l2": "flushed between timed steps (256 MiB write, untimed)", "cameras": "yaw 0 / 2 deg alternating, steady-state mask",
            "parallelism": f"mesh-instance shards x{args.gpus}" if args.gpus > 1 else "single GPU"}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from oxylus_b200 import abi, capi, dist as oxdist, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if multi:
        # NCCL_DEBUG is left exactly as the launcher set it (the driver reads the communicator banner for its rank proof);
        # the JSON line is the last line rank 0 prints
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    capi.load(build_if_missing=False)

    # ---------------- scene ----------------
    total_meshlets = args.meshlets * n_gpus
    scene = synth.make_scene(total_meshlets, config_index=2, width=args.width, height=args.height)
    shard = None
    if multi:
        parts = oxdist.partition_mesh_instances(oxdist.lod0_counts_of(scene), world)
        shard = parts[rank]
    pipe = pipeline.VisibilityPipeline(scene, device=local_rank, shard=shard, auto_id_base=True)
    cams = [scene.camera(0.0), scene.camera(2.0)]
    dev = pipe.device
    w, h = scene.width, scene.height

    # multi-GPU exchange state
    hooks, gathered = {}, {}
    if multi:
        lod0 = oxdist.lod0_counts_of(scene)
        cap = max(int(lod0[f:f + c].sum()) for f, c in parts)
        out = pipe.ctx.out
        # zero-copy torch views of the context's device buffers (no staging copies in the exchange)
        vis_view = pipeline.wrap_device(out.visibility, 3, torch.int32, dev)            # total / early / late
        # survivor allgather: fixed-capacity segments (NCCL has no allgatherv).  Half a shard is ample for this scene
        # (~30 % visible); an overflow is detected from the gathered counts after the run and reported.
        gcap = max(1024, cap // 2)
        ids_view = pipeline.wrap_device(out.visible_meshlet_instances_indices, gcap, torch.int32, dev)
        vis_all = torch.zeros(world * 3, dtype=torch.int32, device=dev)
        ids_all = torch.zeros(world * gcap, dtype=torch.int32, device=dev)

        after_cull_meshes = None  # global id base: computed locally by oxc_cull_meshes (oxc_set_shard_auto), no exchange

        hw_, hh_ = scene.hiz_extent()
        mip0_view = pipeline.wrap_device(out.hiz, hw_ * hh_, torch.int32, dev)  # level 0 starts at offset 0

        def between_passes():
            dist.all_reduce(mip0_view, op=dist.ReduceOp.MAX)  # depths are >= +0: int32 order == float order

        def after_frame():
            oxdist.reduce_visbuffer(pipe.vis64)               # per-pixel max of the packed depth|id image (NVLS all-reduce)
            dist.all_gather_into_tensor(vis_all, vis_view)   # early / late counts of every rank
            dist.all_gather_into_tensor(ids_all, ids_view)   # survivor ids (global), fixed-capacity segments
            gathered["last"] = (ids_all, vis_all)

        hooks = dict(after_cull_meshes=after_cull_meshes, between_passes=between_passes, after_frame=after_frame)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up (also brings the visibility mask to steady state) ----------------
    W = max(4, args.warmup)  # >= 4 so the persistent visibility mask reaches its steady state
    K = max(1, args.steps)
    sampler = ClockSampler(local_rank)  # samples span warm-up + timed region + per-kernel loop (all GPU-busy)
    sampler.start()
    for i in range(W):
        pipe.frame(cams[i % 2], **hooks)
    torch.cuda.synchronize()

    # ---------------- CUDA graphs of one frame per camera ----------------
    graphs = None
    if not args.no_graph and not multi:  # N > 1 launches eagerly: capturing the NCCL exchange steps hung in testing
        try:
            graphs = []
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                pipe.use_torch_stream()
                for cam in cams:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        pipe.use_torch_stream()
                        pipe.frame(cam, **hooks)  # N > 1: the NCCL exchange steps are captured with the kernels
                    graphs.append(g)
            torch.cuda.current_stream().wait_stream(side)
            pipe.use_torch_stream()
            for i in range(2):
                graphs[i % 2].replay()
            torch.cuda.synchronize()
        except Exception as e:  # capture of the collectives unsupported: launch eagerly (still correct, more launch gaps)
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e!r}); timing eager launches\n")
            graphs = None
            pipe.use_torch_stream()
            torch.cuda.synchronize()

    # N > 1: the two halves of the frame around the Hi-Z exchange are captured as CUDA graphs; the NCCL collectives are
    # launched between / after them (capturing the collectives themselves hung in testing)
    half_graphs = None
    if multi and not args.no_graph:
        try:
            half_graphs = []
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for b, cam in enumerate(cams):
                    pipe.select_buffer(b)  # camera index == buffer index: both alternate every step
                    pair = []
                    for part in (pipe.frame_before_exchange, pipe.frame_after_exchange):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=side):
                            pipe.use_torch_stream()
                            part(cam)
                        pair.append(g)
                    half_graphs.append(pair)
            torch.cuda.current_stream().wait_stream(side)
            pipe.use_torch_stream()
            pipe.select_buffer(0)
            torch.cuda.synchronize()
        except Exception as e:
            sys.stderr.write(f"[bench] half-frame graph capture failed ({e!r}); timing eager launches\n")
            half_graphs = None
            pipe.use_torch_stream()
            pipe.select_buffer(0)
            torch.cuda.synchronize()

    # N > 1: trailing exchange (vis-buffer max-reduce, survivor allgather) on a side stream with its own communicator and
    # double-buffered vis / survivor staging, so it overlaps the next frame; buffer b is reused two frames later, after
    # its exchange has completed (event wait on the main stream)
    overlap = None
    if multi and not args.no_overlap:
        overlap = dict(pg=dist.new_group(), cs=torch.cuda.Stream(),
                       ids_stage=[torch.zeros(gcap, dtype=torch.int32, device=dev) for _ in range(2)],
                       cnt_stage=[torch.zeros(3, dtype=torch.int32, device=dev) for _ in range(2)],
                       ids_all=[torch.zeros(world * gcap, dtype=torch.int32, device=dev) for _ in range(2)],
                       cnt_all=[torch.zeros(world * 3, dtype=torch.int32, device=dev) for _ in range(2)],
                       frame_done=[torch.cuda.Event() for _ in range(2)], tail_done=[torch.cuda.Event() for _ in range(2)],
                       pending=[False, False])

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_overlapped(i):
        b = i & 1
        o = overlap
        main = torch.cuda.current_stream()
        if o["pending"][b]:
            main.wait_event(o["tail_done"][b])  # the exchange that last used buffer b (two frames ago) has finished
        pipe.select_buffer(b)
        if half_graphs is not None:
            half_graphs[b][0].replay()
            hooks["between_passes"]()
            half_graphs[b][1].replay()
        else:
            pipe.frame(cams[b], between_passes=hooks["between_passes"])
        o["ids_stage"][b].copy_(ids_view)   # the context's survivor list / counters are overwritten by the next frame
        o["cnt_stage"][b].copy_(vis_view)
        o["frame_done"][b].record(main)
        o["cs"].wait_event(o["frame_done"][b])
        with torch.cuda.stream(o["cs"]):
            dist.all_reduce(pipe.vis64_bufs[b], op=dist.ReduceOp.MAX, group=o["pg"])
            dist.all_gather_into_tensor(o["cnt_all"][b], o["cnt_stage"][b], group=o["pg"])
            dist.all_gather_into_tensor(o["ids_all"][b], o["ids_stage"][b], group=o["pg"])
            o["tail_done"][b].record(o["cs"])
        o["pending"][b] = True

    def step(i, mark=None):
        if graphs is not None and mark is None:
            graphs[i % 2].replay()
        elif overlap is not None and mark is None:
            step_overlapped(i)
        elif half_graphs is not None and mark is None:
            pipe.select_buffer(i & 1)
            half_graphs[i % 2][0].replay()
            hooks["between_passes"]()
            half_graphs[i % 2][1].replay()
            hooks["after_frame"]()
        else:
            pipe.select_buffer(0)
            pipe.frame(cams[i % 2], mark=mark, **hooks)

    if overlap is not None:  # warm the second communicator / side stream outside the timed region
        for i in range(4):
            step_overlapped(i)
        torch.cuda.current_stream().wait_stream(overlap["cs"])
        torch.cuda.synchronize()

    # ---------------- timed region: exactly K steps ----------------
    launches0 = capi.kernel_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xFF)  # L2 flush, outside the per-step event pair
        ev[i][0].record()
        step(i)
        if overlap is not None and i == K - 1:
            torch.cuda.current_stream().wait_stream(overlap["cs"])  # the last steps' exchanges are inside the timed region
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms_per_step = float(np.mean(step_ms))
    if graphs is not None:
        launches_per_step = None  # replayed nodes are counted below from an eager frame
    launches_timed = capi.kernel_launch_count() - launches0
    cnt = pipe.counters()

    # ---------------- per-kernel durations (same steps, eager launches, one CUDA event after every stage) ----------------
    stage_names = ["begin"] + pipeline.STAGES
    stage_acc = {n: [] for n in pipeline.STAGES}
    stage_acc["clear"] = []
    l0 = capi.kernel_launch_count()
    for i in range(K):
        flush.fill_(i & 0xFF)
        marks = {}

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks[name] = e

        s0 = torch.cuda.Event(enable_timing=True)
        s0.record()
        step(i, mark=mark)
        torch.cuda.synchronize()
        stage_acc["clear"].append(s0.elapsed_time(marks["begin"]))
        for a, b in zip(stage_names[:-1], stage_names[1:]):
            stage_acc[b].append(marks[a].elapsed_time(marks[b]))
    launches_per_frame = (capi.kernel_launch_count() - l0) // K
    # keep the GPU busy with the same steps until the sampler has a few readings, then stop it
    t_busy = time.perf_counter()
    while len(sampler.samples) < 5 and time.perf_counter() - t_busy < 2.0:
        step(0)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    stages_ms = {k: float(np.mean(v)) for k, v in stage_acc.items()}
    cnt_late = pipe.counters()

    stages_all = None
    if multi:
        keys = sorted(stages_ms)
        t_st = torch.tensor([stages_ms[k] for k in keys] + [cnt_late["early"] + cnt_late["late"], cnt_late["triangles"]], dtype=torch.float64, device=dev)
        g_st = torch.empty(world * len(t_st), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(g_st, t_st)
        g_st = g_st.view(world, -1).cpu().numpy()
        stages_all = {k: [round(float(g_st[r, i]), 4) for r in range(world)] for i, k in enumerate(keys)}
        stages_all["survivors"] = [int(g_st[r, len(keys)]) for r in range(world)]
        stages_all["triangles"] = [int(g_st[r, len(keys) + 1]) for r in range(world)]
    # max over ranks
    if multi:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
        tot = torch.tensor([cnt["total"], cnt["triangles"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        job_meshlets, job_tris = float(tot[0].item()), float(tot[1].item())
    else:
        job_meshlets, job_tris = float(cnt["total"]), float(cnt["triangles"])
    value = job_meshlets / (ms_per_step * 1e-3)

    # ---------------- roofline of the dominant cull kernel (late pass: every meshlet instance fully tested) ----------------
    peak, peak_src = measured_peak_hbm()
    N_local = cnt_late["total"]
    I_local = shard[1] if shard else scene.mesh_instance_count
    M_bits = scene.max_meshlet_instance_count
    S_late = cnt_late["late"]
    algo_bytes = N_local * 24 + 2 * 4 * ((M_bits + 31) // 32) + 4 * S_late + I_local * 84 + len(scene.meshes) * 128
    t_late = stages_ms["cull_late"] * 1e-3
    achieved = algo_bytes / t_late / 1e9 if t_late > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_cull_late_summary.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_cull_meshlets<HIZ,OCC,LATE> (late pass)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": stages_ms["cull_late"],
                "meshlets_per_s_kernel": N_local / t_late if t_late > 0 else None,
                "note": "algorithmic bytes = N*24 + 8*ceil(M/32) + 4*S + I*84 + U*128 (SURVEY 8d, Hi-Z bytes excluded); "
                        "256 unique meshes => bounds are L2-resident and the kernel is issue-bound (DESIGN.md)"}

    # the other kernels of the step against the same HBM roofline (SURVEY 8d byte formulas; 64 triangles / 49 vertices
    # per meshlet in the synthetic meshes): the step is dominated by the raster, which is instruction-bound like the cull
    def _k(name, algo, ms):
        return {"kernel": name, "algorithmic_bytes": int(algo), "ms": ms, "achieved_gbs": algo / (ms * 1e-3) / 1e9 if ms > 0 else None,
                "frac": algo / (ms * 1e-3) / 1e9 / peak if ms > 0 else None}
    per_meshlet = 16 + 3 * 64 + 4 * 49 + 8 * 49          # Meshlet + micro indices + vertex indices + positions
    hw_, hh_ = scene.hiz_extent()
    roofline["other_kernels"] = [] if any(k not in stages_ms for k in ("cull_early", "raster_early", "hiz", "raster_late")) else [
        _k("k_cull_meshlets<HIZ,OCC,EARLY,ZERO>", N_local * 24 + 2 * 4 * ((M_bits + 31) // 32) + 4 * cnt_late["early"] + I_local * 84, stages_ms["cull_early"]),
        _k("k_raster_visbuffer (early)", cnt_late["early"] * per_meshlet + 8 * w * h, stages_ms["raster_early"]),
        _k("k_hiz_tiles + k_hiz_tail", 4 * hw_ * hh_ + 4 * (4 * hw_ * hh_) // 3, stages_ms["hiz"]),
        _k("k_raster_visbuffer (late)", cnt_late["late"] * per_meshlet, stages_ms["raster_late"]),
    ]

    # ---------------- e2e through the reference-facing host API with HOST buffers ----------------
    e2e = None
    if not args.no_e2e and not multi:
        r = capi.Renderer(local_rank, scene)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
        # per-frame HOST inputs = what RendererInstance::update / render receive every frame: the camera and the (dirty)
        # transforms; the external depth (terrain stand-in) is GPU-resident in the engine, so it is uploaded once
        r.set_external_depth(scene.occluder_depth)
        xf_pinned = pin((len(scene.transforms), 16), torch.float32)
        xf_pinned[...] = scene.transforms["world"]
        # per-frame HOST outputs = the integer results of the path: the R32UI vis image, the survivor ids and the counters.
        # The D32F depth attachment only feeds GPU passes (Hi-Z, shading) and stays device-resident, as in the engine.
        outbufs = [dict(vis32=pin((h, w), torch.int32).view(np.uint32),
                        idx=pin((max(1, scene.max_meshlet_instance_count),), torch.int32).view(np.uint32)) for _ in range(2)]

        def e2e_steps(n):
            """n pipelined frames: frame i's device->host copies overlap frame i+1's kernels (oxr_submit / oxr_wait);
            every frame still pays its own H2D (camera, transforms) and D2H (vis32, survivor ids, counters)."""
            prev, res_ = None, None
            for i in range(n):
                r.update_transforms(xf_pinned)                   # H2D: all transforms (pinned)
                t = r.submit(cams[i % 2], outbufs[i % 2])        # H2D camera; kernels; D2H enqueued on the copy stream
                if prev is not None:
                    res_ = r.wait(prev)                          # frame i-1 is now in host memory
                prev = t
            return r.wait(prev)

        e2e_steps(W)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = e2e_steps(K)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / K
        h2d = 96 + xf_pinned.nbytes
        d2h = outbufs[0]["vis32"].nbytes + outbufs[0]["idx"].nbytes + 12 + 8 + 8
        e2e = {"value": res["total"] / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_s * 1e3, "api": "oxr_update_transforms + oxr_submit / oxr_wait (C++ ox::RendererInstance mirror over the C ABI), pinned host buffers, "
                      "2 frames in flight; H2D camera + all transforms, D2H vis32 image + survivor ids + counters (depth stays on the device)"}
        r.close()
    elif multi:
        e2e = None

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu_baseline = None
    if not args.no_cpu and not multi and rank == 0:
        cores = os.cpu_count() or 1
        sec, last = cpu_frames(scene, args.cpu_frames, cores)
        cpu_baseline = {"value": scene.max_meshlet_instance_count / sec, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{args.cpu_frames} full frames of the same {scene.max_meshlet_instance_count}-meshlet scene "
                                  f"(oracle port of the reference shaders, pthreads x{cores}); {sec * 1e3:.0f} ms/frame",
                        "triangles_per_s": last["triangles"] / sec,
                        "frustum_cull_draw_list_loops": cpu_frustum_loops(scene, cams[0], cores)}

    overlap_check = None
    if overlap is not None:
        # the overlapped steps must deliver what the serial exchange delivers: replay the same 4-frame sequence from a
        # zeroed visibility mask through both paths and compare counts, survivor sets and the reduced image
        torch.cuda.current_stream().wait_stream(overlap["cs"])
        torch.cuda.synchronize()
        pipe.ctx.reset_visibility_mask()
        for i in range(4):
            step_overlapped(i)
        torch.cuda.current_stream().wait_stream(overlap["cs"])
        torch.cuda.synchronize()
        img_o = pipe.vis64_bufs[1].clone()
        cnt_o = overlap["cnt_all"][1].view(world, 3).cpu().numpy()
        ids_o = overlap["ids_all"][1].view(world, gcap).cpu().numpy()
        pipe.ctx.reset_visibility_mask()
        pipe.select_buffer(0)
        for i in range(4):
            pipe.frame(cams[i % 2], **hooks)
        torch.cuda.synchronize()
        cnt_s = vis_all.view(world, 3).cpu().numpy()
        ids_s = ids_all.view(world, gcap).cpu().numpy()
        same_ids = all(np.array_equal(np.sort(ids_o[r, : cnt_o[r, 1] + cnt_o[r, 2]]), np.sort(ids_s[r, : cnt_s[r, 1] + cnt_s[r, 2]])) for r in range(world))
        overlap_check = bool(np.array_equal(cnt_o, cnt_s) and same_ids and torch.equal(img_o, pipe.vis64))

    exchange = None
    if multi:
        def time_op(fn, n=10):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n * 1e3  # us

        op_us = {"all_reduce_max_visbuffer_%dMB" % (pipe.vis64.numel() * 8 >> 20): time_op(lambda: oxdist.reduce_visbuffer(pipe.vis64)),
                 "all_reduce_max_hiz_mip0_%dMB" % (mip0_view.numel() * 4 >> 20): time_op(lambda: dist.all_reduce(mip0_view, op=dist.ReduceOp.MAX)),
                 "all_gather_survivor_ids_%dMB_per_rank" % (gcap * 4 >> 20): time_op(lambda: dist.all_gather_into_tensor(ids_all, ids_view)),
                 "all_gather_counts": time_op(lambda: dist.all_gather_into_tensor(vis_all, vis_view))}
        cnts = vis_all.view(world, 3).cpu().numpy()
        exchange = {"survivor_gather_capacity": int(gcap), "max_survivors_per_rank": int((cnts[:, 1] + cnts[:, 2]).max()),
                    "overflow": bool((cnts[:, 1] + cnts[:, 2]).max() > gcap), "op_us_back_to_back": op_us,
                    "steps": "id base from a local count-only replay (no exchange); all_reduce(MAX) Hi-Z mip 0; all_reduce(MAX) vis buffer; allgather(counts); allgather(survivor ids)"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, scene),
            "triangles_rasterised_per_s": job_tris / (ms_per_step * 1e-3),
            "per_frame": {"meshlet_instances": job_meshlets, "early_survivors": cnt["early"], "late_survivors": cnt["late"],
                          "triangles_rasterised": job_tris},
            "stages_ms": stages_ms, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches_per_frame * K), "gpu_launches_per_step": int(launches_per_frame),
            "cuda_graph": (graphs is not None) or (half_graphs is not None), "exchange_overlapped": overlap is not None, "overlap_check": overlap_check, "wall_s_timed_region": t_wall, "exchange": exchange, "stages_ms_per_rank": stages_all,
        }
        print(json.dumps(line), flush=True)
    pipe.close()
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
