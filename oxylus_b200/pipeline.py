"""Python host for the device-resident frame loop (bench.py `value`, multi-GPU sharding).

Sequencing mirrors RendererInstance::render's geometry section (RendererInstance.cpp:842-884):
    cull_meshes -> cull_meshlets(early) -> vis encode -> generate_hiz -> cull_meshlets(late) -> vis encode
All compute goes through the C ABI (capi.Context) on torch's current CUDA stream; torch only provides device
memory, streams, events and (N > 1) torch.distributed.
"""
import numpy as np
import torch

from . import abi, capi

STAGES = ["cull_meshes", "cull_early", "raster_early", "hiz", "cull_late", "raster_late"]


class _DevView:
    """Zero-copy torch view of a raw device pointer owned by the C-ABI context (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def wrap_device(ptr, n, dtype, device):
    typestr = {torch.int32: "<i4", torch.float32: "<f4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_DevView(ptr, n, typestr), device=device)


class VisibilityPipeline:
    """shard = (first mesh instance, count) restricts the context to a contiguous range of mesh instances (multi-GPU).
    mgpu = dict(rank=, world=, unique_id=, survivor_capacity=0): the product's multi-GPU exchange (oxc_mgpu_*) is set up and
    frame() uses oxc_mgpu_exchange_hiz between the passes; exchange_frame() merges the outputs.
    shard_capacity: meshlet instances this context can hold (default: the whole scene; a shard host passes its own share so
    the 8 B + 4 B per-meshlet buffers are shard-sized while the 1-bit mask still spans the scene)."""

    def __init__(self, scene, device=0, shard=None, alloc_reordered_indices=False, auto_id_base=False, mgpu=None,
                 shard_capacity=None, wide_ids=False):
        self.scene = scene
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        hw, hh = scene.hiz_extent()
        cap = max(1, scene.max_meshlet_instance_count if shard_capacity is None else shard_capacity)
        self.ctx = capi.Context(device, max(1, scene.mesh_instance_count), cap, hw, hh,
                                alloc_reordered_indices=alloc_reordered_indices, stream=0,
                                max_mask_bits=max(1, scene.max_meshlet_instance_count), wide_ids=wide_ids)
        if shard is not None:  # before set_scene: the capacity check is made against the shard's range
            if auto_id_base:
                self.ctx.set_shard_auto(shard[0], shard[1])  # id base from a local count-only replay: no exchange
        self.ctx.set_scene(scene)
        self.mgpu = None
        if mgpu is not None:
            self.mgpu = self.ctx.mgpu_init(mgpu["rank"], mgpu["world"], mgpu["unique_id"], mgpu.get("survivor_capacity", 0))
        self.w, self.h = scene.width, scene.height
        # two packed vis buffers: multi-GPU runs alternate them so the trailing exchange of frame i (side stream) can
        # overlap frame i+1; single-GPU code only ever uses buffer 0
        self.vis64_bufs = [torch.empty((self.h, self.w), dtype=torch.int64, device=self.device) for _ in range(2)]
        self.vis64 = self.vis64_bufs[0]
        self.occluder = None
        if scene.occluder_depth is not None:
            self.occluder = torch.from_numpy(np.ascontiguousarray(scene.occluder_depth)).to(self.device)
        self.id_base = torch.zeros(1, dtype=torch.int32, device=self.device)
        if shard is not None and not auto_id_base:
            self.ctx.set_shard(shard[0], shard[1], self.id_base.data_ptr())
        self.use_torch_stream()

    def select_buffer(self, b):
        self.vis64 = self.vis64_bufs[b & 1]

    def use_torch_stream(self):
        self.ctx.stream = torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        self.ctx.close()

    # --- one frame; `mark(name)` is called after each stage (records a CUDA event in bench.py) ---
    def frame(self, cam, mark=None, after_cull_meshes=None, between_passes=None, after_frame=None):
        c, w, h = self.ctx, self.w, self.h
        v = self.vis64.data_ptr()
        if self.occluder is not None:
            c.clear_visbuffer_with_depth(v, self.occluder.data_ptr(), w, h)  # clear + external depth, one pass
        else:
            c.clear_visbuffer(v, w, h)
        c.clear_hiz()
        if mark:
            mark("begin")
        c.cull_meshes(cam, abi.CULL_TEST_ALL)
        if after_cull_meshes:
            after_cull_meshes()  # multi-GPU: exchange emitted counts -> this rank's global id base
        if mark:
            mark("cull_meshes")
        c.cull_meshlets(cam, abi.CULL_TEST_ALL, True)
        if mark:
            mark("cull_early")
        c.raster_visbuffer(cam, abi.CULL_TEST_ALL, w, h, v)
        if mark:
            mark("raster_early")
        if self.mgpu is not None:
            c.mgpu_exchange_hiz(v, w, h)  # mip-0 texels max-reduced into every peer over NVLink, then the identical pyramid
        elif between_passes:
            # legacy host-driven exchange: only the point-sampled mip 0 travels (max over ranks commutes with sampling)
            c.build_hiz_mip0_packed(v, w, h)
            between_passes()
            c.build_hiz_from_mip0()
        else:
            c.build_hiz_packed(v, w, h)
        if mark:
            mark("hiz")
        c.cull_meshlets(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, True)
        if mark:
            mark("cull_late")
        c.raster_visbuffer(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, w, h, v)
        if mark:
            mark("raster_late")
        if after_frame:
            after_frame()

    # --- the same frame in two halves around the multi-GPU Hi-Z exchange (each half can be a CUDA graph; the NCCL
    #     collectives stay outside the graphs) ---
    def frame_before_exchange(self, cam):
        c, w, h = self.ctx, self.w, self.h
        v = self.vis64.data_ptr()
        if self.occluder is not None:
            c.clear_visbuffer_with_depth(v, self.occluder.data_ptr(), w, h)  # clear + external depth, one pass
        else:
            c.clear_visbuffer(v, w, h)
        c.clear_hiz()
        c.cull_meshes(cam, abi.CULL_TEST_ALL)
        c.cull_meshlets(cam, abi.CULL_TEST_ALL, True)
        c.raster_visbuffer(cam, abi.CULL_TEST_ALL, w, h, v)
        c.build_hiz_mip0_packed(v, w, h)

    def frame_after_exchange(self, cam):
        c, w, h = self.ctx, self.w, self.h
        v = self.vis64.data_ptr()
        c.build_hiz_from_mip0()
        c.cull_meshlets(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, True)
        c.raster_visbuffer(cam, abi.CULL_TEST_ALL | abi.CULL_LATE_PASS, w, h, v)

    def exchange_frame(self, slot=0, stream=None, with_image=True, already_staged=False, vis=None):
        """oxc_mgpu_exchange_frame: vis-buffer max-reduce (in place) + survivor allgather into the context's slot buffers."""
        v = self.vis64 if vis is None else vis
        self.ctx.mgpu_exchange_frame(v.data_ptr() if with_image else None, self.w, self.h, slot,
                                     None if stream is None else stream.cuda_stream, already_staged=already_staged)

    def counters(self):
        vis = self.ctx.visibility()
        return dict(total=int(vis["total"][0]), early=int(vis["early"][0]), late=int(vis["late"][0]),
                    triangles=self.ctx.raster_triangle_count())
