// renderer_instance.cpp — see renderer_instance.hpp.  Host code only: every byte of compute goes through
// the C ABI of include/oxcull.h.
#include "renderer_instance.hpp"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace ox {

namespace {
// OXR_TRACE=1: per-frame device timeline of the pipelined path (frame start / compute end / copy start / copy end),
// printed by oxr_wait relative to the first traced frame.  Debug aid for tools/e2e_probe.py.
struct Trace {
  bool on = false, have_base = false;
  cudaEvent_t base = nullptr;
  cudaEvent_t ev[2][4] = {};
  cudaEvent_t stage[2][8] = {}; // OXR_TRACE=1: after clear+merge, cull_meshes+early cull, early raster, hiz, late cull, late raster
  int cur_slot = 0, n_stage = 0;
  void mark(cudaStream_t s) {
    if (!on || n_stage >= 8) return;
    if (!stage[cur_slot][n_stage]) cudaEventCreate(&stage[cur_slot][n_stage]);
    cudaEventRecord(stage[cur_slot][n_stage++], s);
  }
  Trace() {
    const char* e = std::getenv("OXR_TRACE");
    on = e && e[0] == '1';
  }
} g_trace;

struct Readback {
  OxcMeshletInstanceVisibility visibility;
  uint32_t draw_index_count[2];
  unsigned long long raster_triangles;
};
} // namespace

int RendererInstance::fail(int rc) {
  if (rc != OXC_OK && error_.empty()) error_ = oxc_last_error();
  return rc;
}

RendererInstance::RendererInstance(int device, const OxcCreateInfo& info, uint32_t width, uint32_t height)
    : width_(width), height_(height) {
  if (oxc_create(device, &info, &ctx_) != OXC_OK) {
    error_ = oxc_last_error();
    ctx_ = nullptr;
    return;
  }
  cudaStream_t s = nullptr;
  const size_t px = (size_t)width * height;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&d_vis64_), px * 8) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&d_vis32_), px * 4) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&d_depth_), px * 4) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&d_occluder_), px * 4) != cudaSuccess ||
      cudaMallocHost(&h_pinned_, sizeof(Readback)) != cudaSuccess) {
    error_ = std::string("RendererInstance allocation failed: ") + cudaGetErrorString(cudaGetLastError());
  }
  stream_ = s;
  ids_capacity_ = info.max_meshlet_instances ? info.max_meshlet_instances : 1;
  cudaStream_t cs = nullptr;
  bool ok = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess;
  copy_stream_ = cs;
  {
    cudaEvent_t ew = nullptr;
    ok = ok && cudaEventCreateWithFlags(&ew, cudaEventDisableTiming) == cudaSuccess;
    ev_window_ = ew;
  }
  for (auto& sl : slots_) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ok = ok && cudaMalloc(reinterpret_cast<void**>(&sl.d_vis32), px * 4) == cudaSuccess &&
         cudaMalloc(reinterpret_cast<void**>(&sl.d_depth), px * 4) == cudaSuccess &&
         cudaMalloc(reinterpret_cast<void**>(&sl.d_ids), (size_t)ids_capacity_ * 4) == cudaSuccess &&
         cudaMalloc(&sl.d_counters, sizeof(Readback)) == cudaSuccess &&
         cudaMemset(sl.d_counters, 0, sizeof(Readback)) == cudaSuccess &&
         cudaMallocHost(&sl.h_readback, sizeof(Readback)) == cudaSuccess &&
         cudaEventCreateWithFlags(&e0, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) == cudaSuccess;
    sl.ev_compute = e0;
    sl.ev_copy = e1;
  }
  for (int k = 0; k < 2; k++) {
    ok = ok && cudaMalloc(&d_cam_[k], sizeof(OxcCullCamera)) == cudaSuccess && cudaMallocHost(&h_cam_[k], sizeof(OxcCullCamera)) == cudaSuccess;
  }
  {
    const char* e = std::getenv("OXR_NO_GRAPH");
    use_graphs_ = !(e && e[0] == '1');
  }
  if (!ok && error_.empty()) error_ = std::string("RendererInstance pipeline allocation failed: ") + cudaGetErrorString(cudaGetLastError());
}

RendererInstance::~RendererInstance() {
  if (stream_) cudaStreamSynchronize(static_cast<cudaStream_t>(stream_));
  cudaFree(d_vis64_); cudaFree(d_vis32_); cudaFree(d_depth_); cudaFree(d_occluder_); cudaFree(d_overdraw_);
  if (h_pinned_) cudaFreeHost(h_pinned_);
  if (copy_stream_) cudaStreamSynchronize(static_cast<cudaStream_t>(copy_stream_));
  for (auto& sl : slots_) {
    cudaFree(sl.d_vis32); cudaFree(sl.d_depth); cudaFree(sl.d_ids); cudaFree(sl.d_counters);
    if (sl.h_readback) cudaFreeHost(sl.h_readback);
    if (sl.ev_compute) cudaEventDestroy(static_cast<cudaEvent_t>(sl.ev_compute));
    if (sl.ev_copy) cudaEventDestroy(static_cast<cudaEvent_t>(sl.ev_copy));
  }
  drop_graphs();
  for (int k = 0; k < 2; k++) { cudaFree(d_cam_[k]); if (h_cam_[k]) cudaFreeHost(h_cam_[k]); }
  if (ev_window_) cudaEventDestroy(static_cast<cudaEvent_t>(ev_window_));
  if (copy_stream_) cudaStreamDestroy(static_cast<cudaStream_t>(copy_stream_));
  if (stream_) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
  oxc_destroy(ctx_);
}

// RendererInstance::update (RendererInstance.cpp:1333-1788): table uploads + mask (re)allocation / zero fill
auto RendererInstance::update(const RendererInstanceUpdateInfo& info) -> int {
  if (!ctx_) return OXC_E_STATE;
  error_.clear();
  drop_graphs(); // the scene tables may be reallocated: captured launches would keep the old addresses
  oxc_bind_camera_buffer(ctx_, nullptr);
  return fail(oxc_set_scene(ctx_, &info.scene, stream_));
}

auto RendererInstance::update_transforms(const OxcTransformWorld* transforms, uint32_t first, uint32_t count) -> int {
  if (!ctx_) return OXC_E_STATE;
  return fail(oxc_update_transforms(ctx_, transforms, first, count, stream_));
}

// MainGeometryContext::draw_overdraw (RendererInstance.cpp:771-776; debug view "Overdraw"): the encode pass's fragment counter for
// the frame that was rendered last — both passes' survivor lists are still in place, so this runs after the frame, not inside it
auto RendererInstance::overdraw(const OxcCullCamera& camera, uint32_t* overdraw_host) -> int {
  if (!ctx_ || !error_.empty()) return OXC_E_STATE;
  if (!overdraw_host) return OXC_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const size_t px = (size_t)width_ * height_;
  if (!d_overdraw_ && cudaMalloc(reinterpret_cast<void**>(&d_overdraw_), px * 4) != cudaSuccess) return fail(OXC_E_CUDA);
  int rc;
  if ((rc = oxc_bind_camera_buffer(ctx_, nullptr)) != OXC_OK) return fail(rc); // as render(): the camera comes from the argument
  if ((rc = oxc_clear_overdraw(ctx_, d_overdraw_, width_, height_, s)) != OXC_OK) return fail(rc);
  if ((rc = oxc_raster_overdraw(ctx_, &camera, OXC_CULL_TEST_ALL, width_, height_, d_overdraw_, /*after_frame=*/1, s)) != OXC_OK) return fail(rc);
  if ((rc = oxc_raster_overdraw(ctx_, &camera, OXC_CULL_TEST_ALL | OXC_CULL_LATE_PASS, width_, height_, d_overdraw_, /*after_frame=*/1, s)) != OXC_OK) return fail(rc);
  if (cudaMemcpyAsync(overdraw_host, d_overdraw_, px * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
    return fail(OXC_E_CUDA);
  return OXC_OK;
}

// the reference uploads its material table with the scene (RendererInstance.cpp:1333-1788); the vis-buffer encode reads it for
// the alpha-tested discard (visbuffer_encode.slang:54-66)
auto RendererInstance::set_materials(const OxcMaterialTable* table) -> int {
  if (!ctx_) return OXC_E_STATE;
  error_.clear();
  drop_graphs(); // the raster's launch sequence depends on whether a table is set
  return fail(oxc_set_materials(ctx_, table, stream_));
}

auto RendererInstance::set_external_depth(const float* depth_host) -> int {
  if (!ctx_) return OXC_E_STATE;
  has_external_depth_ = depth_host != nullptr;
  if (!depth_host) return OXC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (cudaMemcpyAsync(d_occluder_, depth_host, (size_t)width_ * height_ * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return fail(OXC_E_CUDA);
  return OXC_OK;
}

// CullGeometry.cpp:61-404
auto RendererInstance::cull_geometry(CullGeometryContext& context) -> int {
  int rc;
  // --- Stage 1: cull_meshes (only on the first cull of a sequence), :68-117
  if (context.init_cull_meshes)
    if ((rc = oxc_cull_meshes(ctx_, &context.cull_camera, context.cull_flags, stream_)) != OXC_OK) return fail(rc);
  // --- Stage 2: cull_meshlets, :124-335.  The plain variant is dispatched with TestFrustum only (:275).
  if (context.use_hiz) rc = oxc_cull_meshlets(ctx_, &context.cull_camera, context.cull_flags, 1, stream_);
  else rc = oxc_cull_meshlets(ctx_, &context.cull_camera, OXC_CULL_TEST_FRUSTUM, 0, stream_);
  if (rc != OXC_OK) return fail(rc);
  // --- Stage 3: cull_triangles, :337-403 (optional here: the raster fuses it)
  if (context.materialize_indices)
    if ((rc = oxc_cull_triangles(ctx_, &context.cull_camera, context.cull_flags, stream_)) != OXC_OK) return fail(rc);
  return OXC_OK;
}

// CullGeometry.cpp:10-59 — source is the depth half of the packed vis buffer
auto RendererInstance::generate_hiz(MainGeometryContext& context) -> int {
  return fail(oxc_build_hiz_packed(ctx_, context.visbuffer_attachment, context.width, context.height, stream_));
}

// DrawGeometry.cpp:104-190 — "vis encode": the HW raster of reordered_indices becomes the fused
// triangle-cull + software raster into the packed image
auto RendererInstance::draw_for_visbuffer(MainGeometryContext& context) -> int {
  return fail(oxc_raster_visbuffer(ctx_, &context.cull_camera, context.cull_flags, context.width, context.height,
                                   context.visbuffer_attachment, 0, stream_));
}

// The geometry section of RendererInstance::render (RendererInstance.cpp:768-884), enqueued on stream_, in two halves: the
// pipelined path sends the previous frame's results to the host between them (flush_pending_copy), and replays each half
// from a CUDA graph.
//   head: attachments cleared, run_geometry_pass(false) up to and including its cull          (:562-588, :842-870)
//   tail: its draw, generate_hiz, run_geometry_pass(true)                                      (:871-884)
int RendererInstance::frame_head(const OxcCullCamera& camera, const float* occluder_depth_host, void* readback, bool readback_on_device) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const size_t px = (size_t)width_ * height_;
  int rc;
  OxcOutputs out;
  if ((rc = oxc_get_outputs(ctx_, &out)) != OXC_OK) return rc;
  Readback* rb = static_cast<Readback*>(readback);
  // attachments: depth cleared to 0, vis buffer to ~0 (RendererInstance.cpp:562-571,629-680), Hi-Z cleared every frame (:579-588)
  if (occluder_depth_host) {
    if (cudaMemcpyAsync(d_occluder_, occluder_depth_host, px * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) return OXC_E_CUDA;
    has_external_depth_ = true;
  }
  if (has_external_depth_) rc = oxc_clear_visbuffer_with_depth(ctx_, d_vis64_, d_occluder_, width_, height_, s); // clear + merge, one pass
  else rc = oxc_clear_visbuffer(ctx_, d_vis64_, width_, height_, s);
  if (rc != OXC_OK) return rc;
  if ((rc = oxc_clear_hiz(ctx_, s)) != OXC_OK) return rc;
  g_trace.mark(s);
  if (readback_on_device) {
    if (cudaMemsetAsync(rb->draw_index_count, 0, sizeof rb->draw_index_count, s) != cudaSuccess) return OXC_E_CUDA;
  } else {
    rb->draw_index_count[0] = rb->draw_index_count[1] = 0;
  }
  // RendererInstance.cpp:796-800: hoisted so the early pass's visibility / dispatch buffers persist into the late pass
  CullGeometryContext cull_geometry_context;
  cull_geometry_context.use_hiz = true;
  cull_geometry_context.init_cull_meshes = true;
  cull_geometry_context.cull_camera = camera;
  cull_geometry_context.materialize_indices = out.reordered_indices != nullptr;
  if ((rc = cull_geometry(cull_geometry_context)) != OXC_OK) return rc; // run_geometry_pass(false), cull half
  g_trace.mark(s);
  return OXC_OK;
}

int RendererInstance::frame_tail(const OxcCullCamera& camera, void* readback) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int rc;
  OxcOutputs out;
  if ((rc = oxc_get_outputs(ctx_, &out)) != OXC_OK) return rc;
  Readback* rb = static_cast<Readback*>(readback);
  const bool materialize = out.reordered_indices != nullptr;
  MainGeometryContext main_geometry_context;
  main_geometry_context.visbuffer_attachment = d_vis64_;
  main_geometry_context.width = width_;
  main_geometry_context.height = height_;
  main_geometry_context.cull_camera = camera;
  // run_geometry_pass(false), draw half (:871-881)
  if (materialize && cudaMemcpyAsync(&rb->draw_index_count[0], &out.draw_cmd->index_count, 4, cudaMemcpyDefault, s) != cudaSuccess) return OXC_E_CUDA;
  main_geometry_context.cull_flags = OXC_CULL_TEST_ALL;
  if ((rc = draw_for_visbuffer(main_geometry_context)) != OXC_OK) return rc;
  g_trace.mark(s);
  if ((rc = generate_hiz(main_geometry_context)) != OXC_OK) return rc; // :883
  g_trace.mark(s);
  // run_geometry_pass(true) (:884): cull_flags |= LatePass, init_cull_meshes = false
  CullGeometryContext cull_geometry_context;
  cull_geometry_context.use_hiz = true;
  cull_geometry_context.init_cull_meshes = false;
  cull_geometry_context.cull_flags = OXC_CULL_TEST_ALL | OXC_CULL_LATE_PASS;
  cull_geometry_context.cull_camera = camera;
  cull_geometry_context.materialize_indices = materialize;
  if ((rc = cull_geometry(cull_geometry_context)) != OXC_OK) return rc;
  g_trace.mark(s);
  if (materialize && cudaMemcpyAsync(&rb->draw_index_count[1], &out.draw_cmd->index_count, 4, cudaMemcpyDefault, s) != cudaSuccess) return OXC_E_CUDA;
  main_geometry_context.cull_flags = cull_geometry_context.cull_flags;
  if ((rc = draw_for_visbuffer(main_geometry_context)) != OXC_OK) return rc;
  g_trace.mark(s);
  return OXC_OK;
}

int RendererInstance::run_frame(const OxcCullCamera& camera, const float* occluder_depth_host, void* readback, bool readback_on_device) {
  int rc;
  if ((rc = frame_head(camera, occluder_depth_host, readback, readback_on_device)) != OXC_OK) return rc;
  if (pending_.active && (rc = flush_pending_copy(true)) != OXC_OK) return rc; // previous frame's results go out now
  return frame_tail(camera, readback);
}

// results of the frame -> the slot's device staging (drained by wait() of the slot's previous ticket)
int RendererInstance::stage_results(int slot, bool want_vis32, bool want_depth, uint32_t n_ids) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  Slot& sl = slots_[slot];
  int rc;
  OxcOutputs out;
  if ((rc = oxc_get_outputs(ctx_, &out)) != OXC_OK) return rc;
  Readback* dc = static_cast<Readback*>(sl.d_counters);
  if (want_vis32 || want_depth)
    if ((rc = oxc_resolve_visbuffer(ctx_, d_vis64_, width_, height_, want_vis32 ? sl.d_vis32 : nullptr, want_depth ? sl.d_depth : nullptr, s)) != OXC_OK)
      return rc;
  if (n_ids && cudaMemcpyAsync(sl.d_ids, out.visible_meshlet_instances_indices, (size_t)n_ids * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return OXC_E_CUDA;
  if (cudaMemcpyAsync(&dc->visibility, out.visibility, sizeof dc->visibility, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return OXC_E_CUDA;
  if (cudaMemcpyAsync(&dc->raster_triangles, out.raster_triangle_count, 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return OXC_E_CUDA;
  return OXC_OK;
}

void RendererInstance::drop_graphs() {
  for (auto& g : graphs_) {
    if (g.head) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g.head));
    if (g.tail) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g.tail));
    g = FrameGraph();
  }
}

// Captures one half of the frame into an executable graph.  `body` enqueues on stream_.  On any failure the capture is ended,
// graphs are switched off for this renderer and the caller falls back to eager launches.
template <typename Body>
int RendererInstance::capture(void** exec_out, Body body) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  *exec_out = nullptr;
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return OXC_E_CUDA; }
  const int rc = body();
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc != OXC_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc != OXC_OK ? rc : OXC_E_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) { cudaGetLastError(); return OXC_E_CUDA; }
  *exec_out = exec;
  return OXC_OK;
}

auto RendererInstance::render(const OxcCullCamera& camera, const float* occluder_depth_host, uint32_t* vis32_host,
                              float* depth_host, uint32_t* visible_indices_host, uint32_t visible_indices_capacity,
                              OxrFrameResult* result) -> int {
  if (!ctx_ || !error_.empty()) return OXC_E_STATE;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const size_t px = (size_t)width_ * height_;
  int rc;
  OxcOutputs out;
  if ((rc = oxc_get_outputs(ctx_, &out)) != OXC_OK) return fail(rc);
  Readback* rb = static_cast<Readback*>(h_pinned_);
  if ((rc = oxc_bind_camera_buffer(ctx_, nullptr)) != OXC_OK) return fail(rc); // the graph path may have left a device camera bound
  if ((rc = run_frame(camera, occluder_depth_host, rb, false)) != OXC_OK) return fail(rc);

  // results back to the host (the engine would hand the attachments to decode_visbuffer, :923-925)
  if (vis32_host || depth_host) {
    if ((rc = oxc_resolve_visbuffer(ctx_, d_vis64_, width_, height_, vis32_host ? d_vis32_ : nullptr,
                                    depth_host ? d_depth_ : nullptr, s)) != OXC_OK) return fail(rc);
    if (vis32_host && cudaMemcpyAsync(vis32_host, d_vis32_, px * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) return fail(OXC_E_CUDA);
    if (depth_host && cudaMemcpyAsync(depth_host, d_depth_, px * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) return fail(OXC_E_CUDA);
  }
  if (cudaMemcpyAsync(&rb->visibility, out.visibility, sizeof rb->visibility, cudaMemcpyDeviceToHost, s) != cudaSuccess) return fail(OXC_E_CUDA);
  if (cudaMemcpyAsync(&rb->raster_triangles, out.raster_triangle_count, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess) return fail(OXC_E_CUDA);
  if (cudaStreamSynchronize(s) != cudaSuccess) {
    error_ = cudaGetErrorString(cudaGetLastError());
    return OXC_E_CUDA;
  }
  if (visible_indices_host) {
    uint32_t n = rb->visibility.early_visible_meshlet_instances + rb->visibility.late_visible_meshlet_instances;
    if (n > visible_indices_capacity) n = visible_indices_capacity;
    if (n && cudaMemcpy(visible_indices_host, out.visible_meshlet_instances_indices, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail(OXC_E_CUDA);
  }
  if (result) {
    result->visibility = rb->visibility;
    result->draw_index_count_early = rb->draw_index_count[0];
    result->draw_index_count_late = rb->draw_index_count[1];
    result->raster_triangles = rb->raster_triangles;
  }
  return OXC_OK;
}

auto RendererInstance::submit(const OxcCullCamera& camera, uint32_t* vis32_host, float* depth_host,
                              uint32_t* visible_indices_host, uint32_t visible_indices_capacity, int* ticket) -> int {
  if (!ctx_ || !error_.empty() || !ticket) return OXC_E_STATE;
  const int slot = (int)(frame_ & 1);
  Slot& sl = slots_[slot];
  if (sl.in_flight) {
    error_ = "oxr_submit: the previous frame of this slot has not been waited";
    return OXC_E_STATE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int rc;
  OxcOutputs out;
  if ((rc = oxc_get_outputs(ctx_, &out)) != OXC_OK) return fail(rc);
  // The compute stream carries no device->host copy: counters are staged device-side like the images and every D2H
  // goes out on the copy stream (flush_pending_copy).
  Readback* dc = static_cast<Readback*>(sl.d_counters);
  if (g_trace.on) {
    if (!g_trace.ev[slot][0])
      for (int k = 0; k < 4; k++) cudaEventCreate(&g_trace.ev[slot][k]);
    if (!g_trace.have_base) { cudaEventCreate(&g_trace.base); cudaEventRecord(g_trace.base, s); g_trace.have_base = true; }
    cudaEventRecord(g_trace.ev[slot][0], s);
    g_trace.cur_slot = slot;
    g_trace.n_stage = 0;
  }
  const bool want_vis32 = vis32_host != nullptr, want_depth = depth_host != nullptr;
  const uint32_t n_ids = visible_indices_host ? (visible_indices_capacity < ids_capacity_ ? visible_indices_capacity : ids_capacity_) : 0;
  // Steady state: both halves of the frame replay from CUDA graphs (one pair per slot).  The camera reaches the kernels through
  // a device buffer (oxc_bind_camera_buffer) that the head graph fills from this slot's pinned copy, so a replay only needs the
  // new camera written there.  Everything else a graph bakes in is part of its key; a mismatch re-captures.
  bool done = false;
  if (use_graphs_ && !g_trace.on && out.reordered_indices == nullptr && d_cam_[slot] && h_cam_[slot]) {
    FrameGraph& g = graphs_[slot];
    const uint32_t key = (has_external_depth_ ? 1u : 0u) | (want_vis32 ? 2u : 0u) | (want_depth ? 4u : 0u);
    *static_cast<OxcCullCamera*>(h_cam_[slot]) = camera;
    if ((rc = oxc_bind_camera_buffer(ctx_, static_cast<const OxcCullCamera*>(d_cam_[slot]))) != OXC_OK) return fail(rc);
    if (!g.head || !g.tail || g.key != key || g.n_ids != n_ids || g.mesh_instance_count != camera.mesh_instance_count) {
      if (g.head) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g.head));
      if (g.tail) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g.tail));
      g = FrameGraph();
      int rc_a = capture(&g.head, [&]() -> int {
        const int r = oxc_load_camera(ctx_, static_cast<OxcCullCamera*>(d_cam_[slot]), static_cast<const OxcCullCamera*>(h_cam_[slot]), s);
        return r != OXC_OK ? r : frame_head(camera, nullptr, dc, true);
      });
      int rc_b = rc_a == OXC_OK ? capture(&g.tail, [&]() -> int {
        const int r = frame_tail(camera, dc);
        return r != OXC_OK ? r : stage_results(slot, want_vis32, want_depth, n_ids);
      }) : rc_a;
      if (rc_a != OXC_OK || rc_b != OXC_OK) { // capture is not available here: eager launches from now on
        drop_graphs();
        use_graphs_ = false;
      } else {
        g.key = key; g.n_ids = n_ids; g.mesh_instance_count = camera.mesh_instance_count;
      }
    }
    if (use_graphs_) {
      if (cudaGraphLaunch(static_cast<cudaGraphExec_t>(g.head), s) != cudaSuccess) return fail(OXC_E_CUDA);
      if (pending_.active && (rc = flush_pending_copy(true)) != OXC_OK) return fail(rc);
      if (cudaGraphLaunch(static_cast<cudaGraphExec_t>(g.tail), s) != cudaSuccess) return fail(OXC_E_CUDA);
      done = true;
    }
  }
  if (!done) {
    if ((rc = oxc_bind_camera_buffer(ctx_, nullptr)) != OXC_OK) return fail(rc);
    if ((rc = run_frame(camera, nullptr, dc, true)) != OXC_OK) return fail(rc);
    if ((rc = stage_results(slot, want_vis32, want_depth, n_ids)) != OXC_OK) return fail(rc);
  }
  if (cudaEventRecord(static_cast<cudaEvent_t>(sl.ev_compute), s) != cudaSuccess) return fail(OXC_E_CUDA);
  // the device -> host copies are issued later, from inside the next frame (or by wait()): see PendingCopy
  if (g_trace.on) cudaEventRecord(g_trace.ev[slot][1], s);
  pending_.active = true;
  pending_.slot = slot;
  pending_.vis32_host = vis32_host;
  pending_.depth_host = depth_host;
  pending_.ids_host = visible_indices_host;
  pending_.n_ids = n_ids;
  sl.in_flight = true;
  *ticket = slot;
  frame_++;
  return OXC_OK;
}

// Enqueues the device->host copies of the pending frame on the copy stream.  behind_window: called from inside the next
// frame after its early cull was launched — the copies then also wait for that point of the compute stream.
int RendererInstance::flush_pending_copy(bool behind_window) {
  if (!pending_.active) return OXC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream_), cs = static_cast<cudaStream_t>(copy_stream_);
  Slot& sl = slots_[pending_.slot];
  const size_t px = (size_t)width_ * height_;
  const int slot = pending_.slot;
  pending_.active = false;
  if (behind_window) {
    if (cudaEventRecord(static_cast<cudaEvent_t>(ev_window_), s) != cudaSuccess) return OXC_E_CUDA;
    if (cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(ev_window_), 0) != cudaSuccess) return OXC_E_CUDA;
  }
  if (cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(sl.ev_compute), 0) != cudaSuccess) return OXC_E_CUDA;
  if (g_trace.on) cudaEventRecord(g_trace.ev[slot][2], cs);
  if (cudaMemcpyAsync(sl.h_readback, sl.d_counters, sizeof(Readback), cudaMemcpyDeviceToHost, cs) != cudaSuccess) return OXC_E_CUDA;
  if (pending_.vis32_host && cudaMemcpyAsync(pending_.vis32_host, sl.d_vis32, px * 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return OXC_E_CUDA;
  if (pending_.depth_host && cudaMemcpyAsync(pending_.depth_host, sl.d_depth, px * 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return OXC_E_CUDA;
  if (pending_.n_ids && cudaMemcpyAsync(pending_.ids_host, sl.d_ids, (size_t)pending_.n_ids * 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess)
    return OXC_E_CUDA;
  if (g_trace.on) cudaEventRecord(g_trace.ev[slot][3], cs);
  if (cudaEventRecord(static_cast<cudaEvent_t>(sl.ev_copy), cs) != cudaSuccess) return OXC_E_CUDA;
  return OXC_OK;
}

auto RendererInstance::wait(int ticket, OxrFrameResult* result) -> int {
  if (ticket < 0 || ticket > 1) return OXC_E_INVALID;
  Slot& sl = slots_[ticket];
  if (!sl.in_flight) return OXC_E_STATE;
  if (pending_.active && pending_.slot == ticket) { // nobody submitted a frame after this one: send its results now
    const int rc = flush_pending_copy(false);
    if (rc != OXC_OK) return fail(rc);
  }
  if (cudaEventSynchronize(static_cast<cudaEvent_t>(sl.ev_copy)) != cudaSuccess) {
    error_ = cudaGetErrorString(cudaGetLastError());
    return OXC_E_CUDA;
  }
  sl.in_flight = false;
  if (g_trace.on && g_trace.ev[ticket][0]) {
    float t[4];
    for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], g_trace.base, g_trace.ev[ticket][k]);
    std::fprintf(stderr, "[oxr trace] slot %d: start %.3f  compute_end %.3f  copy_start %.3f  copy_end %.3f ms | stages us:", ticket, t[0], t[1], t[2], t[3]);
    float prev_t = t[0];
    for (int k = 0; k < 8 && g_trace.stage[ticket][k]; k++) {
      float st;
      if (cudaEventElapsedTime(&st, g_trace.base, g_trace.stage[ticket][k]) != cudaSuccess) break;
      std::fprintf(stderr, " %.0f", (st - prev_t) * 1e3f);
      prev_t = st;
    }
    std::fprintf(stderr, " | tail %.0f\n", (t[1] - prev_t) * 1e3f);
  }
  if (result) {
    const Readback* rb = static_cast<const Readback*>(sl.h_readback);
    result->visibility = rb->visibility;
    result->draw_index_count_early = rb->draw_index_count[0];
    result->draw_index_count_late = rb->draw_index_count[1];
    result->raster_triangles = rb->raster_triangles;
  }
  return OXC_OK;
}

} // namespace ox

// ---- C exports ----
struct OxrRenderer {
  ox::RendererInstance impl;
  OxrRenderer(int device, const OxcCreateInfo& info, uint32_t w, uint32_t h) : impl(device, info, w, h) {}
};

extern "C" {

int oxr_create(int device, const OxcCreateInfo* info, uint32_t width, uint32_t height, OxrRenderer** out) {
  if (!info || !out || !width || !height) return OXC_E_INVALID;
  *out = nullptr;
  OxrRenderer* r = new (std::nothrow) OxrRenderer(device, *info, width, height);
  if (!r) return OXC_E_INVALID;
  if (!r->impl.ok()) {
    const int rc = r->impl.context() ? OXC_E_CUDA : OXC_E_NO_DEVICE;
    delete r;
    return rc;
  }
  *out = r;
  return OXC_OK;
}

void oxr_destroy(OxrRenderer* r) { delete r; }

OxcContext* oxr_context(OxrRenderer* r) { return r ? r->impl.context() : nullptr; }

int oxr_update(OxrRenderer* r, const OxcSceneDesc* scene) {
  if (!r || !scene) return OXC_E_INVALID;
  ox::RendererInstanceUpdateInfo info;
  info.mesh_instance_count = scene->mesh_instance_count;
  info.scene = *scene;
  int rc = r->impl.update(info);
  if (rc == OXC_OK && cudaStreamSynchronize(static_cast<cudaStream_t>(r->impl.stream())) != cudaSuccess) rc = OXC_E_CUDA;
  return rc;
}

int oxr_update_transforms(OxrRenderer* r, const OxcTransformWorld* transforms, uint32_t first, uint32_t count) {
  if (!r || !transforms) return OXC_E_INVALID;
  return r->impl.update_transforms(transforms, first, count);
}

int oxr_overdraw(OxrRenderer* r, const OxcCullCamera* camera, uint32_t* overdraw_host) {
  if (!r || !camera) return OXC_E_INVALID;
  return r->impl.overdraw(*camera, overdraw_host);
}

int oxr_set_materials(OxrRenderer* r, const OxcMaterialTable* table) {
  if (!r) return OXC_E_INVALID;
  return r->impl.set_materials(table);
}

int oxr_set_external_depth(OxrRenderer* r, const float* depth_host) {
  if (!r) return OXC_E_INVALID;
  return r->impl.set_external_depth(depth_host);
}

int oxr_submit(OxrRenderer* r, const OxcCullCamera* camera, uint32_t* vis32_host, float* depth_host,
               uint32_t* visible_indices_host, uint32_t visible_indices_capacity, int* ticket) {
  if (!r || !camera || !ticket) return OXC_E_INVALID;
  return r->impl.submit(*camera, vis32_host, depth_host, visible_indices_host, visible_indices_capacity, ticket);
}

int oxr_wait(OxrRenderer* r, int ticket, OxrFrameResult* result) {
  if (!r) return OXC_E_INVALID;
  return r->impl.wait(ticket, result);
}

int oxr_render(OxrRenderer* r, const OxcCullCamera* camera, const float* occluder_depth_host, uint32_t* vis32_host,
               float* depth_host, uint32_t* visible_indices_host, uint32_t visible_indices_capacity, OxrFrameResult* result) {
  if (!r || !camera) return OXC_E_INVALID;
  return r->impl.render(*camera, occluder_depth_host, vis32_host, depth_host, visible_indices_host, visible_indices_capacity, result);
}

} // extern "C"
