// BodyFitter.fit for HOST-resident targets (the end-to-end entry point): the batch is cut into chunks,
// the host-to-device copies run on a library-owned copy stream ahead of the fits, consecutive chunks are fitted on
// a small ring of library-owned compute streams (so their latency-bound stages overlap), and the (small) results are copied back into the caller's host buffers.  Nothing is
// allocated per call: the staging buffers, the per-batch result buffers and the fit workspace are carved from
// the caller's device workspace; the copy stream and the events are created once per device.
// Reference behaviour: pt/bodyfitter.py:283-549 applied to every chunk (instances are independent; share_beta,
// the only cross-instance option, is rejected here).
#include <mutex>

#include "common.cuh"

namespace sf {
namespace {

// Chunk k is staged into slot k % kSlots and fitted on that slot's own compute stream, so the short dependent
// kernels of one chunk (per-instance solves, ~1 ms of fixed latency per fit) overlap the vertex passes of its
// neighbours, and the copy stream can run up to kSlots - 1 chunks ahead of the oldest fit still in flight.
constexpr int kSlots = 3;

struct HostPipe {
  cudaStream_t copy = nullptr;
  cudaStream_t compute[kSlots] = {};
  cudaEvent_t ready[kSlots] = {};  // H2D into slot i complete
  cudaEvent_t done[kSlots] = {};   // fit of slot i complete
  cudaEvent_t entry = nullptr;     // work queued on the caller's stream before this call
  bool ok = false;
};

constexpr int kMaxDevices = 64;
HostPipe g_pipes[kMaxDevices];
std::mutex g_pipe_mutex;

HostPipe* pipe_for_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  HostPipe& p = g_pipes[dev];
  if (!p.ok) {
    if (cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < kSlots; ++i) {
      if (cudaStreamCreateWithFlags(&p.compute[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&p.ready[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&p.entry, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    p.ok = true;
  }
  return &p;
}

struct HostWs {
  float *stage_v[kSlots], *stage_j[kSlots];
  float *pose, *betas, *trans, *orient, *rel, *kid, *scale;
  void* fit_ws[kSlots];
  size_t fit_ws_bytes, bytes;
};

HostWs carve_host(void* base, const smplfit_model_t* m, int64_t B, int64_t chunk, const smplfit_fit_opts_t* o,
                  int has_joints) {
  HostWs w{};
  Carver c(base);
  const size_t V = m->num_vertices, J = m->num_joints, S = m->num_betas;
  w.fit_ws_bytes = smplfit_fit_workspace_bytes(m, chunk, o, has_joints, 0, 0);
  const int slots = (int)((B + chunk - 1) / chunk < kSlots ? (B + chunk - 1) / chunk : kSlots);
  for (int i = 0; i < kSlots; ++i) {
    const bool used = i < slots;
    w.stage_v[i] = used ? c.take<float>((size_t)chunk * V * 3) : nullptr;
    w.stage_j[i] = (used && has_joints) ? c.take<float>((size_t)chunk * J * 3) : nullptr;
    w.fit_ws[i] = used ? c.take<char>(w.fit_ws_bytes) : nullptr;
  }
  w.pose = c.take<float>((size_t)B * J * 3);
  w.betas = c.take<float>((size_t)B * S);
  w.trans = c.take<float>((size_t)B * 3);
  w.orient = c.take<float>((size_t)B * J * 9);
  w.rel = c.take<float>((size_t)B * J * 9);
  w.kid = c.take<float>((size_t)B);
  w.scale = c.take<float>((size_t)B);
  w.bytes = c.off + 256;
  return w;
}

}  // namespace
}  // namespace sf

using namespace sf;

extern "C" size_t smplfit_fit_host_workspace_bytes(const smplfit_model_t* m, int64_t batch, int64_t chunk,
                                                   const smplfit_fit_opts_t* o, int has_joints) {
  if (!m || !o || batch <= 0 || chunk <= 0) return 0;
  if (chunk > batch) chunk = batch;
  if (smplfit_fit_workspace_bytes(m, chunk, o, has_joints, 0, 0) == 0) return 0;
  return carve_host(nullptr, m, batch, chunk, o, has_joints).bytes;
}

extern "C" int smplfit_fit_host(const smplfit_model_t* m, int64_t batch, int64_t chunk, const float* host_target_vertices,
                                const float* host_target_joints, const smplfit_fit_opts_t* o,
                                float* host_pose_rotvecs, float* host_shape_betas, float* host_trans,
                                float* host_orientations, float* host_rel_orientations, float* host_kid_factor,
                                float* host_scale_corr, void* workspace, size_t workspace_bytes, void* stream) {
  if (!m || !o || !host_target_vertices || !host_shape_betas || !host_trans)
    return fail(SMPLFIT_ERR_ARG, "missing required pointer");
  if (batch <= 0 || chunk <= 0) return fail(SMPLFIT_ERR_ARG, "batch and chunk must be positive");
  if (o->share_beta) return fail(SMPLFIT_ERR_UNSUPPORTED, "share_beta couples the whole batch: use smplfit_fit");
  if (o->want_pose_rotvecs && !host_pose_rotvecs) return fail(SMPLFIT_ERR_ARG, "pose_rotvecs output required");
  if (o->scale_mode != 0 && !host_scale_corr) return fail(SMPLFIT_ERR_ARG, "scale_corr output required");
  if (chunk > batch) chunk = batch;
  const int has_joints = host_target_joints != nullptr;
  HostWs w = carve_host(workspace, m, batch, chunk, o, has_joints);
  if (w.fit_ws_bytes == 0) return fail(SMPLFIT_ERR_UNSUPPORTED, "model not supported by smplfit_fit");
  if (!workspace || w.bytes > workspace_bytes) return fail(SMPLFIT_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_pipe_mutex);  // one pipeline per device: concurrent callers take turns
  HostPipe* p = pipe_for_current_device();
  if (!p) return fail(SMPLFIT_ERR_CUDA, "could not create the copy stream / events");
  const size_t V = m->num_vertices, J = m->num_joints, S = m->num_betas;
  const int64_t n_chunks = (batch + chunk - 1) / chunk;

#define SF_CU(x)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) return fail(SMPLFIT_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e_)); \
  } while (0)

  auto stage = [&](int64_t k) -> cudaError_t {
    const int slot = (int)(k % kSlots);
    const int64_t lo = k * chunk, n = (lo + chunk <= batch) ? chunk : batch - lo;
    cudaError_t e;
    if (k >= kSlots && (e = cudaStreamWaitEvent(p->copy, p->done[slot], 0)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(w.stage_v[slot], host_target_vertices + (size_t)lo * V * 3, (size_t)n * V * 3 * sizeof(float),
                             cudaMemcpyHostToDevice, p->copy)) != cudaSuccess)
      return e;
    if (has_joints && (e = cudaMemcpyAsync(w.stage_j[slot], host_target_joints + (size_t)lo * J * 3,
                                           (size_t)n * J * 3 * sizeof(float), cudaMemcpyHostToDevice, p->copy)) != cudaSuccess)
      return e;
    return cudaEventRecord(p->ready[slot], p->copy);
  };

  // the workspace belongs to the caller: earlier work on the caller's stream may still be using it
  SF_CU(cudaEventRecord(p->entry, st));
  SF_CU(cudaStreamWaitEvent(p->copy, p->entry, 0));
  for (int i = 0; i < kSlots; ++i) SF_CU(cudaStreamWaitEvent(p->compute[i], p->entry, 0));
  // copies are issued kSlots - 1 chunks ahead of the fits (the copy stream executes them back to back)
  for (int64_t k = 0; k < n_chunks && k < kSlots - 1; ++k) SF_CU(stage(k));
  for (int64_t k = 0; k < n_chunks; ++k) {
    const int slot = (int)(k % kSlots);
    const int64_t lo = k * chunk, n = (lo + chunk <= batch) ? chunk : batch - lo;
    if (k + kSlots - 1 < n_chunks) SF_CU(stage(k + kSlots - 1));
    cudaStream_t cs = p->compute[slot];
    SF_CU(cudaStreamWaitEvent(cs, p->ready[slot], 0));
    const int rc = smplfit_fit(m, n, w.stage_v[slot], has_joints ? w.stage_j[slot] : nullptr, nullptr, nullptr, nullptr,
                               nullptr, nullptr, nullptr, nullptr, o, w.pose + (size_t)lo * J * 3, w.betas + (size_t)lo * S,
                               w.trans + (size_t)lo * 3, w.orient + (size_t)lo * J * 9, w.rel + (size_t)lo * J * 9,
                               w.kid + lo, w.scale + lo, w.fit_ws[slot], w.fit_ws_bytes, cs);
    if (rc != SMPLFIT_OK) return rc;
    SF_CU(cudaEventRecord(p->done[slot], cs));
  }
  for (int i = 0; i < kSlots && i < n_chunks; ++i) SF_CU(cudaStreamWaitEvent(st, p->done[i], 0));
  // results: a few hundred bytes per instance, one copy per output for the whole batch
  auto back = [&](float* host, const float* dev, size_t per_instance) -> cudaError_t {
    if (!host) return cudaSuccess;
    return cudaMemcpyAsync(host, dev, (size_t)batch * per_instance * sizeof(float), cudaMemcpyDeviceToHost, st);
  };
  if (o->want_pose_rotvecs) SF_CU(back(host_pose_rotvecs, w.pose, J * 3));
  SF_CU(back(host_shape_betas, w.betas, S));
  SF_CU(back(host_trans, w.trans, 3));
  SF_CU(back(host_orientations, w.orient, J * 9));
  SF_CU(back(host_rel_orientations, w.rel, J * 9));
  if (o->enable_kid) SF_CU(back(host_kid_factor, w.kid, 1));
  if (o->scale_mode != 0) SF_CU(back(host_scale_corr, w.scale, 1));
#undef SF_CU
  return SMPLFIT_OK;
}
