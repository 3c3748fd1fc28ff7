// BodyFitter.fit for HOST-resident targets (the end-to-end entry point): the batch is cut into chunks,
// the host-to-device copies run on a library-owned copy stream ahead of the fits, consecutive chunks are fitted on
// a small ring of library-owned compute streams (so their latency-bound stages overlap), and the (small) results are copied back into the caller's host buffers.  Nothing is
// allocated per call: the staging buffers, the per-batch result buffers and the fit workspace are carved from
// the caller's device workspace; the copy stream and the events are created once per device.
// Reference behaviour: pt/bodyfitter.py:283-549 applied to every chunk (instances are independent; share_beta,
// the only cross-instance option, is rejected here).
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace sf {
namespace {

// Chunk k is staged into slot k % kSlots and fitted on that slot's own compute stream, so the short dependent
// kernels of one chunk (per-instance solves, ~1 ms of fixed latency per fit) overlap the vertex passes of its
// neighbours, and the copy stream can run up to kSlots - 1 chunks ahead of the oldest fit still in flight
// (kSlots = SMPLFIT_B200_HOST_SLOTS, default 6).
constexpr int kMaxSlots = 8;
static bool host_taper() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SMPLFIT_B200_HOST_TAPER");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}
static int host_slots() {
  static int n = 0;
  if (n == 0) {
    const char* e = getenv("SMPLFIT_B200_HOST_SLOTS");
    n = e ? atoi(e) : 6;
    if (n < 2) n = 2;
    if (n > kMaxSlots) n = kMaxSlots;
  }
  return n;
}

struct HostPipe {
  cudaStream_t copy = nullptr;
  cudaStream_t compute[kMaxSlots] = {};
  cudaEvent_t ready[kMaxSlots] = {};  // H2D into slot i complete
  cudaEvent_t done[kMaxSlots] = {};   // fit of slot i complete
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
    for (int i = 0; i < kMaxSlots; ++i) {
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
  float *stage_v[kMaxSlots], *stage_j[kMaxSlots];
  float *pose, *betas, *trans, *orient, *rel, *kid, *scale;
  void* fit_ws[kMaxSlots];
  size_t fit_ws_bytes, bytes;
};

HostWs carve_host(void* base, const smplfit_model_t* m, int64_t B, int64_t chunk, const smplfit_fit_opts_t* o,
                  int has_joints) {
  HostWs w{};
  Carver c(base);
  const size_t V = m->num_vertices, J = m->num_joints, S = m->num_betas;
  w.fit_ws_bytes = smplfit_fit_workspace_bytes(m, chunk, o, has_joints, 0, 0);
  const int kSlots = host_slots();
  const int slots = kSlots;  // (the tapered schedule has more chunks than B / chunk)
  for (int i = 0; i < kMaxSlots; ++i) {
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
  const int kSlots = host_slots();
  // chunk schedule: full chunks, then halving ones down to 64 instances.  The end-to-end time is the copy time plus
  // the latency of the LAST chunk's fit, and a fit has ~0.8 ms of fixed latency plus ~0.7 us per instance.
  std::vector<int64_t> starts;
  {
    int64_t lo = 0;
    while (lo < batch) {
      starts.push_back(lo);
      const int64_t rem = batch - lo;
      int64_t c = rem;
      if (host_taper() && rem > 64) {
        const int64_t half = ((rem / 2 + 31) / 32) * 32;
        c = half < 64 ? 64 : half;
      }
      if (c > chunk) c = chunk;
      if (c > rem) c = rem;
      lo += c;
    }
    starts.push_back(batch);
  }
  const int64_t n_chunks = (int64_t)starts.size() - 1;

#define SF_CU(x)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) return fail(SMPLFIT_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e_)); \
  } while (0)

  // SMPLFIT_B200_HOST_TRACE=1: per-chunk timeline (copy done / fit done, ms after entry) printed to stderr; debugging aid,
  // synchronises the device
  static int trace = -1;
  if (trace < 0) {
    const char* e = getenv("SMPLFIT_B200_HOST_TRACE");
    trace = (e && atoi(e) != 0) ? 1 : 0;
  }
  std::vector<cudaEvent_t> tr_copy, tr_fit;
  cudaEvent_t tr_entry = nullptr, tr_end = nullptr;
  if (trace) {
    cudaEventCreate(&tr_entry);
    cudaEventCreate(&tr_end);
    tr_copy.resize(n_chunks);
    tr_fit.resize(n_chunks);
    for (int64_t k = 0; k < n_chunks; ++k) {
      cudaEventCreate(&tr_copy[k]);
      cudaEventCreate(&tr_fit[k]);
    }
    cudaEventRecord(tr_entry, st);
  }
  auto stage = [&](int64_t k) -> cudaError_t {
    const int slot = (int)(k % kSlots);
    const int64_t lo = starts[k], n = starts[k + 1] - lo;
    cudaError_t e;
    if (k >= kSlots && (e = cudaStreamWaitEvent(p->copy, p->done[slot], 0)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(w.stage_v[slot], host_target_vertices + (size_t)lo * V * 3, (size_t)n * V * 3 * sizeof(float),
                             cudaMemcpyHostToDevice, p->copy)) != cudaSuccess)
      return e;
    if (has_joints && (e = cudaMemcpyAsync(w.stage_j[slot], host_target_joints + (size_t)lo * J * 3,
                                           (size_t)n * J * 3 * sizeof(float), cudaMemcpyHostToDevice, p->copy)) != cudaSuccess)
      return e;
    if (trace) cudaEventRecord(tr_copy[k], p->copy);
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
    const int64_t lo = starts[k], n = starts[k + 1] - lo;
    if (k + kSlots - 1 < n_chunks) SF_CU(stage(k + kSlots - 1));
    cudaStream_t cs = p->compute[slot];
    SF_CU(cudaStreamWaitEvent(cs, p->ready[slot], 0));
    const int rc = smplfit_fit(m, n, w.stage_v[slot], has_joints ? w.stage_j[slot] : nullptr, nullptr, nullptr, nullptr,
                               nullptr, nullptr, nullptr, nullptr, o, w.pose + (size_t)lo * J * 3, w.betas + (size_t)lo * S,
                               w.trans + (size_t)lo * 3, w.orient + (size_t)lo * J * 9, w.rel + (size_t)lo * J * 9,
                               w.kid + lo, w.scale + lo, w.fit_ws[slot], w.fit_ws_bytes, cs);
    if (rc != SMPLFIT_OK) return rc;
    if (trace) cudaEventRecord(tr_fit[k], cs);
    // results of this chunk: a few hundred bytes per instance, copied back on the chunk's stream (the D2H engine is
    // otherwise idle), so that only the last chunk's copy is exposed at the end
    auto back = [&](float* host, const float* dev, size_t per_instance) -> cudaError_t {
      if (!host) return cudaSuccess;
      return cudaMemcpyAsync(host + (size_t)lo * per_instance, dev + (size_t)lo * per_instance,
                             (size_t)n * per_instance * sizeof(float), cudaMemcpyDeviceToHost, cs);
    };
    if (o->want_pose_rotvecs) SF_CU(back(host_pose_rotvecs, w.pose, J * 3));
    SF_CU(back(host_shape_betas, w.betas, S));
    SF_CU(back(host_trans, w.trans, 3));
    SF_CU(back(host_orientations, w.orient, J * 9));
    SF_CU(back(host_rel_orientations, w.rel, J * 9));
    if (o->enable_kid) SF_CU(back(host_kid_factor, w.kid, 1));
    if (o->scale_mode != 0) SF_CU(back(host_scale_corr, w.scale, 1));
    SF_CU(cudaEventRecord(p->done[slot], cs));
  }
  // every chunk's results were copied back on its own stream right after its fit: join the ring
  for (int i = 0; i < kSlots && i < n_chunks; ++i) SF_CU(cudaStreamWaitEvent(st, p->done[i], 0));
  if (trace) {
    cudaEventRecord(tr_end, st);
    cudaEventSynchronize(tr_end);
    float t_end = 0.f;
    cudaEventElapsedTime(&t_end, tr_entry, tr_end);
    fprintf(stderr, "[smplfit_fit_host] %lld chunks, total %.3f ms\n", (long long)n_chunks, t_end);
    for (int64_t k = 0; k < n_chunks; ++k) {
      float tc = 0.f, tf = 0.f;
      cudaEventElapsedTime(&tc, tr_entry, tr_copy[k]);
      cudaEventElapsedTime(&tf, tr_entry, tr_fit[k]);
      fprintf(stderr, "  chunk %2lld n=%5lld copy done %.3f fit done %.3f (fit latency after copy %.3f)\n", (long long)k,
              (long long)(starts[k + 1] - starts[k]), tc, tf, tf - tc);
      cudaEventDestroy(tr_copy[k]);
      cudaEventDestroy(tr_fit[k]);
    }
    cudaEventDestroy(tr_entry);
    cudaEventDestroy(tr_end);
  }
#undef SF_CU
  return SMPLFIT_OK;
}
