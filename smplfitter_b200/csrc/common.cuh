// Shared host-side plumbing of the C-ABI translation units: error string, launch counter,
// workspace carving.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <vector>

#include "../../include/smplfit_b200.h"

namespace sf {

extern thread_local char g_err[512];
// set by a launcher that could not enqueue its kernel (e.g. a tensor-map encode failure); the C-ABI entry point that
// called it turns it into SMPLFIT_ERR_CUDA instead of returning results computed from garbage
extern thread_local const char* g_launch_failure;
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

// Optional per-kernel timing (bench.py's roofline leg): CUDA events around every launch on the
// launching stream, aggregated by kernel name in smplfit_profile_report().
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
extern bool g_prof_on;
extern std::vector<ProfRec> g_prof;

#define SF_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
  do {                                                                             \
    cudaEvent_t ea__ = nullptr, eb__ = nullptr;                                    \
    if (sf::g_prof_on) {                                                           \
      cudaEventCreate(&ea__);                                                      \
      cudaEventCreate(&eb__);                                                      \
      cudaEventRecord(ea__, (stream));                                             \
    }                                                                              \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                    \
    if (sf::g_prof_on) {                                                           \
      cudaEventRecord(eb__, (stream));                                             \
      sf::g_prof.push_back(sf::ProfRec{#kernel, ea__, eb__});                      \
    }                                                                              \
    sf::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
  } while (0)

#define SF_CHECK_LAST()                                                            \
  do {                                                                             \
    if (sf::g_launch_failure != nullptr) {                                         \
      const char* m__ = sf::g_launch_failure;                                      \
      sf::g_launch_failure = nullptr;                                              \
      return sf::fail(SMPLFIT_ERR_CUDA, "launch failed: %s", m__);                 \
    }                                                                              \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) return sf::fail(SMPLFIT_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e__)); \
  } while (0)

inline int roundup(int x, int m) { return (x + m - 1) / m * m; }

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace sf
