// Micro-benchmark: how fast can one warp-per-(segment, 32-instance group) stream the two [3V] x B float arrays of the
// vertex passes, for the two candidate HBM layouts and three ways of moving the data?
//   A  row-major [row][Bp], plain LDG, lane = instance, U rows in flight per warp
//   B1 row-major, per-warp smem staging with one 128-byte cp.async.bulk per row (lanes issue them)
//   B2 row-major, per-warp smem staging with one 2D TMA box {32 floats, ROWS_PER_STAGE rows} per stage
//   C  group-major [group][row][32], per-warp smem staging with ONE contiguous cp.async.bulk per stage
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_layouts stream_layouts.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int WARPS = 12;
constexpr int VSUB = 8;              // vertices per stage
constexpr int ROWS = 3 * VSUB;       // rows per stage per array
constexpr int STAGE_FLOATS = 2 * ROWS * 32;  // both arrays

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct Args {
  const float *t, *vp;   // layout depends on the mode
  float* out;            // [n_work][32]
  int V, Bp, seg_len, n_seg, groups;
};

// mode A: LDG with U vertices in flight
template <int U>
__global__ void __launch_bounds__(WARPS * 32, 1) k_ldg(Args a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_work = a.n_seg * a.groups;
  float acc = 0.f;
  for (int w = blockIdx.x * WARPS + warp; w < n_work; w += gridDim.x * WARPS) {
    const int seg = w % a.n_seg, g = w / a.n_seg;
    const int i0 = seg * a.seg_len, i1 = min(a.V, i0 + a.seg_len);
    const float* tp = a.t + g * 32 + lane;
    const float* vp = a.vp + g * 32 + lane;
    for (int i = i0; i < i1; i += U) {
      float x[U][6];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int r = min(i + u, i1 - 1) * 3 + c;
          x[u][c] = tp[(size_t)r * a.Bp];
          x[u][3 + c] = vp[(size_t)r * a.Bp];
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc += x[u][c];
    }
    a.out[(size_t)w * 32 + lane] = acc;
  }
}

// modes B1 / B2 / C: per-warp NST-stage smem ring
template <int MODE, int NST>
__global__ void __launch_bounds__(WARPS * 32, 1) k_staged(Args a, const __grid_constant__ CUtensorMap mt, const __grid_constant__ CUtensorMap mv) {
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* buf = smem + (size_t)warp * NST * STAGE_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * NST * STAGE_FLOATS) + warp * NST;
  if (lane == 0) for (int s = 0; s < NST; ++s) mbar_init(bars + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int n_work = a.n_seg * a.groups;
  uint32_t phase = 0;
  float acc = 0.f;
  const size_t rows_total = (size_t)3 * a.V;
  for (int w = blockIdx.x * WARPS + warp; w < n_work; w += gridDim.x * WARPS) {
    const int seg = w % a.n_seg, g = w / a.n_seg;
    const int i0 = seg * a.seg_len, i1 = min(a.V, i0 + a.seg_len);
    const int nsub = (i1 - i0 + VSUB - 1) / VSUB;
    auto issue = [&](int k) {
      if (k >= nsub) return;
      const int s = k % NST;
      const int v0 = i0 + k * VSUB, nv = min(VSUB, i1 - v0);
      float* dst = buf + (size_t)s * STAGE_FLOATS;
      const uint32_t bytes = (uint32_t)nv * 3 * 128 * 2;
      if (MODE == 1) {  // one 128-byte bulk copy per row, issued by the lanes
        if (lane == 0) mbar_expect_tx(bars + s, bytes);
        __syncwarp();
        if (lane < nv * 3) {
          bulk_g2s(dst + lane * 32, a.t + (size_t)(v0 * 3 + lane) * a.Bp + g * 32, 128, bars + s);
          bulk_g2s(dst + ROWS * 32 + lane * 32, a.vp + (size_t)(v0 * 3 + lane) * a.Bp + g * 32, 128, bars + s);
        }
      } else if (MODE == 2) {  // 2D tensor-map box (rows beyond the tensor are zero-filled; bytes always full box)
        if (lane == 0) {
          mbar_expect_tx(bars + s, (uint32_t)ROWS * 128 * 2);
          tma_2d(dst, &mt, bars + s, g * 32, v0 * 3);
          tma_2d(dst + ROWS * 32, &mv, bars + s, g * 32, v0 * 3);
        }
      } else {  // group-major: contiguous
        if (lane == 0) {
          mbar_expect_tx(bars + s, bytes);
          bulk_g2s(dst, a.t + ((size_t)g * rows_total + (size_t)v0 * 3) * 32, bytes / 2, bars + s);
          bulk_g2s(dst + ROWS * 32, a.vp + ((size_t)g * rows_total + (size_t)v0 * 3) * 32, bytes / 2, bars + s);
        }
      }
    };
    for (int k = 0; k < NST - 1; ++k) issue(k);
    for (int k = 0; k < nsub; ++k) {
      const int s = k % NST;
      __syncwarp();  // everyone is done with stage (k-1) % NST before it is refilled
      issue(k + NST - 1);
      mbar_wait(bars + s, (phase >> s) & 1u);
      phase ^= 1u << s;
      const float* src = buf + (size_t)s * STAGE_FLOATS;
      const int nv = min(VSUB, i1 - (i0 + k * VSUB));
      for (int r = 0; r < nv * 3; ++r) acc += src[r * 32 + lane] + src[ROWS * 32 + r * 32 + lane];
    }
    a.out[(size_t)w * 32 + lane] = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int V = 6890, Bp = 4096, seg_len = 58, groups = Bp / 32;
  const int n_seg = (V + seg_len - 1) / seg_len;
  const size_t n = (size_t)3 * V * Bp;
  float *t, *vp, *out;
  CK(cudaMalloc(&t, n * 4)); CK(cudaMalloc(&vp, n * 4)); CK(cudaMalloc(&out, (size_t)n_seg * groups * 32 * 4));
  CK(cudaMemset(t, 0, n * 4)); CK(cudaMemset(vp, 0, n * 4));
  Args a{t, vp, out, V, Bp, seg_len, n_seg, groups};
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap mt, mv;
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[2] = {(cuuint64_t)Bp, (cuuint64_t)3 * V};
    cuuint64_t strides[1] = {(cuuint64_t)Bp * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)ROWS};
    cuuint32_t el[2] = {1, 1};
    CUresult r = enc(which ? &mv : &mt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, which ? vp : t, dims, strides, box, el,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = 2.0 * n * 4;
  auto time = [&](const char* name, auto launch) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("%-44s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6);
  };
  const int grid = 148;
  time("A  LDG U=1", [&] { k_ldg<1><<<grid, WARPS * 32>>>(a); });
  time("A  LDG U=2", [&] { k_ldg<2><<<grid, WARPS * 32>>>(a); });
  time("A  LDG U=4", [&] { k_ldg<4><<<grid, WARPS * 32>>>(a); });
  time("A  LDG U=8", [&] { k_ldg<8><<<grid, WARPS * 32>>>(a); });
#define STAGED(MODE, NST, label)                                                                         \
  {                                                                                                      \
    const size_t smem = (size_t)WARPS * NST * STAGE_FLOATS * 4 + WARPS * NST * 8 + 64;                   \
    CK(cudaFuncSetAttribute(k_staged<MODE, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    time(label, [&] { k_staged<MODE, NST><<<grid, WARPS * 32, smem>>>(a, mt, mv); });                    \
  }
  STAGED(1, 2, "B1 row-major, 128B bulk per row, 2 stages");
  STAGED(1, 3, "B1 row-major, 128B bulk per row, 3 stages");
  STAGED(2, 2, "B2 row-major, 2D TMA box, 2 stages");
  STAGED(2, 3, "B2 row-major, 2D TMA box, 3 stages");
  STAGED(3, 2, "C  group-major, contiguous bulk, 2 stages");
  STAGED(3, 3, "C  group-major, contiguous bulk, 3 stages");
  return 0;
}
