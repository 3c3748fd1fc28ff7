// Tensor-core path of the pose-blend-shape contraction
//     v_posed^T[n][b] = v_template_fit[n] + sum_k posedirs_fit[n][k] feat[b][k]
// (pt/bodymodel.py:293, pt/bodyfitter.py:913-916) -- the one genuinely dense GEMM of the path:
// (3V x P) x (P x B), P = 9 (J-1).
//
// sm_100a design: persistent CTAs (one per SM) loop over 128-instance x 256-row tiles.
// D[128][256] accumulates in TMEM (tcgen05.mma, cta_group::1, kind::tf32, M = 128, N = 256, K = 8
// per instruction); the 512 TMEM columns hold two accumulators so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Both operands are K-major fp32 tiles of 16 floats per row (= one
// 64-byte swizzle atom row) brought in by TMA (cp.async.bulk.tensor.2d, SWIZZLE_64B) through a
// 4-stage mbarrier ring (48 KB per stage).  Instances are the M side so that in the epilogue TMEM
// lane == instance: each warp stores 32 consecutive instances of one v_posed^T row = one
// coalesced 128-byte line per column.  Tiles are ordered so CTAs working at the same time share
// the posedirs tile in L2.
//
// Precision: TF32 keeps 10 mantissa bits, not enough for the 1e-4 parity gate, so every
// operand is split x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo = x - hi) and
// the product is accumulated as hi*hi + hi*lo + lo*hi in the fp32 accumulator (error ~2^-21
// relative, i.e. fp32-GEMM grade).  posedirs hi/lo are model constants; the per-call feature
// split is a tiny elementwise kernel.
//
// Warp roles (256 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (warp 4+q owns TMEM lanes 32q..32q+31).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "fwd_fused.cuh"
#include "vposed_tc.cuh"

namespace sf {

namespace {

constexpr int TILE_M = 128;   // instances per tile (TMEM lanes)
constexpr int TILE_N = 256;   // v_posed^T rows (vertex coordinates) per tile
// k-block: 16 floats = 64 bytes per row (SWIZZLE_64B atoms) and a 4-stage ring.  With 32-float k-blocks only two
// 96 KB stages fit and the kernel ran at period = TMA latency (tensor pipe 58 % active, L2->SM traffic at 27 % of its
// peak: latency-, not bandwidth-bound); four 48 KB stages keep three loads in flight.  -DSMPLFIT_TC_K32 restores the
// 128-byte layout.
#ifdef SMPLFIT_TC_K32
constexpr int TILE_K = 32;
constexpr int STAGES = 2;
#else
constexpr int TILE_K = 16;
constexpr int STAGES = 4;
#endif
constexpr int K_PAD = 32;     // K padding of every operand array (model constants are laid out with it)
constexpr int ROW_BYTES = TILE_K * 4;  // bytes per tile row = swizzle span
constexpr int A_BYTES = TILE_M * TILE_K * 4;     // 16 KB per feature tile
constexpr int B_BYTES = TILE_N * TILE_K * 4;     // 32 KB per posedirs tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // F_hi, F_lo, P_hi, P_lo = 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 512;              // two 128 x 256 fp32 accumulators
constexpr int THREADS = 384;                     // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-7 epilogue, 8-11 hi/lo split
constexpr int CONV_THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 8 rows x ROW_BYTES >> 4 in [32,46), version 1 in [46,48),
// layout type in [61,64): SWIZZLE_128B = 2, SWIZZLE_64B = 4.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
  return d;
}
// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both
// K-major, N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TcMaps {
  CUtensorMap f_hi, f_lo, p_hi, p_lo;
};

// SPLIT = true: the operands arrive as plain fp32 (maps.f_hi = features [Bp][Kp], maps.p_hi = P [rows][Kp]) and four
// converter warps split each stage into its tf32-exact hi / lo parts in shared memory (hi in place, lo into the
// neighbouring slot), which cuts the L2 -> SM traffic of a k-block from 48 KB to 24 KB (the pre-split version moves
// 1.79 GB per launch at 8.8 TB/s).  Experiment (SMPLFIT_B200_GEMM_SPLIT=1): correct, but the extra pipeline stage costs
// more than the traffic saves (0.233 vs 0.201 ms) -- the kernel is not L2-bound; neither did 4 x 48 KB stages beat
// 2 x 96 KB ones, so it is not TMA-latency-bound either.  At 569 TFLOP/s of TF32 work it sits at ~70 % of what the
// tensor pipe sustains on this part under power limits (bf16 measures 1624 of 2250 nominal TFLOP/s).
template <bool BIAS, bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_vposed_tc(const __grid_constant__ TcMaps maps, const float* __restrict__ vt, float* __restrict__ out, int M_rows,
            int Bp, int k_blocks, int tiles_m, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;   // [2] accumulator ready for the epilogue
  uint64_t* acc_empty = acc_full + 2;    // [2] accumulator drained (4 epilogue warps arrive)
  uint64_t* conv = acc_empty + 2;        // [STAGES] stage split into hi / lo (SPLIT only; 4 converter warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(conv + STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.f_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.f_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.p_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.p_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&conv[s], CONV_THREADS / 32);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: persistent over this CTA's tiles ----
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b0 = (tile % tiles_m) * TILE_M, n0 = (tile / tiles_m) * TILE_N;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          if (SPLIT) {
            mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
            tma_load_2d(st, &maps.f_hi, &full[s], kb * TILE_K, b0);
            tma_load_2d(st + 2 * A_BYTES, &maps.p_hi, &full[s], kb * TILE_K, n0);
          } else {
            mbar_expect_tx(&full[s], STAGE_BYTES);
            tma_load_2d(st, &maps.f_hi, &full[s], kb * TILE_K, b0);
            tma_load_2d(st + A_BYTES, &maps.f_lo, &full[s], kb * TILE_K, b0);
            tma_load_2d(st + 2 * A_BYTES, &maps.p_hi, &full[s], kb * TILE_K, n0);
            tma_load_2d(st + 2 * A_BYTES + B_BYTES, &maps.p_lo, &full[s], kb * TILE_K, n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int as = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_N);
        uint32_t acc = 0;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(SPLIT ? &conv[s] : &full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t fhi = make_desc(base), flo = make_desc(base + A_BYTES);
          const uint64_t phi = make_desc(base + 2 * A_BYTES), plo = make_desc(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < TILE_K / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per K = 8 step inside the swizzle atom
            mma_tf32(d_tmem, flo + adv, phi + adv, acc);  // small terms first
            acc = 1;
            mma_tf32(d_tmem, fhi + adv, plo + adv, 1);
            mma_tf32(d_tmem, fhi + adv, phi + adv, 1);
          }
          mma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
        }
        mma_commit(&acc_full[as]);
      }
    }
    __syncwarp();
  } else if (SPLIT && warp >= 8) {
    // ---- converter warps: x -> (hi, lo) of every fp32 word of the stage (elementwise, so the swizzled layout does
    // not matter); generic-proxy writes are fenced towards the async proxy before the MMA issuer is released ----
    const int ct = threadIdx.x - 8 * 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        uint8_t* st = smem + s * STAGE_BYTES;
        float4* fh = reinterpret_cast<float4*>(st);
        float4* fl = reinterpret_cast<float4*>(st + A_BYTES);
        float4* phh = reinterpret_cast<float4*>(st + 2 * A_BYTES);
        float4* pl = reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
        for (int q = 0; q < A_BYTES / 16 / CONV_THREADS; ++q) {
          const int i = q * CONV_THREADS + ct;
          const float4 x = fh[i];
          float4 h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          fh[i] = h;
          fl[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
#pragma unroll
        for (int q = 0; q < B_BYTES / 16 / CONV_THREADS; ++q) {
          const int i = q * CONV_THREADS + ct;
          const float4 x = phh[i];
          float4 h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          phh[i] = h;
          pl[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue warps: TMEM -> registers -> + v_template -> coalesced global stores ----
    const int q = warp - 4;  // TMEM lane quadrant (warp id % 4)
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const int b0 = (tile % tiles_m) * TILE_M, n0 = (tile / tiles_m) * TILE_N;
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int b = b0 + q * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < TILE_N; c0 += 32) {
        if (n0 + c0 >= M_rows) break;
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TILE_N + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + c0 + j;
          if (n < M_rows && b < Bp) out[(size_t)n * Bp + b] = BIAS ? __uint_as_float(r[j]) + __ldg(vt + n) : __uint_as_float(r[j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// feat [Bp][Kp] -> hi / lo [Bp][Kt] (Kt = k_blocks * 32, zero padded)
__global__ void k_split_feat(const float* __restrict__ feat, int Bp, int Kp, int Kt, float* __restrict__ hi,
                             float* __restrict__ lo) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Bp * Kt) return;
  const int b = (int)(idx / Kt), k = (int)(idx % Kt);
  const float x = (k < Kp && b < Bp) ? feat[(size_t)b * Kp + k] : 0.f;
  const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  hi[idx] = h;
  lo[idx] = x - h;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint64_t ld = 0) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if (ld == 0) ld = cols;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TILE_K, (cuuint32_t)box_rows};
  cuuint32_t elem[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem,
            CU_TENSOR_MAP_INTERLEAVE_NONE, ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_tc_mode = -1;  // -1: read SMPLFIT_B200_GEMM env on first use; 0 = SIMT, 1 = tcgen05

}  // namespace

static bool tc_enabled() {
  if (g_tc_mode < 0) {
    const char* e = getenv("SMPLFIT_B200_GEMM");
    g_tc_mode = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return g_tc_mode == 1;
}

static int sm_count_tc() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

int tc_tile_k() { return K_PAD; }  // K padding of operand arrays (a multiple of the k-block)
int tc_tile_m() { return TILE_M; }
int tc_tile_n() { return TILE_N; }

// out[n][b] = bias[n] + sum_k (p_hi + p_lo)[n][k] (f_hi + f_lo)[b][k]   (3xTF32, fp32 accumulate in TMEM)
// p_*: [rows_alloc][Kt] (rows_alloc >= rows), f_*: [Bt][Kt] (Bt = roundup(Bp, 128)), out: [rows][Bp]
bool tc_gemm_run(const float* p_hi, const float* p_lo, int rows, int rows_alloc, int Kt, const float* bias,
                 const float* f_hi, const float* f_lo, int Bt, float* out, int Bp, cudaStream_t st) {
  if (!tc_enabled() || !p_hi || !p_lo || !f_hi || !f_lo || Kt % K_PAD != 0 || Bt % TILE_M != 0) return false;
  TcMaps maps;
  if (!make_map(&maps.f_hi, f_hi, Bt, Kt, TILE_M) || !make_map(&maps.f_lo, f_lo, Bt, Kt, TILE_M) ||
      !make_map(&maps.p_hi, p_hi, rows_alloc, Kt, TILE_N) || !make_map(&maps.p_lo, p_lo, rows_alloc, Kt, TILE_N))
    return false;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_vposed_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_vposed_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set = true;
  }
  const int tiles_m = Bt / TILE_M, tiles_n = (rows + TILE_N - 1) / TILE_N, total = tiles_m * tiles_n;
  const int sms = sm_count_tc();
  const int grid = total < sms ? total : sms;
  if (bias) SF_LAUNCH((k_vposed_tc<true, false>), grid, THREADS, SMEM_BYTES, st, maps, bias, out, rows, Bp, Kt / TILE_K, tiles_m, total);
  else SF_LAUNCH((k_vposed_tc<false, false>), grid, THREADS, SMEM_BYTES, st, maps, bias, out, rows, Bp, Kt / TILE_K, tiles_m, total);
  return true;
}

size_t vposed_tc_scratch_bytes(const smplfit_model_t* m, int Bp) {
  const int Kt = roundup(m->num_pose_feats, K_PAD);
  const int Bt = roundup(Bp, TILE_M);
  const size_t tf32 = (size_t)2 * Bt * Kt * sizeof(float) + 512, f16 = vposed_f16_scratch_bytes(m, Bp);
  return tf32 > f16 ? tf32 : f16;
}

// SMPLFIT_B200_GEMM=tf32: the 3xTF32 kernel of this file instead of the fp16-split main loop of fwd_fused.cu (A/B runs)
static bool tf32_forced() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SMPLFIT_B200_GEMM");
    v = (e && strcmp(e, "tf32") == 0) ? 1 : 0;
  }
  return v == 1;
}

static bool vposed_tc_run_with(const smplfit_model_t* m, const float* p_hi, const float* p_lo, const float* bias,
                               const float* feat, float* vposedT, int Bp, int Kp, void* scratch, cudaStream_t st);

bool vposed_tc_run(const smplfit_model_t* m, const float* feat, float* vposedT, int Bp, int Kp, void* scratch,
                   cudaStream_t st) {
  if (tc_enabled() && !tf32_forced() && vposed_f16_run(m, feat, vposedT, Bp, Kp, scratch, st)) return true;
  return vposed_tc_run_with(m, m->posedirs_hi, m->posedirs_lo, m->v_template_fit, feat, vposedT, Bp, Kp, scratch, st);
}

// rows in MODEL vertex order (forward LBS): v_posed^T[v*3+c] = v_template[v][c] + posedirs[v][c] . feat
bool vposed_tc_run_model(const smplfit_model_t* m, const float* feat, float* vposedT, int Bp, int Kp, void* scratch,
                         cudaStream_t st) {
  return vposed_tc_run_with(m, m->posedirs_model_hi, m->posedirs_model_lo, m->v_template, feat, vposedT, Bp, Kp, scratch, st);
}

static bool vposed_tc_run_with(const smplfit_model_t* m, const float* p_hi, const float* p_lo, const float* bias,
                               const float* feat, float* vposedT, int Bp, int Kp, void* scratch, cudaStream_t st) {
  if (!tc_enabled() || p_hi == nullptr || p_lo == nullptr || scratch == nullptr) return false;
  if (encode_fn() == nullptr) return false;
  const int Kt = roundup(m->num_pose_feats, K_PAD);
  const int Bt = roundup(Bp, TILE_M);
  const int rows = 3 * m->num_vertices;
  float* hi = reinterpret_cast<float*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  float* lo = hi + (size_t)Bt * Kt;
  const size_t n = (size_t)Bt * Kt;
  // rows [Bp, Bt) and columns [Kp, Kt) of the split features are written as zeros by the kernel's bound checks
  SF_LAUNCH(k_split_feat, (unsigned)((n + 255) / 256), 256, 0, st, feat, Bp, Kp, Kt, hi, lo);
  return tc_gemm_run(p_hi, p_lo, rows, rows, Kt, bias, hi, lo, Bt, vposedT, Bp, st);
}

bool tensor_maps_available() { return encode_fn() != nullptr; }

// [rows][Bp] instance-minor array, box = {32 instances, box_rows rows}, no swizzle (the vertex streams of pass_lite.cu)
bool make_im_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t Bp, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {Bp, rows};
  cuuint64_t strides[1] = {Bp * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t elem[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the closed-form Gramian's pair term as the same GEMM (pass_lite.cu): available when the constants are present
bool gram_pairs_tc_available(const smplfit_model_t* m) {
  return tc_enabled() && m->gcf_AT_hi != nullptr && m->gcf_AT_lo != nullptr && m->gcf_npairs > 0 && encode_fn() != nullptr;
}

}  // namespace sf
