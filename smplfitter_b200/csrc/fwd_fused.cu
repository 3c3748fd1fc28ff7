// BodyModel.forward (pt/bodymodel.py:121-307) as ONE tensor-core kernel: blend-shape GEMM with the skinning in its
// epilogue, so that the posed template never exists in HBM.
//
//   k_fwd_prep2 : warp per instance -- Rodrigues, level-synchronous kinematic chain, rest joints, joint outputs; writes
//                 the per-joint skinning rows [G | t] as float4 quads ([J*3][Bt], instance-minor) and the GEMM's
//                 feature rows  F[b] = [vec(R_rel[1:] - I) | betas | kid]  split into fp16 hi / lo parts.
//   k_fwd_fused : persistent CTAs over (128 instances) x (64 vertices) tiles.
//       GEMM  D[b][3v+c] = sum_k F[b][k] P[3v+c][k]  with P = 2^s [posedirs | shapedirs | kid_shapedir] (model constants,
//       fp16 hi / lo split on the host): tcgen05.mma kind::f16, M = 128, N = 192, K = 16 per instruction, three products
//       per K step (lo*hi + hi*lo + hi*hi: 22 mantissa bits, fp32-GEMM grade at half the tensor time of 3xTF32),
//       accumulators in TMEM (2 x 192 columns: the epilogue of tile i overlaps the MMAs of tile i+1), operands by TMA
//       (32-element k-blocks = one 64-byte swizzle row) through a 4-stage mbarrier ring.
//       Epilogue (8 warps; TMEM lane == instance): a warp takes 16-vertex chunks: 48 accumulator columns by
//       tcgen05.ld, x = D 2^-s + v_rest (the un-pose-corrected template, added in fp32), the <= 4 joint rows of every
//       vertex blended from a per-slot REGISTER cache (the host orders the 16 vertices of a chunk and assigns joints to
//       slots so that a slot rarely changes; a per-vertex bit mask says which slots reload, from the L2-resident quad
//       table), out = B x, and the chunk leaves through a padded shared-memory tile so that the caller's (B,V,3) rows
//       are written as contiguous 192-byte pieces.
// HBM traffic per instance: the 12 V output bytes and ~1.4 KB of rows / features.  Operand tiles stream from L2.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "fwd_fused.cuh"
#include "linalg.cuh"

namespace sf {

namespace {

constexpr int TILE_M = 128;            // instances per tile (TMEM lanes)
constexpr int CH = FWD_CHUNK;          // vertices per epilogue chunk
constexpr int NCH = 4;                 // chunks per tile
constexpr int TILE_V = CH * NCH;       // 64 vertices
constexpr int TILE_N = 3 * TILE_V;     // 192 accumulator columns
constexpr int KB = 32;                 // fp16 elements per k-block (64-byte rows, SWIZZLE_64B)
constexpr int ROW_BYTES = KB * 2;
constexpr int A_BYTES = TILE_M * ROW_BYTES;   // 8 KB per feature tile
constexpr int B_BYTES = TILE_N * ROW_BYTES;   // 12 KB per constant tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // F_hi, F_lo, P_hi, P_lo = 40 KB
constexpr int PITCH = 3 * CH + 2;             // staging row pitch in floats (even: 8-byte aligned rows)
constexpr int STG_FLOATS = 32 * PITCH;
constexpr int REC_WORDS = 8;                  // per-vertex record: w[4] | pack | v_rest[3]
constexpr int WARP_SMEM = STG_FLOATS * 4;     // per epilogue warp: the staging tile
constexpr int REC_TILE_BYTES = TILE_V * REC_WORDS * 4;  // the records of one tile (2 KB), double buffered per CTA
constexpr uint32_t TMEM_COLS = 512;
// Epilogue width: 8 warps (two per TMEM lane quadrant), 2 chunks per warp and tile.  Measured alternative: 16 warps
// (one chunk each, 640 threads, register file re-divided with setmaxnreg) ran 0.188 vs 0.177 ms -- not kept.
template <int EW>
struct Cfg {
  static constexpr int STAGES = 4;
  static constexpr int CPW = NCH * 4 / EW;
  static constexpr int THREADS = 128 + EW * 32;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EW * WARP_SMEM + 2 * REC_TILE_BYTES + 256 + 1024;
};

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
// the single-thread producer / MMA loops wait most of the time (the epilogue sets the pace): try_wait with a
// suspend-time hint parks the thread in hardware instead of polling, so the wait does not take issue slots from the
// epilogue warps of the same scheduler
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14),
// LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 8 rows x 64 B >> 4 in [32,46), version 1 in [46,48),
// layout type SWIZZLE_64B = 4 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = F16 (format 0), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define SF_TMEM_LD16(r, o, taddr)                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),           \
                 "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]),        \
                 "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])                                       \
               : "r"(taddr))

#define SF_TMEM_LD8(r, o, taddr)                                                                                \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                  \
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), \
                 "=r"(r[o + 6]), "=r"(r[o + 7])                                                                 \
               : "r"(taddr))

struct FusedMaps {
  CUtensorMap f_hi, f_lo, p_hi, p_lo;
};

struct FusedArgs {
  const float4* quads;     // [J*3][Bt] skinning rows per instance, pair layout (see k_fwd_prep2)
  const uint32_t* vrec;    // [tiles_n * TILE_V][REC_WORDS] per-vertex records in processing order
  float* out;              // MODE 0: (B,V,3);  MODE 1: [rows][Bp] instance-minor
  const float* bias;       // MODE 1: [rows] added to every column of the row
  float inv_scale;         // 2^-s
  int V, B, Bt, k_blocks, tiles_m, total_tiles, aligned8;
  int rows, Bp;            // MODE 1
};

// MODE 0: forward LBS (skinning epilogue).  MODE 1: the same GEMM with a plain epilogue, out[n][b] = bias[n] + 2^-s D[b][n]
// (the fit's posed template v_posed^T = v_template + posedirs . vec(R_rel), pt/bodyfitter.py:913-916).
template <int MODE, int EW>
__global__ void __launch_bounds__(Cfg<EW>::THREADS, 1) k_fwd_fused(const __grid_constant__ FusedMaps maps, const FusedArgs a) {
  constexpr int STAGES = Cfg<EW>::STAGES, CPW = Cfg<EW>::CPW, EPI_WARPS = EW;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared array (keeps the shared address space: LDS / STS, not
  // generic LD / ST, for everything derived from it)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* warp_area = smem + STAGES * STAGE_BYTES;
  uint8_t* rec_area = warp_area + EPI_WARPS * WARP_SMEM;  // [2][TILE_V][REC_WORDS] (MODE 0)
  uint64_t* full = reinterpret_cast<uint64_t*>(rec_area + 2 * REC_TILE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;  // [2] accumulator ready for the epilogue
  uint64_t* acc_empty = acc_full + 2;   // [2] accumulator drained (EPI_WARPS arrivals)
  uint64_t* rec_full = acc_empty + 2;   // [2] the tile's records have landed (MODE 0)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rec_full + 2);
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
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], EPI_WARPS);
      mbar_init(&rec_full[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
        const int b0 = (tile % a.tiles_m) * TILE_M, n0 = (tile / a.tiles_m) * TILE_N;
        // (only the smem stages gate these loads: the operands of the next tile are fetched while the epilogue still
        // owns both accumulators; the MMA issuer waits for the accumulator, warp 2 copies the tile's records)
        for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait_parked(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          mbar_expect_tx(&full[s], STAGE_BYTES);
          tma_load_2d(st, &maps.f_hi, &full[s], kb * KB, b0);
          tma_load_2d(st + A_BYTES, &maps.f_lo, &full[s], kb * KB, b0);
          tma_load_2d(st + 2 * A_BYTES, &maps.p_hi, &full[s], kb * KB, n0);
          tma_load_2d(st + 2 * A_BYTES + B_BYTES, &maps.p_lo, &full[s], kb * KB, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
        const int as = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait_parked(&acc_empty[as], aph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_N);
        uint32_t acc = 0;
        for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait_parked(&full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t fhi = make_desc(base), flo = make_desc(base + A_BYTES);
          const uint64_t phi = make_desc(base + 2 * A_BYTES), plo = make_desc(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per K = 16 step inside the swizzle row
            mma_f16(d_tmem, flo + adv, phi + adv, acc);  // small terms first
            acc = 1;
            mma_f16(d_tmem, fhi + adv, plo + adv, 1);
            mma_f16(d_tmem, fhi + adv, phi + adv, 1);
          }
          mma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
        }
        mma_commit(&acc_full[as]);
      }
    }
    __syncwarp();
  } else if (warp == 2 && MODE == 0) {
    if (lane == 0) {
      // ---- record producer: the tile's per-vertex records, into the buffer that goes with its accumulator (free once
      // the epilogue of two tiles ago has released it) ----
      int tcount = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
        const int as = tcount & 1;
        mbar_wait_parked(&acc_empty[as], ((tcount >> 1) & 1) ^ 1);
        mbar_expect_tx(&rec_full[as], REC_TILE_BYTES);
        bulk_g2s(rec_area + as * REC_TILE_BYTES, a.vrec + (size_t)(tile / a.tiles_m) * TILE_V * REC_WORDS, REC_TILE_BYTES,
                 &rec_full[as]);
      }
    }
    __syncwarp();
  }
  } else if (MODE == 1) {
    // ---- epilogue warps: TMEM -> + bias -> coalesced stores (lane == instance: 128-byte lines) ----
    const int e = warp - 4, q = warp & 3, h = e >> 2;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const int tm = tile % a.tiles_m, tn = tile / a.tiles_m;
      const int b = tm * TILE_M + q * 32 + lane;
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int ch = 0; ch < CPW; ++ch) {
        const int c0 = (h * CPW + ch) * 3 * CH;
        const int n0 = tn * TILE_N + c0;
        if (n0 >= a.rows) break;
        uint32_t r[3 * CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TILE_N + c0);
        SF_TMEM_LD16(r, 0, taddr);
        SF_TMEM_LD16(r, 16, taddr + 16);
        SF_TMEM_LD16(r, 32, taddr + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (b < a.Bp) {
          float* dst = a.out + (size_t)n0 * a.Bp + b;
#pragma unroll
          for (int i = 0; i < 3 * CH; ++i)
            if (n0 + i < a.rows) dst[(size_t)i * a.Bp] = fmaf(__uint_as_float(r[i]), a.inv_scale, __ldg(a.bias + n0 + i));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  } else {
    // ---- epilogue warps: skinning out of TMEM ----
    const int e = warp - 4;
    const int q = warp & 3;    // TMEM lane quadrant this warp may read (warp id % 4)
    const int h = e >> 2;      // which CPW chunks of the tile
    float* stg = reinterpret_cast<float*>(warp_area + e * WARP_SMEM);
    // write-out mappings (a chunk row = 3 CH contiguous floats of one instance):
    //   8-byte form: 4 rows x 24 float2 = 3 warp-wide accesses;  4-byte form (odd V): 2 rows x 48 floats = 3 accesses
    int wo8_s[3], wo8_g[3], wo4_s[3], wo4_g[3], wo8_r[3], wo4_r[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int n = t * 32 + lane;
      wo8_r[t] = n / (3 * CH / 2);
      const int e8 = n - wo8_r[t] * (3 * CH / 2);
      wo8_s[t] = wo8_r[t] * PITCH + 2 * e8;
      wo8_g[t] = wo8_r[t] * a.V * 3 + 2 * e8;
      wo4_r[t] = n / (3 * CH);
      const int e4 = n - wo4_r[t] * (3 * CH);
      wo4_s[t] = wo4_r[t] * PITCH + e4;
      wo4_g[t] = wo4_r[t] * a.V * 3 + e4;
    }
    // cached joint rows per skinning slot, as the pairs the blend / apply steps consume:
    //   cq[k][0..3] = (G00,G10) (G01,G11) (G02,G12) (t0,t1);  cq[k][4..5] = (G20,G21) (G22,t2)
    float2 cq[4][6];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 6; ++i) cq[k][i] = make_float2(0.f, 0.f);
    int tcount = 0;
    const float inv_scale = a.inv_scale;
    float* const my_row = stg + lane * PITCH;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const int tm = tile % a.tiles_m, tn = tile / a.tiles_m;
      const int b = tm * TILE_M + q * 32 + lane;  // this lane's instance (< Bt)
      const int vfirst = tn * TILE_V + h * CPW * CH;  // first vertex of this warp's chunks
      // per-vertex records of this warp's share of the tile (shared memory, warp-uniform 16-byte reads one vertex
      // ahead of their use)
      const uint4* rc = reinterpret_cast<const uint4*>(rec_area + as * REC_TILE_BYTES) + (h * CPW * CH) * 2;
      mbar_wait(&rec_full[as], aph);
      uint4 nw4 = rc[0], nm4 = rc[1];
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float4* qb = a.quads + b;
      const int b_row0 = tm * TILE_M + q * 32;
#pragma unroll 1
      for (int ch = 0; ch < CPW; ++ch) {
        const int v0 = vfirst + ch * CH;
        if (v0 >= a.V) break;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TILE_N + (h * CPW + ch) * 3 * CH);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {  // 8 vertices = 24 accumulator columns at a time (register budget)
        uint32_t r[3 * CH / 2];
        SF_TMEM_LD16(r, 0, taddr + half * (3 * CH / 2));
        SF_TMEM_LD8(r, 16, taddr + half * (3 * CH / 2) + 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int uu = 0; uu < CH / 2; ++uu) {
          const int u = half * (CH / 2) + uu;
          const uint4 w4 = nw4, m4 = nm4;
          if (u + 1 < CH || ch + 1 < CPW) {
            nw4 = rc[2 * (ch * CH + u) + 2];
            nm4 = rc[2 * (ch * CH + u) + 3];
          }
          const uint32_t pack = m4.x;
          // slots that take another joint at this vertex (host-replayed cache, warp-uniform, rare); at the first vertex
          // of the warp's share of a tile every slot is (re)loaded: the previous tile belonged to other instances
          const uint32_t rl = (u == 0 && ch == 0) ? 0xFu : ((pack >> 24) & 0xFu);
          if (rl) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (rl & (1u << k)) {
                const float4* src = qb + (size_t)(((pack >> (6 * k)) & 63) * 3) * a.Bt;
                const float4 A = __ldg(src), Bq = __ldg(src + a.Bt), C = __ldg(src + 2 * (size_t)a.Bt);
                cq[k][0] = make_float2(A.x, A.y); cq[k][1] = make_float2(A.z, A.w);
                cq[k][2] = make_float2(Bq.x, Bq.y); cq[k][3] = make_float2(Bq.z, Bq.w);
                cq[k][4] = make_float2(C.x, C.y); cq[k][5] = make_float2(C.z, C.w);
              }
            }
          }
          const float x0 = fmaf(__uint_as_float(r[3 * uu]), inv_scale, __uint_as_float(m4.y));
          const float x1 = fmaf(__uint_as_float(r[3 * uu + 1]), inv_scale, __uint_as_float(m4.z));
          const float x2 = fmaf(__uint_as_float(r[3 * uu + 2]), inv_scale, __uint_as_float(m4.w));
          float2 B2[6];
          {
            const float w = __uint_as_float(w4.x);
            const float2 ww = make_float2(w, w);
#pragma unroll
            for (int i = 0; i < 6; ++i) B2[i] = __fmul2_rn(ww, cq[0][i]);
          }
          const float wk[3] = {__uint_as_float(w4.y), __uint_as_float(w4.z), __uint_as_float(w4.w)};
#pragma unroll
          for (int k = 1; k < 4; ++k) {  // unused slots carry weight 0 (and finite stale rows)
            const float2 ww = make_float2(wk[k - 1], wk[k - 1]);
#pragma unroll
            for (int i = 0; i < 6; ++i) B2[i] = __ffma2_rn(ww, cq[k][i], B2[i]);
          }
          const float2 o01 = __ffma2_rn(B2[0], make_float2(x0, x0),
                                        __ffma2_rn(B2[1], make_float2(x1, x1), __ffma2_rn(B2[2], make_float2(x2, x2), B2[3])));
          const float o2 = fmaf(B2[4].x, x0, fmaf(B2[4].y, x1, fmaf(B2[5].x, x2, B2[5].y)));
          float* dst = my_row + 3 * (pack >> 28);
          dst[0] = o01.x;
          dst[1] = o01.y;
          dst[2] = o2;
        }
        }
        __syncwarp();
        // write-out: rows = instances b_row0 + r, 3*CH contiguous floats each at vertex v0
        float* gbase = a.out + ((size_t)b_row0 * a.V + v0) * 3;
        const int nrows = a.B - b_row0;  // valid rows of this 32-instance block
        if (v0 + CH <= a.V) {
          if (a.aligned8) {
            const size_t gstep = (size_t)4 * a.V * 3;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              float* gp = gbase + wo8_g[t];
              const float* sp = stg + wo8_s[t];
              if (nrows >= 32) {
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  *reinterpret_cast<float2*>(gp + g * gstep) = *reinterpret_cast<const float2*>(sp + 4 * g * PITCH);
              } else {
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  if (4 * g + wo8_r[t] < nrows)
                    *reinterpret_cast<float2*>(gp + g * gstep) = *reinterpret_cast<const float2*>(sp + 4 * g * PITCH);
              }
            }
          } else {
            const size_t gstep = (size_t)2 * a.V * 3;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              float* gp = gbase + wo4_g[t];
              const float* sp = stg + wo4_s[t];
#pragma unroll
              for (int g = 0; g < 16; ++g)
                if (2 * g + wo4_r[t] < nrows) gp[g * gstep] = sp[2 * g * PITCH];
            }
          }
        } else {  // last, partial chunk of the model
          const int width = 3 * (a.V - v0);
          for (int idx = lane; idx < 32 * width; idx += 32) {
            const int rr = idx / width, ee = idx - rr * width;
            if (b_row0 + rr < a.B) gbase[(size_t)rr * a.V * 3 + ee] = stg[rr * PITCH + ee];
          }
        }
        __syncwarp();
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

// ---------------------------------------------------------------------------------------------------------------------
// k_fwd_prep2: one warp per instance, lanes over joints.  Same arithmetic as the reference's Python loops
// (pt/bodymodel.py:223-284): rest joints j = J_t + J_s beta + kid J_kid; glob = glob[parent] rel; position chain.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PREP_WARPS = 4;

struct Prep2Args {
  const float* rot;    // see rot_mode
  const float* betas;  // (B,n_betas) or null
  const float* trans;  // (B,3) or null
  const float* kid;    // (B) or null
  const int32_t* parents;
  const float* J_template;
  const float* J_shapedirs;     // (J,3,S)
  const float* kid_J_shapedir;  // (J,3)
  float4* quads;                // [J*3][Bt] or null (joints only)
  __half* f_hi;                 // [Bt][Kf] or null
  __half* f_lo;
  float* out_joints;            // (B,J,3)
  float* out_orientations;      // (B,J,3,3)
  int rot_mode, n_betas, J, S, B, Bt, Kf, P;
};

__global__ void __launch_bounds__(PREP_WARPS * 32) k_fwd_prep2(const Prep2Args a) {
  __shared__ float s_rel[PREP_WARPS][SMPLFIT_MAX_JOINTS * 9];
  __shared__ float s_glob[PREP_WARPS][SMPLFIT_MAX_JOINTS * 9];
  __shared__ float s_rest[PREP_WARPS][SMPLFIT_MAX_JOINTS * 3];
  __shared__ float s_pos[PREP_WARPS][SMPLFIT_MAX_JOINTS * 3];
  __shared__ int s_par[SMPLFIT_MAX_JOINTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = a.J, S = a.S;
  for (int j = threadIdx.x; j < J; j += blockDim.x) s_par[j] = j == 0 ? -1 : a.parents[j];
  __syncthreads();
  const int b = blockIdx.x * PREP_WARPS + warp;
  if (b >= a.Bt) return;
  const bool live = b < a.B;
  float* rel = s_rel[warp];
  float* glob = s_glob[warp];
  float* rest = s_rest[warp];
  float* pos = s_pos[warp];
  const int nb = a.betas != nullptr ? min(a.n_betas, S) : 0;
  const float kid = (live && a.kid != nullptr) ? a.kid[b] : 0.f;
  float tr[3] = {0.f, 0.f, 0.f};
  if (live && a.trans != nullptr)
    for (int c = 0; c < 3; ++c) tr[c] = a.trans[(size_t)b * 3 + c];
  int maxd = 0;
  for (int j = lane; j < J; j += 32) {
    // rest joint and the given rotation of joint j
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __ldg(a.J_template + j * 3 + c);
      const float* js = a.J_shapedirs + ((size_t)j * 3 + c) * S;
      if (live)
        for (int s = 0; s < nb; ++s) v = fmaf(__ldg(js + s), a.betas[(size_t)b * a.n_betas + s], v);
      v = fmaf(__ldg(a.kid_J_shapedir + j * 3 + c), kid, v);
      rest[j * 3 + c] = v;
    }
    float m[9];
    if (a.rot_mode == 0) {
      float rv[3] = {0.f, 0.f, 0.f};
      if (live)
        for (int c = 0; c < 3; ++c) rv[c] = a.rot[(size_t)b * 3 * J + j * 3 + c];
      rotvec2mat(rv, m);
    } else if (a.rot_mode == 3 || !live) {
      for (int e = 0; e < 9; ++e) m[e] = (e % 4 == 0) ? 1.f : 0.f;
    } else {
      for (int e = 0; e < 9; ++e) m[e] = a.rot[((size_t)b * J + j) * 9 + e];
    }
    float* dstm = a.rot_mode == 2 ? glob : rel;  // mode 2: the global orientations are given
    for (int e = 0; e < 9; ++e) dstm[j * 9 + e] = m[e];
    int d = 0;
    for (int p = s_par[j]; p >= 0; p = s_par[p]) ++d;
    maxd = max(maxd, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxd = max(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
  __syncwarp();
  // kinematic chain, one tree depth at a time
  for (int d = 0; d <= maxd; ++d) {
    for (int j = lane; j < J; j += 32) {
      int dj = 0;
      for (int p = s_par[j]; p >= 0; p = s_par[p]) ++dj;
      if (dj != d) continue;
      const int par = s_par[j];
      if (par < 0) {
        if (a.rot_mode != 2)
          for (int e = 0; e < 9; ++e) glob[e] = rel[e];
        else
          for (int e = 0; e < 9; ++e) rel[e] = glob[e];
        for (int c = 0; c < 3; ++c) pos[c] = rest[c];
      } else {
        if (a.rot_mode != 2) mat3_mul(glob + par * 9, rel + j * 9, glob + j * 9);
        else mat3_tmul(glob + par * 9, glob + j * 9, rel + j * 9);
        const float bone[3] = {rest[j * 3] - rest[par * 3], rest[j * 3 + 1] - rest[par * 3 + 1],
                               rest[j * 3 + 2] - rest[par * 3 + 2]};
        float rb[3];
        mat3_vec(glob + par * 9, bone, rb);
        for (int c = 0; c < 3; ++c) pos[j * 3 + c] = pos[par * 3 + c] + rb[c];
      }
    }
    __syncwarp();
  }
  for (int j = lane; j < J; j += 32) {
    const float* G = glob + j * 9;
    float rj[3];
    mat3_vec(G, rest + j * 3, rj);
    if (a.quads != nullptr) {
      // pair layout of the fused epilogue: (G00,G10,G01,G11) (G02,G12,t0,t1) (G20,G21,G22,t2)
      const float t0 = (pos[j * 3] - rj[0]) + tr[0], t1 = (pos[j * 3 + 1] - rj[1]) + tr[1], t2 = (pos[j * 3 + 2] - rj[2]) + tr[2];
      a.quads[(size_t)(j * 3) * a.Bt + b] = make_float4(G[0], G[3], G[1], G[4]);
      a.quads[(size_t)(j * 3 + 1) * a.Bt + b] = make_float4(G[2], G[5], t0, t1);
      a.quads[(size_t)(j * 3 + 2) * a.Bt + b] = make_float4(G[6], G[7], G[8], t2);
    }
    if (live) {
      for (int e = 0; e < 9; ++e) a.out_orientations[((size_t)b * J + j) * 9 + e] = G[e];
      for (int c = 0; c < 3; ++c) a.out_joints[((size_t)b * J + j) * 3 + c] = pos[j * 3 + c] + tr[c];
    }
  }
  if (a.f_hi == nullptr) return;
  // feature row [vec(R_rel[1:] - I) | betas (zero padded to S) | kid | 0 ...] as fp16 hi / lo, written two halves per lane
  __half2* rh = reinterpret_cast<__half2*>(a.f_hi + (size_t)b * a.Kf);
  __half2* rl = reinterpret_cast<__half2*>(a.f_lo + (size_t)b * a.Kf);
  for (int k2 = lane; k2 < a.Kf / 2; k2 += 32) {
    float x[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int k = 2 * k2 + t;
      float v = 0.f;
      if (live) {
        if (k < a.P) {
          const int e = k % 9;
          v = rel[9 + k] - ((e % 4 == 0) ? 1.f : 0.f);
        } else if (k < a.P + S) {
          const int s = k - a.P;
          v = s < nb ? a.betas[(size_t)b * a.n_betas + s] : 0.f;
        } else if (k == a.P + S) {
          v = kid;
        }
      }
      x[t] = v;
    }
    const __half h0 = __float2half_rn(x[0]), h1 = __float2half_rn(x[1]);
    rh[k2] = __halves2half2(h0, h1);
    rl[k2] = __halves2half2(__float2half_rn(x[0] - __half2float(h0)), __float2half_rn(x[1] - __half2float(h1)));
  }
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
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 [rows][cols] row-major, box = {KB columns, box_rows rows}, SWIZZLE_64B; rows / columns past the array read as zero
bool make_h_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)KB, box_rows};
  cuuint32_t elem[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace

bool fwd_fused_available(const smplfit_model_t* m) {
  return m->fwd_P_hi != nullptr && m->fwd_P_lo != nullptr && m->fwd_vrec != nullptr && m->fwd_kf > 0 &&
         m->fwd_kf % KB == 0 && m->num_joints <= SMPLFIT_MAX_JOINTS && encode_fn() != nullptr;
}

FwdFusedWs fwd_fused_carve(void* base, const smplfit_model_t* m, int64_t B, bool with_vertices) {
  FwdFusedWs w{};
  Carver c(base);
  const size_t Bt = roundup((int)B, TILE_M);
  w.Bt = (int)Bt;
  if (with_vertices) {
    w.quads = c.take<float4>((size_t)3 * m->num_joints * Bt);
    w.f_hi = c.take<__half>(Bt * m->fwd_kf);
    w.f_lo = c.take<__half>(Bt * m->fwd_kf);
  }
  w.bytes = c.off + 256;
  return w;
}

// joints / orientations (always) and, when out_vertices != NULL, the vertices through the fused kernel
int fwd_fused_run(const smplfit_model_t* m, int B, int rot_mode, const float* rot, const float* betas, int n_betas,
                  const float* trans, const float* kid, float* out_vertices, float* out_joints, float* out_orientations,
                  const FwdFusedWs& w, cudaStream_t st) {
  Prep2Args p;
  p.rot = rot; p.betas = betas; p.trans = trans; p.kid = kid; p.parents = m->parents; p.J_template = m->J_template;
  p.J_shapedirs = m->J_shapedirs; p.kid_J_shapedir = m->kid_J_shapedir;
  p.quads = out_vertices ? w.quads : nullptr; p.f_hi = out_vertices ? w.f_hi : nullptr; p.f_lo = out_vertices ? w.f_lo : nullptr;
  p.out_joints = out_joints; p.out_orientations = out_orientations; p.rot_mode = rot_mode; p.n_betas = betas ? n_betas : 0;
  p.J = m->num_joints; p.S = m->num_betas; p.B = B; p.Bt = out_vertices ? w.Bt : B; p.Kf = m->fwd_kf; p.P = m->num_pose_feats;
  SF_LAUNCH(k_fwd_prep2, (p.Bt + PREP_WARPS - 1) / PREP_WARPS, PREP_WARPS * 32, 0, st, p);
  if (!out_vertices) return SMPLFIT_OK;
  const int V = m->num_vertices;
  const int tiles_n = (V + TILE_V - 1) / TILE_V, tiles_m = w.Bt / TILE_M;
  FusedMaps maps;
  const uint64_t p_rows = (uint64_t)tiles_n * TILE_N;
  if (!make_h_map(&maps.f_hi, w.f_hi, (uint64_t)w.Bt, (uint64_t)m->fwd_kf, TILE_M) ||
      !make_h_map(&maps.f_lo, w.f_lo, (uint64_t)w.Bt, (uint64_t)m->fwd_kf, TILE_M) ||
      !make_h_map(&maps.p_hi, m->fwd_P_hi, p_rows, (uint64_t)m->fwd_kf, TILE_N) ||
      !make_h_map(&maps.p_lo, m->fwd_P_lo, p_rows, (uint64_t)m->fwd_kf, TILE_N))
    return fail(SMPLFIT_ERR_CUDA, "tensor map encode failed (forward)");
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_fwd_fused<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<8>::SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      return fail(SMPLFIT_ERR_CUDA, "cannot reserve shared memory for k_fwd_fused");
    }
    attr_set = true;
  }
  FusedArgs fa{};
  fa.quads = w.quads; fa.vrec = m->fwd_vrec; fa.out = out_vertices; fa.inv_scale = ldexpf(1.f, -m->fwd_scale_log2);
  fa.V = V; fa.B = B; fa.Bt = w.Bt; fa.k_blocks = m->fwd_kf / KB; fa.tiles_m = tiles_m; fa.total_tiles = tiles_m * tiles_n;
  fa.aligned8 = (V % 2 == 0) && ((reinterpret_cast<uintptr_t>(out_vertices) & 7) == 0);
  const int grid = fa.total_tiles < sm_count() ? fa.total_tiles : sm_count();
  SF_LAUNCH((k_fwd_fused<0, 8>), grid, Cfg<8>::THREADS, Cfg<8>::SMEM_BYTES, st, maps, fa);
  return SMPLFIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The fit's pose-blend contraction on the same main loop: v_posed^T [3V][Bp] = v_template_fit + posedirs_fit . vec(R_rel)
// with the constants of smplfit_model_t::fit_P_* (rows in the fit's internal vertex order).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
// feat [Bp][Kp] fp32 -> fp16 hi / lo [Bt][Kf], zero padded (rows >= Bp, columns >= Kp)
__global__ void k_split_feat_h(const float* __restrict__ feat, int Bp, int Kp, int Kf, size_t n2, __half2* __restrict__ hi,
                               __half2* __restrict__ lo) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n2) return;
  const int k2 = (int)(idx % (Kf / 2));
  const size_t b = idx / (Kf / 2);
  float x[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int k = 2 * k2 + t;
    x[t] = (k < Kp && b < (size_t)Bp) ? feat[b * Kp + k] : 0.f;
  }
  const __half h0 = __float2half_rn(x[0]), h1 = __float2half_rn(x[1]);
  hi[idx] = __halves2half2(h0, h1);
  lo[idx] = __halves2half2(__float2half_rn(x[0] - __half2float(h0)), __float2half_rn(x[1] - __half2float(h1)));
}
}  // namespace

bool vposed_f16_available(const smplfit_model_t* m) {
  return m->fit_P_hi != nullptr && m->fit_P_lo != nullptr && m->fit_kf > 0 && m->fit_kf % KB == 0 && encode_fn() != nullptr;
}

size_t vposed_f16_scratch_bytes(const smplfit_model_t* m, int Bp) {
  if (!m->fit_P_hi || m->fit_kf <= 0) return 0;
  return (size_t)2 * roundup(Bp, TILE_M) * m->fit_kf * sizeof(__half) + 512;
}

bool vposed_f16_run(const smplfit_model_t* m, const float* feat, float* vposedT, int Bp, int Kp, void* scratch,
                    cudaStream_t st) {
  if (!vposed_f16_available(m) || scratch == nullptr) return false;
  const int Bt = roundup(Bp, TILE_M), Kf = m->fit_kf, rows = 3 * m->num_vertices;
  __half* hi = reinterpret_cast<__half*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  __half* lo = hi + (size_t)Bt * Kf;
  const int tiles_n = (rows + TILE_N - 1) / TILE_N, tiles_m = Bt / TILE_M;
  FusedMaps maps;
  if (!make_h_map(&maps.f_hi, hi, (uint64_t)Bt, (uint64_t)Kf, TILE_M) || !make_h_map(&maps.f_lo, lo, (uint64_t)Bt, (uint64_t)Kf, TILE_M) ||
      !make_h_map(&maps.p_hi, m->fit_P_hi, (uint64_t)tiles_n * TILE_N, (uint64_t)Kf, TILE_N) ||
      !make_h_map(&maps.p_lo, m->fit_P_lo, (uint64_t)tiles_n * TILE_N, (uint64_t)Kf, TILE_N))
    return false;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_fwd_fused<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<8>::SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set = true;
  }
  const size_t n2 = (size_t)Bt * (Kf / 2);
  SF_LAUNCH(k_split_feat_h, (unsigned)((n2 + 255) / 256), 256, 0, st, feat, Bp, Kp, Kf, n2, reinterpret_cast<__half2*>(hi),
            reinterpret_cast<__half2*>(lo));
  FusedArgs fa{};
  fa.out = vposedT; fa.bias = m->v_template_fit; fa.inv_scale = ldexpf(1.f, -m->fit_scale_log2);
  fa.V = m->num_vertices; fa.B = Bp; fa.Bt = Bt; fa.k_blocks = Kf / KB; fa.tiles_m = tiles_m; fa.total_tiles = tiles_m * tiles_n;
  fa.rows = rows; fa.Bp = Bp;
  const int grid = fa.total_tiles < sm_count() ? fa.total_tiles : sm_count();
  SF_LAUNCH((k_fwd_fused<1, 8>), grid, Cfg<8>::THREADS, Cfg<8>::SMEM_BYTES, st, maps, fa);
  return true;
}

}  // namespace sf
