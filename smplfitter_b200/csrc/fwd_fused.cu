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
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 8;
constexpr int CPW = NCH * 4 / EPI_WARPS;      // chunks per epilogue warp and tile (2)
constexpr int THREADS = 128 + EPI_WARPS * 32;
constexpr int PITCH = 3 * CH + 2;             // staging row pitch in floats (even: 8-byte aligned rows)
constexpr int STG_FLOATS = 32 * PITCH;
constexpr int REC_WORDS = 8;                  // per-vertex record: w[4] | pack | v_rest[3]
constexpr int WARP_SMEM = STG_FLOATS * 4 + CPW * CH * REC_WORDS * 4 + 16;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * WARP_SMEM + 256 + 1024;
constexpr uint32_t TMEM_COLS = 512;

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

struct FusedMaps {
  CUtensorMap f_hi, f_lo, p_hi, p_lo;
};

struct FusedArgs {
  const float4* quads;     // [J*3][Bt] skinning rows (G[c][0..2], t[c]) per instance
  const uint32_t* vrec;    // [tiles_n * TILE_V][REC_WORDS] per-vertex records in processing order
  float* out;              // (B,V,3)
  float inv_scale;         // 2^-s
  int V, B, Bt, k_blocks, tiles_m, total_tiles, aligned8;
};

__global__ void __launch_bounds__(THREADS, 1) k_fwd_fused(const __grid_constant__ FusedMaps maps, const FusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* warp_area = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(warp_area + EPI_WARPS * WARP_SMEM);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;  // [2] accumulator ready for the epilogue
  uint64_t* acc_empty = acc_full + 2;   // [2] accumulator drained (EPI_WARPS arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
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
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp >= 4 && lane == 0) {
    uint64_t* rbar = reinterpret_cast<uint64_t*>(warp_area + (warp - 4) * WARP_SMEM + STG_FLOATS * 4 + CPW * CH * REC_WORDS * 4);
    mbar_init(rbar, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      int it = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int b0 = (tile % a.tiles_m) * TILE_M, n0 = (tile / a.tiles_m) * TILE_N;
        for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
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
        mbar_wait(&acc_empty[as], aph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_N);
        uint32_t acc = 0;
        for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
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
  } else if (warp >= 4) {
    // ---- epilogue warps: skinning out of TMEM ----
    const int e = warp - 4;
    const int q = warp & 3;    // TMEM lane quadrant this warp may read (warp id % 4)
    const int h = e >> 2;      // which CPW chunks of the tile
    float* stg = reinterpret_cast<float*>(warp_area + e * WARP_SMEM);
    uint32_t* recs = reinterpret_cast<uint32_t*>(stg + STG_FLOATS);
    uint64_t* rbar = reinterpret_cast<uint64_t*>(recs + CPW * CH * REC_WORDS);
    // write-out mapping: 4 rows x 24 float2 = 96 float2 = 3 warp-wide 8-byte accesses
    int wo_row[3], wo_e[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int n = t * 32 + lane;
      wo_row[t] = n / (3 * CH / 2);
      wo_e[t] = n - wo_row[t] * (3 * CH / 2);
    }
    float4 cq[4][3];  // cached joint rows per skinning slot
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) cq[k][c] = make_float4(0.f, 0.f, 0.f, 0.f);
    int tcount = 0;
    uint32_t rphase = 0;
    const float inv_scale = a.inv_scale;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const int tm = tile % a.tiles_m, tn = tile / a.tiles_m;
      const int b = tm * TILE_M + q * 32 + lane;  // this lane's instance (< Bt)
      const int vfirst = tn * TILE_V + h * CPW * CH;  // first vertex (processing order == model order per chunk)
      // this warp's records of the tile: one bulk copy, hidden behind the wait for the accumulator
      if (lane == 0) {
        mbar_expect_tx(rbar, CPW * CH * REC_WORDS * 4);
        bulk_g2s(recs, a.vrec + (size_t)vfirst * REC_WORDS, CPW * CH * REC_WORDS * 4, rbar);
      }
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(rbar, rphase);
      rphase ^= 1u;
      const float4* qb = a.quads + b;
#pragma unroll 1
      for (int ch = 0; ch < CPW; ++ch) {
        const int v0 = vfirst + ch * CH;
        if (v0 >= a.V) break;
        uint32_t r[3 * CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TILE_N + (h * CPW + ch) * 3 * CH);
        SF_TMEM_LD16(r, 0, taddr);
        SF_TMEM_LD16(r, 16, taddr + 16);
        SF_TMEM_LD16(r, 32, taddr + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint4* rc = reinterpret_cast<const uint4*>(recs + ch * CH * REC_WORDS);
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const uint4 w4 = rc[2 * u];
          const uint4 m4 = rc[2 * u + 1];
          const uint32_t pack = m4.x;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (pack & (1u << (24 + k))) {  // warp-uniform: slot k takes another joint at this vertex
              const int j = (pack >> (6 * k)) & 63;
#pragma unroll
              for (int c = 0; c < 3; ++c) cq[k][c] = __ldg(qb + (size_t)(j * 3 + c) * a.Bt);
            }
          }
          const float x0 = fmaf(__uint_as_float(r[3 * u]), inv_scale, __uint_as_float(m4.y));
          const float x1 = fmaf(__uint_as_float(r[3 * u + 1]), inv_scale, __uint_as_float(m4.z));
          const float x2 = fmaf(__uint_as_float(r[3 * u + 2]), inv_scale, __uint_as_float(m4.w));
          float2 B2[6];
          {
            const float w = __uint_as_float(w4.x);
            const float2 ww = make_float2(w, w);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              B2[2 * c] = __fmul2_rn(ww, make_float2(cq[0][c].x, cq[0][c].y));
              B2[2 * c + 1] = __fmul2_rn(ww, make_float2(cq[0][c].z, cq[0][c].w));
            }
          }
          const float wk[3] = {__uint_as_float(w4.y), __uint_as_float(w4.z), __uint_as_float(w4.w)};
#pragma unroll
          for (int k = 1; k < 4; ++k) {
            if (wk[k - 1] != 0.f) {  // warp-uniform
              const float2 ww = make_float2(wk[k - 1], wk[k - 1]);
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                B2[2 * c] = __ffma2_rn(ww, make_float2(cq[k][c].x, cq[k][c].y), B2[2 * c]);
                B2[2 * c + 1] = __ffma2_rn(ww, make_float2(cq[k][c].z, cq[k][c].w), B2[2 * c + 1]);
              }
            }
          }
          float* dst = stg + lane * PITCH + 3 * (pack >> 28);
#pragma unroll
          for (int c = 0; c < 3; ++c)
            dst[c] = fmaf(B2[2 * c].x, x0, fmaf(B2[2 * c].y, x1, fmaf(B2[2 * c + 1].x, x2, B2[2 * c + 1].y)));
        }
        __syncwarp();
        // write-out: rows = instances tm*128 + q*32 + r, 3*CH contiguous floats each at vertex v0
        const int b_row0 = tm * TILE_M + q * 32;
        if (a.aligned8 && v0 + CH <= a.V) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int rr = 4 * g + wo_row[t];
              const int bb = b_row0 + rr;
              const float2 v = *reinterpret_cast<const float2*>(stg + rr * PITCH + 2 * wo_e[t]);
              if (bb < a.B) *reinterpret_cast<float2*>(a.out + ((size_t)bb * a.V + v0) * 3 + 2 * wo_e[t]) = v;
            }
          }
        } else {
          const int width = 3 * min(CH, a.V - v0);
          for (int idx = lane; idx < 32 * width; idx += 32) {
            const int rr = idx / width, ee = idx - rr * width;
            const int bb = b_row0 + rr;
            if (bb < a.B) a.out[((size_t)bb * a.V + v0) * 3 + ee] = stg[rr * PITCH + ee];
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
#pragma unroll
      for (int c = 0; c < 3; ++c)
        a.quads[(size_t)(j * 3 + c) * a.Bt + b] =
            make_float4(G[c * 3], G[c * 3 + 1], G[c * 3 + 2], (pos[j * 3 + c] - rj[c]) + tr[c]);
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
    if (cudaFuncSetAttribute(k_fwd_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      return fail(SMPLFIT_ERR_CUDA, "cannot reserve shared memory for k_fwd_fused");
    }
    attr_set = true;
  }
  FusedArgs fa;
  fa.quads = w.quads; fa.vrec = m->fwd_vrec; fa.out = out_vertices; fa.inv_scale = ldexpf(1.f, -m->fwd_scale_log2);
  fa.V = V; fa.B = B; fa.Bt = w.Bt; fa.k_blocks = m->fwd_kf / KB; fa.tiles_m = tiles_m; fa.total_tiles = tiles_m * tiles_n;
  fa.aligned8 = (V % 2 == 0) && ((reinterpret_cast<uintptr_t>(out_vertices) & 7) == 0);
  const int grid = fa.total_tiles < sm_count() ? fa.total_tiles : sm_count();
  SF_LAUNCH(k_fwd_fused, grid, THREADS, SMEM_BYTES, st, maps, fa);
  return SMPLFIT_OK;
}

}  // namespace sf
