// The two vertex passes of one BodyFitter.fit iteration as tensor-core kernels: the blend-shape GEMM
//   x[b][v] = v_rest_v + posedirs_v . vec(R_rel(b) - I) (+ shapedirs_v . beta(b))          (pt/bodyfitter.py:913-916)
// runs on tcgen05 with the accumulators in TMEM, and the per-vertex work that consumes x runs in the EPILOGUE, straight
// out of TMEM, so the posed template never exists in HBM (it used to be written once and read twice per iteration):
//   MODE 2 (shape stage, pt/bodyfitter.py:999-1048 restated as in lite_kernels.cuh):  b_v = t_v - (Rb_v x_v + Tb_v),
//          r[s] += S_vs . Rb_v^T b_v,  Sb += b_v,  Y_k += w_vk b_v          -> partials [segment][NS + 3 + 3 slots][Bp]
//   MODE 3 (rotation-stage statistics, _part_sums pt/bodyfitter.py:235-280, x includes S beta):
//          ref_v = Ab_v x_v + tau_v,  M += (t - ct)(ref - ca)^T, sums        -> partials [segment][16][Bp]
// Both partial layouts are the ones k_shape_lite / k_stats_lite write, so the per-instance solve kernels are unchanged.
//
// Structure (same main loop as csrc/fwd_fused.cu): persistent CTAs, each a contiguous range of (128 instances) x
// (64 vertex slots = 2 statistics segments) tiles in instance-tile-major order, so an epilogue warp keeps walking the
// segments of the SAME 32 instances and its register cache of joint rows survives from tile to tile.
//   warp 0 lane 0 : TMA producer (feature tile F hi/lo, constant tile P hi/lo per 32-element k-block; the tile's slot
//                   records and shape directions by bulk copy)
//   warp 1 lane 0 : tcgen05.mma kind::f16, M = 128, N = 192, three products per K step (lo*hi + hi*lo + hi*hi)
//   warps 4..11   : epilogue; TMEM lane == instance; warp (quadrant q, half h) takes segment 2 tn + h of its 32 instances:
//                   targets by coalesced loads one 4-vertex group ahead (instance-minor [3V][Bp] array), 12 accumulator
//                   columns per tcgen05.ld, joint rows [R | T] blended from the per-slot register cache.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "fit_fused.cuh"

namespace sf {

namespace {

constexpr int TILE_M = 128;
constexpr int SEGV = 32;               // vertex slots per segment
constexpr int TILE_V = 2 * SEGV;       // 64 slots per tile
constexpr int TILE_N = 3 * TILE_V;     // 192 accumulator columns
constexpr int KB = 32;                 // fp16 elements per k-block (64-byte rows, SWIZZLE_64B)
constexpr int ROW_BYTES = KB * 2;
constexpr int A_BYTES = TILE_M * ROW_BYTES;
constexpr int B_BYTES = TILE_N * ROW_BYTES;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 40 KB
constexpr int STAGES = 2;              // the epilogue sets the pace: two operand stages keep the MMAs fed
constexpr int REC_WORDS = 8;
constexpr int REC_TILE_BYTES = TILE_V * REC_WORDS * 4;  // 2 KB
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 128 + EPI_WARPS * 32;
constexpr int NSLOT = 12;              // == LITE_NSLOT / N_SLOTS of pt/bodymodel.py
constexpr int YW_FLOATS = NSLOT * 3 * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr int GV = 8;                  // vertices per epilogue step = one staged target block (24 accumulator columns)
constexpr int TRING = 3;               // staged target blocks per segment chain (per half h)
constexpr int TSLOT_FLOATS = 4 * 3 * GV * 32;  // [quadrant][3 GV rows][32 instances]
constexpr int TSLOT_BYTES = TSLOT_FLOATS * 4;  // 12 KB

__host__ __device__ constexpr int sd_tile_bytes(int sdl) { return TILE_V * sdl * 4; }
__host__ __device__ constexpr int fused_smem_bytes(int mode, int sdl) {
  return STAGES * STAGE_BYTES + 2 * TRING * TSLOT_BYTES + 2 * REC_TILE_BYTES +
         (mode == 2 ? 2 * sd_tile_bytes(sdl) + EPI_WARPS * (YW_FLOATS * 4 + 64) : 0) + 256 + 1024;
}

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
// producer / MMA threads wait most of the time (the epilogue sets the pace): the suspend-time hint parks them
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
// K-major SWIZZLE_64B shared-memory matrix descriptor (see fwd_fused.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
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

#define SF_TMEM_LD8(r, o, taddr)                                                                                \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                  \
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), \
                 "=r"(r[o + 6]), "=r"(r[o + 7])                                                                 \
               : "r"(taddr))
#define SF_TMEM_LD16(r, o, taddr)                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),           \
                 "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]),        \
                 "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])                                       \
               : "r"(taddr))

struct FusedMaps {
  CUtensorMap f_hi, f_lo, p_hi, p_lo;
  CUtensorMap t;  // targets [3V][Bp] fp32, box {32 instances, 3 GV rows}
};

struct FitFusedArgs {
  const float* tT;          // [3V][Bp] centred targets, internal vertex order
  const float* vwT;         // [V][Bp] per-vertex weights (MODE 3, WEIGHTED) or null
  const float4* quads;      // [J*3][Bp]: MODE 2 (R[c][0..2], T0[c]) = RT12;  MODE 3 (A[c][0..2], tau[c]) = skin4
  const uint32_t* rec;      // [nseg_pad*32][8]
  const float* sd;          // [nseg_pad*32][sdl]  (MODE 2)
  const int32_t* seg_start;
  const int32_t* seg_part;
  const int32_t* part_flags;
  const int32_t* seg_slots; // [n_segments][NSLOT]  (MODE 2)
  const float* ct0;         // [3J][Bp]  (MODE 3)
  const float* ca0;         // [3J][Bp]  (MODE 3)
  float* aT_out;            // [3V][Bp] or null (MODE 3)
  float* partials;
  float inv_scale;
  int n_segments, Bp, k_blocks, tiles_n, total_tiles, all_segments, ns, sdl;
};

// the four register-cached joint rows of an epilogue warp: q[k][c] = (row c of the 3x3 block, translation c)
struct RowCache {
  float4 q[4][3];
  __device__ __forceinline__ void load(int k, const float4* src, int Bp) {
    q[k][0] = __ldg(src);
    q[k][1] = __ldg(src + Bp);
    q[k][2] = __ldg(src + 2 * (size_t)Bp);
  }
  // blended rows: B2[2c] = (row c .x, .y), B2[2c+1] = (row c .z, .w)
  __device__ __forceinline__ void blend(const float w[4], float2* B2) const {
    {
      const float2 ww = make_float2(w[0], w[0]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        B2[2 * c] = __fmul2_rn(ww, make_float2(q[0][c].x, q[0][c].y));
        B2[2 * c + 1] = __fmul2_rn(ww, make_float2(q[0][c].z, q[0][c].w));
      }
    }
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float2 ww = make_float2(w[k], w[k]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        B2[2 * c] = __ffma2_rn(ww, make_float2(q[k][c].x, q[k][c].y), B2[2 * c]);
        B2[2 * c + 1] = __ffma2_rn(ww, make_float2(q[k][c].z, q[k][c].w), B2[2 * c + 1]);
      }
    }
  }
};

// is segment `seg` processed in this mode?  (same answer in the target producer and in the epilogue warps)
__device__ __forceinline__ bool seg_active(const FitFusedArgs& a, int mode, int seg, int& i0, int& len, int& part, bool& stat) {
  i0 = 0; len = 0; part = 0; stat = true;
  if (seg >= a.n_segments) return false;
  i0 = __ldg(a.seg_start + seg);
  len = __ldg(a.seg_start + seg + 1) - i0;
  part = __ldg(a.seg_part + seg);
  if (len <= 0) return false;
  if (mode == 3) {
    stat = (__ldg(a.part_flags + part) & 1) != 0;
    if (!stat && !(a.aT_out != nullptr && a.all_segments)) return false;
  }
  return true;
}

// H = NSP / 2 (float2 accumulators of r) for MODE 2; unused for MODE 3.  AOUT (MODE 3): the reference vertices are also
// written out (fits without target joints regress the joints from them) and segments without statistics take part.
template <int MODE, int H, bool WEIGHTED, bool AOUT>
__global__ void __launch_bounds__(THREADS, 1) k_fit_fused(const __grid_constant__ FusedMaps maps, const FitFusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int sd_tile = (MODE == 2) ? sd_tile_bytes(a.sdl) : 0;
  uint8_t* t_area = smem + STAGES * STAGE_BYTES;               // [2 halves][TRING][quadrant][3 GV][32] staged targets
  uint8_t* rec_area = t_area + 2 * TRING * TSLOT_BYTES;        // [2][TILE_V][8 words]
  uint8_t* sd_area = rec_area + 2 * REC_TILE_BYTES;            // [2][TILE_V][sdl floats]  (MODE 2)
  uint8_t* yw_area = sd_area + 2 * sd_tile;                    // [EPI_WARPS][YW_FLOATS floats + 64 B lut]  (MODE 2)
  uint64_t* full = reinterpret_cast<uint64_t*>(yw_area + (MODE == 2 ? EPI_WARPS * (YW_FLOATS * 4 + 64) : 0));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2] (EPI_WARPS arrivals)
  uint64_t* rec_full = acc_empty + 2;   // [2]
  uint64_t* t_full = rec_full + 2;      // [2][TRING]
  uint64_t* t_empty = t_full + 2 * TRING;  // [2][TRING] (4 arrivals: the warps of one half)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2 * TRING);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's contiguous range of tiles (linear index = instance tile * tiles_n + vertex tile)
  const int L0 = (int)((long long)a.total_tiles * blockIdx.x / gridDim.x);
  const int L1 = (int)((long long)a.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.f_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.f_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.p_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.p_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.t) : "memory");
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
    for (int s = 0; s < 2 * TRING; ++s) {
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 4);
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
        // ---- TMA producer: GEMM operands ----
        int it = 0;
        for (int L = L0; L < L1; ++L) {
          const int tm = L / a.tiles_n, tn = L - tm * a.tiles_n;
          const int b0 = tm * TILE_M, n0 = tn * TILE_N;
          // (only the smem stages gate these loads: the operands of the next tile are fetched while the epilogue of
          // the current one still owns both accumulators; the MMA issuer waits for the accumulator)
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
        for (int L = L0; L < L1; ++L, ++tcount) {
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
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
              mma_f16(d_tmem, flo + adv, phi + adv, acc);  // small terms first
              acc = 1;
              mma_f16(d_tmem, fhi + adv, plo + adv, 1);
              mma_f16(d_tmem, fhi + adv, phi + adv, 1);
            }
            mma_commit(&empty[s]);
          }
          mma_commit(&acc_full[as]);
        }
      }
      __syncwarp();
    } else if (warp == 2) {
      if (lane == 0) {
        // ---- record + target producer: per half h a ring of TRING blocks of GV vertices x 128 instances, filled as far ahead
        // of the epilogue warps as the ring allows (the HBM latency of the targets is off their critical path) ----
        int cnt[2] = {0, 0}, tcount = 0;
        for (int L = L0; L < L1; ++L, ++tcount) {
          const int tm = L / a.tiles_n, tn = L - tm * a.tiles_n;
          const int b0 = tm * TILE_M;
          {
            // the tile's slot records (+ shape directions), into the buffer that goes with its accumulator (free once
            // the epilogue of two tiles ago has released it)
            const int as = tcount & 1;
            mbar_wait_parked(&acc_empty[as], ((tcount >> 1) & 1) ^ 1);
            mbar_expect_tx(&rec_full[as], (uint32_t)(REC_TILE_BYTES + sd_tile));
            bulk_g2s(rec_area + as * REC_TILE_BYTES, a.rec + (size_t)tn * TILE_V * REC_WORDS, REC_TILE_BYTES, &rec_full[as]);
            if (MODE == 2)
              bulk_g2s(sd_area + as * sd_tile, a.sd + (size_t)tn * TILE_V * a.sdl, (uint32_t)sd_tile, &rec_full[as]);
          }
          int i0[2], len[2], part;
          bool stat, act[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) act[h] = seg_active(a, MODE, 2 * tn + h, i0[h], len[h], part, stat);
          const int nq = min(4, (a.Bp - b0 + 31) / 32);  // instance quadrants of this tile inside the batch
          for (int k = 0; k < SEGV / GV; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (!act[h] || k * GV >= len[h]) continue;
              const int s = cnt[h] % TRING;
              uint64_t* fb = &t_full[h * TRING + s];
              mbar_wait_parked(&t_empty[h * TRING + s], ((cnt[h] / TRING) & 1) ^ 1);
              float* dst = reinterpret_cast<float*>(t_area + (size_t)(h * TRING + s) * TSLOT_BYTES);
              mbar_expect_tx(fb, (uint32_t)(nq * 3 * GV * 32 * 4));
              for (int q = 0; q < nq; ++q) tma_load_2d(dst + q * 3 * GV * 32, &maps.t, fb, b0 + q * 32, 3 * (i0[h] + k * GV));
              ++cnt[h];
            }
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue warps ----
    const int e = warp - 4;
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int h = e >> 2;    // which segment of the tile
    const int Bp = a.Bp;
    const float inv_scale = a.inv_scale;
    float* yw = reinterpret_cast<float*>(yw_area + e * (YW_FLOATS * 4 + 64));
    unsigned char* lut = reinterpret_cast<unsigned char*>(yw + YW_FLOATS);  // joint -> Y slot of the current segment
    if (MODE == 2) {
      lut[lane] = 0;
      lut[32 + lane] = 0;
#pragma unroll
      for (int r = 0; r < NSLOT * 3; ++r) yw[r * 32 + lane] = 0.f;
      __syncwarp();
    }
    RowCache rc;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) rc.q[k][c] = make_float4(0.f, 0.f, 0.f, 0.f);
    int jcur[4] = {0, 0, 0, 0};
    int cur_tm = -1, cur_part = -1, tcount = 0, tcnt = 0;
    float ct[3] = {0.f, 0.f, 0.f}, ca[3] = {0.f, 0.f, 0.f};  // provisional centres of the current part (MODE 3)
    for (int L = L0; L < L1; ++L, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const int tm = L / a.tiles_n, tn = L - tm * a.tiles_n;
      const int seg = 2 * tn + h;
      const int b = tm * TILE_M + q * 32 + lane;
      int i0, len, part;
      bool stat;
      const bool active = seg_active(a, MODE, seg, i0, len, part, stat);  // uniform over the four warps of this half
      const bool live = active && (tm * TILE_M + q * 32 < Bp);             // warp-uniform (Bp % 32 == 0)
      // full reload of the four slots at the first vertex when the cached rows belong to other instances or the chain of
      // segments the host replayed (every second segment, in order) was left (skipped segment)
      const bool fresh = tm != cur_tm;
      if (fresh) {
        cur_part = -1;
        if (MODE == 2) {  // (a NaN instance may have left NaNs in Y rows that are never handed out)
#pragma unroll
          for (int r = 0; r < NSLOT * 3; ++r) yw[r * 32 + lane] = 0.f;
        }
      }
      cur_tm = live ? tm : -1;
      // per-segment state
      float2 r2[H > 0 ? H : 1];
      float Sb[3] = {0.f, 0.f, 0.f}, Yr[4][3];
      float M[9], st[3] = {0.f, 0.f, 0.f}, sa[3] = {0.f, 0.f, 0.f}, W = 0.f;
      int nslots = 0;
      if (MODE == 2) {
#pragma unroll
        for (int s = 0; s < (H > 0 ? H : 1); ++s) r2[s] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) Yr[k][0] = Yr[k][1] = Yr[k][2] = 0.f;
        if (live) {
          const int sj = (lane < NSLOT) ? __ldg(a.seg_slots + seg * NSLOT + lane) : -1;
          nslots = __popc(__ballot_sync(0xffffffffu, sj >= 0));
          if (sj >= 0) lut[sj] = (unsigned char)lane;
          __syncwarp();
        }
      } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) M[r] = 0.f;
        if (live && part != cur_part) {  // consecutive segments of a chain mostly belong to the same part
          cur_part = part;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            ct[c] = a.ct0[(size_t)(part * 3 + c) * Bp + b];
            ca[c] = a.ca0[(size_t)(part * 3 + c) * Bp + b];
          }
        }
      }
      mbar_wait(&rec_full[as], aph);
      mbar_wait(&acc_full[as], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (active) {
        const uint4* recs = reinterpret_cast<const uint4*>(rec_area + as * REC_TILE_BYTES) + (h * SEGV) * 2;
        const float* sds = reinterpret_cast<const float*>(sd_area + as * sd_tile) + (size_t)(h * SEGV) * a.sdl;
        const float4* qb = a.quads + b;
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TILE_N + h * 3 * SEGV);
        const int nblk = (len + GV - 1) / GV;
        for (int g = 0; g < nblk; ++g, ++tcnt) {
          const int ts = tcnt % TRING;
          uint32_t r[3 * GV];
          if (live) {
            SF_TMEM_LD16(r, 0, tbase + g * 3 * GV);
            SF_TMEM_LD8(r, 16, tbase + g * 3 * GV + 16);
          }
          mbar_wait(&t_full[h * TRING + ts], (tcnt / TRING) & 1);
          if (live) {
            const float* tq = reinterpret_cast<const float*>(t_area + (size_t)(h * TRING + ts) * TSLOT_BYTES) + q * 3 * GV * 32 + lane;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // per-vertex work on x = GEMM column triple (scaled, rest position added), t = target
            auto vertex = [&](const float wk[4], const float x[3], const float t[3], int vi) {
              float2 B2[6];
              rc.blend(wk, B2);
              float p[3];
#pragma unroll
              for (int c = 0; c < 3; ++c)
                p[c] = fmaf(B2[2 * c].x, x[0], fmaf(B2[2 * c].y, x[1], fmaf(B2[2 * c + 1].x, x[2], B2[2 * c + 1].y)));
              if (MODE == 2) {
                float bv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  bv[c] = t[c] - p[c];
                  Sb[c] += bv[c];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                  for (int c = 0; c < 3; ++c) Yr[k][c] = fmaf(wk[k], bv[c], Yr[k][c]);
                float z[3];
                z[0] = fmaf(B2[0].x, bv[0], fmaf(B2[2].x, bv[1], B2[4].x * bv[2]));
                z[1] = fmaf(B2[0].y, bv[0], fmaf(B2[2].y, bv[1], B2[4].y * bv[2]));
                z[2] = fmaf(B2[1].x, bv[0], fmaf(B2[3].x, bv[1], B2[5].x * bv[2]));
                // shapedirs[x][s] at x * NSP4 + s (rows padded to 16-byte multiples: read as 16-byte words)
                constexpr int NSP4 = (2 * H + 3) / 4 * 4;
                const float4* sdv = reinterpret_cast<const float4*>(sds + (size_t)vi * a.sdl);
#pragma unroll
                for (int xx = 0; xx < 3; ++xx) {
                  const float2 zz = make_float2(z[xx], z[xx]);
#pragma unroll
                  for (int q4 = 0; q4 < NSP4 / 4; ++q4) {
                    const float4 s4 = sdv[xx * (NSP4 / 4) + q4];
                    r2[2 * q4] = __ffma2_rn(make_float2(s4.x, s4.y), zz, r2[2 * q4]);
                    if (2 * q4 + 1 < H) r2[2 * q4 + 1] = __ffma2_rn(make_float2(s4.z, s4.w), zz, r2[2 * q4 + 1]);
                  }
                }
              } else {
                if (AOUT) {
                  float* ao = a.aT_out + (size_t)((i0 + vi) * 3) * Bp + b;
                  ao[0] = p[0];
                  ao[Bp] = p[1];
                  ao[2 * (size_t)Bp] = p[2];
                }
                if (!AOUT || stat) {  // (without AOUT only segments with statistics are active)
                  const float wv = WEIGHTED ? a.vwT[(size_t)(i0 + vi) * Bp + b] : 1.f;
                  float dt[3], wa[3];
#pragma unroll
                  for (int c = 0; c < 3; ++c) {
                    dt[c] = t[c] - ct[c];
                    wa[c] = WEIGHTED ? wv * (p[c] - ca[c]) : (p[c] - ca[c]);
                    st[c] = WEIGHTED ? fmaf(wv, dt[c], st[c]) : st[c] + dt[c];
                    sa[c] += wa[c];
                  }
#pragma unroll
                  for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                    for (int c = 0; c < 3; ++c) M[rr * 3 + c] = fmaf(dt[rr], wa[c], M[rr * 3 + c]);
                  W += wv;
                }
              }
            };
            // GV / 2 vertices without a slot reload: one branch-free stretch, so the per-vertex chains interleave
            auto fast_half = [&](int u0) {
#pragma unroll
              for (int uu = 0; uu < GV / 2; ++uu) {
                const int u = u0 + uu;  // compile-time after inlining (u0 is a literal at both call sites)
                const int vi = g * GV + u;
                const uint4 w4 = recs[2 * vi], m4 = recs[2 * vi + 1];
                const float wk[4] = {__uint_as_float(w4.x), __uint_as_float(w4.y), __uint_as_float(w4.z), __uint_as_float(w4.w)};
                float t[3], x[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  t[c] = tq[(u * 3 + c) * 32];
                  x[c] = fmaf(__uint_as_float(r[3 * u + c]), inv_scale, __uint_as_float(c == 0 ? m4.y : (c == 1 ? m4.z : m4.w)));
                }
                vertex(wk, x, t, vi);
              }
            };
            const uint32_t pack0 = recs[2 * g * GV + 1].x;
            const int nv = min(GV, len - g * GV);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              const int u_lo = half * (GV / 2), u_hi = min(nv, u_lo + GV / 2);
              if (u_lo >= u_hi) break;
              const bool fast = u_hi - u_lo == GV / 2 && !(fresh && g == 0 && half == 0) && (pack0 & (1u << (29 + half))) == 0;
              if (fast) {
                if (half == 0) fast_half(0);
                else fast_half(GV / 2);
              } else {
                // general vertices (a slot reloads, first vertices of a fresh chain, or a partial half): rolled loop;
                // the accumulator columns go through a small per-thread array so that they can be indexed at run time
                float xs[3 * GV / 2];
                if (half == 0) {
#pragma unroll
                  for (int i = 0; i < 3 * GV / 2; ++i) xs[i] = __uint_as_float(r[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 3 * GV / 2; ++i) xs[i] = __uint_as_float(r[3 * GV / 2 + i]);
                }
#pragma unroll 1
                for (int u = u_lo; u < u_hi; ++u) {
                  const int vi = g * GV + u;
                  const uint4 w4 = recs[2 * vi], m4 = recs[2 * vi + 1];
                  const uint32_t pack = m4.x;
                  const float wk[4] = {__uint_as_float(w4.x), __uint_as_float(w4.y), __uint_as_float(w4.z), __uint_as_float(w4.w)};
                  const uint32_t rl = (fresh && vi == 0) ? 0xFu : ((pack >> 24) & 0xFu);
                  if (rl) {  // warp-uniform
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      if (rl & (1u << k)) {
                        const int jn = (int)((pack >> (6 * k)) & 63u);
                        if (MODE == 2) {
                          float* yp = yw + (size_t)(lut[jcur[k]] * 3) * 32 + lane;
#pragma unroll
                          for (int c = 0; c < 3; ++c) {
                            yp[c * 32] += Yr[k][c];
                            Yr[k][c] = 0.f;
                          }
                        }
                        jcur[k] = jn;
                        rc.load(k, qb + (size_t)(jn * 3) * Bp, Bp);
                      }
                    }
                  }
                  float t[3], x[3];
#pragma unroll
                  for (int c = 0; c < 3; ++c) {
                    t[c] = tq[(u * 3 + c) * 32];
                    x[c] = fmaf(xs[3 * (u - u_lo) + c], inv_scale, __uint_as_float(c == 0 ? m4.y : (c == 1 ? m4.z : m4.w)));
                  }
                  vertex(wk, x, t, vi);
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[h * TRING + ts]);
        }
      }
      // the accumulator is drained: hand it back before the (global-memory) write-out
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (live) {
        if (MODE == 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float* yp = yw + (size_t)(lut[jcur[k]] * 3) * 32 + lane;
#pragma unroll
            for (int c = 0; c < 3; ++c) yp[c * 32] += Yr[k][c];
          }
          const int NL = a.ns + 3 + 3 * NSLOT;
          float* out = a.partials + (size_t)seg * NL * Bp + b;
#pragma unroll
          for (int s = 0; s < 2 * H; ++s)
            if (s < a.ns) out[(size_t)s * Bp] = (s & 1) ? r2[s >> 1].y : r2[s >> 1].x;
          out += (size_t)a.ns * Bp;
          out[0] = Sb[0];
          out[Bp] = Sb[1];
          out[2 * (size_t)Bp] = Sb[2];
          out += (size_t)3 * Bp;
          for (int r = 0; r < nslots * 3; ++r) {  // hand the Y rows out and leave them zeroed for the next segment
            out[(size_t)r * Bp] = yw[r * 32 + lane];
            yw[r * 32 + lane] = 0.f;
          }
          // a stale slot (weight 0 in this segment) may have been flushed (adding 0) into a row past nslots: those rows
          // only ever receive zeros, so they stay zero
          __syncwarp();
        } else if (!AOUT || stat) {
          float* out = a.partials + (size_t)seg * 16 * Bp + b;
#pragma unroll
          for (int r = 0; r < 9; ++r) out[(size_t)r * Bp] = M[r];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            out[(size_t)(9 + c) * Bp] = st[c];
            out[(size_t)(12 + c) * Bp] = sa[c];
          }
          out[(size_t)15 * Bp] = W;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// feature rows of the GEMM: F[b] = [vec(R_rel[1:]) - vec(I) | unknowns (betas, kid) or zeros | 0 ...] split into fp16
// hi / lo, [Bt][Kf] row-major (rows >= B are zero)
__global__ void k_fq_feat(const float* __restrict__ feat, const float* __restrict__ beta, int B, int Bp, int Kp, int P, int ns,
                          int Kf, size_t n2, __half2* __restrict__ hi, __half2* __restrict__ lo) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n2) return;
  const int k2 = (int)(idx % (Kf / 2));
  const size_t b = idx / (Kf / 2);
  float x[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int k = 2 * k2 + t;
    float v = 0.f;
    if (b < (size_t)B) {
      if (k < P) v = feat[b * Kp + k] - ((k % 9) % 4 == 0 ? 1.f : 0.f);
      else if (k < P + ns && beta != nullptr) v = beta[(size_t)(k - P) * Bp + b];
    }
    x[t] = v;
  }
  const __half h0 = __float2half_rn(x[0]), h1 = __float2half_rn(x[1]);
  hi[idx] = __halves2half2(h0, h1);
  lo[idx] = __halves2half2(__float2half_rn(x[0] - __half2float(h0)), __float2half_rn(x[1] - __half2float(h1)));
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

// targets [rows = 3V][Bp] fp32, box = {32 instances, 3 GV rows}; rows / columns past the array read as zero
bool make_t_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t Bp) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {Bp, rows};
  cuuint64_t strides[1] = {Bp * sizeof(float)};
  cuuint32_t box[2] = {32, 3 * GV};
  cuuint32_t elem[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
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

template <int MODE, int H, bool WEIGHTED, bool AOUT = false>
bool launch_t(const FusedMaps& maps, const FitFusedArgs& fa, int grid, int smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_fit_fused<MODE, H, WEIGHTED, AOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set = true;
  }
  SF_LAUNCH((k_fit_fused<MODE, H, WEIGHTED, AOUT>), grid, THREADS, smem, st, maps, fa);
  return true;
}

}  // namespace

bool fit_fused_available(const smplfit_model_t* m) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("SMPLFIT_B200_FIT_FUSED");
    off = (e && atoi(e) == 0) ? 1 : 0;
  }
  return !off && m->fq_P_hi != nullptr && m->fq_P_lo != nullptr && m->fq_rec != nullptr && m->fq_sd != nullptr && m->fq_kf > 0 &&
         m->fq_kf % KB == 0 && m->fq_nseg_pad >= m->n_segments && m->fq_nseg_pad % 2 == 0 && m->seg_slots != nullptr &&
         m->n_slots == NSLOT && m->num_joints <= 64 && m->fit_ns >= 1 && m->fit_ns <= 18 && encode_fn() != nullptr &&
         fused_smem_bytes(2, m->fq_sdl) <= 227 * 1024;
}

size_t fit_fused_scratch_bytes(const smplfit_model_t* m, int Bp) {
  if (m->fq_kf <= 0) return 0;
  return (size_t)2 * roundup(Bp, TILE_M) * m->fq_kf * sizeof(__half) + 512;
}

void fit_fused_feature_rows(const smplfit_model_t* m, int Bp, void* scratch, void** hi, void** lo) {
  const int Bt = roundup(Bp, TILE_M);
  __half* h = reinterpret_cast<__half*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  *hi = h;
  *lo = h + (size_t)Bt * m->fq_kf;
}

bool fit_fused_run(const smplfit_model_t* m, int mode, int B, int Bp, const float* feat, int Kp, const float* beta,
                   const float* tT, const float* vwT, const float* quads, const float* ct0, const float* ca0, float* aT_out,
                   int all_segments, float* partials, void* scratch, bool feats_ready, cudaStream_t st) {
  if (!fit_fused_available(m) || scratch == nullptr) return false;
  const int Bt = roundup(Bp, TILE_M), Kf = m->fq_kf;
  __half* hi = reinterpret_cast<__half*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  __half* lo = hi + (size_t)Bt * Kf;
  const int tiles_n = m->fq_nseg_pad / 2, tiles_m = Bt / TILE_M;
  FusedMaps maps;
  const uint64_t p_rows = (uint64_t)tiles_n * TILE_N;
  if (!make_h_map(&maps.f_hi, hi, (uint64_t)Bt, (uint64_t)Kf, TILE_M) || !make_h_map(&maps.f_lo, lo, (uint64_t)Bt, (uint64_t)Kf, TILE_M) ||
      !make_h_map(&maps.p_hi, m->fq_P_hi, p_rows, (uint64_t)Kf, TILE_N) || !make_h_map(&maps.p_lo, m->fq_P_lo, p_rows, (uint64_t)Kf, TILE_N) ||
      !make_t_map(&maps.t, tT, (uint64_t)3 * m->num_vertices, (uint64_t)Bp))
    return false;
  if (!feats_ready) {  // (k_front_fused / k_shape_out write the rows themselves on the closed-form path)
    const size_t n2 = (size_t)Bt * (Kf / 2);
    SF_LAUNCH(k_fq_feat, (unsigned)((n2 + 255) / 256), 256, 0, st, feat, beta, B, Bp, Kp, m->num_pose_feats, m->fit_ns, Kf, n2,
              reinterpret_cast<__half2*>(hi), reinterpret_cast<__half2*>(lo));
  }
  FitFusedArgs fa{};
  fa.tT = tT; fa.vwT = vwT; fa.quads = reinterpret_cast<const float4*>(quads); fa.rec = m->fq_rec; fa.sd = m->fq_sd;
  fa.seg_start = m->seg_start; fa.seg_part = m->seg_part; fa.part_flags = m->part_flags; fa.seg_slots = m->seg_slots;
  fa.ct0 = ct0; fa.ca0 = ca0; fa.aT_out = aT_out; fa.partials = partials; fa.inv_scale = ldexpf(1.f, -m->fq_scale_log2);
  fa.n_segments = m->n_segments; fa.Bp = Bp; fa.k_blocks = Kf / KB; fa.tiles_n = tiles_n; fa.total_tiles = tiles_m * tiles_n;
  fa.all_segments = all_segments; fa.ns = m->fit_ns; fa.sdl = m->fq_sdl;
  const int grid = fa.total_tiles < sm_count() ? fa.total_tiles : sm_count();
  const int smem = fused_smem_bytes(mode, m->fq_sdl);
  if (mode == 3) {
    if (aT_out != nullptr)
      return vwT ? launch_t<3, 0, true, true>(maps, fa, grid, smem, st) : launch_t<3, 0, false, true>(maps, fa, grid, smem, st);
    return vwT ? launch_t<3, 0, true>(maps, fa, grid, smem, st) : launch_t<3, 0, false>(maps, fa, grid, smem, st);
  }
  switch ((m->fit_ns + 1) / 2) {
    case 1: return launch_t<2, 1, false>(maps, fa, grid, smem, st);
    case 2: return launch_t<2, 2, false>(maps, fa, grid, smem, st);
    case 3: return launch_t<2, 3, false>(maps, fa, grid, smem, st);
    case 4: return launch_t<2, 4, false>(maps, fa, grid, smem, st);
    case 5: return launch_t<2, 5, false>(maps, fa, grid, smem, st);
    case 6: return launch_t<2, 6, false>(maps, fa, grid, smem, st);
    case 7: return launch_t<2, 7, false>(maps, fa, grid, smem, st);
    case 8: return launch_t<2, 8, false>(maps, fa, grid, smem, st);
    case 9: return launch_t<2, 9, false>(maps, fa, grid, smem, st);
    default: return false;
  }
}

}  // namespace sf
