// The unweighted shape pass in closed form ("lite" path).
//
// With  jac_v[:,s] = sum_k w_vk (R_k S_vs + T_ks)   and   b_v = t_v - (Rb_v vp_v + Tb_v0)
// the normal equations of the shape solve (pt/bodyfitter.py:999-1048) need
//   G[s,t] = sum_v jac_vs . jac_vt ,   r[s] = sum_v jac_vs . b_v ,   Sb = sum_v b_v .
// Without per-vertex weights G does not depend on the targets: it is a contraction of the
// per-instance joint transforms with model constants over the joint pairs that share a vertex
// (k_gram_closed; constants gcf_* of include/smplfit_b200.h).  Only r and Sb need a pass over the
// vertices, and that pass is light (k_shape_lite):
//   r[s] = sum_v S_vs . z_v + sum_k T_ks . Y_k ,   z_v = Rb_v^T b_v ,   Y_k = sum_v w_vk b_v ,
// i.e. per vertex a 12-row blend, one 3x3 transpose product, 3 NS FMAs, and a scatter of b_v into
// the (few) joints the vertex is skinned to.  ~5x fewer FP32 operations than the per-vertex Gramian.
#pragma once
#include <cuda.h>

#include "fit_kernels.cuh"
#include "solve_kernels.cuh"

namespace sf {

constexpr int LITE_NSLOT = 12;  // == N_SLOTS of pt/bodymodel.py
constexpr int LITE_VS = 4;      // vertices per staged sub-block
#ifndef SF_LITE_UNROLL
#define SF_LITE_UNROLL 1
#endif
constexpr int LITE_UNROLL = SF_LITE_UNROLL;  // vertices of a sub-block processed per loop iteration (ILP vs registers)

struct LiteArgs {
  const float* tT;       // [3V][Bp]   (read through the tensor map)
  const float* vposedT;  // [3V][Bp]   (read through the tensor map)
  const float* RT12;     // [J*3][Bp] float4: (R[c][0], R[c][1], R[c][2], T0[c])
  const float* rec;      // [V][Rec<NS>::LEN]
  const int32_t* seg_start;
  const int32_t* seg_slots;  // [n_segments][LITE_NSLOT]
  float* partials;           // [n_segments][NS + 3 + 3*LITE_NSLOT][Bp]: r | Sb | Y per slot
  const uint8_t* slot_mask;  // [V] precomputed "slot k must reload its joint rows" bits (smplfit_b200.h), or null
  int n_segments, J, Bp, segs_per_warp;
};

__device__ __forceinline__ void sf_tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          (uint32_t)__cvta_generic_to_shared(dst)),
      "l"(map), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Per-warp two-stage shared-memory ring of the three streams a vertex pass consumes: the warp's 32-instance
// column of the targets and of the posed template (one 2D TMA box {32 instances, 3 LITE_VS rows} each, from the
// instance-minor [3V][Bp] arrays) and the LITE_VS per-vertex records (one bulk copy), all signalling one mbarrier
// per stage.  Stage k + 1 is in flight while stage k is consumed: the global-memory latency is taken off the
// per-vertex critical path (the register-prefetch version stalled on its own rotation moves, profiles/).
// Stage layout (floats): [t: 3 VS x 32][vp: 3 VS x 32][records: VS x REC], padded to a 128-byte multiple.
template <int REC>
struct VertStager {
  static constexpr int BOX = 3 * LITE_VS * 32;
  static constexpr int STAGE = (2 * BOX + LITE_VS * REC + 31) / 32 * 32;
  float* buf;     // [2][STAGE], 128-byte aligned
  uint64_t* bar;  // [2]
  const float* rec_src;
  const CUtensorMap *mt, *mv;
  int i0, i1, col;
  uint32_t phase;
  __device__ __forceinline__ void issue(int k, int lane) {
    const int first = i0 + k * LITE_VS;
    __syncwarp();  // every lane is done reading the stage that is refilled
    if (first < i1 && lane == 0) {
      float* dst = buf + (size_t)(k & 1) * STAGE;
      const uint32_t rec_bytes = (uint32_t)min(LITE_VS, i1 - first) * REC * 4;
      // a TMA box always delivers its full byte count (rows past the tensor are zero-filled)
      sf_mbar_expect_tx(bar + (k & 1), 2u * BOX * 4u + rec_bytes);
      sf_tma_2d(dst, mt, bar + (k & 1), col, first * 3);
      sf_tma_2d(dst + BOX, mv, bar + (k & 1), col, first * 3);
      sf_bulk_g2s(dst + 2 * BOX, rec_src + (size_t)first * REC, rec_bytes, bar + (k & 1));
    }
  }
  __device__ __forceinline__ void wait(int k) {
    sf_mbar_wait(bar + (k & 1), (phase >> (k & 1)) & 1u);
    phase ^= 1u << (k & 1);
  }
  __device__ __forceinline__ const float* stage(int k) const { return buf + (size_t)(k & 1) * STAGE; }
};

__host__ __device__ inline int lite_stage_floats(int rec_len) { return (2 * 3 * LITE_VS * 32 + LITE_VS * rec_len + 31) / 32 * 32; }

__host__ __device__ inline size_t lite_smem_bytes(int J, int rec_len, int warps) {
  return ((size_t)J * 3 * 128 + (size_t)warps * (2 * lite_stage_floats(rec_len) + LITE_NSLOT * 3 * 32)) * sizeof(float) +
         (size_t)warps * (16 + 64) + 16;
}

// Joint rows in registers: the internal vertex order groups vertices by their tuple of skinning joints
// (pt/bodymodel.py), so the four slots' rows change only ~0.1 times per vertex; they are reloaded from the
// staged shared-memory copy on change (warp-uniform branch).  Slot k of a vertex with weight 0 is never loaded.
struct JointCache {
  float4 q[4][3];
  int j[4];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      j[k] = -1;
#pragma unroll
      for (int c = 0; c < 3; ++c) q[k][c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // blended rows: B2[2c] = (row c .x, .y), B2[2c+1] = (row c .z, .w)
  __device__ __forceinline__ void blend(const float w[4], float2* B2) const {
    {
      const float2 ww = make_float2(w[0], w[0]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        B2[2 * c] = sf_mul2(ww, make_float2(q[0][c].x, q[0][c].y));
        B2[2 * c + 1] = sf_mul2(ww, make_float2(q[0][c].z, q[0][c].w));
      }
    }
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float2 ww = make_float2(w[k], w[k]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        B2[2 * c] = sf_fma2(ww, make_float2(q[k][c].x, q[k][c].y), B2[2 * c]);
        B2[2 * c + 1] = sf_fma2(ww, make_float2(q[k][c].z, q[k][c].w), B2[2 * c + 1]);
      }
    }
  }
};

template <int NS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_shape_lite(const LiteArgs a, const __grid_constant__ CUtensorMap map_t, const __grid_constant__ CUtensorMap map_vp) {
  extern __shared__ __align__(128) float s_lite[];
  constexpr int NSP = Rec<NS>::NSP, H = NSP / 2, REC = Rec<NS>::LEN;
  constexpr int NL = NS + 3 + 3 * LITE_NSLOT;
  using Stager = VertStager<REC>;
  constexpr int BOX = Stager::BOX, STAGE = Stager::STAGE;
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  const bool has_mask = a.slot_mask != nullptr;
  const float4* sq = reinterpret_cast<const float4*>(s_lite);  // [J*3][32]
  float* wbase = s_lite + (size_t)a.J * 3 * 128;
  float* yw = wbase + (size_t)WARPS * (2 * STAGE) + (size_t)warp * (LITE_NSLOT * 3 * 32);  // [slot*3+c][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + (size_t)WARPS * (2 * STAGE + LITE_NSLOT * 3 * 32));
  unsigned char* lut = reinterpret_cast<unsigned char*>(bars + 2 * WARPS) + warp * 64;  // joint -> slot of the segment
  Stager rs;
  rs.buf = wbase + (size_t)warp * (2 * STAGE);
  rs.bar = bars + 2 * warp;
  rs.rec_src = a.rec;
  rs.mt = &map_t;
  rs.mv = &map_vp;
  rs.col = g * 32;
  rs.phase = 0;
  if (lane == 0) {
    sf_mbar_init(rs.bar, 1);
    sf_mbar_init(rs.bar + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  {
    // the CTA's 32-instance slice of every (joint, row) quad: 32 x 16 B contiguous per row
    const int n = a.J * 3 * 32;
    const float4* src = reinterpret_cast<const float4*>(a.RT12);
    for (int q = threadIdx.x; q < n; q += WARPS * 32) {
      const int r = q >> 5, l = q & 31;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_lite + (size_t)q * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)r * Bp + g * 32 + l) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  for (int q = 0; q < a.segs_per_warp; ++q) {
    const int seg = (blockIdx.x * a.segs_per_warp + q) * WARPS + warp;
    if (seg >= a.n_segments) break;
    const int i0 = a.seg_start[seg], i1 = a.seg_start[seg + 1];
    rs.i0 = i0;
    rs.i1 = i1;
    rs.issue(0, lane);
    rs.issue(1, lane);
    const int sj = (lane < LITE_NSLOT) ? __ldg(a.seg_slots + seg * LITE_NSLOT + lane) : -1;
    const int nslots = __popc(__ballot_sync(0xffffffffu, sj >= 0));
    if (sj >= 0) lut[sj] = (unsigned char)lane;
    for (int e = 0; e < nslots * 3; ++e) yw[e * 32 + lane] = 0.f;
    __syncwarp();
    float2 r2[H];
#pragma unroll
    for (int e = 0; e < H; ++e) r2[e] = make_float2(0.f, 0.f);
    float Sb[3] = {0.f, 0.f, 0.f};
    float Yr[4][3];  // Y of the cached joint of each slot, flushed to its shared-memory cell when the joint changes
#pragma unroll
    for (int k = 0; k < 4; ++k) Yr[k][0] = Yr[k][1] = Yr[k][2] = 0.f;
    JointCache jc;
    jc.reset();
    const int nsub = (i1 - i0 + LITE_VS - 1) / LITE_VS;
    for (int k = 0; k < nsub; ++k) {
      const int nv = min(LITE_VS, i1 - (i0 + k * LITE_VS));
      // precomputed reload bits of the sub-block's vertices (one byte per lane, broadcast by shuffle below)
      unsigned mk = 0u;
      if (has_mask && lane < nv) mk = __ldg(a.slot_mask + i0 + k * LITE_VS + lane);
      rs.wait(k);
      if (k >= 1) rs.issue(k + 1, lane);  // the stage of sub-block k-1 is free now
      const float* st = rs.stage(k);
#pragma unroll LITE_UNROLL
      for (int u = 0; u < nv; ++u) {
        float t[3], vp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          t[c] = st[(u * 3 + c) * 32 + lane];
          vp[c] = st[BOX + (u * 3 + c) * 32 + lane];
        }
        const float* rec = st + 2 * BOX + u * REC;
        const float4 w4 = *reinterpret_cast<const float4*>(rec);
        const float wk[4] = {w4.x, w4.y, w4.z, w4.w};
        unsigned m;
        int4 j4 = make_int4(0, 0, 0, 0);
        if (has_mask) {
          m = __shfl_sync(0xffffffffu, mk, u);
          if (m) j4 = *reinterpret_cast<const int4*>(rec + 4);
        } else {
          j4 = *reinterpret_cast<const int4*>(rec + 4);
          m = (wk[0] != 0.f && j4.x != jc.j[0] ? 1u : 0u) | (wk[1] != 0.f && j4.y != jc.j[1] ? 2u : 0u) |
              (wk[2] != 0.f && j4.z != jc.j[2] ? 4u : 0u) | (wk[3] != 0.f && j4.w != jc.j[3] ? 8u : 0u);
        }
        const int jk[4] = {j4.x, j4.y, j4.z, j4.w};
        if (m) {  // warp-uniform, ~4 % of the vertices
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (m & (1u << kk)) {
            if (jc.j[kk] >= 0) {
              float* yp = yw + (size_t)(lut[jc.j[kk]] * 3) * 32 + lane;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                yp[c * 32] += Yr[kk][c];
                Yr[kk][c] = 0.f;
              }
            }
            jc.j[kk] = jk[kk];
#pragma unroll
            for (int c = 0; c < 3; ++c) jc.q[kk][c] = sq[(size_t)(jk[kk] * 3 + c) * 32 + lane];
          }
        }
        }
        float2 B2[6];  // B2[2c] = (Rb[c][0], Rb[c][1]), B2[2c+1] = (Rb[c][2], Tb0[c])
        jc.blend(wk, B2);
        float bv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float pos = fmaf(B2[2 * c].x, vp[0], fmaf(B2[2 * c].y, vp[1], fmaf(B2[2 * c + 1].x, vp[2], B2[2 * c + 1].y)));
          bv[c] = t[c] - pos;
          Sb[c] += bv[c];
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int c = 0; c < 3; ++c) Yr[kk][c] = fmaf(wk[kk], bv[c], Yr[kk][c]);
        float z[3];
        z[0] = fmaf(B2[0].x, bv[0], fmaf(B2[2].x, bv[1], B2[4].x * bv[2]));
        z[1] = fmaf(B2[0].y, bv[0], fmaf(B2[2].y, bv[1], B2[4].y * bv[2]));
        z[2] = fmaf(B2[1].x, bv[0], fmaf(B2[3].x, bv[1], B2[5].x * bv[2]));
        // shapedirs[x][s] of the record: 3 NSP floats read as 16-byte words (the record is padded to a 16-byte
        // multiple past them), used as (s, s+1) pairs
        constexpr int NV4 = (3 * NSP + 3) / 4;
        float sdv[NV4 * 4];
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
          const float4 v4 = *reinterpret_cast<const float4*>(rec + 8 + 4 * q);
          sdv[4 * q] = v4.x; sdv[4 * q + 1] = v4.y; sdv[4 * q + 2] = v4.z; sdv[4 * q + 3] = v4.w;
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const float2 zz = make_float2(z[x], z[x]);
#pragma unroll
          for (int sp = 0; sp < H; ++sp)
            r2[sp] = sf_fma2(make_float2(sdv[x * NSP + 2 * sp], sdv[x * NSP + 2 * sp + 1]), zz, r2[sp]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (jc.j[k] >= 0) {
        float* yp = yw + (size_t)(lut[jc.j[k]] * 3) * 32 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) yp[c * 32] += Yr[k][c];
      }
    }
    float* out = a.partials + (size_t)seg * NL * Bp + b;
#pragma unroll
    for (int s = 0; s < NS; ++s) out[(size_t)s * Bp] = (s & 1) ? r2[s >> 1].y : r2[s >> 1].x;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(size_t)(NS + c) * Bp] = Sb[c];
    for (int e = 0; e < nslots * 3; ++e) out[(size_t)(NS + 3 + e) * Bp] = yw[e * 32 + lane];
    __syncwarp();  // lut / yw are rewritten by the next segment
  }
}

// k_lite_reduce: one thread per (instance, joint): Y_k = sum over the (segment, slot) cells of joint k, in double;
// blockIdx.y >= J: row r of [r | Sb] summed over all segments (what the r / Sb entries of the normal equations need:
// done here, next to chains of the same length, instead of as the longest chains of k_gram_entries).
struct LiteReduceArgs {
  const float* partials;
  const int32_t* yj_start;
  const int32_t* yj_entry;
  double* Yd;  // [3J + NS + 3][Bp]
  int NL, NS, Bp, J, n_segments;
};
static __global__ void __launch_bounds__(32) k_lite_reduce(const LiteReduceArgs a) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y;
  if (b >= a.Bp) return;
  if (j >= a.J) {
    const int row = j - a.J;
    const float* pr = a.partials + (size_t)row * a.Bp + b;
    const size_t qs = (size_t)a.NL * a.Bp;
    double acc = 0.0;
    int q = 0;
#pragma unroll 8
    for (; q + 4 <= a.n_segments; q += 4)  // four segment partials added in fp32, then one double accumulation
      acc += (double)((pr[(size_t)q * qs] + pr[(size_t)(q + 1) * qs]) + (pr[(size_t)(q + 2) * qs] + pr[(size_t)(q + 3) * qs]));
    for (; q < a.n_segments; ++q) acc += (double)pr[(size_t)q * qs];
    a.Yd[(size_t)(3 * a.J + row) * a.Bp + b] = acc;
    return;
  }
  double y[3] = {0.0, 0.0, 0.0};
  auto cell = [&](int q) {
    const int e = a.yj_entry[q];
    const int seg = e / LITE_NSLOT, slot = e % LITE_NSLOT;
    return a.partials + ((size_t)seg * a.NL + a.NS + 3 + slot * 3) * a.Bp + b;
  };
  int q = a.yj_start[j];
  const int q1 = a.yj_start[j + 1];
#pragma unroll 4
  for (; q + 2 <= q1; q += 2) {  // two cells added in fp32 before the double accumulation
    const float* p0 = cell(q);
    const float* p1 = cell(q + 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] += (double)(p0[(size_t)c * a.Bp] + p1[(size_t)c * a.Bp]);
  }
  if (q < q1) {
    const float* p = cell(q);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] += (double)p[(size_t)c * a.Bp];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) a.Yd[(size_t)(j * 3 + c) * a.Bp + b] = y[c];
}

// ---------------------------------------------------------------------------------------
// The closed-form Gramian, two kernels (8 warps per CTA, lane = instance, in-CTA tree reduction so each CTA
// writes one partial Gramian [NG][Bp]):
//   k_gram_pairs<NS>: each warp takes LITE_PPW off-diagonal joint pairs k<l (rotation-rotation term)
//        G[e] += sum_ab (R_k^T R_l)[a,b] * Asym_kl[a,b,e]           (constants: warp-uniform 16-byte loads)
//   k_gram_trans<NS>: one CTA per coordinate c; the c-rows of every joint ([R_k[c][0..2] | T_k[c][1..NS]]) of the
//        CTA's 32 instances are staged in shared memory; warps split the joints l
//        Q_l[s] = sum_k (R_k[c] . Bm_kl[:,s] + W_kl/2 T_k[c][s]),   G[s,t] += T_l[c][t] Q_l[s] + T_l[c][s] Q_l[t]
// (the diagonal rotation term sum_k sum_a A_kk[a,a,e] is the constant gcf_G0, added in k_gram_entries).
// ---------------------------------------------------------------------------------------
constexpr int LITE_PPW = 2;  // pairs per warp in k_gram_pairs

struct GramClosedArgs {
  const float* RT;  // [J*(12+3NS)][Bp]
  const int32_t* pairs;
  const float* A;
  const int32_t* lstart;
  const int32_t* lk;
  const float* Bm;  // (cells, 3, NSP4)
  const float* Wh;
  float* out;  // [n_pair_ctas + 3][NG][Bp]
  int npairs, J, Bp, n_pair_ctas;
};

// sum G over the CTA's 8 warps (fixed order); the result is valid in warp 0.  red: [4][N][32] floats.
template <int N>
__device__ __forceinline__ void cta_reduce8(float* G, float* red, int warp, int lane) {
#pragma unroll
  for (int half = 4; half >= 1; half >>= 1) {
    __syncthreads();
    if (warp >= half && warp < 2 * half) {
#pragma unroll
      for (int e = 0; e < N; ++e) red[((size_t)(warp - half) * N + e) * 32 + lane] = G[e];
    }
    __syncthreads();
    if (warp < half) {
#pragma unroll
      for (int e = 0; e < N; ++e) G[e] += red[((size_t)warp * N + e) * 32 + lane];
    }
  }
}

template <int NS>
__global__ void __launch_bounds__(256) k_gram_pairs(const GramClosedArgs a) {
  extern __shared__ __align__(16) float s_red[];  // [4][NGP][32]
  constexpr int NG = NS * (NS + 1) / 2, NGP = (NG + 3) / 4 * 4;
  constexpr int RW = 12 + 3 * NS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const int Bp = a.Bp;
  float G[NGP];
#pragma unroll
  for (int e = 0; e < NGP; ++e) G[e] = 0.f;
  const int p0 = (blockIdx.y * 8 + warp) * LITE_PPW, p1 = min(a.npairs, p0 + LITE_PPW);
  for (int p = p0; p < p1; ++p) {
    const int k = a.pairs[2 * p], l = a.pairs[2 * p + 1];
    float Rk[9], Rl[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      Rk[e] = SF_IM(a.RT, k * RW + e, Bp, b);
      Rl[e] = SF_IM(a.RT, l * RW + e, Bp, b);
    }
    const float4* A4 = reinterpret_cast<const float4*>(a.A) + (size_t)p * 9 * (NGP / 4);
#pragma unroll
    for (int aa = 0; aa < 3; ++aa)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const float rkl = fmaf(Rk[aa], Rl[bb], fmaf(Rk[3 + aa], Rl[3 + bb], Rk[6 + aa] * Rl[6 + bb]));  // (R_k^T R_l)[aa][bb]
#pragma unroll
        for (int e4 = 0; e4 < NGP / 4; ++e4) {
          const float4 c4 = __ldg(A4 + (aa * 3 + bb) * (NGP / 4) + e4);
          G[4 * e4] = fmaf(rkl, c4.x, G[4 * e4]);
          G[4 * e4 + 1] = fmaf(rkl, c4.y, G[4 * e4 + 1]);
          G[4 * e4 + 2] = fmaf(rkl, c4.z, G[4 * e4 + 2]);
          G[4 * e4 + 3] = fmaf(rkl, c4.w, G[4 * e4 + 3]);
        }
      }
  }
  cta_reduce8<NGP>(G, s_red, warp, lane);
  if (warp == 0) {
    float* out = a.out + (size_t)blockIdx.y * NG * Bp + b;
#pragma unroll
    for (int e = 0; e < NG; ++e) out[(size_t)e * Bp] = G[e];
  }
}

// k_pair_feat: the features of the pair term for the tensor-core path (vposed_tc.cu tc_gemm_run):
//   X[b][p*9 + a*3 + c] = (R_k^T R_l)[a][c]   for pair p = (k, l),
// written pre-split into tf32-exact hi / lo parts, [Bt][Kt] row-major (columns >= 9 npairs and rows >= Bp are zero).
// One CTA per 32 instances: the rotations are staged in shared memory transposed to [instance][J*9] (odd stride),
// warp w handles instances 4w..4w+3, lanes run over the pairs so a warp's stores cover 1152 contiguous bytes.
struct PairFeatArgs {
  const float* RT;  // [J*RW][Bp]
  const int32_t* pairs;
  float *hi, *lo;   // [Bt][Kt]
  int npairs, J, RW, Bp, Kt;
};

static __global__ void __launch_bounds__(256) k_pair_feat(const PairFeatArgs a) {
  extern __shared__ __align__(16) float s_R[];  // [32][stride]
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int J9 = a.J * 9, stride = J9 | 1;
  const bool valid = g * 32 < a.Bp;  // Bp % 32 == 0: a group is entirely inside or outside the batch
  if (valid) {
    for (int q = threadIdx.x; q < J9 * 32; q += 256) {
      const int row = q >> 5, l = q & 31;
      const int k = row / 9, e = row % 9;
      s_R[l * stride + row] = a.RT[(size_t)(k * a.RW + e) * a.Bp + g * 32 + l];
    }
  }
  __syncthreads();
  const int K9 = a.npairs * 9;
  // blockIdx.y splits the pair range (each CTA re-stages the rotations: a few KB from L2) so that several CTAs
  // per SM hide the latency.  Each warp builds one instance's slice of the feature row in shared memory (stride-9
  // writes are bank-conflict free) and copies it out with coalesced stores: writing the 9 values of a pair
  // straight to global memory touched every 32-byte sector once per value (226 MB of partial-sector writes for
  // 23 MB of data).
  const int per = (a.npairs + gridDim.y - 1) / gridDim.y;
  const int p_lo = blockIdx.y * per, p_hi = min(a.npairs, p_lo + per);
  const bool last = blockIdx.y == gridDim.y - 1;
  const int n9 = max(0, p_hi - p_lo) * 9;
  float* sbuf = s_R + 32 * stride + (size_t)warp * (2 * per * 9);  // [hi: per*9][lo: per*9]
  for (int ii = 0; ii < 4; ++ii) {
    const int i = warp * 4 + ii;
    float* hi = a.hi + (size_t)(g * 32 + i) * a.Kt;
    float* lo = a.lo + (size_t)(g * 32 + i) * a.Kt;
    if (valid) {
      const float* R = s_R + i * stride;
      for (int p = p_lo + lane; p < p_hi; p += 32) {
        const int k = __ldg(a.pairs + 2 * p), l = __ldg(a.pairs + 2 * p + 1);
        const float* Rk = R + k * 9;
        const float* Rl = R + l * 9;
        float* oh = sbuf + (p - p_lo) * 9;
        float* ol = oh + per * 9;
#pragma unroll
        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            const float x = fmaf(Rk[aa], Rl[cc], fmaf(Rk[3 + aa], Rl[3 + cc], Rk[6 + aa] * Rl[6 + cc]));
            const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            oh[aa * 3 + cc] = h;
            ol[aa * 3 + cc] = x - h;
          }
      }
    }
    __syncwarp();
    for (int q = lane; q < n9; q += 32) {
      hi[p_lo * 9 + q] = valid ? sbuf[q] : 0.f;
      lo[p_lo * 9 + q] = valid ? sbuf[per * 9 + q] : 0.f;
    }
    if (last)
      for (int q = K9 + lane; q < a.Kt; q += 32) hi[q] = lo[q] = 0.f;
    __syncwarp();
  }
}

template <int NS>
__global__ void __launch_bounds__(256) k_gram_trans(const GramClosedArgs a) {
  extern __shared__ __align__(16) float s_rows[];  // [J][3 + NS][32], then reused as the reduction scratch
  constexpr int NG = NS * (NS + 1) / 2, NGP = (NG + 3) / 4 * 4, NSP4 = (NS + 3) / 4 * 4;
  constexpr int RW = 12 + 3 * NS, CW = 3 + NS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x, c = blockIdx.y;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  {
    const int n16 = a.J * CW * 8;  // 16-byte pieces: row = 128 B
    for (int q = threadIdx.x; q < n16; q += 256) {
      const int r = q >> 3, part = q & 7;
      const int k = r / CW, e = r % CW;
      const int row = (e < 3) ? k * RW + c * 3 + e : k * RW + 9 + c * (1 + NS) + 1 + (e - 3);
      const float* src = a.RT + (size_t)row * Bp + g * 32 + part * 4;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rows + r * 32 + part * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  float G[NGP];
#pragma unroll
  for (int e = 0; e < NGP; ++e) G[e] = 0.f;
  for (int l = warp; l < a.J; l += 8) {
    float Q[NSP4];
#pragma unroll
    for (int s = 0; s < NSP4; ++s) Q[s] = 0.f;
    for (int q = a.lstart[l]; q < a.lstart[l + 1]; ++q) {
      const int k = a.lk[q];
      const float* rk = s_rows + (size_t)(k * CW) * 32 + lane;
      const float r0 = rk[0], r1 = rk[32], r2 = rk[64];
      const float wh = __ldg(a.Wh + q);
      const float4* bm = reinterpret_cast<const float4*>(a.Bm + (size_t)q * 3 * NSP4);
#pragma unroll
      for (int s4 = 0; s4 < NSP4 / 4; ++s4) {
        const float4 b0 = __ldg(bm + s4), b1 = __ldg(bm + NSP4 / 4 + s4), b2 = __ldg(bm + 2 * (NSP4 / 4) + s4);
        const float bx[4][3] = {{b0.x, b1.x, b2.x}, {b0.y, b1.y, b2.y}, {b0.z, b1.z, b2.z}, {b0.w, b1.w, b2.w}};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int s = 4 * s4 + u;
          if (s < NS) Q[s] += fmaf(r0, bx[u][0], fmaf(r1, bx[u][1], fmaf(r2, bx[u][2], wh * rk[(3 + s) * 32])));
        }
      }
    }
    float Tl[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) Tl[s] = s_rows[(size_t)(l * CW + 3 + s) * 32 + lane];
    int e = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int t = s; t < NS; ++t) {
        G[e] = fmaf(Tl[t], Q[s], fmaf(Tl[s], Q[t], G[e]));
        ++e;
      }
  }
  cta_reduce8<NGP>(G, s_rows, warp, lane);  // (its leading __syncthreads orders the reuse of s_rows)
  if (warp == 0) {
    float* out = a.out + (size_t)(a.n_pair_ctas + c) * NG * Bp + b;
#pragma unroll
    for (int e = 0; e < NG; ++e) out[(size_t)e * Bp] = G[e];
  }
}

// ---------------------------------------------------------------------------------------
// k_stats_lite<NS, WEIGHTED>: the statistics pass against the skinned current fit (the hot REF == 1 mode of
// k_stats_rec), restructured like k_shape_lite: joint rows as float4 quads (A[c][0..2], tau[c]) cached in
// registers per slot, records staged per warp by bulk async copies, targets / posed template two vertices ahead.
//   ref_v = (sum_k w_vk [A_k | tau_k]) [v_posed_v + S_v beta; 1];   per part: M += (t - ct)(ref - ca)^T, sums.
// Same partial layout as k_stats ([segment][16][Bp]).
// ---------------------------------------------------------------------------------------
struct StatsLiteArgs {
  const float* tT;
  const float* vwT;
  const float* ct0;      // [3J][Bp]
  const float* ca0;      // [3J][Bp]
  const float* vposedT;
  const float* beta;     // [NS][Bp]
  const float* skin4;    // [J*3][Bp] float4
  float* aT_out;         // [3V][Bp] or null
  float* partials;       // [n_segments][16][Bp]
  const float* rec;
  const int32_t* seg_start;
  const int32_t* seg_part;
  const int32_t* part_flags;
  const uint8_t* slot_mask;  // as in LiteArgs, or null
  int n_segments, Bp, J, all_segments, segs_per_warp;
};

__host__ __device__ inline size_t stats_lite_smem_bytes(int J, int rec_len, int warps) {
  return ((size_t)J * 3 * 128 + (size_t)warps * (2 * lite_stage_floats(rec_len))) * sizeof(float) + (size_t)warps * 16 + 16;
}

template <int NS, bool WEIGHTED, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_stats_lite(const StatsLiteArgs a, const __grid_constant__ CUtensorMap map_t, const __grid_constant__ CUtensorMap map_vp) {
  extern __shared__ __align__(128) float s_st[];
  constexpr int NSP = Rec<NS>::NSP, REC = Rec<NS>::LEN;
  using Stager = VertStager<REC>;
  constexpr int BOX = Stager::BOX, STAGE = Stager::STAGE;
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  const bool has_mask = a.slot_mask != nullptr;
  const float4* sq = reinterpret_cast<const float4*>(s_st);  // [J*3][32]
  float* wbase = s_st + (size_t)a.J * 3 * 128;
  Stager rs;
  rs.buf = wbase + (size_t)warp * (2 * STAGE);
  rs.bar = reinterpret_cast<uint64_t*>(wbase + (size_t)WARPS * (2 * STAGE)) + 2 * warp;
  rs.rec_src = a.rec;
  rs.mt = &map_t;
  rs.mv = &map_vp;
  rs.col = g * 32;
  rs.phase = 0;
  if (lane == 0) {
    sf_mbar_init(rs.bar, 1);
    sf_mbar_init(rs.bar + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  {
    const int n = a.J * 3 * 32;
    const float4* src = reinterpret_cast<const float4*>(a.skin4);
    for (int q = threadIdx.x; q < n; q += WARPS * 32) {
      const int r = q >> 5, l = q & 31;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_st + (size_t)q * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)r * Bp + g * 32 + l) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  float beta[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) beta[s] = SF_IM(a.beta, s, Bp, b);
  for (int q = 0; q < a.segs_per_warp; ++q) {
    const int seg = (blockIdx.x * a.segs_per_warp + q) * WARPS + warp;
    if (seg >= a.n_segments) break;
    const int part = a.seg_part[seg];
    const bool stat = (a.part_flags[part] & 1) != 0;
    if (!stat && !(a.aT_out != nullptr && a.all_segments)) continue;
    const int i0 = a.seg_start[seg], i1 = a.seg_start[seg + 1];
    rs.i0 = i0;
    rs.i1 = i1;
    rs.issue(0, lane);
    rs.issue(1, lane);
    float ct[3], ca[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ct[c] = SF_IM(a.ct0, part * 3 + c, Bp, b);
      ca[c] = SF_IM(a.ca0, part * 3 + c, Bp, b);
    }
    float M[9], st[3], sa[3], W = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) M[e] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) st[c] = sa[c] = 0.f;
    JointCache jc;
    jc.reset();
    const int nsub = (i1 - i0 + LITE_VS - 1) / LITE_VS;
    for (int k = 0; k < nsub; ++k) {
      const int first = i0 + k * LITE_VS;
      const int nv = min(LITE_VS, i1 - first);
      // the (rare) per-vertex weights stay on a plain coalesced load, issued before the stage wait
      float wnext = WEIGHTED ? SF_IM(a.vwT, first, Bp, b) : 1.f;
      unsigned mk = 0u;  // precomputed reload bits of the sub-block's vertices
      if (has_mask && lane < nv) mk = __ldg(a.slot_mask + first + lane);
      rs.wait(k);
      if (k >= 1) rs.issue(k + 1, lane);
      const float* sg = rs.stage(k);
#pragma unroll LITE_UNROLL
      for (int u = 0; u < nv; ++u) {
        {
          const int i = first + u;
          float t[3], x[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            t[c] = sg[(u * 3 + c) * 32 + lane];
            x[c] = sg[BOX + (u * 3 + c) * 32 + lane];
          }
          const float wv = wnext;
          if (WEIGHTED && u + 1 < nv) wnext = SF_IM(a.vwT, i + 1, Bp, b);
          const float* rec = sg + 2 * BOX + u * REC;
          const float4 w4 = *reinterpret_cast<const float4*>(rec);
          const float wk[4] = {w4.x, w4.y, w4.z, w4.w};
          unsigned m;
          int4 j4 = make_int4(0, 0, 0, 0);
          if (has_mask) {
            m = __shfl_sync(0xffffffffu, mk, u);
            if (m) j4 = *reinterpret_cast<const int4*>(rec + 4);
          } else {
            j4 = *reinterpret_cast<const int4*>(rec + 4);
            m = (wk[0] != 0.f && j4.x != jc.j[0] ? 1u : 0u) | (wk[1] != 0.f && j4.y != jc.j[1] ? 2u : 0u) |
                (wk[2] != 0.f && j4.z != jc.j[2] ? 4u : 0u) | (wk[3] != 0.f && j4.w != jc.j[3] ? 8u : 0u);
          }
          if (m) {  // warp-uniform, ~4 % of the vertices
            const int jk[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (m & (1u << kk)) {
                jc.j[kk] = jk[kk];
#pragma unroll
                for (int c = 0; c < 3; ++c) jc.q[kk][c] = sq[(size_t)(jk[kk] * 3 + c) * 32 + lane];
              }
            }
          }
          constexpr int NV4 = (3 * NSP + 3) / 4;
          float sdv[NV4 * 4];  // shapedirs[c][s] of the record as 16-byte words
#pragma unroll
          for (int q4 = 0; q4 < NV4; ++q4) {
            const float4 v4 = *reinterpret_cast<const float4*>(rec + 8 + 4 * q4);
            sdv[4 * q4] = v4.x; sdv[4 * q4 + 1] = v4.y; sdv[4 * q4 + 2] = v4.z; sdv[4 * q4 + 3] = v4.w;
          }
          float vs[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float2 y2 = make_float2(x[c], 0.f);
#pragma unroll
            for (int s2 = 0; s2 < NSP; s2 += 2)
              y2 = sf_fma2(make_float2(sdv[c * NSP + s2], sdv[c * NSP + s2 + 1]),
                           make_float2(beta[s2], (s2 + 1 < NS) ? beta[s2 + 1] : 0.f), y2);
            vs[c] = y2.x + y2.y;
          }
          float2 B2[6];
          jc.blend(wk, B2);
          float ref[3];
#pragma unroll
          for (int c = 0; c < 3; ++c)
            ref[c] = fmaf(B2[2 * c].x, vs[0], fmaf(B2[2 * c].y, vs[1], fmaf(B2[2 * c + 1].x, vs[2], B2[2 * c + 1].y)));
          if (a.aT_out != nullptr) {
#pragma unroll
            for (int c = 0; c < 3; ++c) SF_IM(a.aT_out, i * 3 + c, Bp, b) = ref[c];
          }
          if (stat) {
            float dt[3], wa[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              dt[c] = t[c] - ct[c];
              wa[c] = WEIGHTED ? wv * (ref[c] - ca[c]) : (ref[c] - ca[c]);
              st[c] = WEIGHTED ? fmaf(wv, dt[c], st[c]) : st[c] + dt[c];
              sa[c] += wa[c];
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c) M[r * 3 + c] = fmaf(dt[r], wa[c], M[r * 3 + c]);
            W += wv;
          }
        }
      }
    }
    if (stat) {
      float* out = a.partials + (size_t)seg * 16 * Bp + b;
#pragma unroll
      for (int e = 0; e < 9; ++e) out[(size_t)e * Bp] = M[e];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        out[(size_t)(9 + c) * Bp] = st[c];
        out[(size_t)(12 + c) * Bp] = sa[c];
      }
      out[(size_t)15 * Bp] = W;
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_stats_tmpl<WEIGHTED, WARPS>: the statistics pass of the FIRST rotation fit, against the constant template mesh
// (pt/bodyfitter.py:384-394; REF == 0 of k_stats_rec).  Only the targets stream from HBM: per-warp ring of
// TMPL_NST stages of TMPL_VS vertices (one 2D TMA box each).  9 FMAs per vertex: the pass is bound by the latency of
// what it loads, so nothing a segment needs is requested when the segment starts:
//   * the descriptors of the warp's (<= 32) segments are read once, one segment per lane, and handed out by shuffles;
//   * the target ring runs across segment boundaries (a producer cursor NST - 1 boxes ahead of the consumer);
//   * the part centres and the template coordinates of segment q + 1 (<= 96 floats = 3 registers per lane) are
//     requested before segment q is processed.
// ---------------------------------------------------------------------------------------
constexpr int TMPL_VS = 8, TMPL_NST = 3;
__host__ __device__ inline size_t stats_tmpl_smem_bytes(int warps) {
  return (size_t)warps * TMPL_NST * (3 * TMPL_VS * 32) * sizeof(float) + (size_t)warps * TMPL_NST * 8 + 16;
}

struct TmplSeg {
  float ct[3], ca[3], tm[3];
  int i0, i1, seg;
};

template <bool WEIGHTED, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_stats_tmpl(const StatsLiteArgs a, const float* __restrict__ template_fit, const float* __restrict__ ca0_const,
             const __grid_constant__ CUtensorMap map_t) {
  extern __shared__ __align__(128) float s_tm[];
  constexpr int BOX = 3 * TMPL_VS * 32;
  constexpr unsigned FULL = 0xffffffffu;
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp, spw = a.segs_per_warp;
  const int b = g * 32 + lane;
  float* buf = s_tm + (size_t)warp * TMPL_NST * BOX;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_tm + (size_t)WARPS * TMPL_NST * BOX) + TMPL_NST * warp;
  if (lane == 0)
    for (int s = 0; s < TMPL_NST; ++s) sf_mbar_init(bar + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // descriptors of this warp's segments: lane q holds segment q
  int m_i0 = 0, m_i1 = 0, m_part = 0;
  bool m_act = false;
  if (lane < spw) {
    const int seg = (blockIdx.x * spw + lane) * WARPS + warp;
    if (seg < a.n_segments) {
      m_part = a.seg_part[seg];
      m_act = (a.part_flags[m_part] & 1) != 0;
      m_i0 = a.seg_start[seg];
      m_i1 = a.seg_start[seg + 1];
    }
  }
  __syncwarp();
  const unsigned act = __ballot_sync(FULL, m_act && m_i1 > m_i0);
  for (unsigned em = __ballot_sync(FULL, m_act && m_i1 <= m_i0); em; em &= em - 1) {  // empty segment of a live part
    float* out = a.partials + (size_t)((blockIdx.x * spw + (__ffs(em) - 1)) * WARPS + warp) * 16 * Bp + b;
#pragma unroll
    for (int e = 0; e < 16; ++e) out[(size_t)e * Bp] = 0.f;
  }
  // producer cursor over (segment, box)
  unsigned pm = act;
  int pi0 = 0, pnsub = 0, pk = 0, issued = 0;
  auto p_seg = [&]() {
    const int q = pm ? __ffs(pm) - 1 : 0;
    pi0 = __shfl_sync(FULL, m_i0, q);
    pnsub = (__shfl_sync(FULL, m_i1, q) - pi0 + TMPL_VS - 1) / TMPL_VS;
    pk = 0;
  };
  p_seg();
  auto issue = [&]() {
    __syncwarp();
    if (pm != 0) {
      if (lane == 0) {
        const int s = issued % TMPL_NST;
        sf_mbar_expect_tx(bar + s, (uint32_t)BOX * 4u);
        sf_tma_2d(buf + (size_t)s * BOX, &map_t, bar + s, g * 32, (pi0 + pk * TMPL_VS) * 3);
      }
      ++issued;
      if (++pk == pnsub) {
        pm &= pm - 1;
        p_seg();
      }
    }
  };
  auto fetch = [&](int q, TmplSeg& d) {
    d.i0 = __shfl_sync(FULL, m_i0, q);
    d.i1 = __shfl_sync(FULL, m_i1, q);
    const int part = __shfl_sync(FULL, m_part, q);
    d.seg = (blockIdx.x * spw + q) * WARPS + warp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d.ct[c] = SF_IM(a.ct0, part * 3 + c, Bp, b);
      d.ca[c] = __ldg(ca0_const + part * 3 + c);
    }
    const int n3 = (d.i1 - d.i0) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) d.tm[r] = (r * 32 + lane < n3) ? __ldg(template_fit + (size_t)d.i0 * 3 + r * 32 + lane) : 0.f;
  };
  for (int k = 0; k < TMPL_NST - 1; ++k) issue();
  unsigned cm = act;
  int consumed = 0;
  uint32_t phase = 0;
  TmplSeg cur, nxt;
  if (cm) fetch(__ffs(cm) - 1, cur);
  while (cm) {
    cm &= cm - 1;
    if (cm) fetch(__ffs(cm) - 1, nxt);
    const int i0 = cur.i0, i1 = cur.i1;
    const int nsub = (i1 - i0 + TMPL_VS - 1) / TMPL_VS;
    float M[9], st[3], sa[3], W = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) M[e] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) st[c] = sa[c] = 0.f;
    for (int k = 0; k < nsub; ++k) {
      const int first = i0 + k * TMPL_VS;
      const int nv = min(TMPL_VS, i1 - first);
      // the template coordinates of this box: flat element k * 24 + lane of the segment's (prefetched) 96
      float tm;
      if (k < 96 / (3 * TMPL_VS)) {
        const int f = k * 3 * TMPL_VS + lane;
        const float v0 = __shfl_sync(FULL, cur.tm[0], f & 31), v1 = __shfl_sync(FULL, cur.tm[1], f & 31),
                    v2 = __shfl_sync(FULL, cur.tm[2], f & 31);
        tm = (f < 32) ? v0 : ((f < 64) ? v1 : v2);
      } else {  // (segments longer than 32 vertices: not produced by the current tables)
        tm = (lane < nv * 3) ? __ldg(template_fit + (size_t)first * 3 + lane) : 0.f;
      }
      float wv8[TMPL_VS];
      if (WEIGHTED) {
#pragma unroll
        for (int u = 0; u < TMPL_VS; ++u) wv8[u] = (u < nv) ? SF_IM(a.vwT, first + u, Bp, b) : 0.f;
      }
      issue();
      const int s = consumed % TMPL_NST;
      sf_mbar_wait(bar + s, (phase >> s) & 1u);
      phase ^= 1u << s;
      ++consumed;
      const float* sg = buf + (size_t)s * BOX;
#pragma unroll
      for (int u = 0; u < TMPL_VS; ++u) {
        float x[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = __shfl_sync(FULL, tm, u * 3 + c);
        if (u < nv) {
          const float wv = WEIGHTED ? wv8[u] : 1.f;
          float dt[3], wa[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            dt[c] = sg[(u * 3 + c) * 32 + lane] - cur.ct[c];
            wa[c] = WEIGHTED ? wv * (x[c] - cur.ca[c]) : (x[c] - cur.ca[c]);
            st[c] = WEIGHTED ? fmaf(wv, dt[c], st[c]) : st[c] + dt[c];
            sa[c] += wa[c];
          }
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[r * 3 + c] = fmaf(dt[r], wa[c], M[r * 3 + c]);
          W += wv;
        }
      }
    }
    float* out = a.partials + (size_t)cur.seg * 16 * Bp + b;
#pragma unroll
    for (int e = 0; e < 9; ++e) out[(size_t)e * Bp] = M[e];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      out[(size_t)(9 + c) * Bp] = st[c];
      out[(size_t)(12 + c) * Bp] = sa[c];
    }
    out[(size_t)15 * Bp] = W;
    cur = nxt;
  }
}

}  // namespace sf
