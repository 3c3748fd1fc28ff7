// Per-instance "small" stages of BodyFitter.fit: one thread per instance, instance-minor
// arrays ([row][Bp]) so every access is coalesced across the warp.  All 3x3 / SxS algebra is
// register- or local-resident; nothing here touches per-vertex data.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smplfit_b200.h"
#include "linalg.cuh"

namespace sf {

#ifndef SF_IM
#define SF_IM(ptr, row, Bp, b) (ptr)[(size_t)(row) * (size_t)(Bp) + (size_t)(b)]
#endif

__host__ __device__ inline int quad_rows_ns(int ns) { const int nsp = (ns + 1) / 2 * 2; return (14 + 3 * nsp + 3) / 4 * 4; }

struct TreeTables {
  const int32_t* parents;
  const int32_t* part_kind;
  const int32_t* part_copy_src;
  const int32_t* part_flags;
  const int32_t* cas_table;
  const int32_t* cas_count;
  const int32_t* part_seg_begin;
  const float* Jt_ext;  // (J,3,1+NS)
  int J, NS, max_cas;
};

__device__ __forceinline__ void load3(const float* p, int row, int Bp, int b, float* out) {
  out[0] = SF_IM(p, row * 3 + 0, Bp, b);
  out[1] = SF_IM(p, row * 3 + 1, Bp, b);
  out[2] = SF_IM(p, row * 3 + 2, Bp, b);
}
__device__ __forceinline__ void load3_or_const(const float* p, const float* cst, int row, int Bp, int b, float* out) {
  if (p != nullptr) {
    load3(p, row, Bp, b, out);
  } else {
    out[0] = __ldg(cst + row * 3);
    out[1] = __ldg(cst + row * 3 + 1);
    out[2] = __ldg(cst + row * 3 + 2);
  }
}

// Segment partials of one part -> centred cross-covariance about (ct, ca):
//   sum w (t-ct)(a-ca)^T = M + st (ca0-ca)^T + (ct0-ct) sa^T + W (ct0-ct)(ca0-ca)^T
// the 16 statistics of a part: its segment partials (every `step`-th segment starting at `first`) summed in double
__device__ inline void part_sums(const float* partials, const int32_t* seg_begin, int part, int Bp, int b, double* acc,
                                 int first = 0, int step = 1) {
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0;
#pragma unroll 4
  for (int s = seg_begin[part] + first; s < seg_begin[part + 1]; s += step) {  // unrolled: 4 segments' loads in flight
    const float* p = partials + (size_t)s * 16 * Bp + b;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] += (double)p[(size_t)e * Bp];
  }
}
// sums (as float) -> centred cross-covariance about (ct, ca)
__device__ inline void covariance_from_sums(const float* sm, const float* ct0, const float* ca0, const float* ct, const float* ca,
                                            float* A, float mM = 1.f, float mst = 1.f, float msa = 1.f) {
  const float dt[3] = {ct0[0] - ct[0], ct0[1] - ct[1], ct0[2] - ct[2]};
  const float da[3] = {ca0[0] - ca[0], ca0[1] - ca[1], ca0[2] - ca[2]};
  const float W = sm[15];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      A[r * 3 + c] = mM * sm[r * 3 + c] + mst * sm[9 + r] * da[c] + dt[r] * msa * sm[12 + c] + W * dt[r] * da[c];
}
__device__ inline void part_covariance(const float* partials, const int32_t* seg_begin, int part, int Bp, int b,
                                       const float* ct0, const float* ca0, const float* ct, const float* ca,
                                       float* A, float mM = 1.f, float mst = 1.f, float msa = 1.f) {
  double acc[16];
  part_sums(partials, seg_begin, part, Bp, b, acc);
  float sm[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) sm[e] = (float)acc[e];
  covariance_from_sums(sm, ct0, ca0, ct, ca, A, mM, mst, msa);
}

// ---------------------------------------------------------------------------------------
// k_rot_solve: _fit_global_rotations after the vertex pass (pt/bodyfitter.py:1352-1416),
// the left-multiplication onto the running orientations (:422-433), and the pose-dependent
// front part of the next shape solve (:869-911): relative rotations -> pose features,
// joint positions with their shape Jacobian (FK), per-joint translation offsets.
// ---------------------------------------------------------------------------------------
struct RotArgs {
  const float* partials;   // [n_segments][16][Bp]
  const float* tjT;        // [3J][Bp] target joints (given or regressed), centred
  const float* ajT;        // [3J][Bp] reference joints, or null -> aj_const
  const float* aj_const;   // (J,3)
  const float* ca0T;       // [3J][Bp] provisional reference centres of the stats pass, or null -> ca0_const
  const float* ca0_const;  // (J,3)
  const float* jwT;        // [J][Bp] or null
  const float* R_old;      // [9J][Bp] or null (identity)
  float* R_new;            // [9J][Bp]
  float* RT;               // [J*(12+3NS)][Bp] or null (skip the shape front)
  float* RT4;              // optional quad layout of the same rows (fit_kernels.cuh Quad<NS> / CLay<NS>), or null
  int rt4_clay;            // 0: Quad<NS> row order (k_shape_pass_v2), 1: coordinate-major CLay<NS> (k_shape_pass_v3)
  float* RT12;             // optional [J*3][Bp] float4 (R[c][0..2], T0[c]) for k_shape_lite, or null
  float* Pext;             // [J*3*(1+NS)][Bp]
  float* feat;             // [Bp][Kp]
  TreeTables t;
  int B, Bp, Kp;
  // k_front_fused only: the fp16 hi / lo feature rows [vec(R_rel[1:] - I) | 0 ...] of the fused passes (fit_fused.cu),
  // [>= Bp][fq_kf] row-major, and the number of warps that run kinematic-chain columns at a time (shared-memory budget)
  __half* fq_hi;
  __half* fq_lo;
  int fq_kf, fk_warps;
};

// Rotation fit of one body part (device function shared by the kernels below).
__device__ inline void fit_part(const RotArgs& a, int i, int b, float* R, const float* sums = nullptr) {
  const int Bp = a.Bp;
  const int kind = a.t.part_kind[i];
  if (kind == 0 || kind == 4) {
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = (e % 4 == 0) ? 1.f : 0.f;
    return;
  }
  const int n = a.t.cas_count[i];
  const int32_t* cas = a.t.cas_table + i * a.t.max_cas;
  // children-mean centres (center_matrix, pt/bodyfitter.py:125-129, :1352-1353)
  float mt[3] = {0.f, 0.f, 0.f}, ma[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < n; ++k) {
    float tj[3], aj[3];
    load3(a.tjT, cas[k], Bp, b, tj);
    load3_or_const(a.ajT, a.aj_const, cas[k], Bp, b, aj);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mt[c] += tj[c];
      ma[c] += aj[c];
    }
  }
  const float inv = 1.f / (float)n;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    mt[c] *= inv;
    ma[c] *= inv;
  }
  float A[9];
  if (kind == 1) {
    // multi-joint part: Kabsch on its joints alone (pt/bodyfitter.py:1361-1383)
#pragma unroll
    for (int e = 0; e < 9; ++e) A[e] = 0.f;
    for (int k = 0; k < n; ++k) {
      float tj[3], aj[3];
      load3(a.tjT, cas[k], Bp, b, tj);
      load3_or_const(a.ajT, a.aj_const, cas[k], Bp, b, aj);
      const float w = a.jwT ? SF_IM(a.jwT, cas[k], Bp, b) : 1.f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) A[r * 3 + c] = fmaf(tj[r] - mt[r], w * (aj[c] - ma[c]), A[r * 3 + c]);
    }
    proj_so3(A, R);
    return;
  }
  float ct0[3], ca0[3];
  load3(a.tjT, i, Bp, b, ct0);
  load3_or_const(a.ca0T, a.ca0_const, i, Bp, b, ca0);
  if (sums != nullptr) covariance_from_sums(sums, ct0, ca0, mt, ma, A);
  else part_covariance(a.partials, a.t.part_seg_begin, i, Bp, b, ct0, ca0, mt, ma, A);
  if (kind == 3) {  // leaf part: Kabsch on its vertices
    proj_so3(A, R);
    return;
  }
  // bone part: swing aligns the bone, twist from the vertex covariance (pt/bodyfitter.py:1389-1412)
  float bt[3], br[3], t0[3], t1[3], r0[3], r1[3];
  load3(a.tjT, cas[0], Bp, b, t0);
  load3(a.tjT, cas[1], Bp, b, t1);
  load3_or_const(a.ajT, a.aj_const, cas[0], Bp, b, r0);
  load3_or_const(a.ajT, a.aj_const, cas[1], Bp, b, r1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    bt[c] = t1[c] - t0[c];
    br[c] = r1[c] - r0[c];
  }
  const float nt = sqrtf(bt[0] * bt[0] + bt[1] * bt[1] + bt[2] * bt[2]);
  const float nr = sqrtf(br[0] * br[0] + br[1] * br[1] + br[2] * br[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    bt[c] = div_no_nan(bt[c], nt);
    br[c] = div_no_nan(br[c], nr);
  }
  float Rs[9], H[9];
  align_unit_vectors(br, bt, Rs);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) H[r * 3 + c] = Rs[r * 3] * A[c * 3] + Rs[r * 3 + 1] * A[c * 3 + 1] + Rs[r * 3 + 2] * A[c * 3 + 2];
  const float trH = H[0] + H[4] + H[8];
  float Hb[3];
  mat3_vec(H, bt, Hb);
  const float bHb = bt[0] * Hb[0] + bt[1] * Hb[1] + bt[2] * Hb[2];
  const float vee[3] = {H[5] - H[7], H[6] - H[2], H[1] - H[3]};
  const float ang = atan2f(bt[0] * vee[0] + bt[1] * vee[1] + bt[2] * vee[2], trH - bHb);
  float rv[3] = {bt[0] * ang, bt[1] * ang, bt[2] * ang}, Rt[9];
  rotvec2mat(rv, Rt);
  mat3_mul(Rt, Rs, R);
}

// k_rot_fit: one thread per (instance, part).  Fits the part's rotation (toe parts re-fit their
// foot, pt/bodyfitter.py:1414-1416) and left-multiplies it onto the running orientation
// (:422-433).  In-place safe: thread (b, i) only touches R[i][b].
template <int RF_WARPS>
__global__ void __launch_bounds__(RF_WARPS * 32) k_rot_fit(const RotArgs a) {
  // the part's segment partials are summed by RF_WARPS warps (every RF_WARPS-th segment each: the dependent chain of
  // DRAM / L2 round trips is what bounds this kernel), combined in a fixed order, then warp 0 fits the rotation
  __shared__ double s_acc[RF_WARPS][16][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const int i = blockIdx.y;
  const int Bp = a.Bp;
  const int src = (a.t.part_kind[i] == 4) ? a.t.part_copy_src[i] : i;
  const int kind = a.t.part_kind[src];
  const bool vertex_stats = (kind == 2 || kind == 3);  // block-uniform
  if (vertex_stats) {
    double acc[16];
    part_sums(a.partials, a.t.part_seg_begin, src, Bp, b, acc, warp, RF_WARPS);
#pragma unroll
    for (int e = 0; e < 16; ++e) s_acc[warp][e][lane] = acc[e];
  }
  __syncthreads();
  if (warp != 0) return;
  float sums[16];
  if (vertex_stats) {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      double v = s_acc[0][e][lane];
#pragma unroll
      for (int w = 1; w < RF_WARPS; ++w) v += s_acc[w][e][lane];
      sums[e] = (float)v;
    }
  }
  float Rf[9], Rn[9];
  fit_part(a, src, b, Rf, vertex_stats ? sums : nullptr);
  if (a.R_old != nullptr) {
    float Ro[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Ro[e] = SF_IM(a.R_old, i * 9 + e, Bp, b);
    mat3_mul(Rf, Ro, Rn);
  } else {
#pragma unroll
    for (int e = 0; e < 9; ++e) Rn[e] = Rf[e];
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) SF_IM(a.R_new, i * 9 + e, Bp, b) = Rn[e];
}

__device__ __forceinline__ void rt4_store(float* RT4, int j, int nq, int row, int Bp, int b, float v) {
  RT4[(((size_t)(j * nq + (row >> 2))) * Bp + b) * 4 + (row & 3)] = v;
}

// k_front_rel: one thread per (instance, joint): relative rotation -> pose features, and the
// rotation rows of the [R | T_ext] table (pt/bodyfitter.py:869-876, :913).
static __global__ void __launch_bounds__(32) k_front_rel(const RotArgs a) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y;
  if (b >= a.Bp) return;
  const int Bp = a.Bp, NS = a.t.NS, J = a.t.J;
  const int RW = 12 + 3 * NS;
  float Rj[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    Rj[e] = SF_IM(a.R_new, j * 9 + e, Bp, b);
    SF_IM(a.RT, j * RW + e, Bp, b) = Rj[e];
  }
  if (a.RT12 != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int xx = 0; xx < 3; ++xx) a.RT12[((size_t)(j * 3 + c) * Bp + b) * 4 + xx] = Rj[c * 3 + xx];
  }
  if (a.RT4 != nullptr && a.rt4_clay) {
    const int nsp4 = (NS + 3) / 4 * 4, rc = 4 + nsp4, nq = 3 * rc / 4;
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int xx = 0; xx < 3; ++xx) rt4_store(a.RT4, j, nq, c * rc + xx, Bp, b, Rj[c * 3 + xx]);
      for (int sx = NS; sx < nsp4; ++sx) rt4_store(a.RT4, j, nq, c * rc + 4 + sx, Bp, b, 0.f);
    }
  } else if (a.RT4 != nullptr) {
    const int rows = quad_rows_ns(NS), nq = rows / 4, nsp = (NS + 1) / 2 * 2;
#pragma unroll
    for (int e = 0; e < 9; ++e) rt4_store(a.RT4, j, nq, e, Bp, b, Rj[e]);
    rt4_store(a.RT4, j, nq, 9, Bp, b, 0.f);
    rt4_store(a.RT4, j, nq, 13, Bp, b, 0.f);
    for (int c = 0; c < 3; ++c)
      for (int sx = NS; sx < nsp; ++sx) rt4_store(a.RT4, j, nq, 14 + c * nsp + sx, Bp, b, 0.f);
    for (int r = 14 + 3 * nsp; r < rows; ++r) rt4_store(a.RT4, j, nq, r, Bp, b, 0.f);
  }
  if (j > 0) {
    const int par = a.t.parents[j];
    float Rp[9], rel[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rp[e] = SF_IM(a.R_new, par * 9 + e, Bp, b);
    mat3_tmul(Rp, Rj, rel);
#pragma unroll
    for (int e = 0; e < 9; ++e) a.feat[(size_t)b * a.Kp + (j - 1) * 9 + e] = rel[e];
  } else {
    for (int k = 9 * (J - 1); k < a.Kp; ++k) a.feat[(size_t)b * a.Kp + k] = 0.f;
  }
}

// k_front_fk: one thread per (instance, column s of [position | shape Jacobian]): forward
// kinematics of that column down the tree and the per-joint translation offsets
// (pt/bodyfitter.py:880-911).
static __global__ void __launch_bounds__(32) k_front_fk(const RotArgs a) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  const int s = blockIdx.y;
  if (b >= a.Bp) return;
  const int Bp = a.Bp, NS = a.t.NS, J = a.t.J;
  const int TW = 3 * (1 + NS), RW = 12 + 3 * NS;
  // phase 1: the chain P_j = P_parent + R_parent (J_j - J_parent).  No global stores in this loop, so the loads of
  // the next joints' rotations are issued ahead of the dependent adds (the loop is unrolled).
  float P[SMPLFIT_MAX_JOINTS * 3];
  {
    const float* Jt = a.t.Jt_ext;
    P[0] = __ldg(Jt + s); P[1] = __ldg(Jt + (1 + NS) + s); P[2] = __ldg(Jt + 2 * (1 + NS) + s);
  }
#pragma unroll 4
  for (int j = 1; j < J; ++j) {
    const int par = a.t.parents[j];
    const float* Jt = a.t.Jt_ext + (size_t)j * TW;
    const float* Jp = a.t.Jt_ext + (size_t)par * TW;
    const float d0 = __ldg(Jt + s) - __ldg(Jp + s), d1 = __ldg(Jt + (1 + NS) + s) - __ldg(Jp + (1 + NS) + s),
                d2 = __ldg(Jt + 2 * (1 + NS) + s) - __ldg(Jp + 2 * (1 + NS) + s);
    float Rp[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rp[e] = SF_IM(a.R_new, par * 9 + e, Bp, b);
#pragma unroll
    for (int c = 0; c < 3; ++c) P[j * 3 + c] = P[par * 3 + c] + (Rp[c * 3] * d0 + Rp[c * 3 + 1] * d1 + Rp[c * 3 + 2] * d2);
  }
  // phase 2: per-joint outputs (independent of each other)
  for (int j = 0; j < J; ++j) {
    const float* Jt = a.t.Jt_ext + (size_t)j * TW;
    const float j0 = __ldg(Jt + s), j1 = __ldg(Jt + (1 + NS) + s), j2 = __ldg(Jt + 2 * (1 + NS) + s);
    float Rj[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rj[e] = SF_IM(a.R_new, j * 9 + e, Bp, b);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float tv = P[j * 3 + c] - (Rj[c * 3] * j0 + Rj[c * 3 + 1] * j1 + Rj[c * 3 + 2] * j2);
      SF_IM(a.Pext, j * TW + c * (1 + NS) + s, Bp, b) = P[j * 3 + c];
      SF_IM(a.RT, j * RW + 9 + c * (1 + NS) + s, Bp, b) = tv;
      if (a.RT12 != nullptr && s == 0) a.RT12[((size_t)(j * 3 + c) * Bp + b) * 4 + 3] = tv;
      if (a.RT4 != nullptr && a.rt4_clay) {
        const int rc = 4 + (NS + 3) / 4 * 4;
        rt4_store(a.RT4, j, 3 * rc / 4, c * rc + (s == 0 ? 3 : 4 + (s - 1)), Bp, b, tv);
      } else if (a.RT4 != nullptr) {
        const int nsp = (NS + 1) / 2 * 2;
        rt4_store(a.RT4, j, quad_rows_ns(NS) / 4, s == 0 ? 10 + c : 14 + c * nsp + (s - 1), Bp, b, tv);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_front_fused: k_front_rel + k_front_fk + the feature rows of the fused passes in one kernel (the closed-form path:
// no RT4 layouts).  One CTA (FF_WARPS warps, lane = instance) per 32 instances: the orientations are staged in shared
// memory once (the per-column kinematic chains re-read every parent rotation: 11 x from L2 in k_front_fk), joints are
// dealt to the warps for the relative rotations / row tables, the feature rows [vec(R_rel[1:] - I) | 0 ...] are built
// as fp16 hi / lo in a shared-memory tile and copied out with coalesced stores, then the (1 + NS) chain columns are
// dealt to the warps with their joint positions in shared memory.
// ---------------------------------------------------------------------------------------
__host__ __device__ inline size_t front_fused_smem_bytes(int J, int kf, int fk_warps) {
  const size_t tile = (size_t)2 * 32 * (kf + 2) * sizeof(__half);
  const size_t chains = (size_t)fk_warps * J * 3 * 32 * sizeof(float);
  return (size_t)J * 9 * 32 * sizeof(float) + (tile > chains ? tile : chains);
}

template <int FF_WARPS>
__global__ void __launch_bounds__(FF_WARPS * 32) k_front_fused(const RotArgs a) {
  extern __shared__ __align__(16) float s_ff[];
  const int J = a.t.J, NS = a.t.NS, Bp = a.Bp, Kf = a.fq_kf;
  const int RW = 12 + 3 * NS, TW = 3 * (1 + NS), P = 9 * (J - 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x, b = g * 32 + lane;
  float* sR = s_ff;            // [J*9][32]
  float* sX = sR + J * 9 * 32; // feature tile (phase 1), then the chains' joint positions (phase 2)
  for (int row = warp; row < J * 9; row += FF_WARPS) sR[row * 32 + lane] = a.R_new[(size_t)row * Bp + b];
  const int KS = Kf + 2;  // tile row stride in halves: an odd number of 4-byte words, so a warp's 32 rows hit 32 banks
  __half* th = reinterpret_cast<__half*>(sX);
  __half* tl = th + 32 * KS;
  for (int q = threadIdx.x; q < 32 * (Kf - P); q += FF_WARPS * 32) {
    const int i = q / (Kf - P), k = P + q % (Kf - P);
    th[i * KS + k] = __float2half_rn(0.f);
    tl[i * KS + k] = __float2half_rn(0.f);
  }
  __syncthreads();
  // ---- phase 1: per joint (pt/bodyfitter.py:869-876, :913) ----
  for (int j = warp; j < J; j += FF_WARPS) {
    float Rj[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      Rj[e] = sR[(j * 9 + e) * 32 + lane];
      SF_IM(a.RT, j * RW + e, Bp, b) = Rj[e];
    }
    if (a.RT12 != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int xx = 0; xx < 3; ++xx) a.RT12[((size_t)(j * 3 + c) * Bp + b) * 4 + xx] = Rj[c * 3 + xx];
    }
    if (j > 0) {
      const int par = a.t.parents[j];
      float Rp[9], rel[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) Rp[e] = sR[(par * 9 + e) * 32 + lane];
      mat3_tmul(Rp, Rj, rel);
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        a.feat[(size_t)b * a.Kp + (j - 1) * 9 + e] = rel[e];
        const float v = rel[e] - ((e % 4 == 0) ? 1.f : 0.f);
        const __half hv = __float2half_rn(v);
        th[lane * KS + (j - 1) * 9 + e] = hv;
        tl[lane * KS + (j - 1) * 9 + e] = __float2half_rn(v - __half2float(hv));
      }
    } else {
      for (int k = 9 * (J - 1); k < a.Kp; ++k) a.feat[(size_t)b * a.Kp + k] = 0.f;
    }
  }
  __syncthreads();
  for (int i = warp; i < 32; i += FF_WARPS) {  // coalesced copy-out of the feature rows (zero rows past the batch)
    const size_t row = (size_t)(g * 32 + i) * Kf;
    const bool live = g * 32 + i < a.B;
    for (int k = lane; k < Kf; k += 32) {
      a.fq_hi[row + k] = live ? th[i * KS + k] : __float2half_rn(0.f);
      a.fq_lo[row + k] = live ? tl[i * KS + k] : __float2half_rn(0.f);
    }
  }
  __syncthreads();  // the tile region becomes the chains' scratch
  // ---- phase 2: forward kinematics of column s of [position | shape Jacobian] (pt/bodyfitter.py:880-911) ----
  if (warp < a.fk_warps) {
    float* sP = sX + (size_t)warp * J * 3 * 32;  // [J*3][32]
    for (int s = warp; s <= NS; s += a.fk_warps) {
      {
        const float* Jt = a.t.Jt_ext;
        sP[0 * 32 + lane] = __ldg(Jt + s);
        sP[1 * 32 + lane] = __ldg(Jt + (1 + NS) + s);
        sP[2 * 32 + lane] = __ldg(Jt + 2 * (1 + NS) + s);
      }
      for (int j = 1; j < J; ++j) {
        const int par = a.t.parents[j];
        const float* Jt = a.t.Jt_ext + (size_t)j * TW;
        const float* Jp = a.t.Jt_ext + (size_t)par * TW;
        const float d0 = __ldg(Jt + s) - __ldg(Jp + s), d1 = __ldg(Jt + (1 + NS) + s) - __ldg(Jp + (1 + NS) + s),
                    d2 = __ldg(Jt + 2 * (1 + NS) + s) - __ldg(Jp + 2 * (1 + NS) + s);
        const float* Rp = sR + (size_t)(par * 9) * 32 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          sP[(j * 3 + c) * 32 + lane] = sP[(par * 3 + c) * 32 + lane] + (Rp[(c * 3) * 32] * d0 + Rp[(c * 3 + 1) * 32] * d1 + Rp[(c * 3 + 2) * 32] * d2);
      }
      for (int j = 0; j < J; ++j) {
        const float* Jt = a.t.Jt_ext + (size_t)j * TW;
        const float j0 = __ldg(Jt + s), j1 = __ldg(Jt + (1 + NS) + s), j2 = __ldg(Jt + 2 * (1 + NS) + s);
        const float* Rj = sR + (size_t)(j * 9) * 32 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float pj = sP[(j * 3 + c) * 32 + lane];
          const float tv = pj - (Rj[(c * 3) * 32] * j0 + Rj[(c * 3 + 1) * 32] * j1 + Rj[(c * 3 + 2) * 32] * j2);
          SF_IM(a.Pext, j * TW + c * (1 + NS) + s, Bp, b) = pj;
          SF_IM(a.RT, j * RW + 9 + c * (1 + NS) + s, Bp, b) = tv;
          if (a.RT12 != nullptr && s == 0) a.RT12[((size_t)(j * 3 + c) * Bp + b) * 4 + 3] = tv;
        }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_front_from_R: the shape front alone, for orientations given by the caller
// (fit_with_known_pose, pt/bodyfitter.py:617-639).  Reuses k_rot_solve's tail by running it
// with every part marked "none" and R_old = the given orientations.
// ---------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------
// k_shape_solve<NS>: combine the chunk partials in double, add the joint block, centre with
// the covariance identity, regularise, Cholesky-solve, recover the translation
// (pt/bodyfitter.py:1050-1089; general solve :1199-1283 is algebraically the same system),
// then emit what the next vertex pass needs: betas, translation, the reference joints
// (:1093-1098) and the per-joint skinning transforms [R | T0 + T1 x + trans].
// ---------------------------------------------------------------------------------------
struct SolveArgs {
  const float* partials;  // [n_chunks][NACC][Bp]
  const float* Pext;      // [J*3*(1+NS)][Bp]
  const float* RT;        // [J*(12+3NS)][Bp]
  const float* tjT;       // [3J][Bp] or null (no joint block)
  const float* jwT;       // [J][Bp] or null (unit weights)
  const float* beta_ref;  // (B,S) instance-major or null
  const float* kid_ref;   // (B) or null
  float* beta;            // [NS][Bp]
  float* trans;           // [3][Bp]
  float* refj;            // [3J][Bp]
  float* skin;            // [12J][Bp]
  float* skin4;           // optional [J*3][Bp] float4 (A[c][0..2], tau[c]) copy of skin for k_stats_lite, or null
  const double* wS;      // (J,3,NS) D_k = sum_v w_vk S_v (closed-form SA), or null
  const double* wsum;    // (J)      n_k = sum_v w_vk
  int n_chunks, J, S, Bp, B, V, weighted;
  int sa_closed_form;    // 1: partials hold only [G | r | Sb]; SA from (wS, wsum), W = V
  int scale_mode;        // 0 none, 1 scale_target, 2 scale_fit (extra unknown, pt/bodyfitter.py:1171-1176)
  int shared_noreg;      // k_shared_solve: the summed system already contains the regulariser (partial-share path)
  const float* zpartials;  // [n_zchunks][NS+5][Bp] from k_scale_pass
  int n_zchunks;
  float scale_reg;
  float* scale_out;      // [Bp] scale_corr
  float reg, reg2, kid_reg;
  // closed-form path (lite_kernels.cuh): partials = [n_chunks = n_segments][NS+3+3*slots][Bp] hold r | Sb | Y,
  // G comes from gcf_part [n_gcf][NG][Bp] + G0, and r gets sum_k T_ks . Yd_k
  int lite, lite_nl, n_gcf;
  const float* gcf_part;
  const double* G0;
  const double* Yd;  // [3J + NS + 3][Bp]: Y_k, then the segment sums of r and Sb (k_lite_reduce)
  // feature rows of the fused statistics pass (fit_fused.cu): k_shape_out writes the solved unknowns into columns
  // [fq_p, fq_p + NS) as fp16 hi / lo; null = the caller builds the rows itself
  __half* fq_hi;
  __half* fq_lo;
  int fq_kf, fq_p;
};

// gram_entry<NS>: entry e of [G | r | Sb | SA | W] for instance b: chunk partials + joint block (+ closed-form SA),
// in double (e is warp-uniform).
template <int NS>
__device__ __forceinline__ void gram_entry(const SolveArgs& a, double* __restrict__ Gd, int e, int b) {
  constexpr int NG = NS * (NS + 1) / 2;
  constexpr int NACC = NG + NS + 3 + 3 * NS + 1;
  constexpr int TW = 3 * (1 + NS), RW = 12 + 3 * NS;
  const int Bp = a.Bp, J = a.J;
  // decode the entry
  int kind, s = 0, t = 0, c = 0;  // 0 G(s,t), 1 r(s), 2 Sb(c), 3 SA(c,s), 4 W
  if (e < NG) {
    kind = 0;
    int o = e;
    while (o >= NS - s) { o -= NS - s; ++s; }
    t = s + o;
  } else if (e < NG + NS) { kind = 1; s = e - NG; }
  else if (e < NG + NS + 3) { kind = 2; c = e - NG - NS; }
  else if (e < NG + NS + 3 + 3 * NS) { kind = 3; c = (e - NG - NS - 3) / NS; s = (e - NG - NS - 3) % NS; }
  else kind = 4;
  double acc = 0.0;
  const bool from_partials = !(a.sa_closed_form && kind >= 3);
  if (a.lite && kind <= 2) {
    if (kind == 0) {
      acc = a.G0[e];
#pragma unroll 4
      for (int q = 0; q < a.n_gcf; ++q) acc += (double)a.gcf_part[((size_t)q * NG + e) * Bp + b];
    } else {
      // r / Sb: the sum over the segment partials was taken by k_lite_reduce (rows 3J .. 3J + NS + 2 of Yd)
      acc = a.Yd[(size_t)(3 * J + ((kind == 1) ? s : NS + c)) * Bp + b];
      if (kind == 1) {
#pragma unroll 4
        for (int k = 0; k < J; ++k)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            acc += (double)a.RT[(size_t)(k * RW + 9 + cc * (1 + NS) + 1 + s) * Bp + b] * a.Yd[(size_t)(k * 3 + cc) * Bp + b];
      }
    }
  } else if (from_partials) {
#pragma unroll 8
    for (int q = 0; q < a.n_chunks; ++q) acc += (double)a.partials[((size_t)q * NACC + e) * Bp + b];
  } else if (kind == 4) {
    acc = (double)a.V;
  } else {  // SA(c,s) = sum_k (R_k D_k + n_k T_k[:,1:])
#pragma unroll 4
    for (int k = 0; k < J; ++k) {
      const double nk = a.wsum[k];
      if (nk == 0.0) continue;
      const float* rt = a.RT + (size_t)(k * RW) * Bp + b;
      const double* D = a.wS + (size_t)k * 3 * NS;
      acc += (double)rt[(size_t)(c * 3) * Bp] * D[s] + (double)rt[(size_t)(c * 3 + 1) * Bp] * D[NS + s] +
             (double)rt[(size_t)(c * 3 + 2) * Bp] * D[2 * NS + s] + nk * (double)rt[(size_t)(9 + c * (1 + NS) + 1 + s) * Bp];
    }
  }
  if (a.tjT != nullptr) {  // joint block (pt/bodyfitter.py:1050-1058, _gram_block :1598)
    // per joint the three-coordinate dot product is formed in fp32 (the reference forms the whole block in an fp32
    // GEMM and casts to f64 afterwards, :1034-1062); the sum over joints is accumulated in double.  Conversions and
    // FP64 arithmetic are what bounds this kernel: one F2F + one DFMA per joint instead of six + four.
    if (kind == 4) {
      for (int j = 0; j < J; ++j) acc += a.jwT ? (double)SF_IM(a.jwT, j, Bp, b) : 1.0;
    } else if (kind == 0) {
#pragma unroll 4
      for (int j = 0; j < J; ++j) {
        const float* prow = a.Pext + (size_t)(j * TW) * Bp + b;
        float d = 0.f;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          d = fmaf(prow[(size_t)(cc * (1 + NS) + 1 + s) * Bp], prow[(size_t)(cc * (1 + NS) + 1 + t) * Bp], d);
        if (a.jwT) d *= SF_IM(a.jwT, j, Bp, b);
        acc += (double)d;
      }
    } else if (kind == 1) {
#pragma unroll 4
      for (int j = 0; j < J; ++j) {
        const float* prow = a.Pext + (size_t)(j * TW) * Bp + b;
        float d = 0.f;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          d = fmaf(prow[(size_t)(cc * (1 + NS) + 1 + s) * Bp], SF_IM(a.tjT, j * 3 + cc, Bp, b) - prow[(size_t)(cc * (1 + NS)) * Bp], d);
        if (a.jwT) d *= SF_IM(a.jwT, j, Bp, b);
        acc += (double)d;
      }
    } else {
#pragma unroll 4
      for (int j = 0; j < J; ++j) {
        const float* prow = a.Pext + (size_t)(j * TW + c * (1 + NS)) * Bp + b;
        float d = (kind == 2) ? SF_IM(a.tjT, j * 3 + c, Bp, b) - prow[0] : prow[(size_t)(1 + s) * Bp];
        if (a.jwT) d *= SF_IM(a.jwT, j, Bp, b);
        acc += (double)d;
      }
    }
  }
  Gd[(size_t)e * Bp + b] = acc;
}

// k_gram_entries<NS>: one single-warp CTA per (32 instances, entry): blockIdx.y = entry.  (Dealing the entries to the
// warps of one CTA per group, with the solve and the output rows behind CTA barriers, measured slower: 154 us vs
// 66 + 39 + 15 us for the three separate kernels; the entries are latency-bound and want to be spread over all SMs.)
template <int NS>
__global__ void __launch_bounds__(32) k_gram_entries(const SolveArgs a, double* __restrict__ Gd) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  if (b >= a.Bp) return;
  // the r and Sb entries sum one partial per segment (the longest chains of dependent round trips): dispatched first
  constexpr int NG = NS * (NS + 1) / 2, NL = NS + 3;
  const int y = (int)blockIdx.y;
  const int e = (y < NL) ? NG + y : ((y < NL + NG) ? y - NL : y);
  gram_entry<NS>(a, Gd, e, b);
}

// k_shape_solve<NS>: one thread per instance: centre with the covariance identity, regularise,
// Cholesky-solve in double, recover the translation (pt/bodyfitter.py:1060-1089; the general
// solve :1199-1283 is algebraically the same system).
template <int NS>
__device__ __forceinline__ void shape_solve_body(const SolveArgs& a, const double* Gd, int b) {
  constexpr int NG = NS * (NS + 1) / 2;
  const int Bp = a.Bp;
  double G[NS][NS], r[NS], SA[3][NS], Sb[3];
  {
    int o = 0;
    for (int s = 0; s < NS; ++s)
      for (int t = s; t < NS; ++t) {
        const double v = Gd[(size_t)o * Bp + b];
        G[s][t] = v;
        G[t][s] = v;
        ++o;
      }
    for (int s = 0; s < NS; ++s) r[s] = Gd[(size_t)(o++) * Bp + b];
    for (int c = 0; c < 3; ++c) Sb[c] = Gd[(size_t)(o++) * Bp + b];
    for (int c = 0; c < 3; ++c)
      for (int s = 0; s < NS; ++s) SA[c][s] = Gd[(size_t)(o++) * Bp + b];
  }
  const double W = Gd[(size_t)(NG + NS + 3 + 3 * NS) * Bp + b];
  const double Ws = (W == 0.0) ? 1.0 : W;
  double rhs[NS];
  for (int s = 0; s < NS; ++s) {
    double rc = r[s];
    for (int c = 0; c < 3; ++c) rc -= SA[c][s] * Sb[c] / Ws;
    for (int t = 0; t < NS; ++t) {
      double g = G[s][t];
      for (int c = 0; c < 3; ++c) g -= SA[c][s] * SA[c][t] / Ws;
      G[s][t] = g;
    }
    double lam = (s < 2) ? (double)a.reg2 : (double)a.reg;
    double ref = 0.0;
    if (s < a.S) {
      if (a.beta_ref != nullptr && b < a.B) ref = (double)a.beta_ref[(size_t)b * a.S + s];
    } else {  // kid unknown
      lam = (double)a.kid_reg;
      if (a.kid_ref != nullptr && b < a.B) ref = (double)a.kid_ref[b];
    }
    G[s][s] += lam;
    rhs[s] = rc + lam * ref;
  }
  chol_solve<NS>(G, rhs, NS);
  for (int s = 0; s < NS; ++s) SF_IM(a.beta, s, Bp, b) = (float)rhs[s];
  for (int c = 0; c < 3; ++c) {
    double m = Sb[c] / Ws;
    for (int s = 0; s < NS; ++s) m -= SA[c][s] / Ws * rhs[s];
    SF_IM(a.trans, c, Bp, b) = (float)m;
  }
}

// k_shape_solve_par<NS>: the same solve with the rows of the normal matrix dealt to the warps of a CTA (lane = instance,
// warp = row, the matrix in shared memory as [row][col][32 lanes] doubles): centring and regularisation per row, a
// right-looking Cholesky with three CTA barriers per column, column-oriented forward / backward substitution.  The
// dependent chain shrinks from ~NS^3/3 to ~NS^2 FP64 operations per instance and NS times more warps are in flight
// (the thread-per-instance kernel ran one warp per SM: 39 us for 4096 instances at NS = 10).
template <int NS>
__global__ void __launch_bounds__(NS * 32) k_shape_solve_par(const SolveArgs a, const double* __restrict__ Gd) {
  extern __shared__ __align__(16) double s_g[];  // G [NS][NS][32] | rhs [NS][32] | SA [3][NS][32] | Sb [3][32]
  constexpr int NG = NS * (NS + 1) / 2;
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const int Bp = a.Bp;
  double* G = s_g;
  double* rhs = G + NS * NS * 32;
  double* SA = rhs + NS * 32;
  double* Sb = SA + 3 * NS * 32;
#define SG(i, j) G[((i) * NS + (j)) * 32 + lane]
  // stage SA, Sb (every row needs them)
  for (int q = r; q < 3 * NS + 3; q += NS) {
    if (q < 3 * NS) SA[q * 32 + lane] = Gd[(size_t)(NG + NS + 3 + q) * Bp + b];
    else Sb[(q - 3 * NS) * 32 + lane] = Gd[(size_t)(NG + NS + (q - 3 * NS)) * Bp + b];
  }
  const double W = Gd[(size_t)(NG + NS + 3 + 3 * NS) * Bp + b];
  const double Ws = (W == 0.0) ? 1.0 : W;
  __syncthreads();
  {
    // row r: centre with the covariance identity, regularise (pt/bodyfitter.py:1060-1081)
    const double sa0 = SA[(0 * NS + r) * 32 + lane], sa1 = SA[(1 * NS + r) * 32 + lane], sa2 = SA[(2 * NS + r) * 32 + lane];
    for (int t = 0; t <= r; ++t) {  // lower triangle: entry (t, r) of the upper-triangle list, t <= r
      const int o = t * NS - t * (t - 1) / 2 + (r - t);
      double g = Gd[(size_t)o * Bp + b];
      g -= (sa0 * SA[(0 * NS + t) * 32 + lane] + sa1 * SA[(1 * NS + t) * 32 + lane] + sa2 * SA[(2 * NS + t) * 32 + lane]) / Ws;
      SG(r, t) = g;
    }
    double rc = Gd[(size_t)(NG + r) * Bp + b] - (sa0 * Sb[lane] + sa1 * Sb[32 + lane] + sa2 * Sb[64 + lane]) / Ws;
    double lam = (r < 2) ? (double)a.reg2 : (double)a.reg;
    double ref = 0.0;
    if (r < a.S) {
      if (a.beta_ref != nullptr && b < a.B) ref = (double)a.beta_ref[(size_t)b * a.S + r];
    } else {  // kid unknown
      lam = (double)a.kid_reg;
      if (a.kid_ref != nullptr && b < a.B) ref = (double)a.kid_ref[b];
    }
    SG(r, r) += lam;
    rhs[r * 32 + lane] = rc + lam * ref;
  }
  // Cholesky, lower triangle in place (no pivoting; failures surface as NaN like torch.linalg.cholesky_ex)
  for (int j = 0; j < NS; ++j) {
    __syncthreads();
    if (r == j) SG(j, j) = sqrt(SG(j, j));
    __syncthreads();
    if (r > j) SG(r, j) = SG(r, j) / SG(j, j);
    __syncthreads();
    if (r > j) {
      const double lrj = SG(r, j);
      for (int k = j + 1; k <= r; ++k) SG(r, k) -= lrj * SG(k, j);
    }
  }
  // forward substitution L y = rhs (column oriented: once y_i is final every later row subtracts its share)
  for (int i = 0; i < NS; ++i) {
    __syncthreads();
    if (r == i) rhs[i * 32 + lane] = rhs[i * 32 + lane] / SG(i, i);
    __syncthreads();
    if (r > i) rhs[r * 32 + lane] -= SG(r, i) * rhs[i * 32 + lane];
  }
  // backward substitution L^T x = y
  for (int i = NS - 1; i >= 0; --i) {
    __syncthreads();
    if (r == i) rhs[i * 32 + lane] = rhs[i * 32 + lane] / SG(i, i);
    __syncthreads();
    if (r < i) rhs[r * 32 + lane] -= SG(i, r) * rhs[i * 32 + lane];
  }
  __syncthreads();
  SF_IM(a.beta, r, Bp, b) = (float)rhs[r * 32 + lane];
  if (r < 3) {
    double m = Sb[r * 32 + lane] / Ws;
    for (int s = 0; s < NS; ++s) m -= SA[(r * NS + s) * 32 + lane] / Ws * rhs[s * 32 + lane];
    SF_IM(a.trans, r, Bp, b) = (float)m;
  }
#undef SG
}
inline size_t shape_solve_par_smem(int ns) { return (size_t)(ns * ns + ns + 3 * ns + 3) * 32 * sizeof(double); }

template <int NS>
__global__ void __launch_bounds__(32) k_shape_solve(const SolveArgs a, const double* __restrict__ Gd) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Bp) return;
  shape_solve_body<NS>(a, Gd, b);
}

// k_scale_entries<NS>: the NS+5 extra normal-equation entries of the scale column
// [Gz(NS) | Gzz | rz | SAz(3)]: chunk partials of k_scale_pass + the joint block, in double.
template <int NS>
__global__ void __launch_bounds__(32) k_scale_entries(const SolveArgs a, double* __restrict__ Zd) {
  constexpr int NZ = NS + 5;
  constexpr int TW = 3 * (1 + NS);
  const int b = blockIdx.x * 32 + threadIdx.x;
  const int e = blockIdx.y;
  if (b >= a.Bp) return;
  const int Bp = a.Bp, J = a.J;
  double acc = 0.0;
  for (int q = 0; q < a.n_zchunks; ++q) acc += (double)a.zpartials[((size_t)q * NZ + e) * Bp + b];
  if (a.tjT != nullptr) {
    for (int j = 0; j < J; ++j) {
      const double w = a.jwT ? (double)SF_IM(a.jwT, j, Bp, b) : 1.0;
      for (int c = 0; c < 3; ++c) {
        const float* prow = a.Pext + (size_t)(j * TW + c * (1 + NS)) * Bp + b;
        const double tj = (double)SF_IM(a.tjT, j * 3 + c, Bp, b), pj = (double)prow[0];
        const double z = (a.scale_mode == 1) ? -tj : pj;
        if (e < NS) acc += w * z * (double)prow[(size_t)(1 + e) * Bp];
        else if (e == NS) acc += w * z * z;
        else if (e == NS + 1) acc += w * z * (tj - pj);
        else if (e - NS - 2 == c) acc += w * z;
      }
    }
  }
  Zd[(size_t)e * Bp + b] = acc;
}

// k_shape_solve_scale<NS>: the (NS+1)-unknown variant of k_shape_solve (betas [+ kid] + scale
// delta); outputs the *undivided* unknowns like the reference's result dict plus scale_corr = 1 + delta.
template <int NS>
__global__ void __launch_bounds__(32) k_shape_solve_scale(const SolveArgs a, const double* __restrict__ Gd,
                                                          const double* __restrict__ Zd) {
  constexpr int NG = NS * (NS + 1) / 2;
  constexpr int N1 = NS + 1;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Bp) return;
  const int Bp = a.Bp;
  double G[N1][N1], r[N1], SA[3][N1], Sb[3];
  {
    int o = 0;
    for (int s = 0; s < NS; ++s)
      for (int t = s; t < NS; ++t) {
        const double v = Gd[(size_t)o * Bp + b];
        G[s][t] = v;
        G[t][s] = v;
        ++o;
      }
    for (int s = 0; s < NS; ++s) r[s] = Gd[(size_t)(o++) * Bp + b];
    for (int c = 0; c < 3; ++c) Sb[c] = Gd[(size_t)(o++) * Bp + b];
    for (int c = 0; c < 3; ++c)
      for (int s = 0; s < NS; ++s) SA[c][s] = Gd[(size_t)(o++) * Bp + b];
  }
  for (int s = 0; s < NS; ++s) {
    G[s][NS] = Zd[(size_t)s * Bp + b];
    G[NS][s] = G[s][NS];
  }
  G[NS][NS] = Zd[(size_t)NS * Bp + b];
  r[NS] = Zd[(size_t)(NS + 1) * Bp + b];
  for (int c = 0; c < 3; ++c) SA[c][NS] = Zd[(size_t)(NS + 2 + c) * Bp + b];
  const double W = Gd[(size_t)(NG + NS + 3 + 3 * NS) * Bp + b];
  const double Ws = (W == 0.0) ? 1.0 : W;
  double rhs[N1];
  for (int s = 0; s < N1; ++s) {
    double rc = r[s];
    for (int c = 0; c < 3; ++c) rc -= SA[c][s] * Sb[c] / Ws;
    for (int t = 0; t < N1; ++t) {
      double g = G[s][t];
      for (int c = 0; c < 3; ++c) g -= SA[c][s] * SA[c][t] / Ws;
      G[s][t] = g;
    }
    double lam = (s < 2) ? (double)a.reg2 : (double)a.reg;
    double ref = 0.0;
    if (s < a.S) {
      if (a.beta_ref != nullptr && b < a.B) ref = (double)a.beta_ref[(size_t)b * a.S + s];
    } else if (s < NS) {
      lam = (double)a.kid_reg;
      if (a.kid_ref != nullptr && b < a.B) ref = (double)a.kid_ref[b];
    } else {
      lam = (double)a.scale_reg;
    }
    G[s][s] += lam;
    rhs[s] = rc + lam * ref;
  }
  chol_solve<N1>(G, rhs, N1);
  for (int s = 0; s < NS; ++s) SF_IM(a.beta, s, Bp, b) = (float)rhs[s];
  a.scale_out[b] = (float)rhs[NS] + 1.f;
  for (int c = 0; c < 3; ++c) {
    double m = Sb[c] / Ws;
    for (int s = 0; s < N1; ++s) m -= SA[c][s] / Ws * rhs[s];
    SF_IM(a.trans, c, Bp, b) = (float)m;
  }
}

// k_shape_out: one thread per (instance, joint): reference joint (pt/bodyfitter.py:1093-1098)
// and the skinning transform [R | T0 + T1 x + trans] the statistics pass consumes.
__device__ __forceinline__ void shape_out_joint(const SolveArgs& a, int NS, int j, int b) {
  const int Bp = a.Bp;
  const int TW = 3 * (1 + NS), RW = 12 + 3 * NS;
  float x[SMPLFIT_MAX_UNKNOWNS];
  const float inv_sc = (a.scale_mode == 2) ? 1.f / a.scale_out[b] : 1.f;  // pt/bodyfitter.py:1293-1304
  for (int s = 0; s < NS; ++s) x[s] = SF_IM(a.beta, s, Bp, b) * inv_sc;
#pragma unroll
  for (int e = 0; e < 9; ++e) SF_IM(a.skin, j * 12 + e, Bp, b) = SF_IM(a.RT, j * RW + e, Bp, b);
  for (int c = 0; c < 3; ++c) {
    const float tr = SF_IM(a.trans, c, Bp, b);
    const float* prow = a.Pext + (size_t)(j * TW + c * (1 + NS)) * Bp + b;
    const float* trow = a.RT + (size_t)(j * RW + 9 + c * (1 + NS)) * Bp + b;
    float pj = 0.f, tj = 0.f;
    for (int s = 0; s < NS; ++s) {
      pj = fmaf(prow[(size_t)(1 + s) * Bp], x[s], pj);
      tj = fmaf(trow[(size_t)(1 + s) * Bp], x[s], tj);
    }
    SF_IM(a.refj, j * 3 + c, Bp, b) = prow[0] + pj + tr;
    SF_IM(a.skin, j * 12 + 9 + c, Bp, b) = trow[0] + tj + tr;
    if (a.skin4 != nullptr) {
      float4 q;
      q.x = SF_IM(a.RT, j * RW + c * 3, Bp, b);
      q.y = SF_IM(a.RT, j * RW + c * 3 + 1, Bp, b);
      q.z = SF_IM(a.RT, j * RW + c * 3 + 2, Bp, b);
      q.w = trow[0] + tj + tr;
      reinterpret_cast<float4*>(a.skin4)[(size_t)(j * 3 + c) * Bp + b] = q;
    }
  }
}

static __global__ void __launch_bounds__(32) k_shape_out(const SolveArgs a, int NS) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y;
  if (b >= a.Bp) return;
  shape_out_joint(a, NS, j, b);
  if (j == 0 && a.fq_hi != nullptr) {  // the unknowns as GEMM features of the statistics pass (x = ... + S beta)
    for (int s = 0; s < NS; ++s) {
      const float v = (b < a.B) ? SF_IM(a.beta, s, a.Bp, b) : 0.f;
      const __half hv = __float2half_rn(v);
      a.fq_hi[(size_t)b * a.fq_kf + a.fq_p + s] = hv;
      a.fq_lo[(size_t)b * a.fq_kf + a.fq_p + s] = __float2half_rn(v - __half2float(hv));
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_adjust_solve: _fit_global_rotations_dependent (pt/bodyfitter.py:1418-1595, sequential
// form): walk the tree, re-anchor each adjustable part at its recomputed joint position.
// ---------------------------------------------------------------------------------------
struct AdjustArgs {
  const float* partials;  // stats of (target, final reference) about (ct0 = tj_i, ca0 = refj_i)
  const float* tjT;       // [3J][Bp]
  const float* ajT;       // [3J][Bp] reference joints of the joint term (given-joints: == refj)
  const float* refj;      // [3J][Bp] true reference joints (c_a)
  const float* jwT;       // [J][Bp] or null
  const float* R_prev;    // [9J][Bp]
  const float* beta;      // [NS][Bp]
  const float* trans;     // [3][Bp]
  float* R_out;           // [9J][Bp]
  const float* scale;     // [Bp] scale_corr or null
  int scale_mode;         // 1: targets scaled (pt/bodyfitter.py:465-480), 2: reference scaled about trans (:481-496),
                          // 3: reference' = scale * reference + trans (fit_with_known_shape, :774-803)
  TreeTables t;
  int Bp;
  int n_adj;              // upper bound of the adjustable parts (shared-memory sizing of k_adjust_par)
};

static __global__ void __launch_bounds__(32) k_adjust_solve(const AdjustArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Bp) return;
  const int J = a.t.J, NS = a.t.NS, Bp = a.Bp;
  const int TW = 3 * (1 + NS);
  float x[SMPLFIT_MAX_UNKNOWNS];
  for (int s = 0; s < NS; ++s) x[s] = SF_IM(a.beta, s, Bp, b);
  const float sc = (a.scale != nullptr) ? a.scale[b] : 1.f;
  const float st_t = (a.scale_mode == 1) ? sc : 1.f;  // scale of the target side
  const float st_a = (a.scale_mode >= 2) ? sc : 1.f;  // scale of the reference side
  const float tr_a = (a.scale_mode == 3) ? 1.f : (1.f - st_a);  // reference' = st_a * reference + tr_a * trans
  float trv[3];
  for (int c = 0; c < 3; ++c) trv[c] = SF_IM(a.trans, c, Bp, b);
  float pos[SMPLFIT_MAX_JOINTS * 3], rest[SMPLFIT_MAX_JOINTS * 3];
  for (int j = 0; j < J; ++j)
    for (int c = 0; c < 3; ++c) {
      const float* Jt = a.t.Jt_ext + (size_t)j * TW + c * (1 + NS);
      float v = __ldg(Jt);
      for (int s = 0; s < NS; ++s) v = fmaf(__ldg(Jt + 1 + s), x[s], v);
      rest[j * 3 + c] = (a.scale_mode >= 2) ? v * sc : v;  // j = j * scale_corr (pt/bodyfitter.py:1449-1450)
    }
  for (int i = 0; i < J; ++i) {
    const int par = a.t.parents[i];
    if (i == 0) {
      for (int c = 0; c < 3; ++c) pos[c] = rest[c] + SF_IM(a.trans, c, Bp, b);
    } else {
      float Rp[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) Rp[e] = SF_IM(a.R_out, par * 9 + e, Bp, b);
      const float bone[3] = {rest[i * 3] - rest[par * 3], rest[i * 3 + 1] - rest[par * 3 + 1], rest[i * 3 + 2] - rest[par * 3 + 2]};
      float rb[3];
      mat3_vec(Rp, bone, rb);
      for (int c = 0; c < 3; ++c) pos[i * 3 + c] = pos[par * 3 + c] + rb[c];
    }
    const int kind = a.t.part_kind[i];
    float Rn[9];
    if (kind == 4) {  // toes copy the (already adjusted) feet
      const int src = a.t.part_copy_src[i];
#pragma unroll
      for (int e = 0; e < 9; ++e) Rn[e] = SF_IM(a.R_out, src * 9 + e, Bp, b);
    } else if ((a.t.part_flags[i] & 2) == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) Rn[e] = SF_IM(a.R_prev, i * 9 + e, Bp, b);
    } else {
      // the statistics were accumulated on the unscaled data about (ct0, ca0) = (tj_i, refj_i); they
      // are bilinear, so scaling the targets (mode 1) or the reference about trans (mode 2) just
      // rescales them: M' = sc M, st' = st_t st, sa' = st_a sa, centres scale alike.
      float ct0[3], ca[3], A[9];
      load3(a.tjT, i, Bp, b, ct0);
      load3(a.refj, i, Bp, b, ca);
      for (int c = 0; c < 3; ++c) {
        ct0[c] *= st_t;
        ca[c] = st_a * ca[c] + tr_a * trv[c];
      }
      part_covariance(a.partials, a.t.part_seg_begin, i, Bp, b, ct0, ca, pos + i * 3, ca, A, st_t * st_a, st_t, st_a);
      const int n = a.t.cas_count[i];
      const int32_t* cas = a.t.cas_table + i * a.t.max_cas;
      for (int k = 0; k < n; ++k) {
        float tj[3], aj[3];
        load3(a.tjT, cas[k], Bp, b, tj);
        load3(a.ajT, cas[k], Bp, b, aj);
        for (int c = 0; c < 3; ++c) {
          tj[c] *= st_t;
          aj[c] = st_a * aj[c] + tr_a * trv[c];
        }
        const float w = a.jwT ? SF_IM(a.jwT, cas[k], Bp, b) : 1.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            A[r * 3 + c] = fmaf(tj[r] - pos[i * 3 + r], w * (aj[c] - ca[c]), A[r * 3 + c]);
      }
      float Rf[9], Ro[9];
      proj_so3(A, Rf);
#pragma unroll
      for (int e = 0; e < 9; ++e) Ro[e] = SF_IM(a.R_prev, i * 9 + e, Bp, b);
      mat3_mul(Rf, Ro, Rn);
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) SF_IM(a.R_out, i * 9 + e, Bp, b) = Rn[e];
  }
}

// k_adjust_par: the same walk, level-synchronous.  One CTA (ADJ_WARPS warps) per 32 instances, lane = instance;
// joint positions, rest joints and the output orientations live in shared memory; the joints of one tree depth are
// independent given their parents, so each depth is processed in parallel (warp per joint) between two CTA barriers.
// The dependent chain shrinks from J joints / all adjustable parts to the tree depth / adjustable parts per branch.
constexpr int ADJ_WARPS = 8;
__host__ __device__ inline size_t adjust_par_smem_bytes(int J, int n_adj) {
  return (size_t)J * 15 * 32 * sizeof(float) + (size_t)n_adj * 16 * 32 * sizeof(float) + (size_t)(2 * J + 2) * sizeof(int);
}

static __global__ void __launch_bounds__(ADJ_WARPS * 32) k_adjust_par(const AdjustArgs a) {
  extern __shared__ __align__(16) float s_adj[];
  const int J = a.t.J, NS = a.t.NS, Bp = a.Bp;
  const int TW = 3 * (1 + NS);
  float* pos = s_adj;             // [J*3][32]
  float* rest = pos + J * 96;     // [J*3][32]
  float* Ro = rest + J * 96;      // [J*9][32]
  float* ssum = Ro + J * 288;     // [n_adj][16][32] statistics of the adjustable parts (summed once, off the level loop)
  int* depth = reinterpret_cast<int*>(ssum + (size_t)a.n_adj * 512);  // [J], then max depth
  int* slot = depth + J + 1;      // [J] index of an adjustable part into ssum, else -1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  if (threadIdx.x < J) {
    int d = 0;
    for (int p = threadIdx.x; p > 0; p = a.t.parents[p]) ++d;
    depth[threadIdx.x] = d;
  }
  if (threadIdx.x == 32) {  // (another warp than the one that scans the depths below)
    int n = 0;
    for (int j = 0; j < J; ++j) {
      const bool adj = a.t.part_kind[j] != 4 && (a.t.part_flags[j] & 2) != 0;
      slot[j] = (adj && n < a.n_adj) ? n++ : -1;
    }
  }
  __syncthreads();
  // the segment partials of every adjustable part, all warps at once: the level loop below then has no dependent
  // DRAM / L2 round trips left on its critical path (they were ~2/3 of this kernel's 147 us)
  for (int i = warp; i < J; i += ADJ_WARPS) {
    if (slot[i] < 0) continue;
    double acc[16];
    part_sums(a.partials, a.t.part_seg_begin, i, Bp, b, acc);
#pragma unroll
    for (int e = 0; e < 16; ++e) ssum[((size_t)slot[i] * 16 + e) * 32 + lane] = (float)acc[e];
  }
  float x[SMPLFIT_MAX_UNKNOWNS];
  for (int s = 0; s < NS; ++s) x[s] = SF_IM(a.beta, s, Bp, b);
  const float sc = (a.scale != nullptr) ? a.scale[b] : 1.f;
  const float st_t = (a.scale_mode == 1) ? sc : 1.f;  // scale of the target side
  const float st_a = (a.scale_mode >= 2) ? sc : 1.f;  // scale of the reference side
  const float tr_a = (a.scale_mode == 3) ? 1.f : (1.f - st_a);  // reference' = st_a * reference + tr_a * trans
  float trv[3];
  for (int c = 0; c < 3; ++c) trv[c] = SF_IM(a.trans, c, Bp, b);
  // phase 0: rest joints (pt/bodyfitter.py:1441-1450) and the orientations that are simply kept
  for (int i = warp; i < J; i += ADJ_WARPS) {
    for (int c = 0; c < 3; ++c) {
      const float* Jt = a.t.Jt_ext + (size_t)i * TW + c * (1 + NS);
      float v = __ldg(Jt);
      for (int s = 0; s < NS; ++s) v = fmaf(__ldg(Jt + 1 + s), x[s], v);
      rest[(i * 3 + c) * 32 + lane] = (a.scale_mode >= 2) ? v * sc : v;
    }
    if (a.t.part_kind[i] != 4 && (a.t.part_flags[i] & 2) == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) Ro[(i * 9 + e) * 32 + lane] = SF_IM(a.R_prev, i * 9 + e, Bp, b);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int md = 0;
    for (int j = 0; j < J; ++j) md = max(md, depth[j]);
    depth[J] = md;
  }
  __syncthreads();
  const int max_depth = depth[J];
  for (int d = 0; d <= max_depth; ++d) {
    for (int i = warp; i < J; i += ADJ_WARPS) {
      if (depth[i] != d) continue;
      float pi[3];
      if (i == 0) {
        for (int c = 0; c < 3; ++c) pi[c] = rest[c * 32 + lane] + trv[c];
      } else {
        const int par = a.t.parents[i];
        float Rp[9], bone[3];
#pragma unroll
        for (int e = 0; e < 9; ++e) Rp[e] = Ro[(par * 9 + e) * 32 + lane];
        for (int c = 0; c < 3; ++c) bone[c] = rest[(i * 3 + c) * 32 + lane] - rest[(par * 3 + c) * 32 + lane];
        float rb[3];
        mat3_vec(Rp, bone, rb);
        for (int c = 0; c < 3; ++c) pi[c] = pos[(par * 3 + c) * 32 + lane] + rb[c];
      }
      for (int c = 0; c < 3; ++c) pos[(i * 3 + c) * 32 + lane] = pi[c];
      const int kind = a.t.part_kind[i];
      if (kind == 4) {  // toes copy the (already adjusted) feet: the source is an ancestor (pt/bodyfitter.py:1414-1416)
        const int src = a.t.part_copy_src[i];
#pragma unroll
        for (int e = 0; e < 9; ++e) Ro[(i * 9 + e) * 32 + lane] = Ro[(src * 9 + e) * 32 + lane];
      } else if ((a.t.part_flags[i] & 2) != 0) {
        float ct0[3], ca[3], A[9];
        load3(a.tjT, i, Bp, b, ct0);
        load3(a.refj, i, Bp, b, ca);
        for (int c = 0; c < 3; ++c) {
          ct0[c] *= st_t;
          ca[c] = st_a * ca[c] + tr_a * trv[c];
        }
        if (slot[i] >= 0) {
          float sm[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) sm[e] = ssum[((size_t)slot[i] * 16 + e) * 32 + lane];
          covariance_from_sums(sm, ct0, ca, pi, ca, A, st_t * st_a, st_t, st_a);
        } else {
          part_covariance(a.partials, a.t.part_seg_begin, i, Bp, b, ct0, ca, pi, ca, A, st_t * st_a, st_t, st_a);
        }
        const int n = a.t.cas_count[i];
        const int32_t* cas = a.t.cas_table + i * a.t.max_cas;
        for (int k = 0; k < n; ++k) {
          float tj[3], aj[3];
          load3(a.tjT, cas[k], Bp, b, tj);
          load3(a.ajT, cas[k], Bp, b, aj);
          for (int c = 0; c < 3; ++c) {
            tj[c] *= st_t;
            aj[c] = st_a * aj[c] + tr_a * trv[c];
          }
          const float w = a.jwT ? SF_IM(a.jwT, cas[k], Bp, b) : 1.f;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) A[r * 3 + c] = fmaf(tj[r] - pi[r], w * (aj[c] - ca[c]), A[r * 3 + c]);
        }
        float Rf[9], Rold[9], Rn[9];
        proj_so3(A, Rf);
#pragma unroll
        for (int e = 0; e < 9; ++e) Rold[e] = SF_IM(a.R_prev, i * 9 + e, Bp, b);
        mat3_mul(Rf, Rold, Rn);
#pragma unroll
        for (int e = 0; e < 9; ++e) Ro[(i * 9 + e) * 32 + lane] = Rn[e];
      }
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < J * 9 * 32; q += ADJ_WARPS * 32) a.R_out[(size_t)(q >> 5) * Bp + blockIdx.x * 32 + (q & 31)] = Ro[q];
}

// ---------------------------------------------------------------------------------------
// k_output: caller-facing results in the reference layouts (pt/bodyfitter.py:513-549):
// trans + target mean, orientations, relative orientations, pose rotation vectors.
// ---------------------------------------------------------------------------------------
struct OutputArgs {
  const float* R_final;  // [9J][Bp]
  const float* R_rel_src;  // orientations the relative rotations are derived from
  const float* beta;     // [NS][Bp]
  const float* trans;    // [3][Bp]
  const float* mean;     // [3][Bp]
  const int32_t* parents;
  float* pose_rotvecs;   // (B,3J) or null
  float* shape_betas;    // (B,S)
  float* out_trans;      // (B,3)
  float* orientations;   // (B,J,3,3)
  float* rel_orient;     // (B,J,3,3) or null
  float* kid;            // (B) or null
  const float* scale;    // [Bp] or null
  float* scale_corr;     // (B) or null
  int scale_mode;
  int J, S, NS, B, Bp;
};

// one thread per (instance, joint); the joint-0 thread also writes the per-instance scalars
static __global__ void __launch_bounds__(32) k_output(const OutputArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (b >= a.B) return;
  const int J = a.J, Bp = a.Bp;
  if (j == 0) {
    if (a.shape_betas != nullptr)
      for (int s = 0; s < a.S; ++s) a.shape_betas[(size_t)b * a.S + s] = SF_IM(a.beta, s, Bp, b);
    if (a.kid != nullptr) a.kid[b] = SF_IM(a.beta, a.S, Bp, b);
    // pt/bodyfitter.py:513-519: the target mean is added back (scaled in the scale modes)
    const float sc = (a.scale != nullptr) ? a.scale[b] : 1.f;
    const float f = (a.scale_mode == 1) ? sc : (a.scale_mode == 2 ? 1.f / sc : 1.f);
    for (int c = 0; c < 3; ++c) a.out_trans[(size_t)b * 3 + c] = SF_IM(a.trans, c, Bp, b) + SF_IM(a.mean, c, Bp, b) * f;
    if (a.scale_corr != nullptr) a.scale_corr[b] = sc;
  }
  if (a.orientations != nullptr) {
#pragma unroll
    for (int e = 0; e < 9; ++e) a.orientations[((size_t)b * J + j) * 9 + e] = SF_IM(a.R_final, j * 9 + e, Bp, b);
  }
  if (a.rel_orient != nullptr || a.pose_rotvecs != nullptr) {
    float Rs[9], rel[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rs[e] = SF_IM(a.R_rel_src, j * 9 + e, Bp, b);
    if (j == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) rel[e] = Rs[e];
    } else {
      float Rp[9];
      const int par = a.parents[j];
#pragma unroll
      for (int e = 0; e < 9; ++e) Rp[e] = SF_IM(a.R_rel_src, par * 9 + e, Bp, b);
      mat3_tmul(Rp, Rs, rel);
    }
    if (a.rel_orient != nullptr) {
#pragma unroll
      for (int e = 0; e < 9; ++e) a.rel_orient[((size_t)b * J + j) * 9 + e] = rel[e];
    }
    if (a.pose_rotvecs != nullptr) {
      float rv[3];
      mat2rotvec(rel, rv);
      for (int c = 0; c < 3; ++c) a.pose_rotvecs[(size_t)b * 3 * J + j * 3 + c] = rv[c];
    }
  }
}

// ---------------------------------------------------------------------------------------
// fit_scale_and_translation (pt/bodyfitter.py:1628-1681) for fit_with_known_shape: weighted
// first / second moments of targets and reference over vertices (+ joints), reduced per
// (vertex block, instance) then finalised per instance.
// moments layout: [W, St(3), Sa(3), Stt, Saa]
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128) k_moments(const float* __restrict__ tT, const float* __restrict__ aT,
                                                        const float* __restrict__ wT, int n_items, int items_per_warp,
                                                        int n_blocks, int Bp, float* __restrict__ partials) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int blk = warp % n_blocks, g = warp / n_blocks;
  if (g * 32 >= Bp) return;
  const int b = g * 32 + lane;
  float W = 0.f, St[3] = {0.f, 0.f, 0.f}, Sa[3] = {0.f, 0.f, 0.f}, Stt = 0.f, Saa = 0.f;
  const int i0 = blk * items_per_warp, i1 = min(n_items, i0 + items_per_warp);
  for (int i = i0; i < i1; ++i) {
    const float w = wT ? SF_IM(wT, i, Bp, b) : 1.f;
    W += w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = SF_IM(tT, i * 3 + c, Bp, b), x = SF_IM(aT, i * 3 + c, Bp, b);
      St[c] = fmaf(w, t, St[c]);
      Sa[c] = fmaf(w, x, Sa[c]);
      Stt = fmaf(w * t, t, Stt);
      Saa = fmaf(w * x, x, Saa);
    }
  }
  float* out = partials + (size_t)blk * 9 * Bp + b;
  out[0] = W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out[(size_t)(1 + c) * Bp] = St[c];
    out[(size_t)(4 + c) * Bp] = Sa[c];
  }
  out[(size_t)7 * Bp] = Stt;
  out[(size_t)8 * Bp] = Saa;
}

struct ScaleTransArgs {
  const float* vpart;  // [n_vblocks][9][Bp]
  const float* jpart;  // [1][9][Bp] or null (no joints)
  int n_vblocks, Bp, estimate_scale;
  float* scale;  // [Bp]
  float* trans;  // [3][Bp]
};

static __global__ void __launch_bounds__(32) k_scale_trans(const ScaleTransArgs a) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  if (b >= a.Bp) return;
  double m[9];
  for (int e = 0; e < 9; ++e) m[e] = 0.0;
  for (int q = 0; q < a.n_vblocks; ++q)
    for (int e = 0; e < 9; ++e) m[e] += (double)a.vpart[((size_t)q * 9 + e) * a.Bp + b];
  if (a.jpart != nullptr)
    for (int e = 0; e < 9; ++e) m[e] += (double)a.jpart[(size_t)e * a.Bp + b];
  const double W = m[0];
  double mt[3], ma[3], sst = m[7], ssa = m[8];
  for (int c = 0; c < 3; ++c) {
    mt[c] = m[1 + c] / W;
    ma[c] = m[4 + c] / W;
    sst -= W * mt[c] * mt[c];
    ssa -= W * ma[c] * ma[c];
  }
  const double sc = a.estimate_scale ? sqrt(sst / ssa) : 1.0;
  a.scale[b] = (float)sc;
  for (int c = 0; c < 3; ++c) SF_IM(a.trans, c, a.Bp, b) = (float)(mt[c] - sc * ma[c]);
}

// betas (B,n) / kid (B) given by the caller -> [NS][Bp] unknown vector (zero padded), zero trans
static __global__ void __launch_bounds__(32) k_set_shape(const float* __restrict__ betas, int n_betas,
                                                         const float* __restrict__ kid, int S, int NS, int B, int Bp,
                                                         float* __restrict__ beta, float* __restrict__ trans) {
  const int b = blockIdx.x * 32 + threadIdx.x;
  if (b >= Bp) return;
  const bool live = b < B;
  for (int s = 0; s < S; ++s) SF_IM(beta, s, Bp, b) = (live && betas && s < n_betas) ? betas[(size_t)b * n_betas + s] : 0.f;
  if (NS > S) SF_IM(beta, S, Bp, b) = (live && kid) ? kid[b] : 0.f;
  for (int c = 0; c < 3; ++c) SF_IM(trans, c, Bp, b) = 0.f;
}

// ---------------------------------------------------------------------------------------
// share_beta (pt/bodyfitter.py:1266-1274, pt/lstsq.py:24-26, :43-45): the centred, regularised
// normal equations of all instances are summed and solved once; the translation stays per
// instance.  Mirrors the reference's shared branch exactly: the regulariser is added per instance
// *before* the batch sum (i.e. B times) and the regulariser-reference term is not applied.
//   k_center_entries : per instance  Gc = G - SA^T SA / W,  rc = r - SA^T Sb / W   (upper + rhs)
//   k_batch_sum      : deterministic sum over the batch, one CTA per entry
//   k_shared_solve   : one thread: + B * diag(lambda), Cholesky
//   k_shared_apply   : per instance: beta = x, trans = (Sb - SA x) / W
// ---------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(32) k_center_entries(const SolveArgs a, const double* __restrict__ Gd,
                                                       double* __restrict__ Cd) {
  constexpr int NG = NS * (NS + 1) / 2;
  const int b = blockIdx.x * 32 + threadIdx.x;
  if (b >= a.Bp) return;
  const int Bp = a.Bp;
  const bool live = b < a.B;
  double SA[3][NS], Sb[3];
  for (int c = 0; c < 3; ++c) Sb[c] = Gd[(size_t)(NG + NS + c) * Bp + b];
  for (int c = 0; c < 3; ++c)
    for (int s = 0; s < NS; ++s) SA[c][s] = Gd[(size_t)(NG + NS + 3 + c * NS + s) * Bp + b];
  const double W = Gd[(size_t)(NG + NS + 3 + 3 * NS) * Bp + b];
  const double Ws = (W == 0.0) ? 1.0 : W;
  int o = 0;
  for (int s = 0; s < NS; ++s)
    for (int t = s; t < NS; ++t) {
      double g = Gd[(size_t)o * Bp + b];
      for (int c = 0; c < 3; ++c) g -= SA[c][s] * SA[c][t] / Ws;
      Cd[(size_t)o * Bp + b] = live ? g : 0.0;
      ++o;
    }
  for (int s = 0; s < NS; ++s) {
    double rc = Gd[(size_t)(NG + s) * Bp + b];
    for (int c = 0; c < 3; ++c) rc -= SA[c][s] * Sb[c] / Ws;
    Cd[(size_t)(NG + s) * Bp + b] = live ? rc : 0.0;
  }
}

static __global__ void __launch_bounds__(256) k_batch_sum(const double* __restrict__ Cd, int Bp, double* __restrict__ out) {
  __shared__ double red[256];
  const int e = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < Bp; b += 256) acc += Cd[(size_t)e * Bp + b];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[e] = red[0];
}

template <int NS>
__global__ void k_shared_solve(const SolveArgs a, const double* __restrict__ sums, double* __restrict__ x) {
  constexpr int NG = NS * (NS + 1) / 2;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double G[NS][NS], rhs[NS];
  int o = 0;
  for (int s = 0; s < NS; ++s)
    for (int t = s; t < NS; ++t) {
      G[s][t] = sums[o];
      G[t][s] = sums[o];
      ++o;
    }
  for (int s = 0; s < NS; ++s) {
    const double lam = (s >= a.S) ? (double)a.kid_reg : ((s < 2) ? (double)a.reg2 : (double)a.reg);
    if (!a.shared_noreg) G[s][s] += lam * (double)a.B;  // diag(lambda) added per instance, then summed (pt/lstsq.py:16-26)
    rhs[s] = sums[NG + s];
  }
  chol_solve<NS>(G, rhs, NS);
  for (int s = 0; s < NS; ++s) x[s] = rhs[s];
}

template <int NS>
__global__ void __launch_bounds__(32) k_shared_apply(const SolveArgs a, const double* __restrict__ Gd,
                                                     const double* __restrict__ x) {
  constexpr int NG = NS * (NS + 1) / 2;
  const int b = blockIdx.x * 32 + threadIdx.x;
  if (b >= a.Bp) return;
  const int Bp = a.Bp;
  const double W = Gd[(size_t)(NG + NS + 3 + 3 * NS) * Bp + b];
  const double Ws = (W == 0.0) ? 1.0 : W;
  for (int s = 0; s < NS; ++s) SF_IM(a.beta, s, Bp, b) = (float)x[s];
  for (int c = 0; c < 3; ++c) {
    double m = Gd[(size_t)(NG + NS + c) * Bp + b] / Ws;
    for (int s = 0; s < NS; ++s) m -= Gd[(size_t)(NG + NS + 3 + c * NS + s) * Bp + b] / Ws * x[s];
    SF_IM(a.trans, c, Bp, b) = (float)m;
  }
}

// ---------------------------------------------------------------------------------------
// share_beta together with scale estimation (pt/lstsq.py:32-90 lstsq_partial_share): the betas (+ kid) are shared over
// the batch, the scale column stays per instance.  The regulariser enters as extra rows (row e_i, weight lambda_i,
// right-hand side lambda_i ref_i -- so its reference term is lambda_i^2 ref_i on this path).  Per instance the scale
// column is eliminated from the centred, regularised normal equations (Schur complement):
//   c_s = G_sz / G_zz',  c_r = r_z / G_zz',  S_b = G_ss' - G_sz c_s^T,  t_b = r_s' - G_sz c_r        (k_center_entries_scale)
// S_b, t_b are summed over the batch (k_batch_sum, all-reduce hook), solved once (k_shared_solve, no extra regulariser),
// and every instance recovers its own scale unknown delta_b = c_r - c_s . x and translation (k_shared_apply_scale).
// Zd rows [0, NS] are overwritten with c_s, c_r.
// ---------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(32) k_center_entries_scale(const SolveArgs a, const double* __restrict__ Gd,
                                                             double* __restrict__ Zd, double* __restrict__ Cd) {
  constexpr int NG = NS * (NS + 1) / 2;
  constexpr int N1 = NS + 1;
  const int b = blockIdx.x * 32 + threadIdx.x;
  if (b >= a.Bp) return;
  const int Bp = a.Bp;
  const bool live = b < a.B;
  double G[N1][N1], r[N1], SA[3][N1], Sb[3];
  {
    int o = 0;
    for (int s = 0; s < NS; ++s)
      for (int t = s; t < NS; ++t) {
        const double v = Gd[(size_t)o * Bp + b];
        G[s][t] = v;
        G[t][s] = v;
        ++o;
      }
    for (int s = 0; s < NS; ++s) r[s] = Gd[(size_t)(o++) * Bp + b];
    for (int c = 0; c < 3; ++c) Sb[c] = Gd[(size_t)(o++) * Bp + b];
    for (int c = 0; c < 3; ++c)
      for (int s = 0; s < NS; ++s) SA[c][s] = Gd[(size_t)(o++) * Bp + b];
  }
  for (int s = 0; s < NS; ++s) {
    G[s][NS] = Zd[(size_t)s * Bp + b];
    G[NS][s] = G[s][NS];
  }
  G[NS][NS] = Zd[(size_t)NS * Bp + b];
  r[NS] = Zd[(size_t)(NS + 1) * Bp + b];
  for (int c = 0; c < 3; ++c) SA[c][NS] = Zd[(size_t)(NS + 2 + c) * Bp + b];
  const double W = Gd[(size_t)(NG + NS + 3 + 3 * NS) * Bp + b];
  const double Ws = (W == 0.0) ? 1.0 : W;
  double rc[N1], lam[N1], ref[N1];
  for (int s = 0; s < N1; ++s) {
    double v = r[s];
    for (int c = 0; c < 3; ++c) v -= SA[c][s] * Sb[c] / Ws;
    rc[s] = v;
    for (int t = 0; t < N1; ++t) {
      double g = G[s][t];
      for (int c = 0; c < 3; ++c) g -= SA[c][s] * SA[c][t] / Ws;
      G[s][t] = g;
    }
    lam[s] = (s < 2) ? (double)a.reg2 : (double)a.reg;
    ref[s] = 0.0;
    if (s < a.S) {
      if (a.beta_ref != nullptr && live) ref[s] = (double)a.beta_ref[(size_t)b * a.S + s];
    } else if (s < NS) {
      lam[s] = (double)a.kid_reg;
      if (a.kid_ref != nullptr && live) ref[s] = (double)a.kid_ref[b];
    } else {
      lam[s] = (double)a.scale_reg;
    }
  }
  const double gzz = G[NS][NS] + lam[NS];
  const double c_r = rc[NS] / gzz;
  double c_s[NS];
  for (int s = 0; s < NS; ++s) c_s[s] = G[s][NS] / gzz;
  int o = 0;
  for (int s = 0; s < NS; ++s)
    for (int t = s; t < NS; ++t) {
      const double v = G[s][t] + ((s == t) ? lam[s] : 0.0) - G[s][NS] * c_s[t];
      Cd[(size_t)o * Bp + b] = live ? v : 0.0;
      ++o;
    }
  for (int s = 0; s < NS; ++s) {
    const double v = rc[s] + lam[s] * lam[s] * ref[s] - G[s][NS] * c_r;
    Cd[(size_t)(NG + s) * Bp + b] = live ? v : 0.0;
  }
  for (int s = 0; s < NS; ++s) Zd[(size_t)s * Bp + b] = c_s[s];
  Zd[(size_t)NS * Bp + b] = c_r;
}

template <int NS>
__global__ void __launch_bounds__(32) k_shared_apply_scale(const SolveArgs a, const double* __restrict__ Gd,
                                                           const double* __restrict__ Zd, const double* __restrict__ x) {
  constexpr int NG = NS * (NS + 1) / 2;
  const int b = blockIdx.x * 32 + threadIdx.x;
  if (b >= a.Bp) return;
  const int Bp = a.Bp;
  const double W = Gd[(size_t)(NG + NS + 3 + 3 * NS) * Bp + b];
  const double Ws = (W == 0.0) ? 1.0 : W;
  double delta = Zd[(size_t)NS * Bp + b];
  for (int s = 0; s < NS; ++s) delta -= Zd[(size_t)s * Bp + b] * x[s];
  for (int s = 0; s < NS; ++s) SF_IM(a.beta, s, Bp, b) = (float)x[s];
  a.scale_out[b] = (float)delta + 1.f;
  for (int c = 0; c < 3; ++c) {
    double m = Gd[(size_t)(NG + NS + c) * Bp + b] / Ws;
    for (int s = 0; s < NS; ++s) m -= Gd[(size_t)(NG + NS + 3 + c * NS + s) * Bp + b] / Ws * x[s];
    m -= Zd[(size_t)(NS + 2 + c) * Bp + b] / Ws * delta;
    SF_IM(a.trans, c, Bp, b) = (float)m;
  }
}

}  // namespace sf
