// BodyModel.forward (pt/bodymodel.py:121-307) and BodyConverter.convert_vertices
// (pt/bodyconverter.py:129-149) for sm_100a.
//
// forward = (1) k_fwd_prep: one thread per instance -- Rodrigues / rotation chain, shaped rest
//               joints, joint positions (FK), per-joint skinning transforms, pose features;
//           (2) the pose-blend-shape contraction v_posed^T = v_template + posedirs . feat
//               (tensor-core path in vposed_tc.cu, FP32 SIMT fallback);
//           (3) k_fwd_skin: lane = instance; adds the shape/kid blend shapes, blends the
//               <= K skinning transforms per vertex and writes the caller's (B,V,3) layout
//               through a shared-memory transpose so both sides stay coalesced.
#include <stdlib.h>

#include "common.cuh"
#include "fit_kernels.cuh"
#include "lite_kernels.cuh"
#include "solve_kernels.cuh"
#include "vposed_tc.cuh"

namespace sf {

struct FwdPrepArgs {
  const float* rot;    // see rot_mode
  const float* betas;  // (B,n_betas) or null
  const float* trans;  // (B,3) or null
  const float* kid;    // (B) or null
  const int32_t* parents;
  const float* J_template;
  const float* J_shapedirs;     // (J,3,S)
  const float* kid_J_shapedir;  // (J,3)
  float* skin;   // [12J][Bp]
  float* skin4;  // optional [J*3][Bp] float4 (G[c][0..2], t[c]): the same rows as quads for k_fwd_skin_tma
  float* feat;   // [Bp][Kp]
  float* betaT;  // [S+1][Bp] (betas zero-padded to S, then kid)
  float* out_joints;        // (B,J,3)
  float* out_orientations;  // (B,J,3,3)
  int rot_mode, n_betas, J, S, B, Bp, Kp;
};

__global__ void __launch_bounds__(32) k_fwd_prep(const FwdPrepArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Bp) return;
  const int J = a.J, S = a.S, Bp = a.Bp;
  const bool live = b < a.B;
  float x[SMPLFIT_MAX_UNKNOWNS + 16];
  const int nb = min(a.n_betas, S);
  for (int s = 0; s < S; ++s) {
    x[s] = (live && a.betas != nullptr && s < nb) ? a.betas[(size_t)b * a.n_betas + s] : 0.f;
    SF_IM(a.betaT, s, Bp, b) = x[s];
  }
  const float kid = (live && a.kid != nullptr) ? a.kid[b] : 0.f;
  SF_IM(a.betaT, S, Bp, b) = kid;
  float tr[3] = {0.f, 0.f, 0.f};
  if (live && a.trans != nullptr)
    for (int c = 0; c < 3; ++c) tr[c] = a.trans[(size_t)b * 3 + c];
  float glob[SMPLFIT_MAX_JOINTS * 9], pos[SMPLFIT_MAX_JOINTS * 3], rest[SMPLFIT_MAX_JOINTS * 3];
  for (int j = 0; j < J; ++j)
    for (int c = 0; c < 3; ++c) {
      float v = __ldg(a.J_template + j * 3 + c);
      const float* js = a.J_shapedirs + ((size_t)j * 3 + c) * S;
      for (int s = 0; s < nb; ++s) v = fmaf(__ldg(js + s), x[s], v);
      v = fmaf(__ldg(a.kid_J_shapedir + j * 3 + c), kid, v);
      rest[j * 3 + c] = v;
    }
  for (int j = 0; j < J; ++j) {
    const int par = a.parents[j];
    float rel[9];
    float* G = glob + j * 9;
    if (a.rot_mode == 2) {
      for (int e = 0; e < 9; ++e) G[e] = live ? a.rot[((size_t)b * J + j) * 9 + e] : ((e % 4 == 0) ? 1.f : 0.f);
      if (j > 0) mat3_tmul(glob + par * 9, G, rel);
    } else {
      if (a.rot_mode == 0) {
        float rv[3] = {0.f, 0.f, 0.f};
        if (live)
          for (int c = 0; c < 3; ++c) rv[c] = a.rot[(size_t)b * 3 * J + j * 3 + c];
        rotvec2mat(rv, rel);
      } else if (a.rot_mode == 1) {
        for (int e = 0; e < 9; ++e) rel[e] = live ? a.rot[((size_t)b * J + j) * 9 + e] : ((e % 4 == 0) ? 1.f : 0.f);
      } else {
        for (int e = 0; e < 9; ++e) rel[e] = (e % 4 == 0) ? 1.f : 0.f;
      }
      if (j == 0) {
        for (int e = 0; e < 9; ++e) G[e] = rel[e];
      } else {
        mat3_mul(glob + par * 9, rel, G);
      }
    }
    if (j > 0) {
      for (int e = 0; e < 9; ++e) a.feat[(size_t)b * a.Kp + (j - 1) * 9 + e] = rel[e];
      const float bone[3] = {rest[j * 3] - rest[par * 3], rest[j * 3 + 1] - rest[par * 3 + 1], rest[j * 3 + 2] - rest[par * 3 + 2]};
      float rb[3];
      mat3_vec(glob + par * 9, bone, rb);
      for (int c = 0; c < 3; ++c) pos[j * 3 + c] = pos[par * 3 + c] + rb[c];
    } else {
      for (int c = 0; c < 3; ++c) pos[c] = rest[c];
    }
    float rj[3];
    mat3_vec(G, rest + j * 3, rj);
    for (int e = 0; e < 9; ++e) SF_IM(a.skin, j * 12 + e, Bp, b) = G[e];
    for (int c = 0; c < 3; ++c) SF_IM(a.skin, j * 12 + 9 + c, Bp, b) = (pos[j * 3 + c] - rj[c]) + tr[c];
    if (a.skin4 != nullptr) {
      for (int c = 0; c < 3; ++c)
        reinterpret_cast<float4*>(a.skin4)[(size_t)(j * 3 + c) * Bp + b] =
            make_float4(G[c * 3], G[c * 3 + 1], G[c * 3 + 2], (pos[j * 3 + c] - rj[c]) + tr[c]);
    }
    if (live) {
      for (int e = 0; e < 9; ++e) a.out_orientations[((size_t)b * J + j) * 9 + e] = G[e];
      for (int c = 0; c < 3; ++c) a.out_joints[((size_t)b * J + j) * 3 + c] = pos[j * 3 + c] + tr[c];
    }
  }
  for (int k = 9 * (J - 1); k < a.Kp; ++k) a.feat[(size_t)b * a.Kp + k] = 0.f;
}

struct FwdSkinArgs {
  const float* vposedT;  // [3V][Bp] rows in internal order
  const float* betaT;    // [S+1][Bp]
  const float* skin;     // [12J][Bp]
  const float* shapedirs;     // (V,3,S)
  const float* kid_shapedir;  // (V,3)
  const int32_t* skin_idx;
  const float* skin_w;
  const int32_t* inv_order;
  float* out;  // (B,V,3)
  int V, S, B, Bp, skin_k, use_kid, nb;
};

// one warp = 32 instances x 32 consecutive model vertices; 3 warps per CTA
__global__ void __launch_bounds__(96) k_fwd_skin(const FwdSkinArgs a) {
  __shared__ float tile[3][32][97];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vt = blockIdx.x * 3 + warp;
  const int g = blockIdx.y;
  const int v0 = vt * 32;
  if (v0 >= a.V) return;
  const int Bp = a.Bp, b = g * 32 + lane;
  float beta[SMPLFIT_MAX_UNKNOWNS];
#pragma unroll
  for (int s = 0; s < SMPLFIT_MAX_UNKNOWNS; ++s) beta[s] = (s < a.nb) ? SF_IM(a.betaT, s, Bp, b) : 0.f;
  const float kid = a.use_kid ? SF_IM(a.betaT, a.S, Bp, b) : 0.f;
  const int nv = min(32, a.V - v0);
  for (int q = 0; q < nv; ++q) {
    const int v = v0 + q;
    const int i = __ldg(a.inv_order + v);
    float vs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = SF_IM(a.vposedT, i * 3 + c, Bp, b);
      const float* sd = a.shapedirs + ((size_t)v * 3 + c) * a.S;
#pragma unroll
      for (int s = 0; s < SMPLFIT_MAX_UNKNOWNS; ++s)
        if (s < a.nb) x = fmaf(__ldg(sd + s), beta[s], x);
      if (a.use_kid) x = fmaf(__ldg(a.kid_shapedir + v * 3 + c), kid, x);
      vs[c] = x;
    }
    float o[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < a.skin_k; ++k) {
      const int j = __ldg(a.skin_idx + v * a.skin_k + k);
      const float w = __ldg(a.skin_w + v * a.skin_k + k);
      const float* sk = a.skin + (size_t)(j * 12) * Bp + b;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float y = sk[(size_t)(9 + c) * Bp];
        y = fmaf(sk[(size_t)(c * 3 + 0) * Bp], vs[0], y);
        y = fmaf(sk[(size_t)(c * 3 + 1) * Bp], vs[1], y);
        y = fmaf(sk[(size_t)(c * 3 + 2) * Bp], vs[2], y);
        o[c] = fmaf(w, y, o[c]);
      }
    }
    tile[warp][lane][q * 3 + 0] = o[0];
    tile[warp][lane][q * 3 + 1] = o[1];
    tile[warp][lane][q * 3 + 2] = o[2];
  }
  __syncwarp();
  const int width = nv * 3;
  for (int r = 0; r < 32; ++r) {
    const int bb = g * 32 + r;
    if (bb >= a.B) break;
    float* dst = a.out + ((size_t)bb * a.V + v0) * 3;
    for (int e = lane; e < width; e += 32) dst[e] = tile[warp][r][e];
  }
}

// ---------------------------------------------------------------------------------------
// k_fwd_skin_rec: the skinning pass in the record style of the fit kernels.  CTA = 32 instances
// (lane = instance) x 8 warps; the per-joint [R | t] rows of the group are staged in shared memory
// with cp.async; each warp takes blocks of 32 consecutive model vertices, reads one packed record
// per vertex (4 weights, 4 joint ids, v_posed row, shapedirs[3][SP], kid_shapedir[3]) with
// warp-uniform 16-byte loads one vertex ahead, prefetches the v_posed values two ahead, and
// writes the caller's (B,V,3) layout through a padded shared-memory transpose (coalesced rows).
// ---------------------------------------------------------------------------------------
struct FwdSkinRecArgs {
  const float* vposedT;  // [3V][Bp], rows in internal order
  const float* betaT;    // [S+1][Bp]
  const float* skin;     // [12J][Bp]
  const float* rec;      // [V][rec_len]: w4 | idx4 | row(int) pad3 | S[3][SP] | kid[3] pad
  float* out;            // (B,V,3)
  int V, J, S, SP, rec_len, B, Bp, nb, use_kid, blocks_per_warp;
};

template <int SP>
__global__ void __launch_bounds__(256, 2) k_fwd_skin_rec(const FwdSkinRecArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s_skin = sm;                               // [12J][32]
  float* s_tile = sm + (size_t)a.J * 12 * 32;       // [8 warps][32][49]: 16 vertices x 3 + pad
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp, b = g * 32 + lane;
  {
    const int n16 = a.J * 12 * 8;
    for (int q = threadIdx.x; q < n16; q += 256) {
      const int r = q >> 3, part = q & 7;
      const float* src = a.skin + (size_t)r * Bp + g * 32 + part * 4;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_skin + r * 32 + part * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  float beta[SP];
#pragma unroll
  for (int s = 0; s < SP; ++s) beta[s] = (s < a.nb) ? SF_IM(a.betaT, s, Bp, b) : 0.f;
  const float kid = a.use_kid ? SF_IM(a.betaT, a.S, Bp, b) : 0.f;
  float* tile = s_tile + (size_t)warp * 32 * 49;
  const int n_blocks = (a.V + 15) / 16;
  for (int q = 0; q < a.blocks_per_warp; ++q) {
    const int blk = (blockIdx.x * a.blocks_per_warp + q) * 8 + warp;
    if (blk >= n_blocks) break;
    const int v0 = blk * 16, nv = min(16, a.V - v0);
    const float* rec = a.rec + (size_t)v0 * a.rec_len;
    float4 nw = __ldg(reinterpret_cast<const float4*>(rec));
    int4 nj = __ldg(reinterpret_cast<const int4*>(rec + 4));
    int nrow = __float_as_int(__ldg(rec + 8));
    float nx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) nx[c] = SF_IM(a.vposedT, nrow * 3 + c, Bp, b);
    float Sc[12];
    int cj = -1;
    for (int k = 0; k < nv; ++k) {
      const float4 w4 = nw;
      const int4 j4 = nj;
      float x[3] = {nx[0], nx[1], nx[2]};
      const float* sd = rec + 12;
      if (k + 1 < nv) {
        rec += a.rec_len;
        nw = __ldg(reinterpret_cast<const float4*>(rec));
        nj = __ldg(reinterpret_cast<const int4*>(rec + 4));
        nrow = __float_as_int(__ldg(rec + 8));
#pragma unroll
        for (int c = 0; c < 3; ++c) nx[c] = SF_IM(a.vposedT, nrow * 3 + c, Bp, b);
      }
      float vs[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float y = x[c];
#pragma unroll
        for (int s2 = 0; s2 < SP; s2 += 2) {
          const float2 sv = __ldg(reinterpret_cast<const float2*>(sd + c * SP + s2));
          y = fmaf(sv.x, beta[s2], y);
          y = fmaf(sv.y, beta[s2 + 1], y);
        }
        if (a.use_kid) y = fmaf(__ldg(sd + 3 * SP + c), kid, y);
        vs[c] = y;
      }
      if (j4.x != cj) {
        cj = j4.x;
        const float* p = s_skin + (size_t)(cj * 12) * 32 + lane;
#pragma unroll
        for (int e = 0; e < 12; ++e) Sc[e] = p[e * 32];
      }
      float o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        o[c] = w4.x * fmaf(Sc[c * 3], vs[0], fmaf(Sc[c * 3 + 1], vs[1], fmaf(Sc[c * 3 + 2], vs[2], Sc[9 + c])));
      const float wk[3] = {w4.y, w4.z, w4.w};
      const int jk[3] = {j4.y, j4.z, j4.w};
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        if (wk[kk] != 0.f) {
          const float* p = s_skin + (size_t)(jk[kk] * 12) * 32 + lane;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float y = fmaf(p[(c * 3) * 32], vs[0], fmaf(p[(c * 3 + 1) * 32], vs[1], fmaf(p[(c * 3 + 2) * 32], vs[2], p[(9 + c) * 32])));
            o[c] = fmaf(wk[kk], y, o[c]);
          }
        }
      }
      tile[lane * 49 + k * 3 + 0] = o[0];
      tile[lane * 49 + k * 3 + 1] = o[1];
      tile[lane * 49 + k * 3 + 2] = o[2];
    }
    __syncwarp();
    const int width = nv * 3;
    for (int r = 0; r < 32; ++r) {
      const int bb = g * 32 + r;
      if (bb >= a.B) break;
      float* dst = a.out + ((size_t)bb * a.V + v0) * 3;
      for (int e = lane; e < width; e += 32) dst[e] = tile[r * 49 + e];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------
// k_fwd_skin_tma: the skinning pass with the streams staged by TMA (same scheme as k_stats_lite).  v_posed^T comes
// from the forward GEMM in MODEL vertex order (posedirs_model_hi / lo), so a warp's block of FWD_BLK consecutive
// model vertices is FWD_BLK * 3 contiguous rows: per-warp two-stage ring, one 2D tensor-map box {32 instances,
// 3 LITE_VS rows} + one bulk copy of LITE_VS records per stage on one mbarrier.  Joint rows as float4 quads in shared
// memory, cached in registers per skinning slot; results go through a padded shared-memory tile so that the
// caller's (B,V,3) rows are written in 192-byte contiguous pieces.
// ---------------------------------------------------------------------------------------
constexpr int FWD_BLK = 16;  // model vertices per output tile
struct FwdSkinTmaArgs {
  const float* betaT;    // [S+1][Bp]
  const float* skin4;    // [J*3][Bp] float4
  const float* rec;      // [V][rec_len]: w4 | idx4 | (unused) | S[3][SP] | kid[3] pad
  float* out;            // (B,V,3)
  int V, J, S, rec_len, B, Bp, nb, use_kid, blocks_per_warp;
};

__host__ __device__ inline int fwd_stage_floats(int rec_len) { return (3 * LITE_VS * 32 + LITE_VS * rec_len + 31) / 32 * 32; }
__host__ __device__ inline size_t fwd_tma_smem_bytes(int J, int rec_len, int warps) {
  return ((size_t)J * 3 * 128 + (size_t)warps * (2 * fwd_stage_floats(rec_len) + 32 * (FWD_BLK * 3 + 1))) * sizeof(float) +
         (size_t)warps * 16 + 16;
}

template <int SP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_fwd_skin_tma(const FwdSkinTmaArgs a, const __grid_constant__ CUtensorMap map_vp) {
  extern __shared__ __align__(128) float s_fw[];
  constexpr int BOX = 3 * LITE_VS * 32, TW = FWD_BLK * 3 + 1;
  const int STAGE = fwd_stage_floats(a.rec_len);
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp, b = g * 32 + lane;
  const float4* sq = reinterpret_cast<const float4*>(s_fw);  // [J*3][32]
  float* wbase = s_fw + (size_t)a.J * 3 * 128;
  float* stage_buf = wbase + (size_t)warp * (2 * STAGE);
  float* tile = wbase + (size_t)WARPS * (2 * STAGE) + (size_t)warp * (32 * TW);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wbase + (size_t)WARPS * (2 * STAGE + 32 * TW)) + 2 * warp;
  if (lane == 0) {
    sf_mbar_init(bar, 1);
    sf_mbar_init(bar + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  {
    const int n = a.J * 3 * 32;
    const float4* src = reinterpret_cast<const float4*>(a.skin4);
    for (int q = threadIdx.x; q < n; q += WARPS * 32) {
      const int r = q >> 5, l = q & 31;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_fw + (size_t)q * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)r * Bp + g * 32 + l) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  float beta[SP];
#pragma unroll
  for (int s = 0; s < SP; ++s) beta[s] = (s < a.nb) ? SF_IM(a.betaT, s, Bp, b) : 0.f;
  const float kid = a.use_kid ? SF_IM(a.betaT, a.S, Bp, b) : 0.f;
  const int n_blocks = (a.V + FWD_BLK - 1) / FWD_BLK;
  uint32_t phase = 0;
  auto issue = [&](int first, int i1, int k) {
    __syncwarp();
    if (first < i1 && lane == 0) {
      float* dst = stage_buf + (size_t)(k & 1) * STAGE;
      const uint32_t rec_bytes = (uint32_t)min(LITE_VS, i1 - first) * a.rec_len * 4;
      sf_mbar_expect_tx(bar + (k & 1), (uint32_t)BOX * 4u + rec_bytes);
      sf_tma_2d(dst, &map_vp, bar + (k & 1), g * 32, first * 3);
      sf_bulk_g2s(dst + BOX, a.rec + (size_t)first * a.rec_len, rec_bytes, bar + (k & 1));
    }
  };
  JointCache jc;
  jc.reset();
  for (int q = 0; q < a.blocks_per_warp; ++q) {
    const int blk = (blockIdx.x * a.blocks_per_warp + q) * WARPS + warp;
    if (blk >= n_blocks) break;
    const int v0 = blk * FWD_BLK, v1 = min(a.V, v0 + FWD_BLK);
    const int nsub = (v1 - v0 + LITE_VS - 1) / LITE_VS;
    issue(v0, v1, 0);
    issue(v0 + LITE_VS, v1, 1);
    for (int k = 0; k < nsub; ++k) {
      sf_mbar_wait(bar + (k & 1), (phase >> (k & 1)) & 1u);
      phase ^= 1u << (k & 1);
      if (k >= 1) issue(v0 + (k + 1) * LITE_VS, v1, k + 1);
      const float* sg = stage_buf + (size_t)(k & 1) * STAGE;
      const int nv = min(LITE_VS, v1 - (v0 + k * LITE_VS));
#pragma unroll 1
      for (int u = 0; u < nv; ++u) {
        const float* rec = sg + BOX + u * a.rec_len;
        const float4 w4 = *reinterpret_cast<const float4*>(rec);
        const int4 j4 = *reinterpret_cast<const int4*>(rec + 4);
        const float wk[4] = {w4.x, w4.y, w4.z, w4.w};
        const int jk[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (wk[kk] != 0.f && jk[kk] != jc.j[kk]) {  // warp-uniform
            jc.j[kk] = jk[kk];
#pragma unroll
            for (int c = 0; c < 3; ++c) jc.q[kk][c] = sq[(size_t)(jk[kk] * 3 + c) * 32 + lane];
          }
        }
        constexpr int NV4 = (3 * SP + 3 + 3) / 4;
        float sdv[NV4 * 4];  // shapedirs[c][s] then kid_shapedir[c] of the record, read as 16-byte words
#pragma unroll
        for (int q4 = 0; q4 < NV4; ++q4) {
          const float4 v4 = *reinterpret_cast<const float4*>(rec + 12 + 4 * q4);
          sdv[4 * q4] = v4.x; sdv[4 * q4 + 1] = v4.y; sdv[4 * q4 + 2] = v4.z; sdv[4 * q4 + 3] = v4.w;
        }
        float vs[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float2 y2 = make_float2(sg[(u * 3 + c) * 32 + lane], 0.f);
#pragma unroll
          for (int s2 = 0; s2 < SP; s2 += 2)
            y2 = sf_fma2(make_float2(sdv[c * SP + s2], sdv[c * SP + s2 + 1]), make_float2(beta[s2], beta[s2 + 1]), y2);
          float y = y2.x + y2.y;
          if (a.use_kid) y = fmaf(sdv[3 * SP + c], kid, y);
          vs[c] = y;
        }
        float2 B2[6];
        jc.blend(wk, B2);
        const int col = (k * LITE_VS + u) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          tile[lane * TW + col + c] =
              fmaf(B2[2 * c].x, vs[0], fmaf(B2[2 * c].y, vs[1], fmaf(B2[2 * c + 1].x, vs[2], B2[2 * c + 1].y)));
      }
    }
    __syncwarp();
    const int width = (v1 - v0) * 3;
    float* obase = a.out + (size_t)v0 * 3;
    if (width == FWD_BLK * 3) {  // full tile: compile-time row length
#pragma unroll 4
      for (int idx = lane; idx < 32 * FWD_BLK * 3; idx += 32) {
        const int r = idx / (FWD_BLK * 3), e = idx - r * (FWD_BLK * 3);
        const int bb = g * 32 + r;
        if (bb < a.B) obase[(size_t)bb * a.V * 3 + e] = tile[r * TW + e];
      }
    } else {
      for (int idx = lane; idx < 32 * width; idx += 32) {
        const int r = idx / width, e = idx - r * width;
        const int bb = g * 32 + r;
        if (bb < a.B) obase[(size_t)bb * a.V * 3 + e] = tile[r * TW + e];
      }
    }
    __syncwarp();
  }
}

// CSR SpMM of BodyConverter.convert_vertices: out[b][r][:] = sum_k data[k] in[b][indices[k]][:].
// One thread per (instance, output vertex, coordinate); rows hold ~3 non-zeros (barycentric).
__global__ void k_csr_apply(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const float* __restrict__ data, int v_out, int v_in, long long total,
                            const float* __restrict__ in, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % 3);
  const long long rb = idx / 3;
  const int r = (int)(rb % v_out);
  const long long b = rb / v_out;
  const float* src = in + (size_t)b * v_in * 3 + c;
  float acc = 0.f;
  for (int k = indptr[r]; k < indptr[r + 1]; ++k) acc = fmaf(__ldg(data + k), src[(size_t)indices[k] * 3], acc);
  out[idx] = acc;
}

static bool fwd_tma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SMPLFIT_B200_FWD_VARIANT");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

struct FwdWs {
  float *vposedT, *feat, *skin, *skin4, *betaT;
  void* tc_scratch;
  size_t bytes;
};

static FwdWs carve_fwd(void* base, const smplfit_model_t* m, int64_t B) {
  FwdWs w{};
  Carver c(base);
  const size_t Bp = roundup((int)B, 32);
  const int Kp = roundup(m->num_pose_feats, 16);
  w.vposedT = c.take<float>((size_t)3 * m->num_vertices * Bp);
  w.feat = c.take<float>(Bp * Kp);
  w.skin = c.take<float>((size_t)12 * m->num_joints * Bp);
  w.skin4 = c.take<float>((size_t)12 * m->num_joints * Bp);
  w.betaT = c.take<float>((size_t)(m->num_betas + 1) * Bp);
  w.tc_scratch = c.take<char>(vposed_tc_scratch_bytes(m, (int)Bp));
  w.bytes = c.off + 256;
  return w;
}

}  // namespace sf

using namespace sf;

extern "C" size_t smplfit_forward_workspace_bytes(const smplfit_model_t* m, int64_t batch) {
  if (!m || batch <= 0) return 0;
  return carve_fwd(nullptr, m, batch).bytes;
}

extern "C" int smplfit_forward(const smplfit_model_t* m, int64_t batch, int rot_mode, const float* rot,
                               const float* betas, int n_betas, const float* trans, const float* kid,
                               float* out_vertices, float* out_joints, float* out_orientations, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!m || !out_joints || !out_orientations) return fail(SMPLFIT_ERR_ARG, "missing required pointer");
  if (batch <= 0) return fail(SMPLFIT_ERR_ARG, "batch must be positive");
  if (m->num_joints > SMPLFIT_MAX_JOINTS) return fail(SMPLFIT_ERR_UNSUPPORTED, "num_joints > 64");
  if (m->num_betas > SMPLFIT_MAX_UNKNOWNS) return fail(SMPLFIT_ERR_UNSUPPORTED, "num_betas > 17 in forward");
  if (rot_mode < 0 || rot_mode > 3 || (rot_mode != 3 && !rot)) return fail(SMPLFIT_ERR_ARG, "bad rotation input");
  FwdWs w = carve_fwd(workspace, m, batch);
  if (!workspace || w.bytes > workspace_bytes) return fail(SMPLFIT_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int B = (int)batch, Bp = roundup(B, 32), Kp = roundup(m->num_pose_feats, 16);
  FwdPrepArgs p;
  p.rot = rot; p.betas = betas; p.trans = trans; p.kid = kid; p.parents = m->parents;
  p.J_template = m->J_template; p.J_shapedirs = m->J_shapedirs; p.kid_J_shapedir = m->kid_J_shapedir;
  p.skin = w.skin; p.skin4 = w.skin4; p.feat = w.feat; p.betaT = w.betaT; p.out_joints = out_joints;
  p.out_orientations = out_orientations; p.rot_mode = rot_mode; p.n_betas = betas ? n_betas : 0;
  p.J = m->num_joints; p.S = m->num_betas; p.B = B; p.Bp = Bp; p.Kp = Kp;
  SF_LAUNCH(k_fwd_prep, Bp / 32, 32, 0, st, p);
  if (out_vertices != nullptr) {
    // fast path: GEMM with the MODEL-order posedirs copy, then the TMA-staged skinning kernel
    const int SPf = (m->num_betas + 1) / 2 * 2;
    const bool tma_ok = fwd_tma_enabled() && m->posedirs_model_hi != nullptr && m->posedirs_model_lo != nullptr &&
                        m->fwd_rec != nullptr && m->skin_k <= 4 && m->num_betas <= 16 && tensor_maps_available();
    if (tma_ok) {
      const int warps = fwd_tma_smem_bytes(m->num_joints, m->fwd_rec_len, 12) <= 227 * 1024 ? 12 : 8;
      CUtensorMap mv;
      if (fwd_tma_smem_bytes(m->num_joints, m->fwd_rec_len, warps) <= 227 * 1024 &&
          make_im_map(&mv, w.vposedT, (uint64_t)3 * m->num_vertices, (uint64_t)Bp, 3 * LITE_VS) &&
          vposed_tc_run_model(m, w.feat, w.vposedT, Bp, Kp, w.tc_scratch, st)) {
        FwdSkinTmaArgs r;
        r.betaT = w.betaT; r.skin4 = w.skin4; r.rec = m->fwd_rec; r.out = out_vertices; r.V = m->num_vertices;
        r.J = m->num_joints; r.S = m->num_betas; r.rec_len = m->fwd_rec_len; r.B = B; r.Bp = Bp;
        r.nb = betas ? min(n_betas, m->num_betas) : 0; r.use_kid = kid != nullptr; r.blocks_per_warp = 4;
        const int n_blocks = (r.V + FWD_BLK - 1) / FWD_BLK;
        const size_t smem = fwd_tma_smem_bytes(r.J, r.rec_len, warps);
        dim3 grid2((n_blocks + warps * r.blocks_per_warp - 1) / (warps * r.blocks_per_warp), Bp / 32);
#define SF_FWD_TMA(SPV)                                                                                              \
  do {                                                                                                               \
    if (warps == 12) {                                                                                               \
      cudaFuncSetAttribute(k_fwd_skin_tma<SPV, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
      SF_LAUNCH((k_fwd_skin_tma<SPV, 12>), grid2, 12 * 32, smem, st, r, mv);                                         \
    } else {                                                                                                         \
      cudaFuncSetAttribute(k_fwd_skin_tma<SPV, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
      SF_LAUNCH((k_fwd_skin_tma<SPV, 8>), grid2, 8 * 32, smem, st, r, mv);                                           \
    }                                                                                                                \
  } while (0)
        switch (SPf) {
          case 2: SF_FWD_TMA(2); break;
          case 4: SF_FWD_TMA(4); break;
          case 6: SF_FWD_TMA(6); break;
          case 8: SF_FWD_TMA(8); break;
          case 10: SF_FWD_TMA(10); break;
          case 12: SF_FWD_TMA(12); break;
          case 14: SF_FWD_TMA(14); break;
          default: SF_FWD_TMA(16); break;
        }
#undef SF_FWD_TMA
        SF_CHECK_LAST();
        return SMPLFIT_OK;
      }
    }
    if (!vposed_tc_run(m, w.feat, w.vposedT, Bp, Kp, w.tc_scratch, st)) {
      dim3 grid((3 * m->num_vertices + 127) / 128, (Bp + 63) / 64);
      SF_LAUNCH(k_vposed_gemm_simt, grid, 256, 0, st, m->posedirs_fit, m->v_template_fit, w.feat,
                3 * m->num_vertices, Kp, Bp, w.vposedT);
    }
    FwdSkinArgs s;
    s.vposedT = w.vposedT; s.betaT = w.betaT; s.skin = w.skin; s.shapedirs = m->shapedirs;
    s.kid_shapedir = m->kid_shapedir; s.skin_idx = m->skin_idx; s.skin_w = m->skin_w; s.inv_order = m->inv_order;
    s.out = out_vertices; s.V = m->num_vertices; s.S = m->num_betas; s.B = B; s.Bp = Bp; s.skin_k = m->skin_k;
    s.use_kid = kid != nullptr; s.nb = betas ? min(n_betas, m->num_betas) : 0;
    if (m->fwd_rec != nullptr && m->skin_k <= 4 && m->num_betas <= 16) {
      FwdSkinRecArgs r;
      r.vposedT = w.vposedT; r.betaT = w.betaT; r.skin = w.skin; r.rec = m->fwd_rec; r.out = out_vertices;
      r.V = m->num_vertices; r.J = m->num_joints; r.S = m->num_betas; r.SP = (m->num_betas + 1) / 2 * 2;
      r.rec_len = m->fwd_rec_len; r.B = B; r.Bp = Bp; r.nb = s.nb; r.use_kid = s.use_kid; r.blocks_per_warp = 8;
      const int n_blocks = (r.V + 15) / 16;
      const size_t smem = ((size_t)r.J * 12 * 32 + (size_t)8 * 32 * 49) * sizeof(float);
      dim3 grid2((n_blocks + 8 * r.blocks_per_warp - 1) / (8 * r.blocks_per_warp), Bp / 32);
#define SF_FWD_SKIN(SPV)                                                                                       \
  do {                                                                                                         \
    cudaFuncSetAttribute(k_fwd_skin_rec<SPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    SF_LAUNCH(k_fwd_skin_rec<SPV>, grid2, 256, smem, st, r);                                                   \
  } while (0)
      switch (r.SP) {
        case 2: SF_FWD_SKIN(2); break;
        case 4: SF_FWD_SKIN(4); break;
        case 6: SF_FWD_SKIN(6); break;
        case 8: SF_FWD_SKIN(8); break;
        case 10: SF_FWD_SKIN(10); break;
        case 12: SF_FWD_SKIN(12); break;
        case 14: SF_FWD_SKIN(14); break;
        default: SF_FWD_SKIN(16); break;
      }
#undef SF_FWD_SKIN
    } else {
      dim3 grid(((m->num_vertices + 31) / 32 + 2) / 3, Bp / 32);
      SF_LAUNCH(k_fwd_skin, grid, 96, 0, st, s);
    }
  }
  SF_CHECK_LAST();
  return SMPLFIT_OK;
}

extern "C" int smplfit_convert_vertices(const int32_t* indptr, const int32_t* indices, const float* data,
                                        int32_t v_out, int32_t v_in, int64_t batch, const float* in_vertices,
                                        float* out_vertices, void* stream) {
  if (!indptr || !indices || !data || !in_vertices || !out_vertices) return fail(SMPLFIT_ERR_ARG, "NULL pointer");
  if (batch <= 0) return SMPLFIT_OK;
  const long long total = (long long)batch * v_out * 3;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SF_LAUNCH(k_csr_apply, (unsigned)((total + 255) / 256), 256, 0, st, indptr, indices, data, v_out, v_in, total,
            in_vertices, out_vertices);
  SF_CHECK_LAST();
  return SMPLFIT_OK;
}
