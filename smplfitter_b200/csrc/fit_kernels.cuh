// Kernels of BodyFitter.fit for sm_100a.
//
// Data layout ("instance-minor"): every per-instance array lives in HBM as [row][Bp] with the
// batch index fastest (Bp = batch rounded up to 32).  A warp's 32 lanes are 32 instances, so
// every global access of the hot passes is one fully coalesced 128-byte line, model constants
// are warp-uniform broadcasts, and all per-instance accumulators stay in registers with no
// cross-lane reduction.  The caller's (B,V,3) targets are re-laid out once (k_transpose).
//
// Vertices are processed in an internal order (grouped by body part, see masks.py), so a
// statistics segment belongs to one part and is flushed exactly once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smplfit_b200.h"
#include "linalg.cuh"

namespace sf {

#define SF_IM(ptr, row, Bp, b) (ptr)[(size_t)(row) * (size_t)(Bp) + (size_t)(b)]

// ---------------------------------------------------------------------------------------
// k_mean: unweighted mean over cat[vertices, joints] (pt/bodyfitter.py:355-361).  One CTA of MEAN_THREADS per
// instance: the 12 V bytes of an instance are one contiguous row, read with a stride of 3 * MEAN_THREADS elements so
// that each of a thread's three accumulators keeps one coordinate ((tid + k) % 3 since MEAN_THREADS % 3 == 1), six
// independent loads in flight per thread; shuffle + shared-memory reduction in a fixed order.
// ---------------------------------------------------------------------------------------
constexpr int MEAN_THREADS = 256;
static_assert(MEAN_THREADS % 3 == 1, "coordinate phase of the strided accumulators");

static __device__ __forceinline__ void mean_row(const float* __restrict__ row, int n, int tid, float& a0, float& a1, float& a2) {
  constexpr int T = MEAN_THREADS;
  int i = tid;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  for (; i + 5 * T < n; i += 6 * T) {
    a0 += row[i];
    a1 += row[i + T];
    a2 += row[i + 2 * T];
    b0 += row[i + 3 * T];
    b1 += row[i + 4 * T];
    b2 += row[i + 5 * T];
  }
  a0 += b0;
  a1 += b1;
  a2 += b2;
  for (; i + 2 * T < n; i += 3 * T) {
    a0 += row[i];
    a1 += row[i + T];
    a2 += row[i + 2 * T];
  }
  if (i < n) a0 += row[i];
  if (i + T < n) a1 += row[i + T];
}

static __global__ void __launch_bounds__(MEAN_THREADS) k_mean(const float* __restrict__ tv, const float* __restrict__ tj, int V,
                                                              int J, int B, int Bp, float* __restrict__ mean) {
  __shared__ float red[MEAN_THREADS / 32][3];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;  // coordinates tid % 3, (tid + 1) % 3, (tid + 2) % 3
  if (b < B) {
    mean_row(tv + (size_t)b * 3 * V, 3 * V, tid, a0, a1, a2);
    if (tj != nullptr) mean_row(tj + (size_t)b * 3 * J, 3 * J, tid, a0, a1, a2);
  }
  const int c0 = tid % 3;
  float s[3];
  s[c0] = a0;
  s[(c0 + 1) % 3] = a1;
  s[(c0 + 2) % 3] = a2;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = s[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][c] = v;
  }
  __syncthreads();
  if (tid < 3) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < MEAN_THREADS / 32; ++w) v += red[w][tid];
    const float cnt = (float)(V + (tj != nullptr ? J : 0));
    SF_IM(mean, tid, Bp, b) = (b < B) ? v / cnt : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// k_transpose: (B,N,C) instance-major -> [pos(n)*C + c][Bp] instance-minor, optionally
// subtracting the per-instance mean (C == 3) and renumbering rows through `inv_order`.
// 32 instances x 32 items per CTA through a padded shared tile; both sides coalesced.
// ---------------------------------------------------------------------------------------
template <int C>
__global__ void k_transpose(const float* __restrict__ src, int N, int B, int Bp,
                            const int32_t* __restrict__ inv_order, const float* __restrict__ mean,
                            float* __restrict__ dst) {
  __shared__ float tile[32][32 * C + 1];
  const int n0 = blockIdx.x * 32;
  const int g = blockIdx.y;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  const int width = min(32, N - n0) * C;
  for (int r = ty; r < 32; r += 8) {
    const int b = g * 32 + r;
    for (int e = tx; e < 32 * C; e += 32)
      tile[r][e] = (b < B && e < width) ? src[((size_t)b * N + n0) * C + e] : 0.f;
  }
  __syncthreads();
  const int b = g * 32 + tx;
  for (int e = ty; e < width; e += 8) {
    const int n = n0 + e / C, c = e % C;
    const int pos = inv_order ? inv_order[n] : n;
    float v = tile[tx][e];
    if (mean != nullptr) v -= SF_IM(mean, c, Bp, b);
    SF_IM(dst, pos * C + c, Bp, b) = (b < B) ? v : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// k_transpose_v: the vertex re-layout (C == 3) with a larger tile: 32 instances x 64 vertices per CTA (768
// contiguous bytes read per instance), the destination rows of the tile (inv_order) staged once in shared memory,
// the per-instance mean held in registers.  dst[(pos(n)*3 + c)][b] = src[b][n][c] - mean[c][b].
// ---------------------------------------------------------------------------------------
constexpr int TV_N = 64;
static __global__ void __launch_bounds__(256) k_transpose_v(const float* __restrict__ src, int N, int B, int Bp,
                                                            const int32_t* __restrict__ inv_order,
                                                            const float* __restrict__ mean, float* __restrict__ dst) {
  __shared__ float tile[32][TV_N * 3 + 1];
  __shared__ int s_pos[TV_N];
  const int n0 = blockIdx.x * TV_N;
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nn = min(TV_N, N - n0), width = nn * 3;
  if (threadIdx.x < TV_N) s_pos[threadIdx.x] = (threadIdx.x < nn) ? (inv_order ? inv_order[n0 + threadIdx.x] : n0 + threadIdx.x) : 0;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    const int b = g * 32 + r;
    const float* row = src + ((size_t)b * N + n0) * 3;
#pragma unroll
    for (int k = 0; k < TV_N * 3 / 32; ++k) {
      const int e = k * 32 + lane;
      tile[r][e] = (b < B && e < width) ? row[e] : 0.f;
    }
  }
  __syncthreads();
  const int b = g * 32 + lane;
  float m[3] = {0.f, 0.f, 0.f};
  if (mean != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) m[c] = SF_IM(mean, c, Bp, b);
  }
  const bool live = b < B;
  for (int v = warp; v < nn; v += 8) {
    const int pos = s_pos[v];
#pragma unroll
    for (int c = 0; c < 3; ++c) SF_IM(dst, pos * 3 + c, Bp, b) = live ? tile[lane][v * 3 + c] - m[c] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// k_vposed_gemm_simt: v_posed^T[n][b] = v_template_fit[n] + sum_k posedirs_fit[n][k] feat[b][k]
// (pt/bodyfitter.py:913-916).  Plain FP32 shared-memory tiled GEMM (128 x 64 x 16 tiles,
// 8 x 4 register micro-tiles); the tcgen05 kernel in vposed_tc.cu replaces it.
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_vposed_gemm_simt(const float* __restrict__ A, const float* __restrict__ vt,
                                                          const float* __restrict__ F, int M, int Kp, int Bp,
                                                          float* __restrict__ Cout) {
  __shared__ float As[16][128 + 4];
  __shared__ float Fs[16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 128, b0 = blockIdx.y * 64;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads; thread tile 8 (m) x 4 (b)
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < Kp; k0 += 16) {
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const int idx = tid + l * 256;       // 512 float4 = 128 rows x 4
      const int r = idx >> 2, kq = (idx & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * Kp + k0 + kq);
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
    {
      const int r = tid >> 2, kq = (tid & 3) * 4;  // 256 float4 = 64 rows x 4
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < Bp) v = *reinterpret_cast<const float4*>(F + (size_t)(b0 + r) * Kp + k0 + kq);
      Fs[kq + 0][r] = v.x; Fs[kq + 1][r] = v.y; Fs[kq + 2][r] = v.z; Fs[kq + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[8], f[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = Fs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], f[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m < M && b0 + tx * 4 < Bp) {
      const float t = vt[m];
      float4 o = make_float4(acc[i][0] + t, acc[i][1] + t, acc[i][2] + t, acc[i][3] + t);
      *reinterpret_cast<float4*>(Cout + (size_t)m * Bp + b0 + tx * 4) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_regress_csr: the same regression from the regressor's non-zeros (CSR over the internal vertex order): one warp per
// (32 instances, joint); the three coordinate rows of every referenced vertex are coalesced 128-byte loads, four
// non-zeros in flight per step.
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128) k_regress_csr(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                            const float* __restrict__ val, const float* __restrict__ X, int J,
                                                            int Bp, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int g = warp / J, j = warp - g * J;
  if (g * 32 >= Bp) return;
  const int b = g * 32 + lane;
  const int e0 = __ldg(ptr + j), e1 = __ldg(ptr + j + 1);
  float acc[3] = {0.f, 0.f, 0.f};
  int e = e0;
  for (; e + 4 <= e1; e += 4) {
    int i[4];
    float w[4], x[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      i[u] = __ldg(idx + e + u);
      w[u] = __ldg(val + e + u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < 3; ++c) x[u][c] = SF_IM(X, i[u] * 3 + c, Bp, b);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] = fmaf(w[u], x[u][c], acc[c]);
  }
  for (; e < e1; ++e) {
    const int i = __ldg(idx + e);
    const float w = __ldg(val + e);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = fmaf(w, SF_IM(X, i * 3 + c, Bp, b), acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) SF_IM(out, j * 3 + c, Bp, b) = acc[c];
}

// ---------------------------------------------------------------------------------------
// k_regress: joints^T[j*3+c][b] = sum_i Jreg_fit[j][i] X^T[i*3+c][b]
// (pt/bodyfitter.py:1342-1344).  One warp per (instance group, block of 8 joints).
// ---------------------------------------------------------------------------------------
static __global__ void k_regress(const float* __restrict__ Jreg, const float* __restrict__ X, int V, int J, int Bp,
                          float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int jblocks = (J + 7) / 8;
  const int g = warp / jblocks, jb = warp % jblocks;
  if (g * 32 >= Bp) return;
  const int b = g * 32 + lane;
  const int j0 = jb * 8;
  float acc[8][3];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
  for (int i = 0; i < V; ++i) {
    const float x = SF_IM(X, i * 3 + 0, Bp, b), y = SF_IM(X, i * 3 + 1, Bp, b), z = SF_IM(X, i * 3 + 2, Bp, b);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (j0 + q < J) {
        const float w = __ldg(Jreg + (size_t)(j0 + q) * V + i);
        if (w != 0.f) {
          acc[q][0] = fmaf(w, x, acc[q][0]);
          acc[q][1] = fmaf(w, y, acc[q][1]);
          acc[q][2] = fmaf(w, z, acc[q][2]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (j0 + q < J) {
      SF_IM(out, (j0 + q) * 3 + 0, Bp, b) = acc[q][0];
      SF_IM(out, (j0 + q) * 3 + 1, Bp, b) = acc[q][1];
      SF_IM(out, (j0 + q) * 3 + 2, Bp, b) = acc[q][2];
    }
}

// ---------------------------------------------------------------------------------------
// k_stats: per-(segment, instance) sufficient statistics of the rotation stage
// (pt/bodyfitter.py:235-280), accumulated about provisional centres so nothing cancels:
//   M  = sum w (t - ct0)(a - ca0)^T   (9)     st = sum w (t - ct0)   (3)
//   sa = sum w (a - ca0)              (3)     W  = sum w             (1)
// REF: 0 = template mesh (first fit, B_ref = 1), 1 = reference vertices skinned on the fly
// from (v_posed, betas, per-joint [R|t]), 2 = explicit reference vertices a^T.
// One warp per (instance group, segment); lane = instance.
// ---------------------------------------------------------------------------------------
struct StatsArgs {
  const float* tT;        // [3V][Bp] centred targets (internal order)
  const float* vwT;       // [V][Bp] or null
  const float* ct0;       // [3J][Bp] provisional target-side centres (target joints)
  const float* ca0;       // [3J][Bp] provisional reference-side centres (REF != 0)
  const float* ca0_const; // (J,3) for REF == 0
  const float* vposedT;   // [3V][Bp] (REF == 1)
  const float* beta;      // [NS][Bp] (REF == 1)
  const float* skin;      // [12J][Bp] (REF == 1): rows j*12 + (0..8 R, 9..11 t)
  const float* aT_in;     // [3V][Bp] (REF == 2)
  float* aT_out;          // optional [3V][Bp] store of the reference vertices (REF == 1)
  float* partials;        // [n_segments][16][Bp]
  const float* template_mesh;  // (V,3)
  const float* shapedirs;      // (V,3,NS)
  const int32_t* skin_idx;
  const float* skin_w;
  const int32_t* order;
  const int32_t* seg_start;
  const int32_t* seg_part;
  const int32_t* part_flags;
  int n_segments, Bp, ns, skin_k, all_segments;
};

template <int REF, bool WEIGHTED>
__global__ void __launch_bounds__(128) k_stats(const StatsArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int seg = warp % a.n_segments, g = warp / a.n_segments;
  if (g * 32 >= a.Bp) return;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  const int part = a.seg_part[seg];
  const bool stat = (a.part_flags[part] & 1) != 0;
  if (!stat && !(REF == 1 && a.aT_out != nullptr && a.all_segments)) return;
  float ct[3], ca[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ct[c] = SF_IM(a.ct0, part * 3 + c, Bp, b);
    ca[c] = (REF == 0) ? __ldg(a.ca0_const + part * 3 + c) : SF_IM(a.ca0, part * 3 + c, Bp, b);
  }
  float beta[SMPLFIT_MAX_UNKNOWNS];
  if (REF == 1) {
#pragma unroll
    for (int s = 0; s < SMPLFIT_MAX_UNKNOWNS; ++s) beta[s] = (s < a.ns) ? SF_IM(a.beta, s, Bp, b) : 0.f;
  }
  float M[9], st[3], sa[3], W = 0.f;
#pragma unroll
  for (int e = 0; e < 9; ++e) M[e] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) st[c] = sa[c] = 0.f;
  const int i0 = a.seg_start[seg], i1 = a.seg_start[seg + 1];
  for (int i = i0; i < i1; ++i) {
    const int v = a.order[i];
    float ref[3];
    if (REF == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) ref[c] = __ldg(a.template_mesh + v * 3 + c);
    } else if (REF == 2) {
#pragma unroll
      for (int c = 0; c < 3; ++c) ref[c] = SF_IM(a.aT_in, i * 3 + c, Bp, b);
    } else {
      float vs[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float x = SF_IM(a.vposedT, i * 3 + c, Bp, b);
        const float* sd = a.shapedirs + ((size_t)v * 3 + c) * a.ns;
#pragma unroll
        for (int s = 0; s < SMPLFIT_MAX_UNKNOWNS; ++s)
          if (s < a.ns) x = fmaf(__ldg(sd + s), beta[s], x);
        vs[c] = x;
      }
      ref[0] = ref[1] = ref[2] = 0.f;
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
          ref[c] = fmaf(w, y, ref[c]);
        }
      }
      if (a.aT_out != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) SF_IM(a.aT_out, i * 3 + c, Bp, b) = ref[c];
      }
    }
    if (stat) {
      const float w = WEIGHTED ? SF_IM(a.vwT, i, Bp, b) : 1.f;
      float dt[3], wa[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dt[c] = SF_IM(a.tT, i * 3 + c, Bp, b) - ct[c];
        wa[c] = w * (ref[c] - ca[c]);
        st[c] = fmaf(w, dt[c], st[c]);
        sa[c] += wa[c];
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) M[r * 3 + c] = fmaf(dt[r], wa[c], M[r * 3 + c]);
      W += w;
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

// ---------------------------------------------------------------------------------------
// k_shape_pass<NS>: per-(chunk, instance) normal-equation pieces of the shape solve
// (pt/bodyfitter.py:999-1048):  per vertex  Rb = sum_k w_k R_jk,  Tb = sum_k w_k T_jk,
//   pos = Rb v_posed + Tb[:,0],   jac[:,s] = Rb S_v[:,s] + Tb[:,1+s],   b = t - pos,
// accumulating  G = sum w jac^T jac (upper triangle), r = sum w jac^T b, SA = sum w jac,
// Sb = sum w b, W = sum w  in registers.  The per-joint [R | T_ext] rows of the CTA's 32
// instances are staged once in shared memory ([row][lane], conflict-free); warps split the
// CTA's chunks.  Falls back to global/L1 reads when the rows do not fit (SMEM_RT = false).
// ---------------------------------------------------------------------------------------
template <int NS>
struct ShapeAcc {
  static constexpr int NG = NS * (NS + 1) / 2;
  static constexpr int N = NG + NS + 3 + 3 * NS + 1;  // G, r, Sb, SA, W
  static constexpr int N_UNWEIGHTED = NG + NS + 3;   // G, r, Sb (SA in closed form, W = V)
};

struct ShapeArgs {
  const float* tT;       // [3V][Bp]
  const float* vwT;      // [V][Bp] or null (shape-stage weights)
  const float* vposedT;  // [3V][Bp]
  const float* RT;       // [J * RW][Bp], RW = 12 + 3 NS: 9 R, then T_ext[c][0..NS]
  const float* shapedirs;  // (V,3,NS)
  const int32_t* skin_idx;
  const float* skin_w;
  const int32_t* order;
  float* partials;       // [n_chunks][ShapeAcc<NS>::N][Bp]
  const float* rec;      // [V][Rec<NS>::LEN] packed per-vertex records (internal order)
  const float* RT4;      // [J * Quad<NS>::NQ][Bp] float4 quad layout of the per-joint rows (k_shape_pass_v2)
  int V, J, Bp, skin_k, chunk_len, n_chunks, chunks_per_cta;
};

template <int NS, bool WEIGHTED, bool SMEM_RT>
__global__ void __launch_bounds__(256) k_shape_pass(const ShapeArgs a) {
  extern __shared__ __align__(16) float s_rt[];  // [J*RW][32] staged rows, later reused for the CTA reduction
  constexpr int RW = 12 + 3 * NS;
  constexpr int TW = 3 * (1 + NS);
  constexpr int NACC = ShapeAcc<NS>::N;
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  if (SMEM_RT) {
    // stage the 32-instance slice of every [R | T_ext] row with 16-byte cp.async (row = 128 B)
    const int n16 = a.J * RW * 8;
    for (int q = threadIdx.x; q < n16; q += 256) {
      const int r = q >> 3, part = q & 7;
      const float* src = a.RT + (size_t)r * Bp + g * 32 + part * 4;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rt + r * 32 + part * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  const int chunk = blockIdx.x * 8 + warp;
  const bool active = chunk < a.n_chunks;
  float acc[NACC];  // G (upper triangle), r, SA, Sb, W
#pragma unroll
  for (int e = 0; e < NACC; ++e) acc[e] = 0.f;
  constexpr int OG = 0, OR = ShapeAcc<NS>::NG, OSB = OR + NS, OSA = OSB + 3, OW = OSA + 3 * NS;  // [G | r | Sb | SA | W]
  if (active) {
    const int i0 = chunk * a.chunk_len, i1 = min(a.V, i0 + a.chunk_len);
    for (int i = i0; i < i1; ++i) {
      const int v = __ldg(a.order + i);
      float Rb[9], Tb[TW];
#pragma unroll
      for (int e = 0; e < 9; ++e) Rb[e] = 0.f;
#pragma unroll
      for (int e = 0; e < TW; ++e) Tb[e] = 0.f;
      for (int k = 0; k < a.skin_k; ++k) {
        const int j = __ldg(a.skin_idx + v * a.skin_k + k);
        const float w = __ldg(a.skin_w + v * a.skin_k + k);
        if (SMEM_RT) {
          const float* p = s_rt + (size_t)(j * RW) * 32 + lane;
#pragma unroll
          for (int e = 0; e < 9; ++e) Rb[e] = fmaf(w, p[e * 32], Rb[e]);
#pragma unroll
          for (int e = 0; e < TW; ++e) Tb[e] = fmaf(w, p[(9 + e) * 32], Tb[e]);
        } else {
          const float* p = a.RT + (size_t)(j * RW) * Bp + b;
#pragma unroll
          for (int e = 0; e < 9; ++e) Rb[e] = fmaf(w, p[(size_t)e * Bp], Rb[e]);
#pragma unroll
          for (int e = 0; e < TW; ++e) Tb[e] = fmaf(w, p[(size_t)(9 + e) * Bp], Tb[e]);
        }
      }
      float vp[3], bv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) vp[c] = SF_IM(a.vposedT, i * 3 + c, Bp, b);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float pos = fmaf(Rb[c * 3], vp[0], fmaf(Rb[c * 3 + 1], vp[1], fmaf(Rb[c * 3 + 2], vp[2], Tb[c * (1 + NS)])));
        bv[c] = SF_IM(a.tT, i * 3 + c, Bp, b) - pos;
      }
      // jac[c][s] overwrites Tb[c][1+s]
      const float* sd = a.shapedirs + (size_t)v * 3 * NS;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float s0 = __ldg(sd + s), s1 = __ldg(sd + NS + s), s2 = __ldg(sd + 2 * NS + s);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          Tb[c * (1 + NS) + 1 + s] =
              fmaf(Rb[c * 3], s0, fmaf(Rb[c * 3 + 1], s1, fmaf(Rb[c * 3 + 2], s2, Tb[c * (1 + NS) + 1 + s])));
      }
      const float w = WEIGHTED ? SF_IM(a.vwT, i, Bp, b) : 1.f;
      if (WEIGHTED) acc[OW] += w;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float wb = WEIGHTED ? w * bv[c] : bv[c];
        acc[OSB + c] += wb;
        int e = OG;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float js = Tb[c * (1 + NS) + 1 + s];
          const float wj = WEIGHTED ? w * js : js;
          acc[OSA + c * NS + s] += wj;
          acc[OR + s] = fmaf(wj, bv[c], acc[OR + s]);
#pragma unroll
          for (int t = s; t < NS; ++t) {
            acc[e] = fmaf(wj, Tb[c * (1 + NS) + 1 + t], acc[e]);
            ++e;
          }
        }
      }
    }
    if (!WEIGHTED) acc[OW] = (float)(i1 - i0);
  }
  // ---- deterministic tree reduction over the CTA's 8 warps through shared memory ----
  float* red = s_rt;  // [4][NACC][32]
#pragma unroll 1
  for (int half = 4; half >= 1; half >>= 1) {
    __syncthreads();  // staged rows (or the previous round) are dead
    if (warp >= half && warp < 2 * half) {
      float* dst = red + (size_t)(warp - half) * NACC * 32 + lane;
#pragma unroll
      for (int e = 0; e < NACC; ++e) dst[e * 32] = acc[e];
    }
    __syncthreads();
    if (warp < half) {
      const float* src = red + (size_t)warp * NACC * 32 + lane;
#pragma unroll
      for (int e = 0; e < NACC; ++e) acc[e] += src[e * 32];
    }
  }
  if (warp == 0) {
    float* out = a.partials + (size_t)blockIdx.x * NACC * Bp + b;
#pragma unroll
    for (int e = 0; e < NACC; ++e) out[(size_t)e * Bp] = acc[e];
  }
}

// ---------------------------------------------------------------------------------------
// Packed per-vertex record (internal order), LEN floats, 16-byte aligned:
//   [0..3] skin weights, descending (slot 0 = the vertex's dominant joint), zero padded
//   [4..7] joint ids of those slots (int bits)
//   [8 .. 8+3NS) shapedirs[c][s] (with the kid column when enabled)
// One record = 2 cache lines read with warp-uniform 16-byte loads, no index indirection.
// ---------------------------------------------------------------------------------------
template <int NS>
struct Rec {
  static constexpr int NSP = (NS + 1) / 2 * 2;            // shape columns padded to even (float2 pairs)
  static constexpr int LEN = (8 + 3 * NSP + 3) / 4 * 4;   // shapedirs[x][s] at 8 + x * NSP + s
};
// Quad layout of the per-joint rows for the packed-math shape pass: row order
//   0..8 R[c][x], 9 pad, 10..12 T[c][0], 13 pad, 14 + c*NSP + s -> T[c][1+s]; NQ = rows / 4 (padded);
// in memory [joint][quad][instance][4] so one instance's 4 rows are one 16-byte word.
template <int NS>
struct Quad {
  static constexpr int NSP = Rec<NS>::NSP;
  static constexpr int ROWS = (14 + 3 * NSP + 3) / 4 * 4;
  static constexpr int NQ = ROWS / 4;
};
__host__ __device__ inline int quad_rows(int ns) { const int nsp = (ns + 1) / 2 * 2; return (14 + 3 * nsp + 3) / 4 * 4; }
// row index inside a joint's quad block of T_ext[c][col] (col = 0 position, 1+s Jacobian) / R[e]
__host__ __device__ inline int quad_row_T(int ns, int c, int col) {
  const int nsp = (ns + 1) / 2 * 2;
  return col == 0 ? 10 + c : 14 + c * nsp + (col - 1);
}

// ---------------------------------------------------------------------------------------
// k_shape_pass_v2<NS>: packed-math shape pass (Blackwell FFMA2, two FP32 FMAs per issue slot).
// The pass is issue-bound, so the arithmetic is arranged in float2 pairs along the shape index:
//  * per-joint rows live in shared memory as float4 quads per instance (Quad<NS> order):
//    one LDS.128 + two FFMA2 blend four rows; the dominant joint's quads are cached in registers;
//  * jac[c][s,s+1] += (R[c][x], R[c][x]) * shapedirs[x][s,s+1]   (uniform 8-byte record loads);
//  * G[s][t,t+1] += (jac[c][s], jac[c][s]) * jac[c][t,t+1] for pairs t >= s & ~1
//    (one extra lower-triangle entry per odd row, dropped when the partials are written).
// Same outputs and partial layout as k_shape_pass (without per-vertex weights SA = sum_v jac_v is not accumulated: it has the
// closed form sum_k (R_k D_k + n_k T_k[:,1:]) with D_k = sum_v w_vk S_v, n_k = sum_v w_vk, evaluated in k_gram_entries).
// ---------------------------------------------------------------------------------------
// 1-D bulk async copy global -> shared with mbarrier completion (TMA engine, UBLKCP)
__device__ __forceinline__ void sf_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void sf_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void sf_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void sf_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
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

constexpr int REC_SUB = 16;  // vertices per staged record sub-block

// Per-warp double-buffered staging of per-vertex records (contiguous in the internal order) with
// 1-D bulk async copies: sub-block k lives in buffer k & 1; sub-block k+1 is in flight while k is
// processed.  All calls are warp-uniform.
template <int REC>
struct RecStager {
  float* buf;      // [2][REC_SUB * REC]
  uint64_t* bar;   // [2]
  const float* src;
  int i0, i1;
  uint32_t phase;
  __device__ __forceinline__ void issue(int k, int lane) {  // stage sub-block k (if it exists)
    const int first = i0 + k * REC_SUB;
    __syncwarp();
    if (first < i1 && lane == 0) {
      const uint32_t bytes = (uint32_t)min(REC_SUB, i1 - first) * REC * 4;
      sf_mbar_expect_tx(bar + (k & 1), bytes);
      sf_bulk_g2s(buf + (size_t)(k & 1) * REC_SUB * REC, src + (size_t)first * REC, bytes, bar + (k & 1));
    }
  }
  __device__ __forceinline__ void wait(int k) {
    sf_mbar_wait(bar + (k & 1), (phase >> (k & 1)) & 1u);
    phase ^= 1u << (k & 1);
  }
  __device__ __forceinline__ const float* rec(int i) const {
    const int d = i - i0;
    return buf + (size_t)((d / REC_SUB) & 1) * REC_SUB * REC + (size_t)(d % REC_SUB) * REC;
  }
};

__device__ __forceinline__ float2 sf_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 sf_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

template <int NS>
struct PackedG {
  static constexpr int H = Rec<NS>::NSP / 2;  // pairs per row
  __host__ __device__ static constexpr int row_off(int s) {  // first pair slot of row s
    int o = 0;
    for (int r = 0; r < s; ++r) o += H - r / 2;
    return o;
  }
  static constexpr int NPAIRS = row_off(NS);
};

template <int NS, bool WEIGHTED>
__global__ void __launch_bounds__(256, 1) k_shape_pass_v2(const ShapeArgs a) {
  extern __shared__ __align__(16) float s_rt[];  // [J][NQ][32] float4
  constexpr int NSP = Rec<NS>::NSP, H = NSP / 2;
  constexpr int NQ = Quad<NS>::NQ;
  constexpr int REC = Rec<NS>::LEN;
  constexpr int NP = PackedG<NS>::NPAIRS;
  constexpr int NRED = 2 * NP + 2 * H + 3 + (WEIGHTED ? 6 * H + 1 : 0);  // floats tree-reduced per lane
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  if (lane == 0) {
    sf_mbar_init(reinterpret_cast<uint64_t*>(s_rt + (size_t)a.J * NQ * 128 + (size_t)8 * 2 * REC_SUB * REC) + 2 * warp, 1);
    sf_mbar_init(reinterpret_cast<uint64_t*>(s_rt + (size_t)a.J * NQ * 128 + (size_t)8 * 2 * REC_SUB * REC) + 2 * warp + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  {
    // RT4 is [J*NQ][Bp] float4: the CTA's slice of each (joint, quad) row is 32 x 16 B contiguous
    const int n = a.J * NQ * 32;
    const float4* src = reinterpret_cast<const float4*>(a.RT4);
    for (int q = threadIdx.x; q < n; q += 256) {
      const int r = q >> 5, l = q & 31;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rt + (size_t)q * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)r * Bp + g * 32 + l) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  const float4* sq = reinterpret_cast<const float4*>(s_rt);
  RecStager<REC> rs;
  rs.buf = s_rt + (size_t)a.J * NQ * 128 + (size_t)warp * 2 * REC_SUB * REC;
  rs.bar = reinterpret_cast<uint64_t*>(s_rt + (size_t)a.J * NQ * 128 + (size_t)8 * 2 * REC_SUB * REC) + 2 * warp;
  rs.src = a.rec;
  rs.phase = 0;
  const int chunk = blockIdx.x * 8 + warp;
  const bool active = chunk < a.n_chunks;
  float2 G2[NP], r2[H], SA2[WEIGHTED ? 3 * H : 1];
  float Sb[3], Wsum = 0.f;
#pragma unroll
  for (int e = 0; e < NP; ++e) G2[e] = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < H; ++e) r2[e] = make_float2(0.f, 0.f);
  if (WEIGHTED) {
#pragma unroll
    for (int e = 0; e < 3 * H; ++e) SA2[e] = make_float2(0.f, 0.f);
  }
  Sb[0] = Sb[1] = Sb[2] = 0.f;
  if (active) {
    const int i0 = chunk * a.chunk_len, i1 = min(a.V, i0 + a.chunk_len);
    rs.i0 = i0;
    rs.i1 = i1;
    rs.issue(0, lane);
    rs.issue(1, lane);
    rs.wait(0);
    float4 nw = *reinterpret_cast<const float4*>(rs.rec(i0));
    int4 nj = *reinterpret_cast<const int4*>(rs.rec(i0) + 4);
    float nt[3], nvp[3], nvw = 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      nt[c] = SF_IM(a.tT, i0 * 3 + c, Bp, b);
      nvp[c] = SF_IM(a.vposedT, i0 * 3 + c, Bp, b);
    }
    if (WEIGHTED) nvw = SF_IM(a.vwT, i0, Bp, b);
    float4 Cq[NQ];
    int cj = -1;
    for (int i = i0; i < i1; ++i) {
      const float4 w4 = nw;
      const int4 j4 = nj;
      float t[3], vp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t[c] = nt[c];
        vp[c] = nvp[c];
      }
      const float wv = nvw;
      if ((i - i0) % REC_SUB == 0 && i > i0) rs.issue((i - i0) / REC_SUB + 1, lane);  // buffer of the previous sub-block is free
      const float2* sd2 = reinterpret_cast<const float2*>(rs.rec(i) + 8);  // shapedirs[x][sp] pairs (shared memory)
      if (i + 1 < i1) {
        if ((i + 1 - i0) % REC_SUB == 0) rs.wait((i + 1 - i0) / REC_SUB);
        const float* recn = rs.rec(i + 1);
        nw = *reinterpret_cast<const float4*>(recn);
        nj = *reinterpret_cast<const int4*>(recn + 4);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          nt[c] = SF_IM(a.tT, (i + 1) * 3 + c, Bp, b);
          nvp[c] = SF_IM(a.vposedT, (i + 1) * 3 + c, Bp, b);
        }
        if (WEIGHTED) nvw = SF_IM(a.vwT, i + 1, Bp, b);
      }
      if (j4.x != cj) {
        cj = j4.x;
#pragma unroll
        for (int q = 0; q < NQ; ++q) Cq[q] = sq[(size_t)(cj * NQ + q) * 32 + lane];
      }
      // blended rows as pairs: B2[2q] = rows (4q, 4q+1), B2[2q+1] = rows (4q+2, 4q+3)
      float2 B2[2 * NQ];
      {
        const float2 ww = make_float2(w4.x, w4.x);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          B2[2 * q] = sf_mul2(ww, make_float2(Cq[q].x, Cq[q].y));
          B2[2 * q + 1] = sf_mul2(ww, make_float2(Cq[q].z, Cq[q].w));
        }
      }
      const float wk[3] = {w4.y, w4.z, w4.w};
      const int jk[3] = {j4.y, j4.z, j4.w};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (wk[k] != 0.f) {
          const float2 ww = make_float2(wk[k], wk[k]);
          const float4* p = sq + (size_t)(jk[k] * NQ) * 32 + lane;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const float4 v = p[q * 32];
            B2[2 * q] = sf_fma2(ww, make_float2(v.x, v.y), B2[2 * q]);
            B2[2 * q + 1] = sf_fma2(ww, make_float2(v.z, v.w), B2[2 * q + 1]);
          }
        }
      }
      // rows: R[c][x] = row c*3+x; T0[c] = row 10+c; jac pairs start at pair index 7
      auto rowv = [&](int r) -> float { return (r & 1) ? B2[r >> 1].y : B2[r >> 1].x; };
      float bv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float pos = fmaf(rowv(c * 3), vp[0], fmaf(rowv(c * 3 + 1), vp[1], fmaf(rowv(c * 3 + 2), vp[2], rowv(10 + c))));
        bv[c] = t[c] - pos;
      }
      float2* J2 = B2 + 7;  // J2[c * H + sp] = (jac[c][2sp], jac[c][2sp+1])
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        float2 S2[H];
#pragma unroll
        for (int sp = 0; sp < H; ++sp) S2[sp] = sd2[x * H + sp];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float r = rowv(c * 3 + x);
          const float2 rr = make_float2(r, r);
#pragma unroll
          for (int sp = 0; sp < H; ++sp) J2[c * H + sp] = sf_fma2(rr, S2[sp], J2[c * H + sp]);
        }
      }
      if (WEIGHTED) Wsum += wv;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float wb = WEIGHTED ? wv * bv[c] : bv[c];
        Sb[c] += wb;
        const float2 bb = make_float2(wb, wb);
        if (WEIGHTED) {
          const float2 w2 = make_float2(wv, wv);
#pragma unroll
          for (int sp = 0; sp < H; ++sp) SA2[c * H + sp] = sf_fma2(w2, J2[c * H + sp], SA2[c * H + sp]);
        }
#pragma unroll
        for (int sp = 0; sp < H; ++sp) r2[sp] = sf_fma2(J2[c * H + sp], bb, r2[sp]);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float js = (s & 1) ? J2[c * H + (s >> 1)].y : J2[c * H + (s >> 1)].x;
          const float wj = WEIGHTED ? wv * js : js;
          const float2 jj = make_float2(wj, wj);
#pragma unroll
          for (int q = s / 2; q < H; ++q) {
            G2[PackedG<NS>::row_off(s) + q - s / 2] = sf_fma2(jj, J2[c * H + q], G2[PackedG<NS>::row_off(s) + q - s / 2]);
          }
        }
      }
    }
  }
  // ---- tree reduction over the 8 warps (float view of the accumulators) ----
  float* red = s_rt;
  auto fold = [&](bool write, int slot) {
    float* base = red + (size_t)slot * NRED * 32 + lane;
    int o = 0;
#pragma unroll
    for (int e = 0; e < NP; ++e) {
      if (write) { base[(o) * 32] = G2[e].x; base[(o + 1) * 32] = G2[e].y; }
      else { G2[e].x += base[(o) * 32]; G2[e].y += base[(o + 1) * 32]; }
      o += 2;
    }
#pragma unroll
    for (int e = 0; e < H; ++e) {
      if (write) { base[(o) * 32] = r2[e].x; base[(o + 1) * 32] = r2[e].y; }
      else { r2[e].x += base[(o) * 32]; r2[e].y += base[(o + 1) * 32]; }
      o += 2;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (write) base[o * 32] = Sb[c]; else Sb[c] += base[o * 32];
      ++o;
    }
    if (WEIGHTED) {
#pragma unroll
      for (int e = 0; e < 3 * H; ++e) {
        if (write) { base[(o) * 32] = SA2[e].x; base[(o + 1) * 32] = SA2[e].y; }
        else { SA2[e].x += base[(o) * 32]; SA2[e].y += base[(o + 1) * 32]; }
        o += 2;
      }
      if (write) base[o * 32] = Wsum; else Wsum += base[o * 32];
    }
  };
#pragma unroll 1
  for (int half = 4; half >= 1; half >>= 1) {
    __syncthreads();
    if (warp >= half && warp < 2 * half) fold(true, warp - half);
    __syncthreads();
    if (warp < half) fold(false, warp);
  }
  if (warp == 0) {
    // standard partial layout [G upper | r | Sb | SA | W]
    float* out = a.partials + (size_t)blockIdx.x * ShapeAcc<NS>::N * Bp + b;
    int o = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int t2 = s; t2 < NS; ++t2) {
        const float2 gp = G2[PackedG<NS>::row_off(s) + t2 / 2 - s / 2];
        out[(size_t)(o++) * Bp] = (t2 & 1) ? gp.y : gp.x;
      }
#pragma unroll
    for (int s = 0; s < NS; ++s) out[(size_t)(o++) * Bp] = (s & 1) ? r2[s >> 1].y : r2[s >> 1].x;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(size_t)(o++) * Bp] = Sb[c];
    if (WEIGHTED) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int s = 0; s < NS; ++s) out[(size_t)(o++) * Bp] = (s & 1) ? SA2[c * H + (s >> 1)].y : SA2[c * H + (s >> 1)].x;
      out[(size_t)o * Bp] = Wsum;
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_shape_pass_v3<NS, WEIGHTED, ROWSEL>: packed-math shape pass for models whose per-joint rows
// of 32 instances do not fit in shared memory (SMPL-X: 55 joints) and / or many unknowns.
//  * joint-subset staging: vertices are grouped by body part, so the CTA's vertices reference only
//    a handful of joints; the CTA discovers them (bit mask over its records), stages just those
//    joints' rows (cp.async) and reads any overflow joint straight from L2;
//  * coordinate-major rows ("C layout", CLay<NS>): per joint and coordinate c the rows
//    [R[c][0..2], T0[c], T[c][1..]] form QC quads, and the vertex loop handles one coordinate at a
//    time (pos_c, jac_c, its Gram / rhs contribution), so only ~2 QC pairs are live at once;
//  * ROWSEL splits the Gram rows over two launches when the accumulators would not fit in
//    registers (NS >= 14): 1 = rows < SH plus r, Sb (SA, W when weighted), 2 = rows >= SH only.
// Same partial layout as the other shape-pass kernels.
// ---------------------------------------------------------------------------------------
template <int NS>
struct CLay {
  static constexpr int NSP4 = (NS + 3) / 4 * 4;
  static constexpr int RC = 4 + NSP4;   // rows per (joint, coordinate)
  static constexpr int QC = RC / 4;     // quads per (joint, coordinate)
  static constexpr int JQ = 3 * QC;     // quads per joint
  static constexpr int H = Rec<NS>::NSP / 2;  // Jacobian pairs actually used
  __host__ __device__ static constexpr int split_row() {  // first row of the second half
    int s = 0;
    while (s < NS && 2 * PackedG<NS>::row_off(s) < PackedG<NS>::NPAIRS) ++s;
    return s;
  }
};
__host__ __device__ inline int clay_rows_per_joint(int ns) { return 3 * (4 + (ns + 3) / 4 * 4); }

struct ShapeV3Extra {
  int cap_joints;  // joints that fit in the staging area
};

template <int NS, bool WEIGHTED, int ROWSEL>
__global__ void __launch_bounds__(256, 1) k_shape_pass_v3(const ShapeArgs a, const ShapeV3Extra x3) {
  extern __shared__ __align__(16) float s_rt[];
  using L = CLay<NS>;
  constexpr int H = L::H, QC = L::QC, JQ = L::JQ;
  constexpr int REC = Rec<NS>::LEN;
  constexpr int SH = L::split_row();
  constexpr int P0 = (ROWSEL == 2) ? PackedG<NS>::row_off(SH) : 0;
  constexpr int P1 = (ROWSEL == 1) ? PackedG<NS>::row_off(SH) : PackedG<NS>::NPAIRS;
  constexpr int NPS = P1 - P0;  // Gram pairs accumulated by this launch
  constexpr bool AUX = (ROWSEL != 2);  // r, Sb (SA, W) accumulated here
  constexpr int NRED = 2 * NPS + (AUX ? 2 * H + 3 + (WEIGHTED ? 6 * H + 1 : 0) : 0);
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  const int cap = x3.cap_joints;
  float4* sq = reinterpret_cast<float4*>(s_rt);                                   // [cap][JQ][32]
  float* s_after = s_rt + (size_t)cap * JQ * 128;
  RecStager<REC> rs;
  rs.buf = s_after + (size_t)warp * 2 * REC_SUB * REC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_after + (size_t)8 * 2 * REC_SUB * REC);
  rs.bar = bars + 2 * warp;
  rs.src = a.rec;
  rs.phase = 0;
  int* s_slot = reinterpret_cast<int*>(bars + 16);  // [64]
  unsigned* s_mask = reinterpret_cast<unsigned*>(s_slot + 64);  // [2]
  if (threadIdx.x < 2) s_mask[threadIdx.x] = 0u;
  if (lane == 0) {
    sf_mbar_init(rs.bar, 1);
    sf_mbar_init(rs.bar + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int chunk = blockIdx.x * 8 + warp;
  const bool active = chunk < a.n_chunks;
  const int i0 = chunk * a.chunk_len, i1 = active ? min(a.V, i0 + a.chunk_len) : i0;
  {
    // joints referenced (with non-zero weight) by this warp's vertices
    unsigned m0 = 0u, m1 = 0u;
    for (int i = i0 + lane; i < i1; i += 32) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.rec + (size_t)i * REC));
      const int4 j4 = __ldg(reinterpret_cast<const int4*>(a.rec + (size_t)i * REC + 4));
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
      const int jv[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (wv[k] != 0.f) {
          if (jv[k] < 32) m0 |= 1u << jv[k]; else m1 |= 1u << (jv[k] - 32);
        }
    }
    m0 = __reduce_or_sync(0xffffffffu, m0);
    m1 = __reduce_or_sync(0xffffffffu, m1);
    if (lane == 0) {
      atomicOr(&s_mask[0], m0);
      atomicOr(&s_mask[1], m1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int j = threadIdx.x;
    const unsigned lo = s_mask[0], hi = s_mask[1];
    const bool used = j < 32 ? ((lo >> j) & 1u) : ((hi >> (j - 32)) & 1u);
    const int before = j < 32 ? __popc(lo & ((1u << j) - 1u)) : __popc(lo) + __popc(hi & ((1u << (j - 32)) - 1u));
    s_slot[j] = (used && before < cap) ? before : -1;
  }
  __syncthreads();
  {
    const float4* src = reinterpret_cast<const float4*>(a.RT4);
    for (int j = 0; j < a.J; ++j) {
      const int slot = s_slot[j];
      if (slot < 0) continue;
      for (int q = threadIdx.x; q < JQ * 32; q += 256) {
        const int r = q >> 5, l = q & 31;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sq + (size_t)(slot * JQ + r) * 32 + l);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)(j * JQ + r) * Bp + g * 32 + l) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  float2 G2[NPS > 0 ? NPS : 1], r2[AUX ? H : 1], SA2[(AUX && WEIGHTED) ? 3 * H : 1];
  float Sb[3] = {0.f, 0.f, 0.f}, Wsum = 0.f;
#pragma unroll
  for (int e = 0; e < NPS; ++e) G2[e] = make_float2(0.f, 0.f);
  if (AUX) {
#pragma unroll
    for (int e = 0; e < H; ++e) r2[e] = make_float2(0.f, 0.f);
    if (WEIGHTED) {
#pragma unroll
      for (int e = 0; e < 3 * H; ++e) SA2[e] = make_float2(0.f, 0.f);
    }
  }
  if (active) {
    rs.i0 = i0;
    rs.i1 = i1;
    rs.issue(0, lane);
    rs.issue(1, lane);
    rs.wait(0);
    float nt[3], nvp[3], nvw = 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      nt[c] = SF_IM(a.tT, i0 * 3 + c, Bp, b);
      nvp[c] = SF_IM(a.vposedT, i0 * 3 + c, Bp, b);
    }
    if (WEIGHTED) nvw = SF_IM(a.vwT, i0, Bp, b);
    const float4* gsrc = reinterpret_cast<const float4*>(a.RT4);
    for (int i = i0; i < i1; ++i) {
      if ((i - i0) % REC_SUB == 0 && i > i0) {
        rs.wait((i - i0) / REC_SUB);
        rs.issue((i - i0) / REC_SUB + 1, lane);
      }
      const float* rec = rs.rec(i);
      const float4 w4 = *reinterpret_cast<const float4*>(rec);
      const int4 j4 = *reinterpret_cast<const int4*>(rec + 4);
      const float2* sd2 = reinterpret_cast<const float2*>(rec + 8);
      float t[3], vp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t[c] = nt[c];
        vp[c] = nvp[c];
      }
      const float wv = nvw;
      if (i + 1 < i1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          nt[c] = SF_IM(a.tT, (i + 1) * 3 + c, Bp, b);
          nvp[c] = SF_IM(a.vposedT, (i + 1) * 3 + c, Bp, b);
        }
        if (WEIGHTED) nvw = SF_IM(a.vwT, i + 1, Bp, b);
      }
      // row sources of the (up to) four joints: staged copy or global (overflow)
      const float wk[4] = {w4.x, w4.y, w4.z, w4.w};
      const int jk[4] = {j4.x, j4.y, j4.z, j4.w};
      const float4* jp[4];
      int jstride[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int slot = s_slot[jk[k]];
        jp[k] = (slot >= 0) ? (sq + (size_t)slot * JQ * 32 + lane) : (gsrc + (size_t)jk[k] * JQ * Bp + b);
        jstride[k] = (slot >= 0) ? 32 : Bp;
      }
      if (AUX && WEIGHTED) Wsum += wv;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float2 B2[2 * QC];
#pragma unroll
        for (int q = 0; q < 2 * QC; ++q) B2[q] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (wk[k] != 0.f) {
            const float2 ww = make_float2(wk[k], wk[k]);
#pragma unroll
            for (int q = 0; q < QC; ++q) {
              const float4 v = jp[k][(size_t)(c * QC + q) * jstride[k]];
              B2[2 * q] = sf_fma2(ww, make_float2(v.x, v.y), B2[2 * q]);
              B2[2 * q + 1] = sf_fma2(ww, make_float2(v.z, v.w), B2[2 * q + 1]);
            }
          }
        }
        // rows: B2[0] = (R[c][0], R[c][1]), B2[1] = (R[c][2], T0[c]), B2[2 + sp] = jac pair sp
        const float pos = fmaf(B2[0].x, vp[0], fmaf(B2[0].y, vp[1], fmaf(B2[1].x, vp[2], B2[1].y)));
        const float bv = t[c] - pos;
        float2* J2 = B2 + 2;
        const float rr3[3] = {B2[0].x, B2[0].y, B2[1].x};
#pragma unroll
        for (int xx = 0; xx < 3; ++xx) {
          const float2 rr = make_float2(rr3[xx], rr3[xx]);
#pragma unroll
          for (int sp = 0; sp < H; ++sp) J2[sp] = sf_fma2(rr, sd2[xx * H + sp], J2[sp]);
        }
        if (AUX) {
          const float wb = WEIGHTED ? wv * bv : bv;
          Sb[c] += wb;
          const float2 bb = make_float2(wb, wb);
          if (WEIGHTED) {
            const float2 w2 = make_float2(wv, wv);
#pragma unroll
            for (int sp = 0; sp < H; ++sp) SA2[c * H + sp] = sf_fma2(w2, J2[sp], SA2[c * H + sp]);
          }
#pragma unroll
          for (int sp = 0; sp < H; ++sp) r2[sp] = sf_fma2(J2[sp], bb, r2[sp]);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if ((ROWSEL == 1 && s >= SH) || (ROWSEL == 2 && s < SH)) continue;
          const float js = (s & 1) ? J2[s >> 1].y : J2[s >> 1].x;
          const float wj = WEIGHTED ? wv * js : js;
          const float2 jj = make_float2(wj, wj);
#pragma unroll
          for (int q = s / 2; q < H; ++q)
            G2[PackedG<NS>::row_off(s) - P0 + q - s / 2] = sf_fma2(jj, J2[q], G2[PackedG<NS>::row_off(s) - P0 + q - s / 2]);
        }
      }
    }
  }
  // ---- tree reduction over the 8 warps ----
  float* red = s_rt;
  auto fold = [&](bool write, int slot) {
    float* base = red + (size_t)slot * NRED * 32 + lane;
    int o = 0;
#pragma unroll
    for (int e = 0; e < NPS; ++e) {
      if (write) { base[(o) * 32] = G2[e].x; base[(o + 1) * 32] = G2[e].y; }
      else { G2[e].x += base[(o) * 32]; G2[e].y += base[(o + 1) * 32]; }
      o += 2;
    }
    if (AUX) {
#pragma unroll
      for (int e = 0; e < H; ++e) {
        if (write) { base[(o) * 32] = r2[e].x; base[(o + 1) * 32] = r2[e].y; }
        else { r2[e].x += base[(o) * 32]; r2[e].y += base[(o + 1) * 32]; }
        o += 2;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (write) base[o * 32] = Sb[c]; else Sb[c] += base[o * 32];
        ++o;
      }
      if (WEIGHTED) {
#pragma unroll
        for (int e = 0; e < 3 * H; ++e) {
          if (write) { base[(o) * 32] = SA2[e].x; base[(o + 1) * 32] = SA2[e].y; }
          else { SA2[e].x += base[(o) * 32]; SA2[e].y += base[(o + 1) * 32]; }
          o += 2;
        }
        if (write) base[o * 32] = Wsum; else Wsum += base[o * 32];
      }
    }
  };
#pragma unroll 1
  for (int half = 4; half >= 1; half >>= 1) {
    __syncthreads();
    if (warp >= half && warp < 2 * half) fold(true, warp - half);
    __syncthreads();
    if (warp < half) fold(false, warp);
  }
  if (warp == 0) {
    float* out = a.partials + (size_t)blockIdx.x * ShapeAcc<NS>::N * Bp + b;
    int o = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int t2 = s; t2 < NS; ++t2) {
        if (!((ROWSEL == 1 && s >= SH) || (ROWSEL == 2 && s < SH))) {
          const float2 gp = G2[PackedG<NS>::row_off(s) - P0 + t2 / 2 - s / 2];
          out[(size_t)o * Bp] = (t2 & 1) ? gp.y : gp.x;
        }
        ++o;
      }
    if (AUX) {
#pragma unroll
      for (int s = 0; s < NS; ++s) out[(size_t)(o++) * Bp] = (s & 1) ? r2[s >> 1].y : r2[s >> 1].x;
#pragma unroll
      for (int c = 0; c < 3; ++c) out[(size_t)(o++) * Bp] = Sb[c];
      if (WEIGHTED) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int s = 0; s < NS; ++s) out[(size_t)(o++) * Bp] = (s & 1) ? SA2[c * H + (s >> 1)].y : SA2[c * H + (s >> 1)].x;
        out[(size_t)o * Bp] = Wsum;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_stats_rec<NS, REF, WEIGHTED>: the statistics pass in the same style (records, one-ahead
// prefetch, per-joint skinning transforms of the CTA's 32 instances staged in shared memory,
// dominant joint cached in registers).  8 warps per CTA, SEGS_PER_WARP segments per warp.
// ---------------------------------------------------------------------------------------
struct StatsRecArgs {
  const float* tT;
  const float* vwT;
  const float* ct0;
  const float* ca0;
  const float* ca0_const;
  const float* vposedT;
  const float* beta;
  const float* skin;          // [12J][Bp]
  const float* aT_in;
  float* aT_out;
  float* partials;
  const float* rec;           // [V][Rec<NS>::LEN]
  const float* template_fit;  // (V,3) template mesh, internal order (REF == 0)
  const int32_t* seg_start;
  const int32_t* seg_part;
  const int32_t* part_flags;
  int n_segments, Bp, J, all_segments, segs_per_warp;
};

template <int NS, int REF, bool WEIGHTED>
__global__ void __launch_bounds__(256) k_stats_rec(const StatsRecArgs a) {
  extern __shared__ __align__(16) float s_skin[];  // [12J][32], then per-warp record buffers + mbarriers (REF == 1)
  constexpr int REC = Rec<NS>::LEN;
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Bp = a.Bp;
  const int b = g * 32 + lane;
  RecStager<REC> rs;
  rs.buf = s_skin + (size_t)a.J * 12 * 32 + (size_t)warp * 2 * REC_SUB * REC;
  rs.bar = reinterpret_cast<uint64_t*>(s_skin + (size_t)a.J * 12 * 32 + (size_t)8 * 2 * REC_SUB * REC) + 2 * warp;
  rs.src = a.rec;
  rs.phase = 0;
  if (REF == 1) {
    if (lane == 0) {
      sf_mbar_init(rs.bar, 1);
      sf_mbar_init(rs.bar + 1, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
  float beta[NS];
  if (REF == 1) {
#pragma unroll
    for (int s = 0; s < NS; ++s) beta[s] = SF_IM(a.beta, s, Bp, b);
  }
  for (int q = 0; q < a.segs_per_warp; ++q) {
    const int seg = (blockIdx.x * a.segs_per_warp + q) * 8 + warp;
    if (seg >= a.n_segments) break;
    const int part = a.seg_part[seg];
    const bool stat = (a.part_flags[part] & 1) != 0;
    if (!stat && !(REF == 1 && a.aT_out != nullptr && a.all_segments)) continue;
    const int i0 = a.seg_start[seg], i1 = a.seg_start[seg + 1];
    rs.i0 = i0;
    rs.i1 = i1;
    if (REF == 1) {
      rs.issue(0, lane);
      rs.issue(1, lane);
    }
    float ct[3], ca[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ct[c] = SF_IM(a.ct0, part * 3 + c, Bp, b);
      ca[c] = (REF == 0) ? __ldg(a.ca0_const + part * 3 + c) : SF_IM(a.ca0, part * 3 + c, Bp, b);
    }
    float M[9], st[3], sa[3], W = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) M[e] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) st[c] = sa[c] = 0.f;
    float nt[3], nx[3], nvw = 1.f;  // nx: v_posed (REF 1), explicit reference (REF 2), template (REF 0)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      nt[c] = SF_IM(a.tT, i0 * 3 + c, Bp, b);
      nx[c] = (REF == 0) ? __ldg(a.template_fit + i0 * 3 + c)
                         : SF_IM((REF == 1 ? a.vposedT : a.aT_in), i0 * 3 + c, Bp, b);
    }
    if (WEIGHTED) nvw = SF_IM(a.vwT, i0, Bp, b);
    float Sc[12];
    int cj = -1;
    for (int i = i0; i < i1; ++i) {
      if (REF == 1 && (i - i0) % REC_SUB == 0) {  // entering sub-block k (warp-uniform)
        const int k = (i - i0) / REC_SUB;
        rs.wait(k);
        if (k >= 1) rs.issue(k + 1, lane);  // the buffer of sub-block k-1 is free now
      }
      float t[3], x[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t[c] = nt[c];
        x[c] = nx[c];
      }
      const float wv = nvw;
      if (i + 1 < i1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          nt[c] = SF_IM(a.tT, (i + 1) * 3 + c, Bp, b);
          nx[c] = (REF == 0) ? __ldg(a.template_fit + (i + 1) * 3 + c)
                             : SF_IM((REF == 1 ? a.vposedT : a.aT_in), (i + 1) * 3 + c, Bp, b);
        }
        if (WEIGHTED) nvw = SF_IM(a.vwT, i + 1, Bp, b);
      }
      float ref[3];
      if (REF == 1) {
        const float* rec = rs.rec(i);  // shared-memory broadcasts
        const float4 w4 = *reinterpret_cast<const float4*>(rec);
        const int4 j4 = *reinterpret_cast<const int4*>(rec + 4);
        float vs[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float y = x[c];
#pragma unroll
          for (int s2 = 0; s2 < Rec<NS>::NSP; s2 += 2) {
            const float2 sv = *reinterpret_cast<const float2*>(rec + 8 + c * Rec<NS>::NSP + s2);
            y = fmaf(sv.x, beta[s2], y);
            if (s2 + 1 < NS) y = fmaf(sv.y, beta[s2 + 1], y);
          }
          vs[c] = y;
        }
        if (j4.x != cj) {
          cj = j4.x;
          const float* p = s_skin + (size_t)(cj * 12) * 32 + lane;
#pragma unroll
          for (int e = 0; e < 12; ++e) Sc[e] = p[e * 32];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
          ref[c] = w4.x * fmaf(Sc[c * 3], vs[0], fmaf(Sc[c * 3 + 1], vs[1], fmaf(Sc[c * 3 + 2], vs[2], Sc[9 + c])));
        const float wk[3] = {w4.y, w4.z, w4.w};
        const int jk[3] = {j4.y, j4.z, j4.w};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (wk[k] != 0.f) {
            const float* p = s_skin + (size_t)(jk[k] * 12) * 32 + lane;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float y = fmaf(p[(c * 3) * 32], vs[0], fmaf(p[(c * 3 + 1) * 32], vs[1], fmaf(p[(c * 3 + 2) * 32], vs[2], p[(9 + c) * 32])));
              ref[c] = fmaf(wk[k], y, ref[c]);
            }
          }
        }
        if (a.aT_out != nullptr) {
#pragma unroll
          for (int c = 0; c < 3; ++c) SF_IM(a.aT_out, i * 3 + c, Bp, b) = ref[c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) ref[c] = x[c];
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
// k_scale_pass<NS, MODE>: extra accumulations of the final shape solve when a scale factor is
// estimated (pt/bodyfitter.py:1171-1176): the design matrix gets one more column z, with
// z = -target (MODE 1, scale_target) or z = pos (MODE 2, scale_fit).  Per (chunk, instance):
//   Gz[s] = sum w z.jac[:,s],  Gzz = sum w z.z,  rz = sum w z.b,  SAz[c] = sum w z_c.
// Only used once per fit and only in scale mode, so it simply recomputes pos / jac with global
// reads of the per-joint rows (no shared-memory staging).  Layout of the partials: [Gz | Gzz | rz | SAz].
// ---------------------------------------------------------------------------------------
template <int NS, int MODE>
__global__ void __launch_bounds__(128) k_scale_pass(const ShapeArgs a) {
  constexpr int RW = 12 + 3 * NS;
  constexpr int TW = 3 * (1 + NS);
  constexpr int NZ = NS + 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int chunk = warp % a.n_chunks, g = warp / a.n_chunks;
  const int Bp = a.Bp;
  if (g * 32 >= Bp) return;
  const int b = g * 32 + lane;
  float Gz[NS], Gzz = 0.f, rz = 0.f, SAz[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < NS; ++s) Gz[s] = 0.f;
  const int i0 = chunk * a.chunk_len, i1 = min(a.V, i0 + a.chunk_len);
  for (int i = i0; i < i1; ++i) {
    const int v = __ldg(a.order + i);
    float Rb[9], Tb[TW];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rb[e] = 0.f;
#pragma unroll
    for (int e = 0; e < TW; ++e) Tb[e] = 0.f;
    for (int k = 0; k < a.skin_k; ++k) {
      const int j = __ldg(a.skin_idx + v * a.skin_k + k);
      const float w = __ldg(a.skin_w + v * a.skin_k + k);
      if (w == 0.f) continue;
      const float* p = a.RT + (size_t)(j * RW) * Bp + b;
#pragma unroll
      for (int e = 0; e < 9; ++e) Rb[e] = fmaf(w, p[(size_t)e * Bp], Rb[e]);
#pragma unroll
      for (int e = 0; e < TW; ++e) Tb[e] = fmaf(w, p[(size_t)(9 + e) * Bp], Tb[e]);
    }
    float vp[3], t[3], pos[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      vp[c] = SF_IM(a.vposedT, i * 3 + c, Bp, b);
      t[c] = SF_IM(a.tT, i * 3 + c, Bp, b);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      pos[c] = fmaf(Rb[c * 3], vp[0], fmaf(Rb[c * 3 + 1], vp[1], fmaf(Rb[c * 3 + 2], vp[2], Tb[c * (1 + NS)])));
    const float* sd = a.shapedirs + (size_t)v * 3 * NS;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float s0 = __ldg(sd + s), s1 = __ldg(sd + NS + s), s2 = __ldg(sd + 2 * NS + s);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        Tb[c * (1 + NS) + 1 + s] = fmaf(Rb[c * 3], s0, fmaf(Rb[c * 3 + 1], s1, fmaf(Rb[c * 3 + 2], s2, Tb[c * (1 + NS) + 1 + s])));
    }
    const float w = a.vwT ? SF_IM(a.vwT, i, Bp, b) : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float z = (MODE == 1) ? -t[c] : pos[c];
      const float wz = w * z;
      SAz[c] += wz;
      Gzz = fmaf(wz, z, Gzz);
      rz = fmaf(wz, t[c] - pos[c], rz);
#pragma unroll
      for (int s = 0; s < NS; ++s) Gz[s] = fmaf(wz, Tb[c * (1 + NS) + 1 + s], Gz[s]);
    }
  }
  float* out = a.partials + (size_t)chunk * NZ * Bp + b;
#pragma unroll
  for (int s = 0; s < NS; ++s) out[(size_t)s * Bp] = Gz[s];
  out[(size_t)NS * Bp] = Gzz;
  out[(size_t)(NS + 1) * Bp] = rz;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[(size_t)(NS + 2 + c) * Bp] = SAz[c];
}

}  // namespace sf
