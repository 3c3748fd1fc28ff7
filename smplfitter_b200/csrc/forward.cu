// BodyModel.forward (pt/bodymodel.py:121-307) and BodyConverter.convert_vertices
// (pt/bodyconverter.py:129-149) for sm_100a.
//
// forward = the fused tensor-core kernel of fwd_fused.cu (blend-shape GEMM with the skinning in its epilogue) whenever
// the model carries its constants (<= 4 skinning influences per vertex, <= 64 joints; any number of betas).
// Generic form for other models (this file):
//           (1) k_fwd_prep: one thread per instance -- Rodrigues / rotation chain, shaped rest
//               joints, joint positions (FK), per-joint skinning transforms, pose features;
//           (2) the pose-blend-shape contraction v_posed^T = v_template + posedirs . feat
//               (tensor-core path in vposed_tc.cu, FP32 SIMT fallback);
//           (3) k_fwd_skin: lane = instance; adds the shape/kid blend shapes, blends the
//               <= K skinning transforms per vertex and writes the caller's (B,V,3) layout
//               through a shared-memory transpose so both sides stay coalesced.
#include <stdlib.h>

#include "common.cuh"
#include "fit_kernels.cuh"
#include "fwd_fused.cuh"
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
  float* feat;   // [Bp][Kp]
  float* betaT;  // [S+1][Bp] (betas zero-padded to S, then kid)
  float* out_joints;        // (B,J,3)
  float* out_orientations;  // (B,J,3,3)
  int rot_mode, n_betas, nb, J, S, B, Bp, Kp;
};

__global__ void __launch_bounds__(32) k_fwd_prep(const FwdPrepArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.Bp) return;
  const int J = a.J, S = a.S, Bp = a.Bp;
  const bool live = b < a.B;
  float x[SMPLFIT_MAX_UNKNOWNS];
  const int nb = a.nb;  // <= SMPLFIT_MAX_UNKNOWNS (checked by the caller)
  for (int s = 0; s < S; ++s) {
    const float v = (live && a.betas != nullptr && s < nb) ? a.betas[(size_t)b * a.n_betas + s] : 0.f;
    if (s < SMPLFIT_MAX_UNKNOWNS) x[s] = v;
    SF_IM(a.betaT, s, Bp, b) = v;
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

// The same SpMM with the instance's input vertices staged in shared memory (one CTA per instance): the (B,V_in,3) array
// is read once with coalesced 16-byte loads, the ~3 gathers per output element hit shared memory, the output row is
// written coalesced.  Used when V_in * 12 bytes fit (SMPL -> SMPL-X: 82.7 KB).
__global__ void __launch_bounds__(512) k_csr_apply_smem(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                        const float* __restrict__ data, int v_out, int v_in,
                                                        const float* __restrict__ in, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_in[];  // [v_in][3]
  const size_t b = blockIdx.x;
  const float* src = in + b * (size_t)v_in * 3;
  const int n = v_in * 3;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) reinterpret_cast<float4*>(s_in)[i] = __ldg(s4 + i);
    for (int i = (n / 4) * 4 + threadIdx.x; i < n; i += blockDim.x) s_in[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_in[i] = __ldg(src + i);
  }
  __syncthreads();
  float* dst = out + b * (size_t)v_out * 3;
  // a thread per output vertex: the row's indices / weights are fetched once for its three coordinates
  for (int r = threadIdx.x; r < v_out; r += blockDim.x) {
    const int k0 = __ldg(indptr + r), k1 = __ldg(indptr + r + 1);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = k0; k < k1; ++k) {
      const float w = __ldg(data + k);
      const float* v = s_in + __ldg(indices + k) * 3;
      a0 = fmaf(w, v[0], a0);
      a1 = fmaf(w, v[1], a1);
      a2 = fmaf(w, v[2], a2);
    }
    dst[r * 3] = a0;
    dst[r * 3 + 1] = a1;
    dst[r * 3 + 2] = a2;
  }
}

struct FwdWs {
  float *vposedT, *feat, *skin, *betaT;
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
  w.betaT = c.take<float>((size_t)(m->num_betas + 1) * Bp);
  w.tc_scratch = c.take<char>(vposed_tc_scratch_bytes(m, (int)Bp));
  w.bytes = c.off + 256;
  return w;
}

}  // namespace sf

using namespace sf;

extern "C" size_t smplfit_forward_workspace_bytes(const smplfit_model_t* m, int64_t batch) {
  if (!m || batch <= 0) return 0;
  if (fwd_fused_available(m)) return fwd_fused_carve(nullptr, m, batch, true).bytes;
  return carve_fwd(nullptr, m, batch).bytes;
}

extern "C" int smplfit_forward(const smplfit_model_t* m, int64_t batch, int rot_mode, const float* rot,
                               const float* betas, int n_betas, const float* trans, const float* kid,
                               float* out_vertices, float* out_joints, float* out_orientations, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!m || !out_joints || !out_orientations) return fail(SMPLFIT_ERR_ARG, "missing required pointer");
  if (batch <= 0) return fail(SMPLFIT_ERR_ARG, "batch must be positive");
  if (m->num_joints > SMPLFIT_MAX_JOINTS) return fail(SMPLFIT_ERR_UNSUPPORTED, "num_joints > 64");
  if (rot_mode < 0 || rot_mode > 3 || (rot_mode != 3 && !rot)) return fail(SMPLFIT_ERR_ARG, "bad rotation input");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (fwd_fused_available(m)) {
    FwdFusedWs fw = fwd_fused_carve(workspace, m, batch, out_vertices != nullptr);
    if (out_vertices != nullptr && (!workspace || fw.bytes > workspace_bytes)) return fail(SMPLFIT_ERR_WORKSPACE, "workspace too small");
    if (int e = fwd_fused_run(m, (int)batch, rot_mode, rot, betas, n_betas, trans, kid, out_vertices, out_joints,
                              out_orientations, fw, st))
      return e;
    SF_CHECK_LAST();
    return SMPLFIT_OK;
  }
  // generic path: the betas actually used (pt/bodymodel.py:251: min(given, S)) must fit the register arrays
  const int nb_used = betas ? (n_betas < m->num_betas ? n_betas : m->num_betas) : 0;
  if (nb_used > SMPLFIT_MAX_UNKNOWNS)
    return fail(SMPLFIT_ERR_UNSUPPORTED, "more than 17 betas need the fused forward path (<= 4 skinning influences per vertex)");
  FwdWs w = carve_fwd(workspace, m, batch);
  if (!workspace || w.bytes > workspace_bytes) return fail(SMPLFIT_ERR_WORKSPACE, "workspace too small");
  const int B = (int)batch, Bp = roundup(B, 32), Kp = roundup(m->num_pose_feats, 16);
  FwdPrepArgs p;
  p.rot = rot; p.betas = betas; p.trans = trans; p.kid = kid; p.parents = m->parents;
  p.J_template = m->J_template; p.J_shapedirs = m->J_shapedirs; p.kid_J_shapedir = m->kid_J_shapedir;
  p.skin = w.skin; p.feat = w.feat; p.betaT = w.betaT; p.out_joints = out_joints;
  p.out_orientations = out_orientations; p.rot_mode = rot_mode; p.n_betas = betas ? n_betas : 0;
  p.J = m->num_joints; p.S = m->num_betas; p.nb = nb_used; p.B = B; p.Bp = Bp; p.Kp = Kp;
  SF_LAUNCH(k_fwd_prep, Bp / 32, 32, 0, st, p);
  if (out_vertices != nullptr) {
    if (!vposed_tc_run(m, w.feat, w.vposedT, Bp, Kp, w.tc_scratch, st)) {
      dim3 grid((3 * m->num_vertices + 127) / 128, (Bp + 63) / 64);
      SF_LAUNCH(k_vposed_gemm_simt, grid, 256, 0, st, m->posedirs_fit, m->v_template_fit, w.feat,
                3 * m->num_vertices, Kp, Bp, w.vposedT);
    }
    FwdSkinArgs s;
    s.vposedT = w.vposedT; s.betaT = w.betaT; s.skin = w.skin; s.shapedirs = m->shapedirs;
    s.kid_shapedir = m->kid_shapedir; s.skin_idx = m->skin_idx; s.skin_w = m->skin_w; s.inv_order = m->inv_order;
    s.out = out_vertices; s.V = m->num_vertices; s.S = m->num_betas; s.B = B; s.Bp = Bp; s.skin_k = m->skin_k;
    s.use_kid = kid != nullptr; s.nb = nb_used;
    dim3 grid(((m->num_vertices + 31) / 32 + 2) / 3, Bp / 32);
    SF_LAUNCH(k_fwd_skin, grid, 96, 0, st, s);
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
  const size_t smem = (size_t)v_in * 3 * sizeof(float);
  if (smem <= 200 * 1024 && batch <= 0x7fffffffLL) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_csr_apply_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH(k_csr_apply_smem, (unsigned)batch, 512, smem, st, indptr, indices, data, v_out, v_in, in_vertices, out_vertices);
  } else {
    SF_LAUNCH(k_csr_apply, (unsigned)((total + 255) / 256), 256, 0, st, indptr, indices, data, v_out, v_in, total,
              in_vertices, out_vertices);
  }
  SF_CHECK_LAST();
  return SMPLFIT_OK;
}
