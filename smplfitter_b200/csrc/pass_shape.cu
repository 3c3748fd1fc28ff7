// Instantiations + launchers of the shape pass and the shape solve for NS = 2..17 unknowns.
#include <stdlib.h>

#include "common.cuh"
#include "passes.cuh"

namespace sf {

#define SF_NS_SWITCH(NSV, CALL)                      \
  switch (NSV) {                                     \
    case 2: { constexpr int NS = 2; CALL; } break;   \
    case 3: { constexpr int NS = 3; CALL; } break;   \
    case 4: { constexpr int NS = 4; CALL; } break;   \
    case 5: { constexpr int NS = 5; CALL; } break;   \
    case 6: { constexpr int NS = 6; CALL; } break;   \
    case 7: { constexpr int NS = 7; CALL; } break;   \
    case 8: { constexpr int NS = 8; CALL; } break;   \
    case 9: { constexpr int NS = 9; CALL; } break;   \
    case 10: { constexpr int NS = 10; CALL; } break; \
    case 11: { constexpr int NS = 11; CALL; } break; \
    case 12: { constexpr int NS = 12; CALL; } break; \
    case 13: { constexpr int NS = 13; CALL; } break; \
    case 14: { constexpr int NS = 14; CALL; } break; \
    case 15: { constexpr int NS = 15; CALL; } break; \
    case 16: { constexpr int NS = 16; CALL; } break; \
    case 17: { constexpr int NS = 17; CALL; } break; \
    default: break;                                  \
  }

static size_t rt_smem_bytes(int J, int ns) { return (size_t)J * (12 + 3 * ns) * 32 * sizeof(float); }

bool shape_pass_uses_records(const smplfit_model_t* m) {
  return m->fit_rec != nullptr && m->skin_k <= 4 && rt_smem_bytes(m->num_joints, m->fit_ns) <= 200 * 1024;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static size_t rec_stage_bytes(int ns) {
  const int nsp = (ns + 1) / 2 * 2, rec = (8 + 3 * nsp + 3) / 4 * 4;
  return (size_t)8 * 2 * REC_SUB * rec * sizeof(float) + 16 * 8 + 64 * 4 + 64;
}

size_t rt4_floats(const smplfit_model_t* m, int Bp) {
  const size_t a = (size_t)m->num_joints * quad_rows_ns(m->fit_ns), b = (size_t)m->num_joints * clay_rows_per_joint(m->fit_ns);
  return (a > b ? a : b) * (size_t)Bp;
}

ShapePlan plan_shape_pass(const smplfit_model_t* m, int groups) {
  ShapePlan p;
  const bool rec_ok = m->fit_rec != nullptr && m->skin_k <= 4;
  const size_t v2_smem = (size_t)m->num_joints * (quad_rows_ns(m->fit_ns) / 4) * 512 + rec_stage_bytes(m->fit_ns);
  p.kind = 0;
  p.cap_joints = 0;
  if (rec_ok && v2_smem <= 210 * 1024 && m->fit_ns <= 12) {
    p.kind = 2;
  } else if (rec_ok) {
    p.kind = 3;
    const size_t per_joint = (size_t)(clay_rows_per_joint(m->fit_ns) / 4) * 512;
    const size_t budget = (size_t)200 * 1024 - rec_stage_bytes(m->fit_ns);
    p.cap_joints = (int)(budget / per_joint);
    if (p.cap_joints > m->num_joints) p.cap_joints = m->num_joints;
  }
  p.use_rec = p.kind != 0;
  p.warps = 8;
  // CTAs per instance group chosen so that the grid fills whole waves of the SM count
  const int V = m->num_vertices, sms = num_sms();
  const int c_min = (V + p.warps * 256 - 1) / (p.warps * 256), c_max = (V + p.warps * 32 - 1) / (p.warps * 32);
  double best = -1.0;
  int best_c = c_min;
  for (int c = c_min; c <= c_max; ++c) {
    const long long ctas = (long long)groups * c;
    const long long waves = (ctas + sms - 1) / sms;
    const double eff = (double)ctas / (double)(waves * sms) - 1e-4 * c;  // prefer fewer, longer chunks on ties
    if (eff > best) { best = eff; best_c = c; }
  }
  p.chunk_len = (V + p.warps * best_c - 1) / (p.warps * best_c);
  p.n_chunks = (V + p.chunk_len - 1) / p.chunk_len;
  p.n_partials = (p.n_chunks + p.warps - 1) / p.warps;
  return p;
}

int max_shape_partials(const smplfit_model_t* m) {
  const int V = m->num_vertices;
  return (V + 8 * 32 - 1) / (8 * 32) + 1;
}

template <int NS, bool WEIGHTED>
static void shape_v2_launch(const ShapeArgs& a, int groups, const ShapePlan& p, cudaStream_t st) {
  const size_t smem_rt = (size_t)a.J * Quad<NS>::NQ * 32 * 16 + (size_t)8 * 2 * REC_SUB * Rec<NS>::LEN * sizeof(float) + 16 * 8 + 16;
  const size_t smem_red = (size_t)4 * (2 * PackedG<NS>::NPAIRS + 8 * (Rec<NS>::NSP / 2) + 4) * 32 * sizeof(float);
  const size_t smem = smem_rt > smem_red ? smem_rt : smem_red;
  cudaFuncSetAttribute(k_shape_pass_v2<NS, WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SF_LAUNCH((k_shape_pass_v2<NS, WEIGHTED>), dim3(p.n_partials, groups), 256, smem, st, a);
}

template <int NS, bool WEIGHTED>
static void shape_pass_t(const ShapeArgs& sa, int groups, const ShapePlan& p, cudaStream_t st) {
  const size_t smem_rt = rt_smem_bytes(sa.J, NS);
  const size_t smem_red = (size_t)4 * ShapeAcc<NS>::N * 32 * sizeof(float);
  ShapeArgs a = sa;
  a.chunk_len = p.chunk_len;
  a.n_chunks = p.n_chunks;
  a.chunks_per_cta = p.warps;
  if (p.kind == 3) {
    launch_shape_pass_v3(a, NS, groups, p, st);
    return;
  }
  if (p.use_rec) {
    shape_v2_launch<NS, WEIGHTED>(a, groups, p, st);
    return;
  }
  dim3 grid(p.n_partials, groups);
  if (smem_rt <= 200 * 1024) {
    const size_t smem = smem_rt > smem_red ? smem_rt : smem_red;
    cudaFuncSetAttribute(k_shape_pass<NS, WEIGHTED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH((k_shape_pass<NS, WEIGHTED, true>), grid, 256, smem, st, a);
  } else {
    cudaFuncSetAttribute(k_shape_pass<NS, WEIGHTED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_red);
    SF_LAUNCH((k_shape_pass<NS, WEIGHTED, false>), grid, 256, smem_red, st, a);
  }
}

void launch_shape_pass(const ShapeArgs& a, int ns, int groups, const ShapePlan& p, cudaStream_t st) {
  if (a.vwT) { SF_NS_SWITCH(ns, (shape_pass_t<NS, true>(a, groups, p, st))); }
  else { SF_NS_SWITCH(ns, (shape_pass_t<NS, false>(a, groups, p, st))); }
}

template <int NS>
static void shape_solve_t(const SolveArgs& so, double* Gd, int groups, cudaStream_t st) {
  constexpr int NACC = ShapeAcc<NS>::N;
  SF_LAUNCH(k_gram_entries<NS>, dim3(groups, NACC), 32, 0, st, so, Gd);
  if constexpr (NS >= 3) {  // (trans is written by rows 0..2)
    const size_t smem = shape_solve_par_smem(NS);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_shape_solve_par<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH(k_shape_solve_par<NS>, groups, NS * 32, smem, st, so, (const double*)Gd);
  } else {
    SF_LAUNCH(k_shape_solve<NS>, groups, 32, 0, st, so, Gd);
  }
  SF_LAUNCH(k_shape_out, dim3(groups, so.J), 32, 0, st, so, NS);
}

void launch_shape_solve(const SolveArgs& a, double* Gd, int ns, int groups, cudaStream_t st) {
  SF_NS_SWITCH(ns, (shape_solve_t<NS>(a, Gd, groups, st)));
}

}  // namespace sf
