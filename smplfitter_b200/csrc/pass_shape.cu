// Instantiations + launchers of the shape pass and the shape solve for NS = 2..17 unknowns.
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

template <int NS, bool WEIGHTED>
static void shape_pass_t(const ShapeArgs& sa, int groups, bool use_rec, cudaStream_t st) {
  const size_t smem_rt = rt_smem_bytes(sa.J, NS);
  const size_t smem_red = (size_t)4 * ShapeAcc<NS>::N * 32 * sizeof(float);
  ShapeArgs a = sa;
  a.chunks_per_cta = 8;
  dim3 grid((sa.n_chunks + 7) / 8, groups);
  if (use_rec) {
    const size_t smem = smem_rt > smem_red ? smem_rt : smem_red;
    cudaFuncSetAttribute(k_shape_pass_rec<NS, WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH((k_shape_pass_rec<NS, WEIGHTED>), grid, 256, smem, st, a);
  } else if (smem_rt <= 200 * 1024) {
    const size_t smem = smem_rt > smem_red ? smem_rt : smem_red;
    cudaFuncSetAttribute(k_shape_pass<NS, WEIGHTED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH((k_shape_pass<NS, WEIGHTED, true>), grid, 256, smem, st, a);
  } else {
    cudaFuncSetAttribute(k_shape_pass<NS, WEIGHTED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_red);
    SF_LAUNCH((k_shape_pass<NS, WEIGHTED, false>), grid, 256, smem_red, st, a);
  }
}

void launch_shape_pass(const ShapeArgs& a, int ns, int groups, bool use_rec, cudaStream_t st) {
  if (a.vwT) { SF_NS_SWITCH(ns, (shape_pass_t<NS, true>(a, groups, use_rec, st))); }
  else { SF_NS_SWITCH(ns, (shape_pass_t<NS, false>(a, groups, use_rec, st))); }
}

template <int NS>
static void shape_solve_t(const SolveArgs& so, double* Gd, int groups, cudaStream_t st) {
  constexpr int NACC = ShapeAcc<NS>::N;
  SF_LAUNCH(k_gram_entries<NS>, dim3(groups, NACC), 32, 0, st, so, Gd);
  SF_LAUNCH(k_shape_solve<NS>, groups, 32, 0, st, so, Gd);
  SF_LAUNCH(k_shape_out, dim3(groups, so.J), 32, 0, st, so, NS);
}

void launch_shape_solve(const SolveArgs& a, double* Gd, int ns, int groups, cudaStream_t st) {
  SF_NS_SWITCH(ns, (shape_solve_t<NS>(a, Gd, groups, st)));
}

}  // namespace sf
