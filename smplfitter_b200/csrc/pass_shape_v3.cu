// Instantiations + launcher of k_shape_pass_v3 (large models: joint-subset staging, coordinate split).
#include "common.cuh"
#include "passes.cuh"

namespace sf {

template <int NS, bool WEIGHTED, int ROWSEL>
static void v3_launch(const ShapeArgs& a, int groups, const ShapePlan& p, cudaStream_t st) {
  using L = CLay<NS>;
  const int nsp = Rec<NS>::NSP;
  (void)nsp;
  // reduction area: 4 x NRED x 32 floats must fit in the staging area
  const size_t red = (size_t)4 * (2 * PackedG<NS>::NPAIRS + 8 * L::H + 4) * 32 * sizeof(float);
  size_t stage = (size_t)p.cap_joints * L::JQ * 512;
  if (stage < red) stage = red;
  const size_t smem = stage + (size_t)8 * 2 * REC_SUB * Rec<NS>::LEN * sizeof(float) + 16 * 8 + 64 * 4 + 64;
  ShapeV3Extra x3;
  x3.cap_joints = (int)(stage / ((size_t)L::JQ * 512));
  if (x3.cap_joints > p.cap_joints && p.cap_joints > 0) x3.cap_joints = p.cap_joints;
  cudaFuncSetAttribute(k_shape_pass_v3<NS, WEIGHTED, ROWSEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SF_LAUNCH((k_shape_pass_v3<NS, WEIGHTED, ROWSEL>), dim3(p.n_partials, groups), 256, smem, st, a, x3);
}

template <int NS, bool WEIGHTED>
static void v3_dispatch(const ShapeArgs& a, int groups, const ShapePlan& p, cudaStream_t st) {
  if constexpr (NS >= 14) {
    v3_launch<NS, WEIGHTED, 1>(a, groups, p, st);
    v3_launch<NS, WEIGHTED, 2>(a, groups, p, st);
  } else {
    v3_launch<NS, WEIGHTED, 0>(a, groups, p, st);
  }
}

#define SF_V3_CASE(N)                                                              \
  case N:                                                                          \
    if (a.vwT) v3_dispatch<N, true>(a, groups, p, st); else v3_dispatch<N, false>(a, groups, p, st); \
    break;

void launch_shape_pass_v3(const ShapeArgs& a, int ns, int groups, const ShapePlan& p, cudaStream_t st) {
  switch (ns) {
    SF_V3_CASE(2) SF_V3_CASE(3) SF_V3_CASE(4) SF_V3_CASE(5) SF_V3_CASE(6) SF_V3_CASE(7) SF_V3_CASE(8) SF_V3_CASE(9)
    SF_V3_CASE(10) SF_V3_CASE(11) SF_V3_CASE(12) SF_V3_CASE(13) SF_V3_CASE(14) SF_V3_CASE(15) SF_V3_CASE(16)
    SF_V3_CASE(17)
    default: break;
  }
}

}  // namespace sf
