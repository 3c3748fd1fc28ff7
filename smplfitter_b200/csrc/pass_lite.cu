// Launchers of the closed-form ("lite") unweighted shape path: k_shape_lite (vertex pass for r, Sb, Y),
// k_lite_reduce (Y per joint) and k_gram_closed (Gramian from the joint transforms and model constants).
#include <stdlib.h>

#include "common.cuh"
#include "lite_kernels.cuh"
#include "passes.cuh"
#include "vposed_tc.cuh"

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

static constexpr size_t kSmemMax = 227 * 1024;

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// warps per CTA of the vertex kernels (one CTA per SM): 12 when the staging fits, else 8, else unavailable
static int env_int(const char* name, int dflt);
static int lite_warps(const smplfit_model_t* m) {
  if (lite_smem_bytes(m->num_joints, m->fit_rec_len, 12) <= kSmemMax) return 12;
  if (lite_smem_bytes(m->num_joints, m->fit_rec_len, 8) <= kSmemMax) return 8;
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

bool lite_available(const smplfit_model_t* m) {
  return m->fit_rec != nullptr && m->skin_k <= 4 && m->seg_slots != nullptr && m->yj_start != nullptr &&
         m->gcf_A != nullptr && m->gcf_lstart != nullptr && m->gcf_G0 != nullptr && m->n_slots == LITE_NSLOT &&
         m->fit_wS != nullptr && lite_warps(m) != 0 && tensor_maps_available();
}

bool lite_enabled(const smplfit_model_t* m) {
  static int v = -1;
  if (v < 0) v = env_int("SMPLFIT_B200_SHAPE_VARIANT", 6) == 6 ? 1 : 0;
  return v == 1 && lite_available(m);
}

bool stats_lite_enabled(const smplfit_model_t* m) {
  static int v = -1;
  if (v < 0) v = env_int("SMPLFIT_B200_STATS_VARIANT", 1) == 1 ? 1 : 0;
  return v == 1 && m->fit_rec != nullptr && m->skin_k <= 4 && tensor_maps_available() &&
         stats_lite_smem_bytes(m->num_joints, m->fit_rec_len, 8) <= kSmemMax;
}

int lite_rows(int ns) { return ns + 3 + 3 * LITE_NSLOT; }

// partial Gramians of the pair term: one (the tcgen05 GEMM) or one per CTA of the SIMT fallback k_gram_pairs
static int pair_ctas(const smplfit_model_t* m) {
  if (gram_pairs_tc_available(m)) return 1;
  return (m->gcf_npairs + 8 * LITE_PPW - 1) / (8 * LITE_PPW);
}
int gram_closed_blocks(const smplfit_model_t* m) { return pair_ctas(m) + 3; }
// hi / lo pair features [Bt][Kt] of the tensor-core path
size_t gram_pairs_scratch_floats(const smplfit_model_t* m, int Bp) {
  if (!gram_pairs_tc_available(m)) return 0;
  return (size_t)2 * roundup(Bp, tc_tile_m()) * roundup(9 * m->gcf_npairs, tc_tile_k()) + 64;
}

// segments per warp that fills whole waves of the SM count best (one CTA per SM)
static int pick_spw(int n_segments, int warps, int groups) {
  const long long slots = sm_count();
  double best = -1.0;
  int best_spw = 1;
  for (int spw = 1; spw <= 4; ++spw) {
    const long long ctas = (long long)groups * ((n_segments + warps * spw - 1) / (warps * spw));
    const long long waves = (ctas + slots - 1) / slots;
    const double eff = (double)ctas / (double)(waves * slots) + 0.01 * spw;  // prefer longer CTAs on near-ties
    if (eff > best) { best = eff; best_spw = spw; }
  }
  return best_spw;
}

// tensor maps of the two instance-minor [3V][Bp] streams: box = {32 instances, 3 LITE_VS rows}
static bool stream_maps(const float* tT, const float* vposedT, int V, int Bp, CUtensorMap* mt, CUtensorMap* mv) {
  return make_im_map(mt, tT, (uint64_t)3 * V, (uint64_t)Bp, 3 * LITE_VS) && make_im_map(mv, vposedT, (uint64_t)3 * V, (uint64_t)Bp, 3 * LITE_VS);
}

template <int NS, int WARPS>
static void lite_launch_t(LiteArgs a, int V, int groups, cudaStream_t st) {
  CUtensorMap mt, mv;
  if (!stream_maps(a.tT, a.vposedT, V, a.Bp, &mt, &mv)) {
    g_launch_failure = "cuTensorMapEncodeTiled failed for the vertex streams";
    return;
  }
  const size_t smem = lite_smem_bytes(a.J, Rec<NS>::LEN, WARPS);
  a.segs_per_warp = pick_spw(a.n_segments, WARPS, groups);
  cudaFuncSetAttribute(k_shape_lite<NS, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((a.n_segments + WARPS * a.segs_per_warp - 1) / (WARPS * a.segs_per_warp), groups);
  SF_LAUNCH((k_shape_lite<NS, WARPS>), grid, WARPS * 32, smem, st, a, mt, mv);
}

template <int NS>
static void gram_closed_t(const smplfit_model_t* m, int groups, int Bp, const float* RT, float* gcf_part,
                          float* pair_scratch, cudaStream_t st) {
  constexpr int NG = NS * (NS + 1) / 2, NGP = (NG + 3) / 4 * 4;
  struct { int Bp; } a{Bp};
  GramClosedArgs ga;
  ga.RT = RT; ga.pairs = m->gcf_pairs; ga.A = m->gcf_A; ga.lstart = m->gcf_lstart; ga.lk = m->gcf_lk; ga.Bm = m->gcf_Bm;
  ga.Wh = m->gcf_Wh; ga.out = gcf_part; ga.npairs = m->gcf_npairs; ga.J = m->num_joints; ga.Bp = a.Bp;
  ga.n_pair_ctas = pair_ctas(m);
  const size_t red = (size_t)4 * NGP * 32 * sizeof(float);
  const size_t rows = (size_t)m->num_joints * (3 + NS) * 32 * sizeof(float);
  const size_t smem_t = rows > red ? rows : red;
  if (red > 48 * 1024) cudaFuncSetAttribute(k_gram_pairs<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red);
  if (smem_t > 48 * 1024) cudaFuncSetAttribute(k_gram_trans<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
  bool pairs_done = false;
  if (gram_pairs_tc_available(m) && pair_scratch != nullptr) {
    // pair term on the tensor cores: features (R_k^T R_l) x constants A^T, 3xTF32 (vposed_tc.cu)
    const int Kt = roundup(9 * m->gcf_npairs, tc_tile_k()), Bt = roundup(a.Bp, tc_tile_m());
    PairFeatArgs pf;
    pf.RT = RT; pf.pairs = m->gcf_pairs; pf.hi = pair_scratch; pf.lo = pair_scratch + (size_t)Bt * Kt;
    pf.npairs = m->gcf_npairs; pf.J = m->num_joints; pf.RW = 12 + 3 * NS; pf.Bp = a.Bp; pf.Kt = Kt;
    constexpr int PF_SPLIT = 4;  // CTAs per instance group (pair ranges)
    const int pf_per = (m->gcf_npairs + PF_SPLIT - 1) / PF_SPLIT;
    const size_t smem_f = ((size_t)32 * ((m->num_joints * 9) | 1) + (size_t)8 * 2 * pf_per * 9) * sizeof(float);
    if (smem_f > 48 * 1024) cudaFuncSetAttribute(k_pair_feat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
    SF_LAUNCH(k_pair_feat, dim3(Bt / 32, PF_SPLIT), 256, smem_f, st, pf);
    pairs_done = tc_gemm_run(m->gcf_AT_hi, m->gcf_AT_lo, NG, roundup(NG, tc_tile_n()), Kt, nullptr, pf.hi, pf.lo, Bt,
                             gcf_part, a.Bp, st);
  }
  if (!pairs_done && ga.n_pair_ctas > 0) {
    if (gram_pairs_tc_available(m)) {
      // the workspace holds one pair block on this path: the SIMT kernel cannot take over
      g_launch_failure = "tcgen05 pair GEMM could not be launched";
    } else {
      SF_LAUNCH(k_gram_pairs<NS>, dim3(groups, ga.n_pair_ctas), 256, red, st, ga);
    }
  }
  SF_LAUNCH(k_gram_trans<NS>, dim3(groups, 3), 256, smem_t, st, ga);
}

template <int NS>
static void lite_t(const LiteArgs& a, const smplfit_model_t* m, int groups, double* Yd, cudaStream_t st) {
  if (lite_warps(m) == 12) lite_launch_t<NS, 12>(a, m->num_vertices, groups, st);
  else lite_launch_t<NS, 8>(a, m->num_vertices, groups, st);
  LiteReduceArgs ra;
  ra.partials = a.partials; ra.yj_start = m->yj_start; ra.yj_entry = m->yj_entry; ra.Yd = Yd; ra.NL = lite_rows(NS);
  ra.NS = NS; ra.Bp = a.Bp; ra.J = m->num_joints; ra.n_segments = m->n_segments;
  SF_LAUNCH(k_lite_reduce, dim3(groups, m->num_joints + NS + 3), 32, 0, st, ra);
}

void launch_lite_reduce(const LiteArgs& a, const smplfit_model_t* m, int groups, double* Yd, cudaStream_t st) {
  LiteReduceArgs ra;
  ra.partials = a.partials; ra.yj_start = m->yj_start; ra.yj_entry = m->yj_entry; ra.Yd = Yd; ra.NL = lite_rows(m->fit_ns);
  ra.NS = m->fit_ns; ra.Bp = a.Bp; ra.J = m->num_joints; ra.n_segments = m->n_segments;
  SF_LAUNCH(k_lite_reduce, dim3(groups, m->num_joints + m->fit_ns + 3), 32, 0, st, ra);
}

void launch_shape_lite(const LiteArgs& a, const smplfit_model_t* m, int groups, double* Yd, cudaStream_t st) {
  SF_NS_SWITCH(m->fit_ns, (lite_t<NS>(a, m, groups, Yd, st)));
}

void launch_gram_closed(const smplfit_model_t* m, int groups, int Bp, const float* RT, float* gcf_part, float* pair_scratch,
                        cudaStream_t st) {
  SF_NS_SWITCH(m->fit_ns, (gram_closed_t<NS>(m, groups, Bp, RT, gcf_part, pair_scratch, st)));
}

template <int NS, bool WEIGHTED, int WARPS>
static void stats_lite_launch_t(StatsLiteArgs a, int V, int groups, cudaStream_t st) {
  CUtensorMap mt, mv;
  if (!stream_maps(a.tT, a.vposedT, V, a.Bp, &mt, &mv)) {
    g_launch_failure = "cuTensorMapEncodeTiled failed for the vertex streams";
    return;
  }
  const size_t smem = stats_lite_smem_bytes(a.J, Rec<NS>::LEN, WARPS);
  a.segs_per_warp = pick_spw(a.n_segments, WARPS, groups);
  cudaFuncSetAttribute(k_stats_lite<NS, WEIGHTED, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((a.n_segments + WARPS * a.segs_per_warp - 1) / (WARPS * a.segs_per_warp), groups);
  SF_LAUNCH((k_stats_lite<NS, WEIGHTED, WARPS>), grid, WARPS * 32, smem, st, a, mt, mv);
}

template <int NS>
static void stats_lite_t(const StatsLiteArgs& a, const smplfit_model_t* m, int groups, cudaStream_t st) {
  const bool w12 = stats_lite_smem_bytes(m->num_joints, m->fit_rec_len, 12) <= kSmemMax;
  if (a.vwT != nullptr) {
    if (w12) stats_lite_launch_t<NS, true, 12>(a, m->num_vertices, groups, st); else stats_lite_launch_t<NS, true, 8>(a, m->num_vertices, groups, st);
  } else {
    if (w12) stats_lite_launch_t<NS, false, 12>(a, m->num_vertices, groups, st); else stats_lite_launch_t<NS, false, 8>(a, m->num_vertices, groups, st);
  }
}

void launch_stats_lite(const StatsLiteArgs& a, const smplfit_model_t* m, int groups, cudaStream_t st) {
  SF_NS_SWITCH(m->fit_ns, (stats_lite_t<NS>(a, m, groups, st)));
}

// first rotation fit: statistics against the constant template (REF == 0), TMA-staged
void launch_stats_tmpl(const StatsLiteArgs& a0, const smplfit_model_t* m, const float* template_fit, const float* ca0_const,
                       int groups, cudaStream_t st) {
  StatsLiteArgs a = a0;
  CUtensorMap mt;
  if (!make_im_map(&mt, a.tT, (uint64_t)3 * m->num_vertices, (uint64_t)a.Bp, 3 * TMPL_VS)) {
    g_launch_failure = "cuTensorMapEncodeTiled failed for the target stream";
    return;
  }
  constexpr int WARPS = 16;
  const size_t smem = stats_tmpl_smem_bytes(WARPS);
  a.segs_per_warp = pick_spw(a.n_segments, WARPS, groups);
  dim3 grid((a.n_segments + WARPS * a.segs_per_warp - 1) / (WARPS * a.segs_per_warp), groups);
  if (a.vwT != nullptr) {
    cudaFuncSetAttribute(k_stats_tmpl<true, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH((k_stats_tmpl<true, WARPS>), grid, WARPS * 32, smem, st, a, template_fit, ca0_const, mt);
  } else {
    cudaFuncSetAttribute(k_stats_tmpl<false, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH((k_stats_tmpl<false, WARPS>), grid, WARPS * 32, smem, st, a, template_fit, ca0_const, mt);
  }
}

}  // namespace sf
