// Instantiations + launchers of the statistics pass (rotation stage).
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

template <int NS, int REF, bool WEIGHTED>
static void stats_rec_t(const StatsRecArgs& ra, int groups, cudaStream_t st) {
  StatsRecArgs a = ra;
  a.segs_per_warp = 2;
  const size_t smem = (REF == 1) ? ((size_t)a.J * 12 * 32 + (size_t)8 * 2 * REC_SUB * Rec<NS>::LEN) * sizeof(float) + 16 * 8 + 16 : 0;
  dim3 grid((a.n_segments + 8 * a.segs_per_warp - 1) / (8 * a.segs_per_warp), groups);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_stats_rec<NS, REF, WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SF_LAUNCH((k_stats_rec<NS, REF, WEIGHTED>), grid, 256, smem, st, a);
}

template <int REF, bool WEIGHTED>
static void stats_legacy_t(const StatsArgs& s, int groups, cudaStream_t st) {
  const long long warps = (long long)s.n_segments * groups;
  SF_LAUNCH((k_stats<REF, WEIGHTED>), (int)((warps + 3) / 4), 128, 0, st, s);
}

void launch_stats(const StatsArgs& legacy, const StatsRecArgs& rec, int ns, int ref_mode, bool weighted, bool use_rec,
                  int groups, cudaStream_t st) {
  if (use_rec) {
    if (ref_mode == 0) {
      if (weighted) stats_rec_t<2, 0, true>(rec, groups, st); else stats_rec_t<2, 0, false>(rec, groups, st);
    } else if (ref_mode == 2) {
      if (weighted) stats_rec_t<2, 2, true>(rec, groups, st); else stats_rec_t<2, 2, false>(rec, groups, st);
    } else if (weighted) {
      SF_NS_SWITCH(ns, (stats_rec_t<NS, 1, true>(rec, groups, st)));
    } else {
      SF_NS_SWITCH(ns, (stats_rec_t<NS, 1, false>(rec, groups, st)));
    }
    return;
  }
  if (ref_mode == 0) {
    if (weighted) stats_legacy_t<0, true>(legacy, groups, st); else stats_legacy_t<0, false>(legacy, groups, st);
  } else if (ref_mode == 1) {
    if (weighted) stats_legacy_t<1, true>(legacy, groups, st); else stats_legacy_t<1, false>(legacy, groups, st);
  } else {
    if (weighted) stats_legacy_t<2, true>(legacy, groups, st); else stats_legacy_t<2, false>(legacy, groups, st);
  }
}

}  // namespace sf
