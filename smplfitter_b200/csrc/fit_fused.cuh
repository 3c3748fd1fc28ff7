// Fused fit passes (fit_fused.cu): blend-shape GEMM on tcgen05 with the shape-stage (mode 2) or statistics (mode 3)
// vertex pass in its epilogue.  Partials in the layouts of k_shape_lite / k_stats_lite.
#pragma once
#include <cuda_runtime.h>

#include "../../include/smplfit_b200.h"

namespace sf {
bool fit_fused_available(const smplfit_model_t* m);
size_t fit_fused_scratch_bytes(const smplfit_model_t* m, int Bp);
// feat [Bp][Kp] = vec(R_rel[1:]) (k_front_rel); beta [NS][Bp] or null (mode 2: null -> x without the shape offsets);
// quads: mode 2 RT12, mode 3 skin4; ct0 / ca0 / aT_out / vwT: mode 3 only.  false = not launched (caller falls back).
// feats_ready: the fp16 hi / lo feature rows in `scratch` (fit_fused_feature_rows) are already up to date
// (k_front_fused wrote the pose part and zeroed the rest, k_shape_out wrote the unknowns): no feature kernel is launched.
bool fit_fused_run(const smplfit_model_t* m, int mode, int B, int Bp, const float* feat, int Kp, const float* beta,
                   const float* tT, const float* vwT, const float* quads, const float* ct0, const float* ca0, float* aT_out,
                   int all_segments, float* partials, void* scratch, bool feats_ready, cudaStream_t st);
// the feature rows inside `scratch`: [roundup(Bp,128)][fq_kf] halves each
void fit_fused_feature_rows(const smplfit_model_t* m, int Bp, void* scratch, void** hi, void** lo);
}  // namespace sf
