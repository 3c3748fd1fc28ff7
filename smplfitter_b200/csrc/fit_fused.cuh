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
bool fit_fused_run(const smplfit_model_t* m, int mode, int B, int Bp, const float* feat, int Kp, const float* beta,
                   const float* tT, const float* vwT, const float* quads, const float* ct0, const float* ca0, float* aT_out,
                   int all_segments, float* partials, void* scratch, cudaStream_t st);
}  // namespace sf
