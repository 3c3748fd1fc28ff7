// Tensor-core (tcgen05) path of the pose-blend-shape contraction; declared here, defined in
// vposed_tc.cu.  Returns false when the path is unavailable for the given shapes so the
// caller falls back to the FP32 SIMT GEMM of fit_kernels.cuh.
#pragma once
#include <cuda_runtime.h>

#include "../../include/smplfit_b200.h"

namespace sf {
size_t vposed_tc_scratch_bytes(const smplfit_model_t* m, int Bp);
bool vposed_tc_run(const smplfit_model_t* m, const float* feat, float* vposedT, int Bp, int Kp, void* scratch,
                   cudaStream_t st);
}  // namespace sf
