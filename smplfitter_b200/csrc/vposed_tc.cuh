// Tensor-core (tcgen05) path of the pose-blend-shape contraction; declared here, defined in
// vposed_tc.cu.  Returns false when the path is unavailable for the given shapes so the
// caller falls back to the FP32 SIMT GEMM of fit_kernels.cuh.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/smplfit_b200.h"

namespace sf {
size_t vposed_tc_scratch_bytes(const smplfit_model_t* m, int Bp);
bool vposed_tc_run(const smplfit_model_t* m, const float* feat, float* vposedT, int Bp, int Kp, void* scratch,
                   cudaStream_t st);
bool vposed_tc_run_model(const smplfit_model_t* m, const float* feat, float* vposedT, int Bp, int Kp, void* scratch,
                         cudaStream_t st);
// the generic kernel behind it: out[n][b] = bias[n] (nullable) + sum_k p[n][k] f[b][k] with both operands pre-split into
// tf32-exact hi / lo parts; p_*: [rows_alloc >= rows][Kt], f_*: [Bt][Kt], Kt % tc_tile_k() == 0, Bt % tc_tile_m() == 0
bool tc_gemm_run(const float* p_hi, const float* p_lo, int rows, int rows_alloc, int Kt, const float* bias,
                 const float* f_hi, const float* f_lo, int Bt, float* out, int Bp, cudaStream_t st);
int tc_tile_k();
int tc_tile_m();
int tc_tile_n();
bool gram_pairs_tc_available(const smplfit_model_t* m);
bool tensor_maps_available();
bool make_im_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t Bp, uint32_t box_rows);
}  // namespace sf
