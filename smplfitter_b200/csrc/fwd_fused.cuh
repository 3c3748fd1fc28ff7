// Fused forward LBS (tcgen05 GEMM + skinning epilogue); defined in fwd_fused.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/smplfit_b200.h"

namespace sf {

constexpr int FWD_CHUNK = 16;  // vertices per epilogue chunk: the host orders vertices / assigns slots per chunk (pt/bodymodel.py)

struct FwdFusedWs {
  float4* quads;  // [J*3][Bt]
  __half* f_hi;   // [Bt][Kf]
  __half* f_lo;
  int Bt;
  size_t bytes;
};

bool fwd_fused_available(const smplfit_model_t* m);
FwdFusedWs fwd_fused_carve(void* base, const smplfit_model_t* m, int64_t B, bool with_vertices);
int fwd_fused_run(const smplfit_model_t* m, int B, int rot_mode, const float* rot, const float* betas, int n_betas,
                  const float* trans, const float* kid, float* out_vertices, float* out_joints, float* out_orientations,
                  const FwdFusedWs& w, cudaStream_t st);

// the fit's v_posed^T GEMM on the same fp16-split main loop (feat: [Bp][Kp] fp32 rows vec(R_rel[1:]))
bool vposed_f16_available(const smplfit_model_t* m);
size_t vposed_f16_scratch_bytes(const smplfit_model_t* m, int Bp);
bool vposed_f16_run(const smplfit_model_t* m, const float* feat, float* vposedT, int Bp, int Kp, void* scratch, cudaStream_t st);

}  // namespace sf
