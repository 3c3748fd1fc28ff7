// Launchers of the templated vertex-pass / solve kernels (instantiated in pass_shape.cu and
// pass_stats.cu so the translation units build in parallel).
#pragma once
#include <cuda_runtime.h>

#include "fit_kernels.cuh"
#include "solve_kernels.cuh"

namespace sf {
// true when the record kernels apply (<= 4 influences per vertex and the per-joint rows of 32
// instances fit in shared memory)
bool shape_pass_uses_records(const smplfit_model_t* m);
// chunking of the shape pass for a given batch: balanced against the SM count at launch time
struct ShapePlan {
  bool use_rec;      // record-based kernels (v2 or v3)
  int kind;          // 0 = generic k_shape_pass, 2 = k_shape_pass_v2 / rec (all joints staged), 3 = k_shape_pass_v3
  int cap_joints;    // v3: joints that fit in the staging area
  int warps, chunk_len, n_chunks, n_partials;
};
size_t rt4_floats(const smplfit_model_t* m, int Bp);  // size of the quad-layout row buffer
void launch_shape_pass_v3(const ShapeArgs& a, int ns, int groups, const ShapePlan& p, cudaStream_t st);
ShapePlan plan_shape_pass(const smplfit_model_t* m, int groups);
int max_shape_partials(const smplfit_model_t* m);
void launch_shape_pass(const ShapeArgs& a, int ns, int groups, const ShapePlan& p, cudaStream_t st);
void launch_shape_solve(const SolveArgs& a, double* Gd, int ns, int groups, cudaStream_t st);
// closed-form ("lite") shape path (pass_lite.cu): availability = tables present; enabled = not switched off by
// SMPLFIT_B200_SHAPE_VARIANT (6 = lite, the default; 4 / 5 / 0 / 1 select the per-vertex Gram kernels)
struct LiteArgs;
bool lite_available(const smplfit_model_t* m);
bool lite_enabled(const smplfit_model_t* m);
int lite_rows(int ns);                          // rows per segment of the lite partials
int gram_closed_blocks(const smplfit_model_t* m);
size_t gram_pairs_scratch_floats(const smplfit_model_t* m, int Bp);  // pair features of the tensor-core pair term
// the vertex pass (r, Sb, Y) + its per-joint reduction, and -- independent of it -- the Gramian from the joint
// transforms (pair term on tcgen05 + translation terms); fit.cu runs the latter on a side stream
void launch_shape_lite(const LiteArgs& a, const smplfit_model_t* m, int groups, double* Yd, cudaStream_t st);
// the per-joint reduction alone (the fused pass of fit_fused.cu has written the partials)
void launch_lite_reduce(const LiteArgs& a, const smplfit_model_t* m, int groups, double* Yd, cudaStream_t st);
void launch_gram_closed(const smplfit_model_t* m, int groups, int Bp, const float* RT, float* gcf_part, float* pair_scratch,
                        cudaStream_t st);
// statistics pass against the skinned current fit in the same style (SMPLFIT_B200_STATS_VARIANT=0 selects k_stats_rec)
struct StatsLiteArgs;
bool stats_lite_enabled(const smplfit_model_t* m);
void launch_stats_lite(const StatsLiteArgs& a, const smplfit_model_t* m, int groups, cudaStream_t st);
void launch_stats_tmpl(const StatsLiteArgs& a, const smplfit_model_t* m, const float* template_fit, const float* ca0_const,
                       int groups, cudaStream_t st);
// scale modes (final solve only): extra vertex pass + (NS+1)-unknown solve (pass_scale.cu)
int scale_chunks(const smplfit_model_t* m);
void launch_scale_pass(const ShapeArgs& a, int ns, int mode, int groups, cudaStream_t st);
void launch_shape_solve_scale(const SolveArgs& a, double* Gd, double* Zd, int ns, int groups, cudaStream_t st);
// share_beta: Cd = [NG+NS][Bp] doubles, sums = NG+NS doubles, x = NS doubles (device scratch)
void launch_shape_solve_shared(const SolveArgs& a, double* Gd, double* Cd, double* sums, double* x, int ns, int groups,
                               cudaStream_t st);
// share_beta together with scale estimation (partial share: the scale unknown stays per instance)
void launch_shape_solve_shared_scale(const SolveArgs& a, double* Gd, double* Zd, double* Cd, double* sums, double* x, int ns,
                                     int groups, cudaStream_t st);
void set_share_beta_allreduce(smplfit_allreduce_fn fn, void* user, int64_t global_batch);
bool share_beta_allreduce_installed();
void launch_stats(const StatsArgs& legacy, const StatsRecArgs& rec, int ns, int ref_mode, bool weighted, bool use_rec,
                  int groups, cudaStream_t st);
}  // namespace sf
