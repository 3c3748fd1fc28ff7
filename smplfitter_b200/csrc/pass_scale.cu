// Instantiations + launchers of the scale-mode kernels (scale_target / scale_fit, final solve only).
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

int scale_chunks(const smplfit_model_t* m) { return (m->num_vertices + 127) / 128; }

template <int NS>
static void scale_pass_t(const ShapeArgs& sa, int mode, int groups, cudaStream_t st) {
  ShapeArgs a = sa;
  a.chunk_len = 128;
  a.n_chunks = (a.V + 127) / 128;
  const long long warps = (long long)a.n_chunks * groups;
  const int blocks = (int)((warps + 3) / 4);
  if (mode == 1) SF_LAUNCH((k_scale_pass<NS, 1>), blocks, 128, 0, st, a);
  else SF_LAUNCH((k_scale_pass<NS, 2>), blocks, 128, 0, st, a);
}

void launch_scale_pass(const ShapeArgs& a, int ns, int mode, int groups, cudaStream_t st) {
  SF_NS_SWITCH(ns, (scale_pass_t<NS>(a, mode, groups, st)));
}

template <int NS>
static void solve_scale_t(const SolveArgs& so, double* Gd, double* Zd, int groups, cudaStream_t st) {
  SF_LAUNCH(k_gram_entries<NS>, dim3(groups, ShapeAcc<NS>::N), 32, 0, st, so, Gd);
  SF_LAUNCH(k_scale_entries<NS>, dim3(groups, NS + 5), 32, 0, st, so, Zd);
  SF_LAUNCH(k_shape_solve_scale<NS>, groups, 32, 0, st, so, (const double*)Gd, (const double*)Zd);
  SF_LAUNCH(k_shape_out, dim3(groups, so.J), 32, 0, st, so, NS);
}

void launch_shape_solve_scale(const SolveArgs& a, double* Gd, double* Zd, int ns, int groups, cudaStream_t st) {
  SF_NS_SWITCH(ns, (solve_scale_t<NS>(a, Gd, Zd, groups, st)));
}

// share_beta over several ranks: hook called on the local batch sums before the shared solve (smplfit_b200.h)
static smplfit_allreduce_fn g_ar_fn = nullptr;
static void* g_ar_user = nullptr;
static int64_t g_ar_total = 0;
void set_share_beta_allreduce(smplfit_allreduce_fn fn, void* user, int64_t global_batch) {
  g_ar_fn = fn;
  g_ar_user = user;
  g_ar_total = global_batch;
}

bool share_beta_allreduce_installed() { return g_ar_fn != nullptr; }

template <int NS>
static void solve_shared_t(const SolveArgs& so, double* Gd, double* Cd, double* sums, double* x, int groups,
                           cudaStream_t st) {
  constexpr int NG = NS * (NS + 1) / 2;
  SF_LAUNCH(k_gram_entries<NS>, dim3(groups, ShapeAcc<NS>::N), 32, 0, st, so, Gd);
  SF_LAUNCH(k_center_entries<NS>, groups, 32, 0, st, so, (const double*)Gd, Cd);
  SF_LAUNCH(k_batch_sum, NG + NS, 256, 0, st, (const double*)Cd, so.Bp, sums);
  SolveArgs sg = so;
  if (g_ar_fn != nullptr) {
    g_ar_fn(sums, NG + NS, reinterpret_cast<void*>(st), g_ar_user);
    sg.B = (int)g_ar_total;  // diag(lambda) is added once per instance of the WHOLE batch
  }
  SF_LAUNCH(k_shared_solve<NS>, 1, 32, 0, st, sg, (const double*)sums, x);
  SF_LAUNCH(k_shared_apply<NS>, groups, 32, 0, st, so, (const double*)Gd, (const double*)x);
  SF_LAUNCH(k_shape_out, dim3(groups, so.J), 32, 0, st, so, NS);
}

template <int NS>
static void solve_shared_scale_t(const SolveArgs& so, double* Gd, double* Zd, double* Cd, double* sums, double* x, int groups,
                                 cudaStream_t st) {
  constexpr int NG = NS * (NS + 1) / 2;
  SF_LAUNCH(k_gram_entries<NS>, dim3(groups, ShapeAcc<NS>::N), 32, 0, st, so, Gd);
  SF_LAUNCH(k_scale_entries<NS>, dim3(groups, NS + 5), 32, 0, st, so, Zd);
  SF_LAUNCH(k_center_entries_scale<NS>, groups, 32, 0, st, so, (const double*)Gd, Zd, Cd);
  SF_LAUNCH(k_batch_sum, NG + NS, 256, 0, st, (const double*)Cd, so.Bp, sums);
  if (g_ar_fn != nullptr) g_ar_fn(sums, NG + NS, reinterpret_cast<void*>(st), g_ar_user);
  SolveArgs sg = so;
  sg.shared_noreg = 1;  // the regulariser rows are inside the per-instance Schur complements
  SF_LAUNCH(k_shared_solve<NS>, 1, 32, 0, st, sg, (const double*)sums, x);
  SF_LAUNCH(k_shared_apply_scale<NS>, groups, 32, 0, st, so, (const double*)Gd, (const double*)Zd, (const double*)x);
  SF_LAUNCH(k_shape_out, dim3(groups, so.J), 32, 0, st, so, NS);
}

void launch_shape_solve_shared_scale(const SolveArgs& a, double* Gd, double* Zd, double* Cd, double* sums, double* x, int ns,
                                     int groups, cudaStream_t st) {
  SF_NS_SWITCH(ns, (solve_shared_scale_t<NS>(a, Gd, Zd, Cd, sums, x, groups, st)));
}

void launch_shape_solve_shared(const SolveArgs& a, double* Gd, double* Cd, double* sums, double* x, int ns, int groups,
                               cudaStream_t st) {
  SF_NS_SWITCH(ns, (solve_shared_t<NS>(a, Gd, Cd, sums, x, groups, st)));
}

}  // namespace sf
