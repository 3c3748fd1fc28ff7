// C-ABI entry points of BodyFitter.fit / fit_with_known_pose: workspace carving and the
// launch sequence.  Kernels live in fit_kernels.cuh (vertex passes) and solve_kernels.cuh
// (per-instance solves).  Reference: /root/reference/src/smplfitter/pt/bodyfitter.py:283-549.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "fit_fused.cuh"
#include "fit_kernels.cuh"
#include "lite_kernels.cuh"
#include "passes.cuh"
#include "solve_kernels.cuh"
#include "vposed_tc.cuh"

namespace sf {

thread_local char g_err[512] = "";
thread_local const char* g_launch_failure = nullptr;
std::atomic<long long> g_launches{0};
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

struct FitWs {
  double* Gd;
  float *mean, *tT, *tjT, *vwT, *jwT, *vposedT, *R, *R2, *RT, *Pext, *feat, *gpart, *beta, *trans, *refj,
      *skin, *spart, *aT, *ajT, *initjT, *RT4, *zpart, *scale, *mpart, *RT12, *gcfpart, *skin4, *pairfeat;
  double *Zd, *Cd, *sums, *Yd;
  void* tc_scratch;
  void* fq_scratch;
  size_t bytes;
};

// final adjustment: level-parallel kernel unless switched off (SMPLFIT_B200_ADJUST=seq) or a copy part's source is not
// an ancestor-side joint of lower depth (never the case for the SMPL family, where toes copy their parent foot)
static void launch_adjust(const AdjustArgs& aa, int groups, cudaStream_t st) {
  static int par = -1;
  if (par < 0) {
    const char* e = getenv("SMPLFIT_B200_ADJUST");
    par = (e && strcmp(e, "seq") == 0) ? 0 : 1;
  }
  const size_t smem = adjust_par_smem_bytes(aa.t.J, aa.n_adj);
  if (par && smem <= 200 * 1024) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_adjust_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SF_LAUNCH(k_adjust_par, groups, ADJ_WARPS * 32, smem, st, aa);
  } else {
    SF_LAUNCH(k_adjust_solve, groups, 32, 0, st, aa);
  }
}

static int moment_blocks(int V) { return (V + 255) / 256; }
static int shape_nacc(int ns) { return ns * (ns + 1) / 2 + ns + 3 * ns + 3 + 1; }

static FitWs carve(void* base, const smplfit_model_t* m, int64_t B, int has_joints, int has_vw, int has_jw,
                   int has_init) {
  FitWs w{};
  Carver c(base);
  const size_t Bp = roundup((int)B, 32);
  const int V = m->num_vertices, J = m->num_joints, NS = m->fit_ns;
  const int Kp = roundup(m->num_pose_feats, 16);
  w.mean = c.take<float>(3 * Bp);
  w.tT = c.take<float>((size_t)3 * V * Bp);
  w.tjT = c.take<float>((size_t)3 * J * Bp);
  w.vwT = has_vw ? c.take<float>((size_t)V * Bp) : nullptr;
  w.jwT = has_jw ? c.take<float>((size_t)J * Bp) : nullptr;
  w.vposedT = c.take<float>((size_t)3 * V * Bp);
  w.R = c.take<float>((size_t)9 * J * Bp);
  w.R2 = c.take<float>((size_t)9 * J * Bp);
  w.RT = c.take<float>((size_t)J * (12 + 3 * NS) * Bp);
  w.Pext = c.take<float>((size_t)J * 3 * (1 + NS) * Bp);
  w.RT4 = c.take<float>(rt4_floats(m, (int)Bp));
  w.feat = c.take<float>((size_t)Bp * Kp);
  {
    const size_t legacy = (size_t)max_shape_partials(m) * shape_nacc(NS);
    const size_t lite = lite_available(m) ? (size_t)m->n_segments * lite_rows(NS) : 0;
    w.gpart = c.take<float>((legacy > lite ? legacy : lite) * Bp);
  }
  w.RT12 = lite_available(m) ? c.take<float>((size_t)12 * J * Bp) : nullptr;
  w.gcfpart = lite_available(m) ? c.take<float>((size_t)gram_closed_blocks(m) * (NS * (NS + 1) / 2) * Bp) : nullptr;
  w.Yd = lite_available(m) ? c.take<double>((size_t)(3 * J + NS + 3) * Bp) : nullptr;  // Y_k | sums of r | Sb
  w.pairfeat = (lite_available(m) && gram_pairs_scratch_floats(m, (int)Bp) > 0) ? c.take<float>(gram_pairs_scratch_floats(m, (int)Bp)) : nullptr;
  w.Gd = c.take<double>((size_t)shape_nacc(NS) * Bp);
  w.beta = c.take<float>((size_t)NS * Bp);
  w.trans = c.take<float>(3 * Bp);
  w.refj = c.take<float>((size_t)3 * J * Bp);
  w.skin = c.take<float>((size_t)12 * J * Bp);
  w.skin4 = c.take<float>((size_t)12 * J * Bp);
  w.spart = c.take<float>((size_t)m->n_segments * 16 * Bp);
  const bool need_aT = true;  // (fit_with_known_shape always materialises the reference)
  w.aT = need_aT ? c.take<float>((size_t)3 * V * Bp) : nullptr;
  w.ajT = need_aT ? c.take<float>((size_t)3 * J * Bp) : nullptr;
  w.initjT = has_init ? c.take<float>((size_t)3 * J * Bp) : nullptr;
  w.zpart = c.take<float>((size_t)scale_chunks(m) * (NS + 5) * Bp);
  w.Zd = c.take<double>((size_t)(NS + 5) * Bp);
  w.Cd = c.take<double>((size_t)(NS * (NS + 1) / 2 + NS) * Bp);
  w.sums = c.take<double>((size_t)(NS * (NS + 1) / 2 + 2 * NS + 8));
  w.scale = c.take<float>(Bp);
  w.mpart = c.take<float>((size_t)(moment_blocks(V) + 1) * 9 * Bp);
  w.tc_scratch = c.take<char>(vposed_tc_scratch_bytes(m, (int)Bp));
  w.fq_scratch = fit_fused_available(m) ? c.take<char>(fit_fused_scratch_bytes(m, (int)Bp)) : nullptr;
  w.bytes = c.off + 256;
  return w;
}

static TreeTables tables(const smplfit_model_t* m) {
  TreeTables t;
  t.parents = m->parents;
  t.part_kind = m->part_kind;
  t.part_copy_src = m->part_copy_src;
  t.part_flags = m->part_flags;
  t.cas_table = m->cas_table;
  t.cas_count = m->cas_count;
  t.part_seg_begin = m->part_seg_begin;
  t.Jt_ext = m->fit_Jt_ext;
  t.J = m->num_joints;
  t.NS = m->fit_ns;
  t.max_cas = m->max_cas;
  return t;
}

static int check_model(const smplfit_model_t* m) {
  if (!m) return fail(SMPLFIT_ERR_ARG, "model is NULL");
  if (m->num_joints > SMPLFIT_MAX_JOINTS) return fail(SMPLFIT_ERR_UNSUPPORTED, "num_joints > 64");
  if (m->fit_ns < 2 || m->fit_ns > SMPLFIT_MAX_UNKNOWNS)
    return fail(SMPLFIT_ERR_UNSUPPORTED, "number of shape unknowns must be in [2, 17]");
  return SMPLFIT_OK;
}

struct FitCtx {
  const smplfit_model_t* m;
  FitWs w;
  int B, Bp, groups, Kp, n_chunks;
  cudaStream_t st;
  const float *vwT_shape, *jwT_shape;  // shape-stage weights (pt/bodyfitter.py:1018-1028)
  bool has_joints, use_rec;
  bool lite;  // closed-form Gramian + light vertex pass (unweighted shape stage)
  bool fused;         // vertex passes in the epilogue of the blend-shape GEMM (fit_fused.cu): no v_posed^T in HBM
  bool vposed_valid;  // w.vposedT holds the posed template of the current orientations (non-fused consumers)
  bool feats_ready;   // the fused passes' feature rows are maintained by k_front_fused / k_shape_out
  ShapePlan plan;
};

// decide the shape path once the weights are known; wires the matching row layouts into RotArgs
static void choose_shape_path(FitCtx& c, RotArgs& ra) {
  c.lite = lite_enabled(c.m) && c.vwT_shape == nullptr;
  ra.RT12 = c.lite ? c.w.RT12 : nullptr;
  ra.RT4 = c.lite ? nullptr : c.w.RT4;
  ra.rt4_clay = c.plan.kind == 3;
}

static void set_feature_rows(FitCtx& c, SolveArgs& so);

static void run_gemm(FitCtx& c) {
  const smplfit_model_t* m = c.m;
  c.vposed_valid = true;
  if (vposed_tc_run(m, c.w.feat, c.w.vposedT, c.Bp, c.Kp, c.w.tc_scratch, c.st)) return;
  dim3 grid((3 * m->num_vertices + 127) / 128, (c.Bp + 63) / 64);
  SF_LAUNCH(k_vposed_gemm_simt, grid, 256, 0, c.st, m->posedirs_fit, m->v_template_fit, c.w.feat,
            3 * m->num_vertices, c.Kp, c.Bp, c.w.vposedT);
}

static void run_shape(FitCtx& c, int scale_mode, const float* beta_ref, const float* kid_ref,
                      const smplfit_fit_opts_t* o) {
  const smplfit_model_t* m = c.m;
  // the closed-form Gramian needs only the joint transforms (not the vertices)
  if (c.lite) launch_gram_closed(m, c.groups, c.Bp, c.w.RT, c.w.gcfpart, c.w.pairfeat, c.st);
  bool shape_fused = false;
  if (c.lite && c.fused && scale_mode == 0)  // (the scale pass of the final solve reads v_posed^T)
    shape_fused = fit_fused_run(m, 2, c.B, c.Bp, c.w.feat, c.Kp, nullptr, c.w.tT, nullptr, c.w.RT12, nullptr, nullptr, nullptr, 0,
                                c.w.gpart, c.w.fq_scratch, c.feats_ready, c.st);
  if (!shape_fused) run_gemm(c);
  ShapeArgs sa;
  sa.tT = c.w.tT; sa.vwT = c.vwT_shape; sa.vposedT = c.w.vposedT; sa.RT = c.w.RT;
  sa.shapedirs = m->fit_shapedirs; sa.skin_idx = m->skin_idx; sa.skin_w = m->skin_w; sa.order = m->order;
  sa.partials = c.w.gpart; sa.rec = m->fit_rec; sa.RT4 = c.w.RT4; sa.V = m->num_vertices; sa.J = m->num_joints; sa.Bp = c.Bp;
  sa.skin_k = m->skin_k; sa.chunk_len = c.plan.chunk_len; sa.n_chunks = c.plan.n_chunks; sa.chunks_per_cta = c.plan.warps;
  SolveArgs so{};
  if (c.lite) {
    LiteArgs la;
    la.tT = c.w.tT; la.vposedT = c.w.vposedT; la.RT12 = c.w.RT12; la.rec = m->fit_rec; la.seg_start = m->seg_start;
    la.seg_slots = m->seg_slots; la.partials = c.w.gpart; la.n_segments = m->n_segments; la.J = m->num_joints;
    la.Bp = c.Bp; la.segs_per_warp = 1; la.slot_mask = m->fit_slot_mask;
    if (shape_fused) launch_lite_reduce(la, m, c.groups, c.w.Yd, c.st);
    else launch_shape_lite(la, m, c.groups, c.w.Yd, c.st);
    so.lite = 1; so.lite_nl = lite_rows(m->fit_ns); so.n_gcf = gram_closed_blocks(m); so.gcf_part = c.w.gcfpart;
    so.G0 = m->gcf_G0; so.Yd = c.w.Yd;
  } else {
    launch_shape_pass(sa, m->fit_ns, c.groups, c.plan, c.st);
  }
  set_feature_rows(c, so);
  so.partials = c.w.gpart; so.Pext = c.w.Pext; so.RT = c.w.RT;
  so.tjT = c.has_joints ? c.w.tjT : nullptr; so.jwT = c.jwT_shape;
  so.beta_ref = beta_ref; so.kid_ref = kid_ref;
  so.beta = c.w.beta; so.trans = c.w.trans; so.refj = c.w.refj; so.skin = c.w.skin; so.skin4 = c.w.skin4;
  so.wS = m->fit_wS; so.wsum = m->fit_wsum;
  so.n_chunks = c.lite ? m->n_segments : c.plan.n_partials; so.J = m->num_joints; so.S = m->num_betas; so.Bp = c.Bp; so.B = c.B;
  so.V = m->num_vertices; so.weighted = c.vwT_shape != nullptr;
  so.sa_closed_form = ((c.use_rec || c.lite) && c.vwT_shape == nullptr && m->fit_wS != nullptr) ? 1 : 0;
  so.reg = o->beta_regularizer; so.reg2 = o->beta_regularizer2; so.kid_reg = o->kid_regularizer;
  so.scale_mode = scale_mode; so.zpartials = c.w.zpart; so.n_zchunks = scale_chunks(m);
  so.scale_reg = o->scale_regularizer; so.scale_out = c.w.scale;
  if (scale_mode != 0) {
    ShapeArgs sz = sa;
    sz.partials = c.w.zpart;
    launch_scale_pass(sz, m->fit_ns, scale_mode, c.groups, c.st);
    // the scale pass writes its own partial layout: point the solve at it
    if (o->share_beta) {
      const int ne = m->fit_ns * (m->fit_ns + 1) / 2 + m->fit_ns;
      launch_shape_solve_shared_scale(so, c.w.Gd, c.w.Zd, c.w.Cd, c.w.sums, c.w.sums + ne, m->fit_ns, c.groups, c.st);
    } else {
      launch_shape_solve_scale(so, c.w.Gd, c.w.Zd, m->fit_ns, c.groups, c.st);
    }
  } else if (o->share_beta) {
    const int ne = m->fit_ns * (m->fit_ns + 1) / 2 + m->fit_ns;
    launch_shape_solve_shared(so, c.w.Gd, c.w.Cd, c.w.sums, c.w.sums + ne, m->fit_ns, c.groups, c.st);
  } else {
    launch_shape_solve(so, c.w.Gd, m->fit_ns, c.groups, c.st);
  }
}

// statistics of (targets, reference) for the rotation stage; ref_mode as in k_stats
static void run_stats(FitCtx& c, int ref_mode, const float* ca0T, const float* aT_in, float* aT_out) {
  const smplfit_model_t* m = c.m;
  StatsArgs s;
  s.tT = c.w.tT; s.vwT = c.w.vwT; s.ct0 = c.w.tjT; s.ca0 = ca0T; s.ca0_const = m->J_template;
  s.vposedT = c.w.vposedT; s.beta = c.w.beta; s.skin = c.w.skin; s.aT_in = aT_in; s.aT_out = aT_out;
  s.partials = c.w.spart; s.template_mesh = m->template_mesh; s.shapedirs = m->fit_shapedirs;
  s.skin_idx = m->skin_idx; s.skin_w = m->skin_w; s.order = m->order; s.seg_start = m->seg_start;
  s.seg_part = m->seg_part; s.part_flags = m->part_flags; s.n_segments = m->n_segments; s.Bp = c.Bp;
  s.ns = m->fit_ns; s.skin_k = m->skin_k; s.all_segments = (aT_out != nullptr);
  StatsRecArgs r;
  r.tT = s.tT; r.vwT = s.vwT; r.ct0 = s.ct0; r.ca0 = s.ca0; r.ca0_const = s.ca0_const; r.vposedT = s.vposedT;
  r.beta = s.beta; r.skin = s.skin; r.aT_in = aT_in; r.aT_out = aT_out; r.partials = s.partials; r.rec = m->fit_rec;
  r.template_fit = m->template_mesh_fit; r.seg_start = s.seg_start; r.seg_part = s.seg_part;
  r.part_flags = s.part_flags; r.n_segments = s.n_segments; r.Bp = c.Bp; r.J = m->num_joints;
  r.all_segments = s.all_segments; r.segs_per_warp = 2;
  if (ref_mode == 0 && stats_lite_enabled(m) && m->template_mesh_fit != nullptr) {
    StatsLiteArgs l{};
    l.tT = c.w.tT; l.vwT = c.w.vwT; l.ct0 = c.w.tjT; l.partials = c.w.spart; l.seg_start = m->seg_start;
    l.seg_part = m->seg_part; l.part_flags = m->part_flags; l.n_segments = m->n_segments; l.Bp = c.Bp;
    l.J = m->num_joints;
    launch_stats_tmpl(l, m, m->template_mesh_fit, m->J_template, c.groups, c.st);
    return;
  }
  if (ref_mode == 1 && c.fused &&
      fit_fused_run(m, 3, c.B, c.Bp, c.w.feat, c.Kp, c.w.beta, c.w.tT, c.w.vwT, c.w.skin4, c.w.tjT, ca0T, aT_out,
                    aT_out != nullptr, c.w.spart, c.w.fq_scratch, c.feats_ready, c.st))
    return;
  if (ref_mode == 1 && !c.vposed_valid) run_gemm(c);
  if (ref_mode == 1 && stats_lite_enabled(m)) {
    StatsLiteArgs l;
    l.tT = c.w.tT; l.vwT = c.w.vwT; l.ct0 = c.w.tjT; l.ca0 = ca0T; l.vposedT = c.w.vposedT; l.beta = c.w.beta;
    l.skin4 = c.w.skin4; l.aT_out = aT_out; l.partials = c.w.spart; l.rec = m->fit_rec; l.seg_start = m->seg_start;
    l.seg_part = m->seg_part; l.part_flags = m->part_flags; l.n_segments = m->n_segments; l.Bp = c.Bp;
    l.J = m->num_joints; l.all_segments = (aT_out != nullptr); l.segs_per_warp = 1;
    l.slot_mask = m->fit_slot_mask;
    launch_stats_lite(l, m, c.groups, c.st);
    return;
  }
  const bool use_rec = m->fit_rec != nullptr && m->skin_k <= 4 && m->template_mesh_fit != nullptr;
  launch_stats(s, r, m->fit_ns, ref_mode, c.w.vwT != nullptr, use_rec, c.groups, c.st);
}

static void run_regress(FitCtx& c, const float* X, float* out) {
  const smplfit_model_t* m = c.m;
  if (m->jreg_ptr != nullptr && m->jreg_idx != nullptr && m->jreg_val != nullptr) {
    const long long nw = (long long)c.groups * m->num_joints;
    SF_LAUNCH(k_regress_csr, (int)((nw + 3) / 4), 128, 0, c.st, m->jreg_ptr, m->jreg_idx, m->jreg_val, X, m->num_joints, c.Bp, out);
    return;
  }
  const int jblocks = (m->num_joints + 7) / 8;
  const long long warps = (long long)c.groups * jblocks;
  SF_LAUNCH(k_regress, (int)((warps + 3) / 4), 128, 0, c.st, m->J_regressor_fit, X, m->num_vertices,
            m->num_joints, c.Bp, out);
}

// rotation fit per (instance, part), then the pose-dependent front of the next shape solve
// feature-row pointers of the fused passes into a SolveArgs (k_shape_out writes the unknowns' columns)
static void set_feature_rows(FitCtx& c, SolveArgs& so) {
  so.fq_hi = so.fq_lo = nullptr;
  so.fq_kf = so.fq_p = 0;
  if (!c.feats_ready) return;
  void *hi, *lo;
  fit_fused_feature_rows(c.m, c.Bp, c.w.fq_scratch, &hi, &lo);
  so.fq_hi = reinterpret_cast<__half*>(hi); so.fq_lo = reinterpret_cast<__half*>(lo);
  so.fq_kf = c.m->fq_kf; so.fq_p = c.m->num_pose_feats;
}

static void run_rot(FitCtx& c, const RotArgs& ra, bool fit) {
  const int J = c.m->num_joints;
  c.vposed_valid = false;
  if (fit) {
    // warps per (instance group, part): the segment partials of a part are summed by several warps when parts are long
    const int spp = c.m->n_segments / (J > 0 ? J : 1);
    if (spp >= 8) SF_LAUNCH(k_rot_fit<4>, dim3(c.groups, J), 128, 0, c.st, ra);
    else if (spp >= 4) SF_LAUNCH(k_rot_fit<2>, dim3(c.groups, J), 64, 0, c.st, ra);
    else SF_LAUNCH(k_rot_fit<1>, dim3(c.groups, J), 32, 0, c.st, ra);
  }
  c.feats_ready = false;
  if (c.fused && ra.RT4 == nullptr && c.w.fq_scratch != nullptr) {
    // closed-form path: relative rotations, row tables, kinematic-chain columns and the fused passes' feature rows in
    // one kernel
    // 12 warps (one kinematic-chain column each for SMPL) when the grid is a single wave, 8 (smaller CTAs, two per SM)
    // when there are more instance groups than SMs
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    int ffw = (c.groups <= sms && c.m->fit_ns + 1 > 8) ? 12 : 8;
    if (ffw == 12 && front_fused_smem_bytes(J, c.m->fq_kf, 12) > 200 * 1024) ffw = 8;  // (measured: only pays when all 12 chains fit)
    int fkw = ffw;
    while (fkw > 1 && front_fused_smem_bytes(J, c.m->fq_kf, fkw) > 200 * 1024) --fkw;
    const size_t smem = front_fused_smem_bytes(J, c.m->fq_kf, fkw);
    if (smem <= 200 * 1024) {
      RotArgs rf = ra;
      void *hi, *lo;
      fit_fused_feature_rows(c.m, c.Bp, c.w.fq_scratch, &hi, &lo);
      rf.fq_hi = reinterpret_cast<__half*>(hi); rf.fq_lo = reinterpret_cast<__half*>(lo);
      rf.fq_kf = c.m->fq_kf; rf.fk_warps = fkw;
      if (ffw == 12) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_front_fused<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        SF_LAUNCH(k_front_fused<12>, c.groups, 12 * 32, smem, c.st, rf);
      } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_front_fused<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        SF_LAUNCH(k_front_fused<8>, c.groups, 8 * 32, smem, c.st, rf);
      }
      c.feats_ready = true;
      return;
    }
  }
  SF_LAUNCH(k_front_rel, dim3(c.groups, J), 32, 0, c.st, ra);
  SF_LAUNCH(k_front_fk, dim3(c.groups, c.m->fit_ns + 1), 32, 0, c.st, ra);
}

template <int C>
static void run_transpose(FitCtx& c, const float* src, int N, const int32_t* inv, const float* mean, float* dst) {
  if (C == 3 && N >= 4 * TV_N) {  // vertex arrays: the larger-tile kernel
    SF_LAUNCH(k_transpose_v, dim3((N + TV_N - 1) / TV_N, c.groups), 256, 0, c.st, src, N, c.B, c.Bp, inv, mean, dst);
    return;
  }
  dim3 grid((N + 31) / 32, c.groups), block(32, 8);
  SF_LAUNCH(k_transpose<C>, grid, block, 0, c.st, src, N, c.B, c.Bp, inv, mean, dst);
}

}  // namespace sf

using namespace sf;

extern "C" const char* smplfit_version(void) { return "smplfit_b200 0.1.0 (sm_100a)"; }
extern "C" const char* smplfit_last_error(void) { return g_err; }
extern "C" int smplfit_profile(int enable) {
  g_prof_on = enable != 0;
  return SMPLFIT_OK;
}
extern "C" int smplfit_profile_report(char* out, size_t cap) {
  // aggregate by kernel name: "name\tlaunches\ttotal_ms\n"
  struct Agg { const char* name; int n; double ms; };
  std::vector<Agg> agg;
  for (auto& r : g_prof) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    bool found = false;
    for (auto& a : agg)
      if (a.name == r.name || strcmp(a.name, r.name) == 0) { a.n++; a.ms += ms; found = true; break; }
    if (!found) agg.push_back(Agg{r.name, 1, (double)ms});
  }
  g_prof.clear();
  size_t off = 0;
  if (out && cap) out[0] = 0;
  for (auto& a : agg) {
    char line[256];
    const int len = snprintf(line, sizeof(line), "%s\t%d\t%.6f\n", a.name, a.n, a.ms);
    if (out && off + len + 1 < cap) { memcpy(out + off, line, len + 1); off += len; }
  }
  return SMPLFIT_OK;
}
extern "C" size_t smplfit_struct_size(int which) {
  return which == 0 ? sizeof(smplfit_model_t) : which == 1 ? sizeof(smplfit_fit_opts_t) : 0;
}
extern "C" int64_t smplfit_launch_count(int reset) {
  const long long v = reset ? g_launches.exchange(0) : g_launches.load();
  return (int64_t)v;
}

extern "C" int smplfit_set_share_beta_allreduce(smplfit_allreduce_fn fn, void* user, int64_t global_batch) {
  if (fn != nullptr && global_batch <= 0) return fail(SMPLFIT_ERR_ARG, "global_batch must be positive");
  set_share_beta_allreduce(fn, user, global_batch);
  return SMPLFIT_OK;
}

extern "C" size_t smplfit_fit_workspace_bytes(const smplfit_model_t* m, int64_t batch, const smplfit_fit_opts_t* o,
                                              int has_joints, int has_vw, int has_jw) {
  (void)o;
  if (check_model(m) != SMPLFIT_OK || batch <= 0) return 0;
  return carve(nullptr, m, batch, has_joints, has_vw, has_jw, /*has_init=*/1).bytes;
}

// the launch sequence of one fit, enqueued kernel by kernel on `stream`
static int fit_direct(const smplfit_model_t* m, int64_t batch, const float* target_vertices,
                      const float* target_joints, const float* vertex_weights, const float* joint_weights,
                      const float* beta_reg_reference, const float* kid_reg_reference,
                      const float* init_vertices, const float* init_joints, const float* init_orientations,
                      const smplfit_fit_opts_t* o, float* out_pose_rotvecs, float* out_shape_betas,
                      float* out_trans, float* out_orientations, float* out_rel_orientations,
                      float* out_kid_factor, float* out_scale_corr, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (int e = check_model(m)) return e;
  if (!o || !target_vertices || !out_shape_betas || !out_trans || !out_orientations)
    return fail(SMPLFIT_ERR_ARG, "missing required pointer");
  if (batch <= 0 || batch > (1 << 24)) return fail(SMPLFIT_ERR_ARG, "batch out of range");
  if (o->num_iter < 1) return fail(SMPLFIT_ERR_ARG, "num_iter must be >= 1");
  if (o->scale_mode < 0 || o->scale_mode > 2) return fail(SMPLFIT_ERR_ARG, "bad scale_mode");
  if (o->scale_mode != 0 && !out_scale_corr) return fail(SMPLFIT_ERR_ARG, "scale_corr output required");
  if ((o->enable_kid != 0) != (m->fit_ns == m->num_betas + 1))
    return fail(SMPLFIT_ERR_ARG, "enable_kid does not match the fitter tables");
  const bool has_init = init_vertices != nullptr;
  if (has_init && (!init_joints || !init_orientations)) return fail(SMPLFIT_ERR_ARG, "incomplete initial guess");
  const bool has_joints = target_joints != nullptr;
  FitCtx c;
  c.m = m;
  c.B = (int)batch;
  c.Bp = roundup(c.B, 32);
  c.groups = c.Bp / 32;
  c.Kp = roundup(m->num_pose_feats, 16);
  c.n_chunks = (m->num_vertices + m->chunk_len - 1) / m->chunk_len;
  c.st = reinterpret_cast<cudaStream_t>(stream);
  c.has_joints = has_joints;
  c.plan = plan_shape_pass(m, c.groups);
  c.use_rec = c.plan.use_rec;
  c.fused = fit_fused_available(m);
  c.vposed_valid = false;
  c.feats_ready = false;
  c.w = carve(workspace, m, batch, has_joints, vertex_weights != nullptr, joint_weights != nullptr, has_init);
  if (!workspace || c.w.bytes > workspace_bytes) return fail(SMPLFIT_ERR_WORKSPACE, "workspace too small");
  FitWs& w = c.w;
  const int V = m->num_vertices, J = m->num_joints;

  // -- re-layout + centring (pt/bodyfitter.py:355-361) --
  SF_LAUNCH(k_mean, c.Bp, MEAN_THREADS, 0, c.st, target_vertices, target_joints, V, J, c.B, c.Bp, w.mean);
  run_transpose<3>(c, target_vertices, V, m->inv_order, w.mean, w.tT);
  if (vertex_weights) run_transpose<1>(c, vertex_weights, V, m->inv_order, nullptr, w.vwT);
  if (joint_weights) run_transpose<1>(c, joint_weights, J, nullptr, nullptr, w.jwT);
  if (has_joints) run_transpose<3>(c, target_joints, J, nullptr, w.mean, w.tjT);
  else run_regress(c, w.tT, w.tjT);
  // shape-stage weight rule (pt/bodyfitter.py:1018-1028)
  c.vwT_shape = o->shape_weights ? w.vwT : nullptr;
  c.jwT_shape = (o->shape_weights && has_joints) ? w.jwT : nullptr;

  RotArgs ra{};
  ra.partials = w.spart; ra.tjT = w.tjT; ra.jwT = w.jwT; ra.R_new = w.R; ra.RT = w.RT; ra.Pext = w.Pext;
  ra.feat = w.feat; ra.t = tables(m); ra.B = c.B; ra.Bp = c.Bp; ra.Kp = c.Kp;
  choose_shape_path(c, ra);
  // -- first rotation fit (pt/bodyfitter.py:363-394) --
  if (has_init) {
    run_transpose<3>(c, init_vertices, V, m->inv_order, nullptr, w.aT);
    run_transpose<3>(c, init_joints, J, nullptr, nullptr, w.initjT);
    dim3 g9((9 * J + 31) / 32, c.groups), blk(32, 8);
    SF_LAUNCH(k_transpose<1>, g9, blk, 0, c.st, init_orientations, 9 * J, c.B, c.Bp, (const int32_t*)nullptr,
              (const float*)nullptr, w.R2);  // [9J][Bp]
    run_stats(c, 2, w.initjT, w.aT, nullptr);
    const float* aj = w.initjT;
    if (!has_joints) {
      run_regress(c, w.aT, w.ajT);
      aj = w.ajT;
    }
    ra.ajT = aj; ra.aj_const = nullptr; ra.ca0T = w.initjT; ra.ca0_const = nullptr; ra.R_old = w.R2;
  } else {
    run_stats(c, 0, nullptr, nullptr, nullptr);
    ra.ajT = nullptr; ra.aj_const = has_joints ? m->J_template : m->template_joints_regressed;
    ra.ca0T = nullptr; ra.ca0_const = m->J_template; ra.R_old = nullptr;
  }
  run_rot(c, ra, true);

  // -- alternate shape and rotation fits (pt/bodyfitter.py:399-461) --
  const float* R_final = w.R;
  for (int it = 0; it < o->num_iter; ++it) {
    const bool last = (it == o->num_iter - 1);
    run_shape(c, last ? o->scale_mode : 0, beta_reg_reference, kid_reg_reference, o);
    if (last && !o->final_adjust_rots) break;
    float* aT_out = has_joints ? nullptr : w.aT;
    run_stats(c, 1, w.refj, nullptr, aT_out);
    const float* aj = w.refj;
    if (!has_joints) {
      run_regress(c, w.aT, w.ajT);
      aj = w.ajT;
    }
    if (!last) {
      ra.ajT = aj; ra.aj_const = nullptr; ra.ca0T = w.refj; ra.ca0_const = nullptr; ra.R_old = w.R; ra.R_new = w.R;
      run_rot(c, ra, true);
    } else {
      AdjustArgs aa;
      aa.partials = w.spart; aa.tjT = w.tjT; aa.ajT = aj; aa.refj = w.refj; aa.jwT = w.jwT; aa.R_prev = w.R;
      aa.beta = w.beta; aa.trans = w.trans; aa.R_out = w.R2; aa.t = tables(m); aa.Bp = c.Bp; aa.n_adj = (m->n_adjustable > 0 && m->n_adjustable <= J) ? m->n_adjustable : J;
      aa.scale = o->scale_mode ? w.scale : nullptr; aa.scale_mode = o->scale_mode;
      launch_adjust(aa, c.Bp / 32, c.st);
      R_final = w.R2;
    }
  }
  OutputArgs oa;
  oa.R_final = R_final;
  oa.R_rel_src = (o->want_pose_rotvecs || o->want_rel_orient) ? R_final : w.R;
  oa.beta = w.beta; oa.trans = w.trans; oa.mean = w.mean; oa.parents = m->parents;
  oa.pose_rotvecs = o->want_pose_rotvecs ? out_pose_rotvecs : nullptr;
  oa.shape_betas = out_shape_betas; oa.out_trans = out_trans; oa.orientations = out_orientations;
  oa.rel_orient = out_rel_orientations; oa.kid = o->enable_kid ? out_kid_factor : nullptr;
  oa.scale = o->scale_mode ? w.scale : nullptr; oa.scale_corr = o->scale_mode ? out_scale_corr : nullptr;
  oa.scale_mode = o->scale_mode;
  oa.J = J; oa.S = m->num_betas; oa.NS = m->fit_ns; oa.B = c.B; oa.Bp = c.Bp;
  SF_LAUNCH(k_output, dim3(c.Bp / 32, J), 32, 0, c.st, oa);
  SF_CHECK_LAST();
  return SMPLFIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// CUDA-graph replay of repeated fits.  A fit is ~40 dependent launches; at small batches the host cannot enqueue them
// as fast as the GPU runs them.  A call whose arguments (model tables, options, batch, every pointer, workspace) were
// seen before is captured once into a graph and replayed from then on with a single cudaGraphLaunch (the kernels, their
// tensor maps and pointers are baked into the nodes; the data they point to is read at run time as usual).  Calls that
// do not repeat pay one hash.  Not used while profiling, inside a caller's own capture, with the share_beta
// all-reduce hook installed, or when SMPLFIT_B200_GRAPH=0.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct GraphEntry {
  uint64_t key = 0;
  cudaGraphExec_t exec = nullptr;
  long long kernels = 0;
  unsigned long long last_use = 0;
  int seen = 0;
};
constexpr int kGraphSlots = 32;
GraphEntry g_graphs[kGraphSlots];
unsigned long long g_graph_clock = 0;
std::atomic<long long> g_graph_replays{0};
std::atomic<long long> g_graph_stat[4];  // first sights, captures started, captures failed, instantiated
std::mutex g_graph_mutex;

inline void fnv(uint64_t& h, const void* p, size_t n) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
}
bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SMPLFIT_B200_GRAPH");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}
}  // namespace

extern "C" int64_t smplfit_graph_replays(int reset) {
  const long long v = reset ? g_graph_replays.exchange(0) : g_graph_replays.load();
  return (int64_t)v;
}
extern "C" void smplfit_graph_stats(int64_t* out4) {
  for (int i = 0; i < 4; ++i) out4[i] = (int64_t)g_graph_stat[i].load();
}

extern "C" int smplfit_fit(const smplfit_model_t* m, int64_t batch, const float* target_vertices,
                           const float* target_joints, const float* vertex_weights, const float* joint_weights,
                           const float* beta_reg_reference, const float* kid_reg_reference,
                           const float* init_vertices, const float* init_joints, const float* init_orientations,
                           const smplfit_fit_opts_t* o, float* out_pose_rotvecs, float* out_shape_betas,
                           float* out_trans, float* out_orientations, float* out_rel_orientations,
                           float* out_kid_factor, float* out_scale_corr, void* workspace, size_t workspace_bytes,
                           void* stream) {
  auto direct_on = [&](void* s) {
    return fit_direct(m, batch, target_vertices, target_joints, vertex_weights, joint_weights, beta_reg_reference,
                      kid_reg_reference, init_vertices, init_joints, init_orientations, o, out_pose_rotvecs, out_shape_betas,
                      out_trans, out_orientations, out_rel_orientations, out_kid_factor, out_scale_corr, workspace,
                      workspace_bytes, s);
  };
  auto direct = [&]() { return direct_on(stream); };
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (!m || !o || !graphs_enabled() || g_prof_on || share_beta_allreduce_installed() ||
      cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
    return direct();
  uint64_t key = 1469598103934665603ull;
  fnv(key, m, sizeof(*m));
  fnv(key, o, sizeof(*o));
  const void* ptrs[] = {target_vertices, target_joints, vertex_weights, joint_weights, beta_reg_reference, kid_reg_reference,
                        init_vertices, init_joints, init_orientations, out_pose_rotvecs, out_shape_betas, out_trans,
                        out_orientations, out_rel_orientations, out_kid_factor, out_scale_corr, workspace};
  fnv(key, ptrs, sizeof(ptrs));
  fnv(key, &batch, sizeof(batch));
  fnv(key, &workspace_bytes, sizeof(workspace_bytes));
  int dev = 0;
  cudaGetDevice(&dev);
  fnv(key, &dev, sizeof(dev));
  if (key == 0) key = 1;
  std::lock_guard<std::mutex> lock(g_graph_mutex);
  GraphEntry* hit = nullptr;
  GraphEntry* victim = &g_graphs[0];
  for (auto& e : g_graphs) {
    if (e.key == key) { hit = &e; break; }
    if (e.last_use < victim->last_use) victim = &e;
  }
  ++g_graph_clock;
  if (hit && hit->exec) {
    hit->last_use = g_graph_clock;
    if (cudaGraphLaunch(hit->exec, st) == cudaSuccess) {
      g_launches.fetch_add(hit->kernels, std::memory_order_relaxed);
      g_graph_replays.fetch_add(1, std::memory_order_relaxed);
      return SMPLFIT_OK;
    }
    cudaGetLastError();
    cudaGraphExecDestroy(hit->exec);
    *hit = GraphEntry{};
    return direct();
  }
  if (!hit) {  // first sight: remember the call, run it directly
    g_graph_stat[0]++;
    if (victim->exec) cudaGraphExecDestroy(victim->exec);
    *victim = GraphEntry{};
    victim->key = key;
    victim->seen = 1;
    victim->last_use = g_graph_clock;
    return direct();
  }
  // second sight: capture the launch sequence, instantiate, launch
  hit->last_use = g_graph_clock;
  hit->seen++;
  const long long before = g_launches.load();
  g_graph_stat[1]++;
  // the sequence is captured on a private stream (the caller's may be the legacy default stream, which cannot capture);
  // the instantiated graph is then launched on the caller's stream
  static thread_local cudaStream_t cap = nullptr;
  static thread_local int cap_dev = -1;
  if (cap == nullptr || cap_dev != dev) {
    if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      cap = nullptr;
      return direct();
    }
    cap_dev = dev;
  }
  if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return direct();
  }
  const int rc = direct_on(cap);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
  const long long kernels = g_launches.load() - before;
  if (rc != SMPLFIT_OK || ce != cudaSuccess || graph == nullptr) {
    g_graph_stat[2]++;
    if (getenv("SMPLFIT_B200_GRAPH_TRACE"))
      fprintf(stderr, "smplfit graph capture failed: rc=%d (%s) end=%s\n", rc, g_err, cudaGetErrorString(ce));
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    g_launches.store(before);
    *hit = GraphEntry{};
    return rc != SMPLFIT_OK ? rc : direct();  // (nothing was executed by the failed capture)
  }
  cudaGraphExec_t exec = nullptr;
  if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess || exec == nullptr) {
    cudaGetLastError();
    cudaGraphDestroy(graph);
    g_launches.store(before);
    *hit = GraphEntry{};
    return direct();
  }
  cudaGraphDestroy(graph);
  g_graph_stat[3]++;
  hit->exec = exec;
  hit->kernels = kernels;
  if (cudaGraphLaunch(exec, st) != cudaSuccess) {
    cudaGetLastError();
    cudaGraphExecDestroy(exec);
    g_launches.store(before);
    *hit = GraphEntry{};
    return direct();
  }
  g_graph_replays.fetch_add(1, std::memory_order_relaxed);
  return SMPLFIT_OK;
}

extern "C" int smplfit_fit_known_pose(const smplfit_model_t* m, int64_t batch, const float* glob_rotmats,
                                      const float* target_vertices, const float* target_joints,
                                      const float* vertex_weights, const float* joint_weights,
                                      const float* beta_reg_reference, const float* kid_reg_reference,
                                      const smplfit_fit_opts_t* o, float* out_shape_betas, float* out_trans,
                                      float* out_rel_orientations, float* out_kid_factor, float* out_scale_corr,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_model(m)) return e;
  if (!o || !glob_rotmats || !target_vertices || !out_shape_betas || !out_trans)
    return fail(SMPLFIT_ERR_ARG, "missing required pointer");
  if (batch <= 0 || batch > (1 << 24)) return fail(SMPLFIT_ERR_ARG, "batch out of range");
  if (o->scale_mode < 0 || o->scale_mode > 2) return fail(SMPLFIT_ERR_ARG, "bad scale_mode");
  if (o->scale_mode != 0 && !out_scale_corr) return fail(SMPLFIT_ERR_ARG, "scale_corr output required");
  if ((o->enable_kid != 0) != (m->fit_ns == m->num_betas + 1))
    return fail(SMPLFIT_ERR_ARG, "enable_kid does not match the fitter tables");
  const bool has_joints = target_joints != nullptr;
  FitCtx c;
  c.m = m;
  c.B = (int)batch;
  c.Bp = roundup(c.B, 32);
  c.groups = c.Bp / 32;
  c.Kp = roundup(m->num_pose_feats, 16);
  c.n_chunks = (m->num_vertices + m->chunk_len - 1) / m->chunk_len;
  c.st = reinterpret_cast<cudaStream_t>(stream);
  c.has_joints = has_joints;
  c.plan = plan_shape_pass(m, c.groups);
  c.use_rec = c.plan.use_rec;
  c.fused = fit_fused_available(m);
  c.vposed_valid = false;
  c.feats_ready = false;
  c.w = carve(workspace, m, batch, has_joints, vertex_weights != nullptr, joint_weights != nullptr, 1);
  if (!workspace || c.w.bytes > workspace_bytes) return fail(SMPLFIT_ERR_WORKSPACE, "workspace too small");
  FitWs& w = c.w;
  const int V = m->num_vertices, J = m->num_joints;
  SF_LAUNCH(k_mean, c.Bp, MEAN_THREADS, 0, c.st, target_vertices, target_joints, V, J, c.B, c.Bp, w.mean);
  run_transpose<3>(c, target_vertices, V, m->inv_order, w.mean, w.tT);
  if (vertex_weights) run_transpose<1>(c, vertex_weights, V, m->inv_order, nullptr, w.vwT);
  if (joint_weights) run_transpose<1>(c, joint_weights, J, nullptr, nullptr, w.jwT);
  if (has_joints) run_transpose<3>(c, target_joints, J, nullptr, w.mean, w.tjT);
  c.vwT_shape = o->shape_weights ? w.vwT : nullptr;
  c.jwT_shape = (o->shape_weights && has_joints) ? w.jwT : nullptr;
  run_transpose<1>(c, glob_rotmats, 9 * J, nullptr, nullptr, w.R);
  RotArgs ra{};
  ra.partials = nullptr; ra.tjT = nullptr; ra.ajT = nullptr; ra.aj_const = nullptr; ra.ca0T = nullptr;
  ra.ca0_const = nullptr; ra.jwT = nullptr; ra.R_old = nullptr; ra.R_new = w.R; ra.RT = w.RT; ra.Pext = w.Pext;
  ra.feat = w.feat; ra.t = tables(m); ra.B = c.B; ra.Bp = c.Bp; ra.Kp = c.Kp;
  choose_shape_path(c, ra);
  run_rot(c, ra, false);
  run_shape(c, o->scale_mode, beta_reg_reference, kid_reg_reference, o);
  // orientations output is not part of this method's result; reuse the scratch R2 for it
  OutputArgs oa;
  oa.R_final = w.R; oa.R_rel_src = w.R; oa.beta = w.beta; oa.trans = w.trans; oa.mean = w.mean;
  oa.parents = m->parents; oa.pose_rotvecs = nullptr; oa.shape_betas = out_shape_betas; oa.out_trans = out_trans;
  oa.orientations = nullptr; oa.rel_orient = out_rel_orientations;
  oa.kid = o->enable_kid ? out_kid_factor : nullptr;
  // pt/bodyfitter.py:644: the mean is added back unscaled here; scale_corr is still reported
  oa.scale = o->scale_mode ? w.scale : nullptr; oa.scale_corr = o->scale_mode ? out_scale_corr : nullptr;
  oa.scale_mode = 0;
  oa.J = J; oa.S = m->num_betas; oa.NS = m->fit_ns; oa.B = c.B; oa.Bp = c.Bp;
  SF_LAUNCH(k_output, dim3(c.Bp / 32, J), 32, 0, c.st, oa);
  SF_CHECK_LAST();
  return SMPLFIT_OK;
}

// Test hook: the pose-blend-shape contraction alone (tests/test_gpu_gemm.py compares the tcgen05
// path with the FP32 SIMT kernel and a float64 host product).  feat is [Bp][Kp] row-major.
extern "C" size_t smplfit_debug_vposed_scratch_bytes(const smplfit_model_t* m, int Bp) {
  return vposed_tc_scratch_bytes(m, Bp) + 256;
}
extern "C" int smplfit_debug_vposed(const smplfit_model_t* m, const float* feat, int Bp, int use_tc, float* out,
                                    void* scratch, void* stream) {
  if (!m || !feat || !out || Bp <= 0 || Bp % 32) return fail(SMPLFIT_ERR_ARG, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Kp = roundup(m->num_pose_feats, 16);
  if (use_tc) {
    if (!vposed_tc_run(m, feat, out, Bp, Kp, scratch, st)) return fail(SMPLFIT_ERR_UNSUPPORTED, "tcgen05 path unavailable");
  } else {
    dim3 grid((3 * m->num_vertices + 127) / 128, (Bp + 63) / 64);
    SF_LAUNCH(k_vposed_gemm_simt, grid, 256, 0, st, m->posedirs_fit, m->v_template_fit, feat, 3 * m->num_vertices,
              Kp, Bp, out);
  }
  SF_CHECK_LAST();
  return SMPLFIT_OK;
}

extern "C" int smplfit_fit_known_shape(const smplfit_model_t* m, int64_t batch, const float* shape_betas, int n_betas,
                                       const float* kid_factor, const float* target_vertices,
                                       const float* target_joints, const float* vertex_weights,
                                       const float* joint_weights, const float* init_vertices,
                                       const float* init_joints, const float* init_orientations,
                                       const smplfit_fit_opts_t* o, float* out_pose_rotvecs, float* out_trans,
                                       float* out_orientations, float* out_rel_orientations, float* out_scale_corr,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_model(m)) return e;
  if (!o || !target_vertices || !init_vertices || !init_joints || !init_orientations || !out_trans || !out_orientations)
    return fail(SMPLFIT_ERR_ARG, "missing required pointer");
  if (batch <= 0 || batch > (1 << 24)) return fail(SMPLFIT_ERR_ARG, "batch out of range");
  if (o->num_iter < 1) return fail(SMPLFIT_ERR_ARG, "num_iter must be >= 1");
  if (o->scale_mode != 0 && o->scale_mode != 2) return fail(SMPLFIT_ERR_ARG, "fit_with_known_shape only knows scale_fit");
  if (o->scale_mode != 0 && !out_scale_corr) return fail(SMPLFIT_ERR_ARG, "scale_corr output required");
  if (kid_factor != nullptr && m->fit_ns != m->num_betas + 1)
    return fail(SMPLFIT_ERR_UNSUPPORTED, "kid_factor needs a fitter built with enable_kid=True");
  const bool has_joints = target_joints != nullptr;
  FitCtx c;
  c.m = m;
  c.B = (int)batch;
  c.Bp = roundup(c.B, 32);
  c.groups = c.Bp / 32;
  c.Kp = roundup(m->num_pose_feats, 16);
  c.n_chunks = (m->num_vertices + m->chunk_len - 1) / m->chunk_len;
  c.st = reinterpret_cast<cudaStream_t>(stream);
  c.has_joints = has_joints;
  c.plan = plan_shape_pass(m, c.groups);
  c.use_rec = c.plan.use_rec;
  c.fused = fit_fused_available(m);
  c.vposed_valid = false;
  c.feats_ready = false;
  c.w = carve(workspace, m, batch, has_joints, vertex_weights != nullptr, joint_weights != nullptr, 1);
  if (!workspace || c.w.bytes > workspace_bytes) return fail(SMPLFIT_ERR_WORKSPACE, "workspace too small");
  FitWs& w = c.w;
  const int V = m->num_vertices, J = m->num_joints;
  SF_LAUNCH(k_mean, c.Bp, MEAN_THREADS, 0, c.st, target_vertices, target_joints, V, J, c.B, c.Bp, w.mean);
  run_transpose<3>(c, target_vertices, V, m->inv_order, w.mean, w.tT);
  if (vertex_weights) run_transpose<1>(c, vertex_weights, V, m->inv_order, nullptr, w.vwT);
  if (joint_weights) run_transpose<1>(c, joint_weights, J, nullptr, nullptr, w.jwT);
  if (has_joints) run_transpose<3>(c, target_joints, J, nullptr, w.mean, w.tjT);
  else run_regress(c, w.tT, w.tjT);
  c.vwT_shape = nullptr;
  c.jwT_shape = nullptr;
  SF_LAUNCH(k_set_shape, c.groups, 32, 0, c.st, shape_betas, n_betas, kid_factor, m->num_betas, m->fit_ns, c.B, c.Bp,
            w.beta, w.trans);

  // -- first rotation fit against the forward of the initial pose (pt/bodyfitter.py:721-737) --
  run_transpose<3>(c, init_vertices, V, m->inv_order, nullptr, w.aT);
  run_transpose<3>(c, init_joints, J, nullptr, nullptr, w.initjT);
  run_transpose<1>(c, init_orientations, 9 * J, nullptr, nullptr, w.R2);
  RotArgs ra{};
  ra.partials = w.spart; ra.tjT = w.tjT; ra.jwT = w.jwT; ra.R_new = w.R; ra.RT = w.RT; ra.Pext = w.Pext;
  ra.feat = w.feat; ra.t = tables(m); ra.B = c.B; ra.Bp = c.Bp; ra.Kp = c.Kp;
  c.lite = false;  // no shape solve in this entry point; keep the quad rows off too
  ra.RT12 = nullptr; ra.RT4 = nullptr; ra.rt4_clay = 0;
  run_stats(c, 2, w.initjT, w.aT, nullptr);
  const float* aj = w.initjT;
  if (!has_joints) {
    run_regress(c, w.aT, w.ajT);
    aj = w.ajT;
  }
  ra.ajT = aj; ra.aj_const = nullptr; ra.ca0T = w.initjT; ra.ca0_const = nullptr; ra.R_old = w.R2;
  run_rot(c, ra, true);

  // forward with the known betas for the current orientations = shape front + k_shape_out (trans = 0)
  SolveArgs so{};
  so.partials = nullptr; so.Pext = w.Pext; so.RT = w.RT; so.tjT = nullptr; so.jwT = nullptr; so.beta_ref = nullptr;
  so.kid_ref = nullptr; so.beta = w.beta; so.trans = w.trans; so.refj = w.refj; so.skin = w.skin; so.skin4 = w.skin4; so.wS = nullptr;
  so.wsum = nullptr; so.n_chunks = 0; so.J = J; so.S = m->num_betas; so.Bp = c.Bp; so.B = c.B; so.V = V;
  so.weighted = 0; so.sa_closed_form = 0; so.scale_mode = 0; so.zpartials = nullptr; so.n_zchunks = 0;
  so.scale_reg = 0.f; so.scale_out = w.scale; so.reg = so.reg2 = so.kid_reg = 0.f;
  const float* R_final = w.R;
  for (int it = 0; it < o->num_iter; ++it) {
    const bool last = (it == o->num_iter - 1);
    set_feature_rows(c, so);
    SF_LAUNCH(k_shape_out, dim3(c.groups, J), 32, 0, c.st, so, m->fit_ns);
    run_stats(c, 1, w.refj, nullptr, w.aT);  // reference vertices skinned on the fly, stored for the moments
    aj = w.refj;
    if (!has_joints) {
      run_regress(c, w.aT, w.ajT);
      aj = w.ajT;
    }
    if (!last) {
      ra.ajT = aj; ra.aj_const = nullptr; ra.ca0T = w.refj; ra.ca0_const = nullptr; ra.R_old = w.R; ra.R_new = w.R;
      run_rot(c, ra, true);
      continue;
    }
    // fit_scale_and_translation (pt/bodyfitter.py:764-772, :1628-1681)
    const bool vweights = has_joints ? (vertex_weights && joint_weights) : (vertex_weights != nullptr);
    const int nvb = moment_blocks(V);
    {
      const long long warps = (long long)nvb * c.groups;
      SF_LAUNCH(k_moments, (int)((warps + 3) / 4), 128, 0, c.st, w.tT, w.aT, vweights ? w.vwT : nullptr, V, 256, nvb,
                c.Bp, w.mpart);
      if (has_joints)
        SF_LAUNCH(k_moments, (c.groups + 3) / 4, 128, 0, c.st, w.tjT, w.refj, vweights ? w.jwT : nullptr, J, J, 1, c.Bp,
                  w.mpart + (size_t)nvb * 9 * c.Bp);
    }
    ScaleTransArgs sta;
    sta.vpart = w.mpart; sta.jpart = has_joints ? w.mpart + (size_t)nvb * 9 * c.Bp : nullptr; sta.n_vblocks = nvb;
    sta.Bp = c.Bp; sta.estimate_scale = o->scale_mode == 2; sta.scale = w.scale; sta.trans = w.trans;
    SF_LAUNCH(k_scale_trans, c.groups, 32, 0, c.st, sta);
    if (o->final_adjust_rots) {
      AdjustArgs aa;
      aa.partials = w.spart; aa.tjT = w.tjT; aa.ajT = aj; aa.refj = w.refj; aa.jwT = w.jwT; aa.R_prev = w.R;
      aa.beta = w.beta; aa.trans = w.trans; aa.R_out = w.R2; aa.t = tables(m); aa.Bp = c.Bp; aa.n_adj = (m->n_adjustable > 0 && m->n_adjustable <= J) ? m->n_adjustable : J;
      aa.scale = w.scale; aa.scale_mode = 3;
      launch_adjust(aa, c.Bp / 32, c.st);
      R_final = w.R2;
    }
  }
  OutputArgs oa;
  oa.R_final = R_final; oa.R_rel_src = R_final; oa.beta = w.beta; oa.trans = w.trans; oa.mean = w.mean;
  oa.parents = m->parents; oa.pose_rotvecs = o->want_pose_rotvecs ? out_pose_rotvecs : nullptr;
  oa.shape_betas = nullptr; oa.out_trans = out_trans; oa.orientations = out_orientations;
  oa.rel_orient = (o->want_pose_rotvecs || o->want_rel_orient) ? out_rel_orientations : nullptr; oa.kid = nullptr;
  oa.scale = w.scale; oa.scale_corr = o->scale_mode ? out_scale_corr : nullptr; oa.scale_mode = 0;
  oa.J = J; oa.S = m->num_betas; oa.NS = m->fit_ns; oa.B = c.B; oa.Bp = c.Bp;
  SF_LAUNCH(k_output, dim3(c.Bp / 32, J), 32, 0, c.st, oa);
  SF_CHECK_LAST();
  return SMPLFIT_OK;
}
