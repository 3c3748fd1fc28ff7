/*
 * smplfit_b200 -- C ABI of the Blackwell-native SMPL fitter hot path (libsmplfit_b200.so).
 *
 * The reference (isarandi/smplfitter) is pure Python: it has no FFI seam, so the drop-in
 * boundary is its `smplfitter.pt` Python API and this library is what a maintainer would bind
 * *behind* those methods.  Every entry point below names the reference
 * method it replaces (paths relative to /root/reference/src/smplfitter/):
 *
 *   smplfit_forward            <- pt/bodymodel.py:121-307     BodyModel.forward
 *   smplfit_fit                <- pt/bodyfitter.py:283-549    BodyFitter.fit (+ the stages it calls:
 *                                   _part_sums :235, _fit_shape :840, _fit_shape_gram :960,
 *                                   _fit_shape_general :1104, _fit_global_rotations :1321,
 *                                   _fit_global_rotations_dependent :1418) and pt/rotation.py
 *   smplfit_fit_host           <- the same fit() for host-resident inputs (chunked H2D / fit overlap)
 *   smplfit_fit_known_pose     <- pt/bodyfitter.py:552-653    BodyFitter.fit_with_known_pose
 *   smplfit_fit_known_shape    <- pt/bodyfitter.py:656-838    BodyFitter.fit_with_known_shape
 *   smplfit_convert_vertices   <- pt/bodyconverter.py:129-149 BodyConverter.convert_vertices
 *
 * Conventions: plain pointers and sizes only (no torch types).  All pointers are DEVICE
 * pointers unless the name says host; float arrays are contiguous float32 in the layouts of
 * the reference tensors; work is enqueued on `stream` (a cudaStream_t passed as void*) and the
 * calls never synchronise.  Outputs and the workspace are allocated by the caller (sizes from
 * the *_workspace_bytes functions).  Return value: 0 = OK, negative = error (message via
 * smplfit_last_error()).  Numerical failure never raises: NaNs propagate, as in the reference
 * (pt/bodyfitter.py:1083 ignores cholesky_ex info; pt/rotation.py:8-11 divide_no_nan).
 */
#ifndef SMPLFIT_B200_H_
#define SMPLFIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMPLFIT_OK 0
#define SMPLFIT_ERR_ARG (-1)
#define SMPLFIT_ERR_UNSUPPORTED (-2)
#define SMPLFIT_ERR_WORKSPACE (-3)
#define SMPLFIT_ERR_CUDA (-4)

#define SMPLFIT_MAX_JOINTS 64
#define SMPLFIT_MAX_UNKNOWNS 17 /* betas (+ kid) solved by the register-resident Gram kernels */

/* Model constants: reference layouts (pt/bodymodel.py:80-93) plus the derived device tables
 * built once on the host by smplfitter_b200/masks.py (pt/bodyfitter.py:36-233 semantics). */
typedef struct smplfit_model {
  int32_t num_vertices; /* V */
  int32_t num_joints;   /* J */
  int32_t num_betas;    /* S */
  int32_t num_pose_feats; /* P = 9 (J-1) */
  int32_t skin_k;       /* influences per vertex in the sparse skin table */
  int32_t is_smpl_family;
  int32_t n_used;       /* vertices taking part in part statistics (first n_used of `order`) */
  int32_t n_segments;   /* statistics segments (<= seg_len vertices of one part each) */
  int32_t chunk_len;    /* vertices per shape-pass chunk */
  int32_t max_cas;      /* row length of cas_table */
  int32_t n_adjustable; /* parts re-fitted by the final adjustment (part_flags bit 1); 0 = unknown (use J) */
  int32_t reserved1;
  const float* v_template;     /* (V,3) pose-corrected template */
  const float* shapedirs;      /* (V,3,S) */
  const float* posedirs;       /* (V,3,P) */
  const float* kid_shapedir;   /* (V,3) */
  const float* J_template;     /* (J,3) */
  const float* J_shapedirs;    /* (J,3,S) */
  const float* kid_J_shapedir; /* (J,3) */
  const float* J_regressor;    /* (J,V) post-LBS joint regressor */
  const float* template_mesh;  /* (V,3) zero-pose zero-shape mesh (pt/bodyfitter.py:49) */
  const int32_t* parents;      /* (J) parents[0] = -1 */
  const int32_t* skin_idx;     /* (V,K) */
  const float* skin_w;         /* (V,K) */
  const int32_t* order;        /* (V) internal vertex order -> model vertex (used vertices, grouped by part, first) */
  const int32_t* seg_start;    /* (n_segments+1) offsets into `order` */
  const int32_t* seg_part;     /* (n_segments) */
  const int32_t* part_seg_begin; /* (J+1) segment range per part */
  const int32_t* part_kind;    /* (J) 0 none, 1 multi-joint, 2 bone, 3 leaf, 4 copy */
  const int32_t* part_copy_src;/* (J) */
  const int32_t* part_flags;   /* (J) bit0 = has statistics, bit1 = adjustable in the final pass */
  const int32_t* cas_table;    /* (J,max_cas) children-and-self joint lists, -1 padded */
  const int32_t* cas_count;    /* (J) */
  const int32_t* inv_order;    /* (V) model vertex -> internal position */
  const float* posedirs_fit;   /* (3V, Ppad) posedirs rows in internal order, K zero-padded to Ppad = roundup(P,16) */
  const float* v_template_fit; /* (3V) v_template in internal order */
  /* fitter-level tables (depend on enable_kid): NS = S (+1 with the kid blend shape) unknowns */
  int32_t fit_ns;
  int32_t fit_reserved;
  const float* fit_shapedirs;  /* (V,3,NS): shapedirs with the kid column appended (pt/bodyfitter.py:1139-1149) */
  const float* fit_Jt_ext;     /* (J,3,1+NS): [J_template | J_shapedirs | kid_J_shapedir] (pt/bodyfitter.py:52-58) */
  const float* template_joints_regressed; /* (J,3) J_regressor @ template_mesh (no-joints first fit) */
  const float* J_regressor_fit; /* (J,V) J_regressor with columns in internal order */
  const float* posedirs_hi;     /* (3V, Kt) tf32-exact high part of posedirs_fit, Kt = roundup(P,32) */
  const float* posedirs_lo;     /* (3V, Kt) posedirs_fit - posedirs_hi */
  const float* template_mesh_fit; /* (V,3) template_mesh in internal order */
  const float* fit_rec;         /* (V, rec_len) packed per-vertex records, internal order: 4 skin weights (descending),
                                   4 joint ids (int bits), shapedirs[3][NS]; NULL when skin_k > 4 */
  const double* fit_wS;         /* (J,3,NS) sum_v w_vk S_v  (closed-form SA of the unweighted shape solve) */
  const double* fit_wsum;       /* (J)      sum_v w_vk */
  const float* fwd_rec;         /* (V, fwd_rec_len) forward records in MODEL order: 4 skin weights, 4 joint ids,
                                   v_posed row (int) + 3 pad, shapedirs[3][SP], kid_shapedir[3]; NULL when skin_k > 4 */
  int32_t fit_rec_len;          /* floats per record = roundup(8 + 3 NSP, 4), NSP = NS rounded up to even; shapedirs[x][s] at 8 + x NSP + s */
  int32_t fwd_rec_len;          /* roundup(12 + 3 SP + 3, 4), SP = S rounded up to even */
  /* ---- closed-form Gramian of the unweighted shape solve (k_shape_lite / k_gram_closed) ----
   * With jac_v[:,s] = sum_k w_vk (R_k S_vs + T_ks) the normal matrix sum_v jac_v^T jac_v is a contraction of the
   * per-instance joint transforms with model constants over the joint pairs (k,l) that share a vertex:
   *   A_kl[a,b,s,t] = sum_v w_vk w_vl S_vs[a] S_vt[b],  Bm_kl[a,s] = sum_v w_vk w_vl S_vs[a],  W_kl = sum_v w_vk w_vl.
   * All NULL / 0 disables the path (the per-vertex Gram kernels are used instead). */
  const int32_t* seg_slots;     /* (n_segments, n_slots) joints whose skin weight is non-zero in the segment, -1 padded */
  const int32_t* yj_start;      /* (J+1) CSR over joints of the (segment, slot) cells holding that joint */
  const int32_t* yj_entry;      /* segment * n_slots + slot */
  const int32_t* gcf_pairs;     /* (gcf_npairs, 2) off-diagonal pairs k < l */
  const float* gcf_A;           /* (gcf_npairs, 9, NGP) A_kl[a,b,e] + A_kl[b,a,e], e = upper-triangle index of (s,t), NGP = roundup(NG, 4) */
  const double* gcf_G0;         /* (NG) sum_k sum_a A_kk[a,a,e] (R_k^T R_k = I: instance independent) */
  const int32_t* gcf_lstart;    /* (J+1) CSR over joints l of the partners k (k = l included) */
  const int32_t* gcf_lk;        /* partner joint k of each CSR cell */
  const float* gcf_Bm;          /* (cells, 3, roundup(NS, 4)) Bm_kl, zero padded */
  const float* gcf_Wh;          /* (cells) W_kl / 2 */
  int32_t n_slots;              /* slots per segment (12) */
  int32_t gcf_npairs;
  const float* gcf_AT_hi;       /* (roundup(NG, 256), roundup(9 gcf_npairs, 32)) gcf_A transposed to [e][pair*9 + a*3 + b], zero
                                   padded, tf32-exact high part: the pair term as a tcgen05 GEMM against (R_k^T R_l) features */
  const float* gcf_AT_lo;       /* same shape, gcf_A^T - gcf_AT_hi */
  const float* posedirs_model_hi; /* (3V, Kt) posedirs rows in MODEL vertex order, tf32-exact high part (forward LBS) */
  const float* posedirs_model_lo; /* (3V, Kt) remainder */
  const float* posedirs_model_f32; /* (3V, Kt) the same rows in full precision (split into hi / lo inside the GEMM kernel) */
  const uint8_t* fit_slot_mask; /* (V) internal order: bit k = skinning slot k of this vertex (fit_rec) has a non-zero weight and a
                                   joint other than the previous one of that slot within the statistics segment, i.e. the
                                   kernels' per-slot register cache of joint rows must reload; NULL = compare at run time */
  /* ---- fused forward LBS (k_fwd_fused: blend-shape GEMM with the skinning in its epilogue) ----
   * Constants of the GEMM  D[b][3p+c] = sum_k F[b][k] P[3p+c][k]:  P = 2^fwd_scale_log2 [posedirs | shapedirs | kid_shapedir]
   * with rows in PROCESSING order p (64-vertex tiles of 16-vertex chunks; a chunk holds 16 consecutive model vertices,
   * re-ordered so that consecutive vertices share skinning joints), split into fp16 hi / lo parts (raw IEEE half bits).
   * All NULL / 0 disables the path (the GEMM + k_fwd_skin kernels are used instead). */
  const uint16_t* fwd_P_hi;     /* (roundup(V,64)*3, fwd_kf) */
  const uint16_t* fwd_P_lo;     /* same shape: P - hi */
  const uint32_t* fwd_vrec;     /* (roundup(V,64), 8) per processed vertex: 4 slot weights (float bits) | pack | v_rest[3] (float
                                   bits; zero-pose zero-shape position = v_template + posedirs vec(I)); pack = 4 x 6-bit joint of
                                   slot k (bits 6k..6k+5) | reload mask (bits 24-27: slot k takes another joint here) |
                                   model-local vertex index inside the chunk (bits 28-31) */
  int32_t fwd_kf;               /* K of the GEMM: roundup(P + S + 1, 32); feature order [pose | betas | kid] */
  int32_t fwd_scale_log2;       /* s: the constants are scaled by 2^s into fp16's normal range */
  /* ---- the fit's pose-blend contraction v_posed^T = v_template_fit + posedirs_fit . vec(R_rel) on the same fp16-split
   * tensor-core main loop: 2^fit_scale_log2 posedirs_fit as fp16 hi / lo (raw half bits), rows in internal order ---- */
  const uint16_t* fit_P_hi;     /* (roundup(3V,192), fit_kf) */
  const uint16_t* fit_P_lo;
  int32_t fit_kf;               /* roundup(P, 32) */
  int32_t fit_scale_log2;
  /* ---- fused fit passes (k_fit_fused, csrc/fit_fused.cu): the blend-shape GEMM with the shape-stage / statistics
   * vertex pass in its epilogue, so that the posed template never exists in HBM.  Statistics segment q (<= 32 vertices,
   * seg_start / seg_part above) owns the padded vertex slots [32 q, 32 q + 32); a GEMM tile is two segments.
   * Fitter-level tables (they depend on NS).  All NULL / 0 disables the path. */
  const uint16_t* fq_P_hi;      /* (fq_nseg_pad * 96, fq_kf) 2^fq_scale_log2 [posedirs | fit_shapedirs] rows in slot order, fp16 high part */
  const uint16_t* fq_P_lo;      /* same shape: remainder */
  const uint32_t* fq_rec;       /* (fq_nseg_pad * 32, 8) per slot: 4 skin weights (float bits, descending) | pack | v_rest[3];
                                   pack = 4 x 6-bit joint of slot k | reload mask (bits 24-27: slot k takes another joint here,
                                   along the chain of segments of the same parity) | valid (bit 28) | bits 29 / 30, first slot of every 8-slot
                                   block only: some slot reloads within the first / second four slots of the block */
  const float* fq_sd;           /* (fq_nseg_pad * 32, fq_sdl) shapedirs[x][s] at x * NSP4 + s, NSP4 = NS rounded up to a multiple of 4 */
  int32_t fq_kf;                /* K of the GEMM: roundup(P + NS, 32); feature order [vec(R_rel[1:] - I) | unknowns] */
  int32_t fq_scale_log2;
  int32_t fq_sdl;               /* 3 NSP4 */
  int32_t fq_nseg_pad;          /* n_segments rounded up to even */
  /* ---- the post-LBS joint regressor in CSR form over the internal vertex order (fits without target joints regress
   * them from the vertices, pt/bodyfitter.py:1342-1344; the licensed regressors have ~1 % non-zeros).  NULL: dense. */
  const int32_t* jreg_ptr;      /* (J+1) */
  const int32_t* jreg_idx;      /* (nnz) internal vertex positions */
  const float* jreg_val;        /* (nnz) */
} smplfit_model_t;

/* Options of BodyFitter.fit (pt/bodyfitter.py:283-302). */
typedef struct smplfit_fit_opts {
  int32_t num_iter;
  int32_t final_adjust_rots;
  int32_t enable_kid;          /* fitter built with enable_kid (extra unknown, general solve) */
  int32_t want_pose_rotvecs;   /* requested_keys contains 'pose_rotvecs' */
  int32_t want_rel_orient;     /* ... or 'relative_orientations' */
  int32_t shape_weights;       /* 1: the shape stage honours vertex/joint weights (pt/bodyfitter.py:1018-1028) */
  int32_t scale_mode;          /* 0 none, 1 scale_target, 2 scale_fit (last solve only) */
  int32_t share_beta;          /* 1: one set of betas for the whole batch (pt/bodyfitter.py:1266-1274) */
  float beta_regularizer;
  float beta_regularizer2;
  float kid_regularizer;       /* resolved on the host (None -> beta_regularizer) */
  float scale_regularizer;
} smplfit_fit_opts_t;

const char* smplfit_version(void);
/* sizeof(smplfit_model_t) (which = 0) / sizeof(smplfit_fit_opts_t) (which = 1): lets a binding verify its struct mirror. */
size_t smplfit_struct_size(int which);
const char* smplfit_last_error(void);

/* -- BodyModel.forward (pt/bodymodel.py:121-307) --------------------------------------
 * rot_mode: 0 = pose_rotvecs (B,3J); 1 = rel_rotmats (B,J,3,3); 2 = glob_rotmats (B,J,3,3);
 *           3 = identity (no rotation input).  betas (B,n_betas) may be NULL/0; trans (B,3)
 * and kid (B) may be NULL.  out_vertices may be NULL (return_vertices=False). */
size_t smplfit_forward_workspace_bytes(const smplfit_model_t* m, int64_t batch);
int smplfit_forward(const smplfit_model_t* m, int64_t batch, int rot_mode, const float* rot,
                    const float* betas, int n_betas, const float* trans, const float* kid,
                    float* out_vertices, float* out_joints, float* out_orientations,
                    void* workspace, size_t workspace_bytes, void* stream);

/* -- BodyFitter.fit (pt/bodyfitter.py:283-549) ----------------------------------------
 * target_vertices (B,V,3); target_joints (B,J,3) or NULL; vertex_weights (B,V) / joint_weights
 * (B,J) or NULL; beta_reg_reference (B,S) or NULL (initial_shape_betas); kid_reg_reference (B)
 * or NULL.  init_* are the forward of the initial guess (reference vertices (B,V,3), joints
 * (B,J,3), orientations (B,J,3,3)) or all NULL for the template start (pt/bodyfitter.py:363-394).
 * Outputs (any may be NULL except betas/trans/orientations): pose_rotvecs (B,3J),
 * shape_betas (B,S), trans (B,3), orientations (B,J,3,3), relative_orientations (B,J,3,3),
 * kid_factor (B), scale_corr (B). */
size_t smplfit_fit_workspace_bytes(const smplfit_model_t* m, int64_t batch, const smplfit_fit_opts_t* o,
                                   int has_joints, int has_vw, int has_jw);
int smplfit_fit(const smplfit_model_t* m, int64_t batch, const float* target_vertices,
                const float* target_joints, const float* vertex_weights, const float* joint_weights,
                const float* beta_reg_reference, const float* kid_reg_reference,
                const float* init_vertices, const float* init_joints, const float* init_orientations,
                const smplfit_fit_opts_t* opts, float* out_pose_rotvecs, float* out_shape_betas,
                float* out_trans, float* out_orientations, float* out_rel_orientations,
                float* out_kid_factor, float* out_scale_corr, void* workspace,
                size_t workspace_bytes, void* stream);

/* -- share_beta across ranks: with one process per GPU and the batch sharded over the ranks, the only
 * cross-instance term of the fit is the batch sum of the centred normal equations of every shape solve
 * (pt/lstsq.py:24-26).  When a hook is installed, every share_beta solve calls it between its local batch sum and the
 * shared Cholesky solve: `device_buf` holds `count` doubles (upper triangle + right-hand side) that the hook must
 * replace by their sum over all ranks, enqueued on `stream` (e.g. one ncclAllReduce; the Python host passes a ctypes
 * callback around torch.distributed.all_reduce).  `global_batch` = instances over all ranks (the regulariser is added
 * once per instance before the sum).  fn = NULL removes the hook.  Process-wide; install it before the fits. */
typedef void (*smplfit_allreduce_fn)(double* device_buf, int64_t count, void* stream, void* user);
int smplfit_set_share_beta_allreduce(smplfit_allreduce_fn fn, void* user, int64_t global_batch);

/* -- BodyFitter.fit for HOST-resident targets (the end-to-end form of pt/bodyfitter.py:283-549: what a caller holding
 * numpy / CPU tensors pays for `fitter.fit(torch.as_tensor(x).cuda(), ...)` followed by `.cpu()` on the results).
 * host_* pointers are HOST memory (page-locked for the copies to overlap the fits; pageable memory works but
 * serialises); `workspace` is DEVICE memory of smplfit_fit_host_workspace_bytes.  The batch is processed in
 * chunks of `chunk` instances: the H2D copy of chunk k+1 runs on a library-owned copy stream while chunk k is
 * fitted on library-owned compute streams; each chunk's results are copied back right after its fit (page-locked
 * output buffers keep these copies asynchronous) and everything is joined into `stream`: the host buffers are valid
 * once `stream` has been synchronised.
 * No vertex/joint weights, initial guesses or share_beta here (per-instance options go through smplfit_fit).
 * host_orientations / host_rel_orientations / host_kid_factor / host_scale_corr may be NULL. */
size_t smplfit_fit_host_workspace_bytes(const smplfit_model_t* m, int64_t batch, int64_t chunk,
                                        const smplfit_fit_opts_t* opts, int has_joints);
int smplfit_fit_host(const smplfit_model_t* m, int64_t batch, int64_t chunk, const float* host_target_vertices,
                     const float* host_target_joints, const smplfit_fit_opts_t* opts, float* host_pose_rotvecs,
                     float* host_shape_betas, float* host_trans, float* host_orientations,
                     float* host_rel_orientations, float* host_kid_factor, float* host_scale_corr,
                     void* workspace, size_t workspace_bytes, void* stream);

/* -- BodyFitter.fit_with_known_pose (pt/bodyfitter.py:552-653): shape/translation only.
 * glob_rotmats (B,J,3,3) are the global orientations of the known pose. */
int smplfit_fit_known_pose(const smplfit_model_t* m, int64_t batch, const float* glob_rotmats,
                           const float* target_vertices, const float* target_joints,
                           const float* vertex_weights, const float* joint_weights,
                           const float* beta_reg_reference, const float* kid_reg_reference,
                           const smplfit_fit_opts_t* opts, float* out_shape_betas, float* out_trans,
                           float* out_rel_orientations, float* out_kid_factor, float* out_scale_corr,
                           void* workspace, size_t workspace_bytes, void* stream);

/* -- BodyFitter.fit_with_known_shape (pt/bodyfitter.py:656-838): pose and translation (and optionally a
 * scale of the fitted mesh) for known betas.  shape_betas (B,n_betas) (first min(n,S) used), kid_factor (B)
 * or NULL (requires a fitter built with enable_kid when given); init_* = forward of the initial pose with these
 * betas (vertices (B,V,3), joints (B,J,3), orientations (B,J,3,3)), required.  opts: num_iter,
 * final_adjust_rots, scale_mode (0 or 2 = scale_fit), want_*.  Workspace: smplfit_fit_workspace_bytes. */
int smplfit_fit_known_shape(const smplfit_model_t* m, int64_t batch, const float* shape_betas, int n_betas,
                            const float* kid_factor, const float* target_vertices, const float* target_joints,
                            const float* vertex_weights, const float* joint_weights, const float* init_vertices,
                            const float* init_joints, const float* init_orientations, const smplfit_fit_opts_t* opts,
                            float* out_pose_rotvecs, float* out_trans, float* out_orientations,
                            float* out_rel_orientations, float* out_scale_corr, void* workspace,
                            size_t workspace_bytes, void* stream);

/* -- BodyConverter.convert_vertices (pt/bodyconverter.py:129-149): CSR (V_out x V_in) applied to
 * every instance: out (B,V_out,3) = M @ in (B,V_in,3). */
int smplfit_convert_vertices(const int32_t* indptr, const int32_t* indices, const float* data,
                             int32_t v_out, int32_t v_in, int64_t batch, const float* in_vertices,
                             float* out_vertices, void* stream);

/* smplfit_fit captures a call whose arguments repeat (same model tables, options, batch, pointers, workspace) into a
 * CUDA graph on its second occurrence and replays it afterwards (one cudaGraphLaunch instead of ~40 kernel launches:
 * what bounds small batches is the host's launch rate).  Number of graph launches since the last reset; the kernels
 * inside a replayed graph are counted by smplfit_launch_count as usual.  SMPLFIT_B200_GRAPH=0 disables the cache. */
int64_t smplfit_graph_replays(int reset);
/* diagnostics: {calls seen for the first time, captures started, captures that failed, graphs instantiated} */
void smplfit_graph_stats(int64_t* out4);

/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t smplfit_launch_count(int reset);

/* Per-kernel device timing for benchmarking: while enabled every launch is bracketed by CUDA
 * events on its stream; smplfit_profile_report synchronises on them and writes one
 * "kernel\tlaunches\ttotal_ms\n" line per kernel into `out` (host buffer). */
int smplfit_profile(int enable);
int smplfit_profile_report(char* out, size_t cap);

/* Test hook: v_posed^T [3V][Bp] = v_template + posedirs . feat for feat [Bp][roundup(P,16)] (Bp % 32 == 0);
 * use_tc = 1 forces the tcgen05 kernel (error if unavailable), 0 the FP32 SIMT kernel. */
size_t smplfit_debug_vposed_scratch_bytes(const smplfit_model_t* m, int Bp);
int smplfit_debug_vposed(const smplfit_model_t* m, const float* feat, int Bp, int use_tc, float* out,
                         void* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMPLFIT_B200_H_ */
