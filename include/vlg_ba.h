/*
 * vlg_ba.h -- C ABI of the B200-native bundle adjuster (libvlgba.so).
 *
 * Drop-in boundary for the Levenberg-Marquardt Euclidean bundle-adjustment hot path of
 * caomw/BundleAdjustmentMatlab (VLG, toolbox/bundle).  Every entry point states the
 * reference interface it replaces (file:line relative to the reference tree).  Plain C:
 * pointers and sizes only, host pointers in and out (caller-owned), device memory owned by
 * the context.  Every call returns 0 on success or a negative VLG_BA_E* code;
 * vlg_ba_last_error() gives the text.  There is NO CPU fallback: without a CUDA device
 * vlg_ba_create() fails with VLG_BA_ECUDA.
 *
 * Array conventions (all double, MATLAB column-major, exactly as the reference's mex
 * files receive them):
 *   K  4 x m   [fx fy cx cy]' per camera             (bundle_euclid.m:5)
 *   a  num_a x m, num_a = 6 + num_variableK: [w; Te; K-part]   (bundle_euclid.m:89-96)
 *   b  3 x n   point coordinates                      (bundle_euclid.m:99)
 *   X  2 x n x m measured image points, visible n x m (bundle_euclid.m:81,102)
 * Sparse form of (X, visible): an observation list in the reference's traversal order,
 * ascending i + n*j (mex_bundle_1_XABeUVWeAeB.c:192-196): obs_xy 2 x nobs, obs_pt[nobs]
 * (point i), obs_cam[nobs] (camera j).
 * Block outputs on the list: W is (num_a x 3) column-major per observation, in list order
 * (the reference's dense W(:,:,i,j), mex_bundle_1_XABeUVWeAeB.c:165, restricted to visible
 * cells); U num_a x num_a x m; V 3 x 3 x n; eA num_a x m; eB 3 x n.
 */
#ifndef VLG_BA_H
#define VLG_BA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLG_BA_OK        0
#define VLG_BA_EINVAL   -1   /* bad argument / shape */
#define VLG_BA_ECUDA    -2   /* CUDA runtime error or no device */
#define VLG_BA_ESTATE   -3   /* call out of order (e.g. stage2 before stage1) */
#define VLG_BA_ENOMEM   -4
#define VLG_BA_ENCCL    -5   /* NCCL missing or failed */
#define VLG_BA_ENUM     -6   /* numerical breakdown (non-finite cost, PCG breakdown) */

/* CHOL: dense blocked Cholesky of the assembled S.  PCG: block-Jacobi PCG on the implicit Schur complement
 * (two sweeps over W per iteration).  PCG_EXPLICIT: the same PCG on the assembled dense S, one symmetric
 * lower-triangle matvec per iteration -- cheaper than the sweeps when 4 (6m)^2 bytes << 304 bytes x nobs.
 * AUTO: CHOL when m <= chol_max_cams, else PCG_EXPLICIT when 6 (6m)^2 < 304 nobs and S fits in 8 GB, else PCG. */
enum { VLG_BA_SOLVER_AUTO = 0, VLG_BA_SOLVER_CHOL = 1, VLG_BA_SOLVER_PCG = 2, VLG_BA_SOLVER_PCG_EXPLICIT = 3 };
enum { VLG_BA_RTABLE_HOST_LIBM = 0, VLG_BA_RTABLE_DEVICE = 1 };
/* EUCLID: bundle_euclid.m (a = [w; Te; K-part], num_a = 6 + num_variableK).  PROJECTIVE: bundle_projective.m
 * (a = vec(P), P 3x4, num_a = 12, no K, no rotation table; lambda /10 on accept, x10 on reject,
 * bundle_projective.m:187-205; K and num_variableK are ignored). */
enum { VLG_BA_MODEL_EUCLID = 0, VLG_BA_MODEL_PROJECTIVE = 1 };
enum { VLG_BA_ORDER_CHUNKED = 0, VLG_BA_ORDER_REFERENCE = 1 };

/* Options.  Defaults (vlg_ba_opts_default) are the constants hard-coded in
 * bundle_euclid.m:111-123 and mex_bundle_1_XABeUVWeAeB.c:23,52. */
typedef struct vlg_ba_opts {
    int    num_variableK;     /* 0 'fix_calibration', 1 'fix_principal', 4 default  (bundle_euclid.m:49,66-69) */
    int    fix_structure;     /* bundle_euclid.m:140-144 */
    int    fix_motion;        /* bundle_euclid.m:145-149 */
    double lambda0;           /* 1e-3   bundle_euclid.m:111 */
    double nu0;               /* 2      bundle_euclid.m:112 */
    int    max_iter;          /* 20     bundle_euclid.m:117 */
    int    max_iter2;         /* 10     bundle_euclid.m:118 */
    double rel_tol;           /* 1e-3   bundle_euclid.m:123 */
    double abs_tol;           /* 1e-20  bundle_euclid.m:123 */
    int    backsub_all_rows;  /* 0 = reference behaviour: only 6 camera rows enter db (mex_bundle_3_db_new.c:113-120) */
    int    solver;            /* VLG_BA_SOLVER_*: dense Cholesky of S, or block-Jacobi PCG on the implicit S */
    int    chol_max_cams;     /* AUTO picks Cholesky when m <= this (default 600: measured crossover, profiles/crossover_chol_pcg_r01.txt) */
    double pcg_rtol;          /* relative residual stop for PCG (default 1e-8: one-step cost within ~1e-12 of the exact solve) */
    int    pcg_max_iter;      /* default 1000 */
    int    rtable;            /* VLG_BA_RTABLE_*: who evaluates vl_rodrigues' sin/cos (host libm = bit parity with the CPU reference) */
    int    order;             /* VLG_BA_ORDER_*: REFERENCE = U/eA accumulated in the reference's exact order (slow), CHUNKED = fixed 256-observation chunks (default) */
    int    device;            /* CUDA device ordinal; -1 = current device */
    int    verbose;           /* 'verbose': print "iter k: error= a -> b" (bundle_euclid.m:221-224) */
    int    pcg_deflate;       /* 1 (default): deflate the 4 gauge directions (world translation, scale) in PCG */
    int    pcg_cluster;       /* 1 (default): PCG preconditions with the inverses of 128/num_a-camera diagonal blocks of S
                                 (cluster-Jacobi) instead of per-camera blocks */
    int    model;             /* VLG_BA_MODEL_* (default EUCLID) */
    int    pcg_autotune;      /* 0 (default): the assembled-S matvec is cut into equal pieces per SM -- results are bit-reproducible
                                 from run to run.  n > 0: over the first n solves of a problem the cut is re-weighted by the measured
                                 per-SM streaming rate (SMs differ by +-4 % with position): ~4 % faster PCG iterations, results
                                 still deterministic for a given cut but the cut depends on the measurement */
} vlg_ba_opts;

typedef struct vlg_ba_ctx vlg_ba_ctx;

/* What one trip of the LM loop (bundle_euclid.m:139-241) did. */
typedef struct vlg_ba_trial_info {
    double old_cost;      /* e'e          bundle_euclid.m:209 */
    double new_cost;      /* e_new'e_new  bundle_euclid.m:210 */
    double denom;         /* dp'(lambda dp + g)  bundle_euclid.m:217 */
    double rho;
    double lambda_used;
    double lambda_next;
    double nu_next;
    int    accepted;
    int    solver_used;   /* VLG_BA_SOLVER_CHOL, VLG_BA_SOLVER_PCG or VLG_BA_SOLVER_PCG_EXPLICIT */
    int    pcg_iters;
    double pcg_relres;
    float  ms_stage1;     /* residual + Jacobian + U,V,W,eA,eB       (mex1) */
    float  ms_schur;      /* damping, V*^-1, S / e_ / preconditioner (bundle_euclid.m:162-184 + mex2) */
    float  ms_solve;      /* da                                       (bundle_euclid.m:193) */
    float  ms_stage3;     /* db, update, new residual                 (mex3 + :205-210) */
} vlg_ba_trial_info;

void vlg_ba_opts_default(vlg_ba_opts *opts);
const char *vlg_ba_version(void);

/* Context: one per host thread / GPU. */
int  vlg_ba_create(const vlg_ba_opts *opts, vlg_ba_ctx **out);
void vlg_ba_destroy(vlg_ba_ctx *ctx);
const char *vlg_ba_last_error(const vlg_ba_ctx *ctx);      /* ctx may be NULL: this host thread's last failure without a context (create, stateless entries) */

/* Multi-GPU: one context per rank holding a point shard.  `unique_id` is the
 * 128-byte ncclUniqueId obtained from vlg_ba_nccl_unique_id() on rank 0 and broadcast by
 * the caller.  After this call the per-camera sums and scalars are all-reduced over ranks. */
int  vlg_ba_nccl_unique_id(void *unique_id_128);
int  vlg_ba_set_comm(vlg_ba_ctx *ctx, int rank, int nranks, const void *unique_id_128);
/* Optional, after set_problem_* on every rank: peer-memory paths over NVLink (ranks = processes on one node with
 * peer access).  (1) the per-iteration PCG vector is all-reduced through mailboxes instead of ncclAllReduce (one kernel:
 * stores into every peer's mailbox, flags riding with the data, sum in rank order); (2) on the assembled-S path each
 * rank pulls the column block of S it multiplies from the peers' shares and sums it (instead of every rank multiplying
 * its own full share).  Each rank exports VLG_BA_P2P_HANDLE_BYTES bytes (two CUDA IPC handles: mailbox, S), the host
 * gathers them in rank order (any transport) and every rank imports the nranks x VLG_BA_P2P_HANDLE_BYTES bytes. */
#define VLG_BA_P2P_HANDLE_BYTES 128
int  vlg_ba_p2p_export(vlg_ba_ctx *ctx, void *ipc_handles_128);
int  vlg_ba_p2p_import(vlg_ba_ctx *ctx, const void *ipc_handles /* nranks x VLG_BA_P2P_HANDLE_BYTES */);

/* Reprojection-error map of the CURRENT state on the observation list (toolbox/test/error_reproj.m:72-84) and the
 * statistics remove_outlier() derives from it (toolbox/geometry/incr_reconstruction.m:363-390):
 *   err[t]   = || x(1:2) - x_reproj(1:2) ||  of observation t (list order);  depth[t] = x_reproj(3) before division;
 *   *mean_err = sum(err) / nobs  (error_reproj's `err`);  *n_bad_depth = #{ depth < 0 or depth > depth_max };
 *   *max_sq_err / *argmax = largest squared error among the observations that pass the depth test and its index
 *   (first occurrence, the reference's strict '>' in list order).  err / depth may be NULL.  All points are treated as
 *   finite (the reference skips points with X(4,i) != 1). */
int  vlg_ba_reproj_errors(vlg_ba_ctx *ctx, double depth_max, double *err, double *depth, double *mean_err,
                          double *max_sq_err, int64_t *argmax, int64_t *n_bad_depth);

/* Problem definition.  Replaces the argument packing of bundle_euclid.m:81-102 and the
 * positional inputs of mex_bundle_1_XABeUVWeAeB.c:76-83.  `pivot` (m doubles or NULL) is
 * the 'fix_pivot' mask (bundle_euclid.m:150-154); non-zero = camera held fixed. */
int  vlg_ba_set_problem_dense(vlg_ba_ctx *ctx, int m, int n, const double *K, const double *a,
                              const double *b, const double *X, const double *visible,
                              const double *pivot);
int  vlg_ba_set_problem_sparse(vlg_ba_ctx *ctx, int m, int n, const double *K, const double *a,
                               const double *b, int64_t nobs, const double *obs_xy,
                               const int32_t *obs_pt, const int32_t *obs_cam, const double *pivot);
/* num_vis of bundle_euclid.m:82 (the divisor of error_, :229-231).  Default: the number of list entries summed over all
 * ranks -- `visible` is treated as a 0/1 mask (the reference's sum(visible(:)) differs only for a non-binary array).
 * A positive value overrides that count until the next set_problem_*; <= 0 restores the default. */
int  vlg_ba_set_num_vis(vlg_ba_ctx *ctx, double num_vis);

int64_t vlg_ba_nobs(const vlg_ba_ctx *ctx);
/* Observation list as compacted from a dense problem (visibility indexing, bit-exact). */
int  vlg_ba_get_obs(vlg_ba_ctx *ctx, double *obs_xy, int32_t *obs_pt, int32_t *obs_cam);

/* LM state (a, b, lambda, nu, iter, iter2): bundle_euclid.m:111-119.  NULL = leave/skip. */
int  vlg_ba_set_state(vlg_ba_ctx *ctx, const double *a, const double *b, double lambda, double nu);
int  vlg_ba_get_state(vlg_ba_ctx *ctx, double *a, double *b, double *lambda, double *nu,
                      int *iter, int *iter2);

/* Stage 1 == mex_bundle_1_XABeUVWeAeB (mex_bundle_1_XABeUVWeAeB.c:72-337) at the current
 * (a, b), followed by the fix_* zeroing of bundle_euclid.m:140-154.  Results stay on the
 * device; the getters copy what the caller asks for (NULL = skip). */
int  vlg_ba_stage1(vlg_ba_ctx *ctx, double *cost);
int  vlg_ba_get_blocks(vlg_ba_ctx *ctx, double *U, double *V, double *W, double *eA, double *eB);
/* Per-observation X_hat (2), A (2 x num_a), B (2 x 3), e (2) in list order: recomputed by a
 * diagnostic launch of the same device code (mex1 outputs pout[0..3] restricted to visible cells). */
int  vlg_ba_get_jacobians(vlg_ba_ctx *ctx, double *X_hat, double *A, double *B, double *e);

/* Stage 2 == damping + pinv(V*) + Y + mex_bundle_2_Se_ (bundle_euclid.m:162-192,
 * mex_bundle_2_Se_.c:72-155) for the given lambda, then da (bundle_euclid.m:193).
 * S (num_a*m)^2 is only formed (and returned) on the Cholesky path. */
int  vlg_ba_stage2(vlg_ba_ctx *ctx, double lambda);
int  vlg_ba_get_reduced(vlg_ba_ctx *ctx, double *Vinv, double *S, double *e_, double *da);
/* Teacher forcing: overwrite da before stage 3 (e.g. with the oracle's pinv(S)*e_). */
int  vlg_ba_set_da(vlg_ba_ctx *ctx, const double *da);

/* Stage 3 == mex_bundle_3_db_new (mex_bundle_3_db_new.c:100-166) + the new cost
 * (bundle_euclid.m:205-210) + dp'(lambda dp + g) (bundle_euclid.m:215-217). */
int  vlg_ba_stage3(vlg_ba_ctx *ctx, double lambda, double *new_cost, double *denom);
int  vlg_ba_get_update(vlg_ba_ctx *ctx, double *db, double *a_new, double *b_new);

/* One whole trip of the while loop, bundle_euclid.m:139-241, including accept/reject and
 * the lambda/nu update; state advances inside the context. */
int  vlg_ba_trial_step(vlg_ba_ctx *ctx, vlg_ba_trial_info *info);

/* The loop control of bundle_euclid.m:111-123 for callers that drive vlg_ba_trial_step
 * themselves: vlg_ba_lm_reset() puts the state back to (a, b, lambda0, nu0, iter=1, iter2=0,
 * error_=[]) (a, b may be NULL = keep), vlg_ba_lm_continue() evaluates the while condition. */
int  vlg_ba_lm_reset(vlg_ba_ctx *ctx, const double *a, const double *b);
int  vlg_ba_lm_continue(const vlg_ba_ctx *ctx);

/* The whole of bundle_euclid.m:111-267 on the device.  error_ must hold max_iter doubles;
 * *n_error receives its length (may be 0: bundle_euclid.m:119).  Outputs as :256-267;
 * Xe4 (n doubles or NULL) is Xe(4,:) passed through. */
int  vlg_ba_solve(vlg_ba_ctx *ctx, double *K_, double *Te_, double *w_, double *Xe_,
                  const double *Xe4, double *error_, int *n_error);

/* B independent single-camera motion-only bundle adjustments side by side (toolbox/geometry/estimate_camera.m:247-253:
 * bundle_euclid(K, T, Omega, X, x, 'fix_structure', 'fix_calibration', 'visibility', inlier') on ONE camera, once per camera
 * added in incr_reconstruction.m:223-348).  The context holds all B cameras (opts.fix_structure = 1; camera j sees its own
 * points); with the structure fixed the cameras decouple, and every camera runs its own loop of bundle_euclid.m:111-249
 * -- own lambda, nu, accept decisions, stop rule -- exactly as if bundle_euclid had been called on it alone, but one
 * ROUND of kernels advances all cameras that have not stopped.  a_out: num_a x m final parameters; error_: max_iter
 * doubles per camera (camera j at error_ + j*max_iter), n_error[j] its length (0: no accepted step); *rounds: trial
 * rounds run.  Any of the outputs may be NULL. */
int  vlg_ba_solve_cameras_independent(vlg_ba_ctx *ctx, double *a_out, double *error_, int *n_error, int *rounds);

/* End-to-end convenience used by the mex wrapper and bench.py's e2e leg: host buffers in,
 * one LM trial step, host buffers out (H2D of a, b and the observation list, D2H of a_new,
 * b_new and the costs all inside the call). */
int  vlg_ba_trial_step_host(vlg_ba_ctx *ctx, const double *a, const double *b,
                            const double *obs_xy, double lambda,
                            double *a_new, double *b_new, vlg_ba_trial_info *info);

/* ---- dense drop-ins: the exact argument layout of the reference's three mexFunctions -------
 * Used by mex/mex_bundle_*.c so that the reference's bundle_euclid.m runs unmodified on the GPU.  Output pointers may be
 * NULL.  The three stage entries keep one context for mex1/mex3 and one for mex2 per host thread and re-use them while
 * (m, n, num_a) and the list of contributing cells stay the same (bundle_euclid.m:139,192,204 calls them once per
 * trial step on one visibility pattern); vlg_ba_dense_release() frees them (the mex wrappers register it with mexAtExit),
 * vlg_ba_dense_cache_stats() reports re-uses and rebuilds.
 *   mex1: [X_hat A B e U V W eA eB] = mex_bundle_1_XABeUVWeAeB(K, a, b, X, visible)
 *         (mex_bundle_1_XABeUVWeAeB.c:76-83 inputs, :136-175 outputs)
 *   mex2: [S e_] = mex_bundle_2_Se_(Y, W, U_, eA, eB)          (mex_bundle_2_Se_.c:21-27,:59-66)
 *   mex3: [db a_new b_new X_hat] = mex_bundle_3_db_new(W, da, eB, V_inv, K, a, b, X, visible)
 *         (mex_bundle_3_db_new.c:18-29,:67-86)
 * num_a = 12 with K = NULL selects the projective model and makes the same three entries the
 * drop-ins for mex_bundle_proj_1_XABeUVWeAeB(a, b, X, visible) (mex_bundle_proj_1_XABeUVWeAeB.c:95-98),
 * mex_bundle_proj_2_Se_ (mex_bundle_proj_2_Se_.c:22-26) and
 * mex_bundle_proj_3_db_new(W, da, eB, V_inv, a, b, X, visible) (mex_bundle_proj_3_db_new.c:42-49). */
int  vlg_ba_mex1_dense(int m, int n, int num_a, const double *K, const double *a, const double *b,
                       const double *X, const double *visible, double *X_hat, double *A, double *B,
                       double *e, double *U, double *V, double *W, double *eA, double *eB);
int  vlg_ba_mex2_dense(int m, int n, int num_a, const double *Y, const double *W, const double *U_,
                       const double *eA, const double *eB, double *S, double *e_);
int  vlg_ba_mex3_dense(int m, int n, int num_a, const double *W, const double *da, const double *eB,
                       const double *Vinv, const double *K, const double *a, const double *b,
                       const double *X, const double *visible, double *db, double *a_new,
                       double *b_new, double *X_hat);
void vlg_ba_dense_release(void);
void vlg_ba_dense_cache_stats(int64_t *hits, int64_t *builds);
/* [K_ Te_ w_ Xe_ error_] = bundle_euclid(K, Te, w, Xe, x, ...) (bundle_euclid.m:1-269) in one
 * call: K 4xm, Te 3xm, w 3xm, Xe 4xn, x 3xnxm, visible nxm or NULL (derive from x,
 * bundle_euclid.m:50), pivot m or NULL; error_ holds max_iter doubles. */
int  vlg_ba_bundle_euclid(const vlg_ba_opts *opts, int m, int n, const double *K, const double *Te,
                          const double *w, const double *Xe, const double *x, const double *visible,
                          const double *pivot, double *K_, double *Te_, double *w_, double *Xe_,
                          double *error_, int *n_error);

/* The same on an observation list (no dense n x m array anywhere): obs_xy 2 x nobs, obs_pt / obs_cam
 * 0-based, in the reference's traversal order (ascending i + n*j).  Replaces the dense `x` (3 x n x m)
 * and `visibility` (n x m) arguments of bundle_euclid.m:9,18 for problems the dense interface cannot
 * express.  mex/mex_bundle_euclid_gpu_sparse.c + mex/bundle_euclid_gpu_sparse.m. */
int  vlg_ba_bundle_euclid_sparse(const vlg_ba_opts *opts, int m, int n,
                                 const double *K, const double *Te, const double *w, const double *Xe,
                                 int64_t nobs, const double *obs_xy, const int32_t *obs_pt, const int32_t *obs_cam,
                                 const double *pivot /*m or NULL*/,
                                 double *K_, double *Te_, double *w_, double *Xe_, double *error_, int *n_error);

/* [Pp_ Xp_ error_] = bundle_projective(Pp, Xp, x, ...)  (toolbox/bundle/bundle_projective.m:1-229, called from
 * mview_reconstruction.m:148 and multi_view.m:190): Pp 3x4xm, Xp 4xn, x 3xnxm, visible nxm or NULL (derived
 * from x, bundle_projective.m:38); opts->fix_structure / fix_motion / verbose as the option strings. */
int  vlg_ba_bundle_projective(const vlg_ba_opts *opts, int m, int n, const double *Pp, const double *Xp,
                              const double *x, const double *visible,
                              double *Pp_, double *Xp_, double *error_, int *n_error);

/* Self-test of stage 1's quotients: the kernels share one correctly rounded reciprocal per denominator and finish every
 * quotient with one exact remainder step (csrc/ba_math.cuh, "Quotients") instead of the reference's 34 independent
 * divisions per observation (mex_bundle_1_XABeUVWeAeB.c:39-40, reproject_point.h:55-56).  Runs >= nsamples random
 * quotients (generic operands, reprojection-like, forward-difference-like) both ways on `device` (-1: current) and
 * counts results that differ bitwise from __ddiv_rn; 0 is the only acceptable answer. */
int  vlg_ba_selftest_quotients(int device, int64_t nsamples, uint64_t seed, int64_t *mismatches);

/* Introspection for tests and benches. */
int  vlg_ba_get_schur_structure(vlg_ba_ctx *ctx, int64_t *n_blocks, int32_t *blk_j, int32_t *blk_k);
int64_t vlg_ba_kernel_launches(const vlg_ba_ctx *ctx);
/* Average device time (ms) and launch count of a named kernel group since the last reset:
 * "stage1_cam", "stage1_pt", "vinv", "schur", "chol", "pcg_sweep_pt", "pcg_sweep_cam", "stage3". */
int  vlg_ba_kernel_time(vlg_ba_ctx *ctx, const char *name, double *avg_ms, int64_t *count);
int  vlg_ba_reset_timers(vlg_ba_ctx *ctx, int enable);
/* CUDA-event stopwatch on the context's own stream (bench.py times its region with it). */
int  vlg_ba_timer_start(vlg_ba_ctx *ctx);
int  vlg_ba_timer_stop(vlg_ba_ctx *ctx, float *elapsed_ms);

/* Diagnostic, no GPU needed: the host-side plan of the assembled-S symmetric matvec (csrc/ba_pcg.cuh, "Work decomposition")
 * for an Np x Np system (Np a multiple of 32) on G persistent CTAs over the strips [J0, J1) of 32 columns, with optional
 * per-CTA speed weights (NULL = equal).  sizes[8] <- {tiles, fragments, row-list entries, column-list entries, row blocks,
 * rows per cell, strips per cell, cells}; the arrays may be NULL (call once for the sizes).  tiles4: 4 ints per tile
 * {strip, first row, rows | flags << 16, fragment | strip-in-cell << 20}, flags: 1 first / 2 last tile of a run in a strip,
 * 4 first / 8 last tile of a fragment.  CPU tests replay the plan against a dense product. */
int  vlg_ba_symv_plan(int Np, int G, int J0, int J1, const double *speed, int32_t *tiles4, int32_t *tile_ptr,
                      int32_t *row_ptr, int32_t *row_list, int32_t *col_ptr, int32_t *col_list, int64_t *sizes);
/* The same with an occupancy map occ[ceil(Np/256)][Np/32] (row-major, 1 = the 256-row slot x 32-column strip holds a
 * non-zero block of S): tiles of empty slots are left out, as build_problem does from the block structure -- a scene whose
 * cameras share points only with their neighbours has a banded S, and its matvec streams the band, not the triangle. */
int  vlg_ba_symv_plan_occ(int Np, int G, int J0, int J1, const double *speed, const unsigned char *occ, int32_t *tiles4,
                          int32_t *tile_ptr, int32_t *row_ptr, int32_t *row_list, int32_t *col_ptr, int32_t *col_list, int64_t *sizes);
/* Bytes of S one assembled-S matvec of this rank streams (kept tiles only; 0 on the other solver paths). */
int64_t vlg_ba_symv_bytes(const vlg_ba_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
