/* libscregclust_b200 -- C ABI of the B200-native scregclust hot path.
 *
 * Drop-in boundary for the `.Call` routines of the reference R package
 * (R/RcppExports.R:4-117 <-> src/RcppExports.cpp:14-117) plus the three pure-R
 * numeric helpers the hot path uses (R/utils.R:13-57, 533-541), and the batched
 * session API that runs Step 1 / Step 2a / Step 2b of R/scregclust.R:1222-1803
 * for every module and penalization value of a cycle in one call.
 *
 * Conventions (identical to what R passes):
 *   - all matrices are FP64, column-major, tightly packed (leading dimension =
 *     number of rows); the library never retains or frees caller pointers;
 *   - module ids are 1-based, -1 = noise module; `update_order` and
 *     `prior_idx` are 0-based (src/allocation.cpp:4-6);
 *   - every function returns 0 on success or a negative status; the message of
 *     the last failure on the calling thread is available from scrc_last_error().
 *     Nothing throws across the ABI and nothing calls exit().
 *   - there is no CPU fallback: without a usable CUDA device every compute
 *     entry point fails with SCRC_ERR_CUDA.
 */
#ifndef SCREGCLUST_B200_H
#define SCREGCLUST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; the bridge maps them to the reference's Rcpp::stop() texts */
#define SCRC_OK 0
#define SCRC_ERR_COOP_DIMS (-1)      /* "COOP LASSO: Matrix dimensions of y and x need to be positive." optim.cpp:77 */
#define SCRC_ERR_COOP_ROWS (-2)      /* "y and x need to have the same number of rows."                optim.cpp:81 */
#define SCRC_ERR_COOP_BETA0 (-3)     /* "beta_0 needs to be of size p x m"                             optim.cpp:88 */
#define SCRC_ERR_COOP_RHO (-4)       /* "rho_0 > 0 needs to hold"                                      optim.cpp:116 */
#define SCRC_ERR_COOP_ALPHA (-5)     /* "0 < alpha_0 < 2 needs to hold"                                optim.cpp:121 */
#define SCRC_ERR_NNLS_DIMS (-6)      /* "NNLS: Matrix dimensions of y and x need to be positive."      optim.cpp:409 */
#define SCRC_ERR_ALLOC_TOO_LARGE (-7)/* "alloc_array: requested allocation too large"                  utils.cpp:20 */
#define SCRC_ERR_RESET_DIMS (-8)     /* "reset_array: input has wrong dimensions"                      utils.cpp:52 */
#define SCRC_ERR_ARG (-20)           /* any other invalid argument */
#define SCRC_ERR_UNSUPPORTED (-21)
#define SCRC_ERR_COMM (-30)          /* NCCL unavailable or a collective failed */
#define SCRC_ERR_INTERRUPTED (-31)   /* scrc_ctx_request_cancel(): the bridge raises R's interrupt condition */
#define SCRC_ERR_INTERNAL (-50)
#define SCRC_ERR_CUDA (-100)         /* CUDA failures: -100 - cudaError_t */

int scrc_version(void);
const char* scrc_last_error(void);
/* number of kernels of this library launched so far by the process */
unsigned long long scrc_launch_count(void);
/* selects the CUDA device used by the drop-in entry points (sessions carry their own) */
int scrc_set_device(int device);

/* ------------------------------------------------------------------ drop-ins */

/* coop_lasso(y, x, lambda, weights, beta_0, rho_0, alpha_0, n_update, eps_corr, max_iter,
 *            eps_rel, eps_abs, verbose) -> list(beta, iterations)
 * R/RcppExports.R:66-68, src/optim.cpp:63-320.
 * y: n x m, x: n_x x p, weights: p, beta0: p x m or NULL; beta_out: p x m; iterations_out: 1 int
 * (max_iter + 1 when the iteration limit was hit, optim.cpp:295-305). */
int scrc_coop_lasso(const double* y, int64_t n, int64_t m, const double* x, int64_t n_x, int64_t p,
                    double lambda, const double* weights, const double* beta0, int64_t beta0_rows,
                    int64_t beta0_cols, double rho0, double alpha0, int n_update, double eps_corr,
                    int max_iter, double eps_rel, double eps_abs, double* beta_out, int* iterations_out);

/* coef_nnls(x, y, eps, max_iter) -> list(beta, iterations)
 * R/RcppExports.R:93-95, src/optim.cpp:403-549.
 * x: nobs x nvar, y: nobs_y x m; beta_out: nvar x m; iterations_out: m ints
 * (max_iter, and a zero column, for columns that never converged). */
int scrc_coef_nnls(const double* x, int64_t nobs, int64_t nvar, const double* y, int64_t nobs_y, int64_t m,
                   double eps, int max_iter, double* beta_out, int* iterations_out);

/* allocate_into_modules(resid_array, resid_var, prior_indicator, k_, update_order,
 *                       prior_baseline, prior_weight) -> integer[T]
 * R/RcppExports.R:4-6, src/allocation.cpp:8-100.
 * sq_resid: K x n2 x T squared residuals; resid_var: T x K; prior_indicator as CSR
 * (prior_off: T+1 offsets into prior_idx, both may be NULL when there is no prior);
 * k_in / k_out: T (1-based, -1 noise); update_order: T (0-based permutation).
 * votes_out (T x K, row-major by gene: votes_out[t*K + k]) may be NULL. */
int scrc_allocate_into_modules(const double* sq_resid, int K, int64_t n2, int64_t T, const double* resid_var,
                               const int* prior_off, const int* prior_idx, const int* k_in,
                               const int* update_order, double prior_baseline, double prior_weight,
                               int* k_out, int* votes_out);

/* alloc_array(input, n_cl) / reset_array(arr, input): src/utils.cpp:12-34, 43-63.
 * Host-side broadcasts kept for API completeness -- the session API never
 * materialises the K x n_obs x n_genes array. */
int scrc_alloc_array(const double* input, int64_t n_obs, int64_t n_genes, int64_t n_cl, double* out);
int scrc_reset_array(double* arr, int64_t n_cl, int64_t n_obs, int64_t n_genes, const double* input,
                     int64_t in_rows, int64_t in_cols);

/* fast_cor(x, y): R/utils.R:533-541.  x: n x px, y: n x py, out: px x py. */
int scrc_fast_cor(const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out);
/* coef_ols(y, x): R/utils.R:13-37.  y: n x m, x: n x p, beta_out: p x m. */
int scrc_coef_ols(const double* y, const double* x, int64_t n, int64_t p, int64_t m, double* beta_out);
/* coef_ridge(y, x, lambda): R/utils.R:49-57. */
int scrc_coef_ridge(const double* y, const double* x, int64_t n, int64_t p, int64_t m, double lambda,
                    double* beta_out);
/* kmeanspp(x, n_cluster, n_init_clusterings, n_max_iter): R/utils.R:351-375 -- k-means++ seeding (kmeanspp_init,
 * R/utils.R:275-322) followed by stats::kmeans (Hartigan-Wong) per initialisation; the clustering with the smallest
 * tot.withinss is returned.  x: n x p column-major (host or device pointer).  R's RNG stays with the caller:
 *   first_centers[i]            the value of sample(n, size = 1L) of initialisation i (1-based)
 *   uniforms[i * (K-1) + r]     the unif_rand() consumed by the r-th sample.int(prob = ps) of initialisation i
 * drawn in R as  for (i in seq_len(n_init)) { first[i] <- sample(n, 1L); u[i, ] <- runif(K - 1L) }  -- the same
 * stream the reference consumes.  cluster_out: n, 1-based; tot_withinss_out / ifault_out: per initialisation (may be
 * NULL; ifault as in stats::kmeans: 0 ok, 2 did not converge in n_max_iter, 4 Quick-TRANSfer steps exceeded; an empty
 * cluster is an error, as in R).  Distances are sequential sums in R's order; see DESIGN.md for what is pinned. */
int scrc_kmeanspp(const double* x, int64_t n, int64_t p, int n_cluster, int n_init_clusterings, int n_max_iter,
                  const int* first_centers, const double* uniforms, int* cluster_out, double* tot_withinss_out,
                  int* ifault_out);

/* crossprod(x, y) = x'y on the FP64 tensor pipe (compute_xtx, src/optim.cpp:17-27, when y == x). */
int scrc_crossprod(const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out);

/* Measurement hook: times `reps` launches of the TN GEMM kernel C = A'B (A: k x m, B: k x n,
 * device-resident pseudo-random data) with CUDA events on the launching stream and returns
 * the average milliseconds per launch.  symmetric != 0 uses the SYRK-style mirrored epilogue. */
int scrc_bench_gemm(int64_t k, int64_t m, int64_t n, int symmetric, int reps, double* ms_per_launch);

/* Per-kernel-class timing with CUDA events on the launching stream (off by default).
 * Classes: 0 dense GEMMs (Gram / correlation / initial fit), 1 ADMM eigenbasis GEMMs,
 * 2 ADMM elementwise kernels, 3 NNLS, 4 fused residual/vote kernel, 5 other.
 * ms: accumulated kernel time, work: algorithmic flops (classes 0, 1) or bytes (2, 4),
 * launches: timed launch groups.  Each array has 6 entries; any may be NULL. */
int scrc_profile_enable(int on);
int scrc_profile_reset(void);
int scrc_profile_get(double* ms, double* work, unsigned long long* launches);
/* Device-side wall clock: both calls synchronise the current device and record a CUDA event,
 * so the interval covers every stream of this library (used by bench.py). */
int scrc_timer_begin(void);
int scrc_timer_end(double* ms);

/* ------------------------------------------------------------------ session API */

typedef struct scrc_ctx scrc_ctx;

/* Solver knobs surfaced by scregclust() (R/scregclust.R:193-196) and the ADMM
 * internals that stay at coop_lasso's defaults (src/optim.cpp:67-69). */
typedef struct scrc_options {
    int max_optim_iter;      /* 10000 */
    double tol_coop_rel;     /* 1e-8  */
    double tol_coop_abs;     /* 1e-12 */
    double tol_nnls;         /* 1e-4  */
    double rho0;             /* 0.2   */
    double alpha0;           /* 1.5   */
    int n_update;            /* 2     */
    double eps_corr;         /* 0.2   */
} scrc_options;
void scrc_default_options(scrc_options* o);

/* Uploads the four scaled data splits (R/scregclust.R:998-1009: z1_reg_scaled n1 x P,
 * z2_reg_scaled n2 x P, z1_target_centered n1 x T, z2_target_centered n2 x T) and
 * z1_target_sds (T) once, then builds on the device: the Gram matrices X'X, X'Y
 * (optim.cpp:17-27,99,414,421), their eigenbasis, the OLS / ridge initial fit,
 * res_sds and beta_adapt_sq (R/scregclust.R:1049-1113). */
int scrc_ctx_create(scrc_ctx** out, int device, const double* z1_reg, const double* z2_reg,
                    const double* z1_target, const double* z2_target, const double* z1_target_sds,
                    int64_t n1, int64_t n2, int64_t P, int64_t T, int K);
void scrc_ctx_destroy(scrc_ctx* ctx);

/* Same session, but from the raw p x n expression matrix (genes x cells, column-major, as the
 * reference's `expression` argument): sample split with explicit `split_indices`, optional
 * per-stratum centring with split-1 means (split_sample, R/scregclust.R:2280-2326 -- including
 * its column re-ordering by stratum when center is set), removal of genes that are constant in
 * either split (R/scregclust.R:899-964) and the scaling of R/scregclust.R:996-1009, all on the
 * device after one upload.
 *   is_regulator[p]   : 1 = regulator, 0 = target, anything else = gene in neither matrix
 *   split_indices[n]  : 1 / 2 = data split, anything else (R's NA) = cell not used
 *   stratification[n] : integer stratum labels, or NULL for a single stratum
 *   keep_gene[p]      : out, 1 where the gene survived the constant filter (may be NULL)
 *   dims_out[4]       : out, n1, n2, P, T of the session (may be NULL)
 * `expr` may be a host or a device pointer.  The random draw of split_indices when the user gives
 * none stays with the caller (it consumes R's RNG stream, R/scregclust.R:2283-2301). */
int scrc_ctx_create_raw(scrc_ctx** out, int device, const double* expr, int64_t p, int64_t n,
                        const int* is_regulator, const int* split_indices, const int* stratification,
                        int center, int K, unsigned char* keep_gene, int64_t* dims_out);
/* The staged inputs of a session: z1_reg_scaled (n1 x P), z2_reg_scaled (n2 x P),
 * z1_target_centered (n1 x T), z2_target_centered (n2 x T), z1_target_sds (T); any may be NULL. */
int scrc_ctx_get_inputs(scrc_ctx* ctx, double* z1_reg, double* z2_reg, double* z1_target, double* z2_target,
                        double* z1_target_sds);

/* res_sds (T), init_df (1), beta_init (P x T) -- any pointer may be NULL. */
int scrc_ctx_get_init(scrc_ctx* ctx, double* res_sds, double* init_df, double* beta_init);
/* cross_corr1 = fast_cor(z1_target_centered, z1_reg_scaled), T x P (R/scregclust.R:1116). */
int scrc_ctx_cross_corr(scrc_ctx* ctx, double* out);
/* k_start_cl <- kmeanspp(cross_corr1, n_modules, n_initializations, 100L) (R/scregclust.R:1131-1136) on the
 * session's own cross_corr1, which never leaves the device; arguments as scrc_kmeanspp. */
int scrc_ctx_kmeanspp(scrc_ctx* ctx, int n_init_clusterings, int n_max_iter, const int* first_centers,
                      const double* uniforms, int* cluster_out, double* tot_withinss_out, int* ifault_out);
/* G_rr (P x P) and G_rt (P x T), any may be NULL. */
int scrc_ctx_get_gram(scrc_ctx* ctx, double* g_rr, double* g_rt);

/* Outputs of one batched cycle.  Every pointer may be NULL.  Layouts per fit f:
 * [f][k][p] / [f][k][t] are exactly R's P x K / T x K column-major matrices. */
typedef struct scrc_cycle_out {
    int* k_new;                /* F x T   first argmax of the votes, 1-based (allocation.cpp:88-94) */
    int* votes;                /* F x T x K ([f][t][k]) */
    unsigned char* models;     /* F x K x P  selected regulators (R/scregclust.R:1384) */
    double* weights;           /* F x K x P  (R/scregclust.R:1385) */
    double* signs;             /* F x K x P  NaN where unselected (R/scregclust.R:1386) */
    double* r2_test;           /* F x K x T  (R/scregclust.R:1629-1631) */
    double* resid_var;         /* F x K x T  (R/scregclust.R:1636-1638) */
    double* best_r2;           /* F x T      (R/scregclust.R:1632) */
    int* best_r2_idx;          /* F x T      1-based which.max (R/scregclust.R:1646) */
    int* admm_iterations;      /* F x K      0 = empty module (R/scregclust.R:1345) */
    int* nnls_iterations;      /* F x K x T  0 = module without regulators */
    int* n_selected;           /* F x K      regulators per module */
    long long* admm_sweeps;    /* 1          batched ADMM sweeps executed */
} scrc_cycle_out;

/* One cycle (Step 1, Step 2a and -- unless `skip_allocation` -- Step 2b of
 * R/scregclust.R:1303-1739) for F independent fits at once.  Fit f uses penalty
 * lambda[f] and the module assignment k_in[f*T .. f*T+T) (1-based, -1 noise).
 * Fits are typically the penalization grid (R/scregclust.R:1223) but any mix of
 * (lambda, configuration) pairs is allowed, e.g. the final configurations of a
 * limit cycle (R/scregclust.R:1307-1311). */
int scrc_fit_cycle(scrc_ctx* ctx, int F, const double* lambda, const int* k_in, const scrc_options* opt,
                   int skip_allocation, scrc_cycle_out* out);

/* Leave-one-regulator-out importance and module R2 of F final configurations
 * (R/scregclust.R:1823-1912, compute_predictive_r2 = TRUE).  k_final: F x T module assignments,
 * models / signs: F x K x P exactly as returned by scrc_fit_cycle for those configurations.
 * r2_module: F x K, importance: F x K x P ([f][k][p] = R's P x K matrix), NaN where the
 * reference leaves NA (modules without targets or with fewer than two regulators). */
int scrc_importance(scrc_ctx* ctx, int F, const int* k_final, const unsigned char* models, const double* signs,
                    const scrc_options* opt, double* r2_module, double* importance);
/* The same, plus the per-target matrices `r2_removed[[m]][[j]]` (R/scregclust.R:1890-1898): for every
 * (fit, module) in order that has targets and more than one regulator, an m_j x (s_j + 1) column-major
 * block (rows = the module's targets in increasing index, column 0 = "None", column r = r-th selected
 * regulator removed), packed one after the other into r2_removed (capacity in doubles; the caller
 * knows m_j and s_j from k_final and models).  R divides t(ssq) by a length-m_j vector there, which
 * recycles down the columns of the (s_j + 1) x m_j matrix; the values reproduce that. */
int scrc_importance_ex(scrc_ctx* ctx, int F, const int* k_final, const unsigned char* models, const double* signs,
                       const scrc_options* opt, double* r2_module, double* importance, double* r2_removed,
                       int64_t r2_removed_capacity);

/* Network prior for Step 2b (src/allocation.cpp:39-56,74-77; R/scregclust.R:584-615,1666-1675).
 * prior_off / prior_idx: CSR form of `prior_indicator` over the T target genes (0-based neighbours);
 * update_order: F x T, the 0-based permutation R draws with sample() for each fit and cycle
 * (R/scregclust.R:1666) -- the caller owns the RNG, so the result matches a serial R run only if it
 * replays R's draws.  With prior_weight == 0 this is identical to scrc_fit_cycle.
 * aggregate != 0 selects the `allocate_per_obs = FALSE` branch of Step 2b instead (R/scregclust.R:1678-1728): genes
 * are reassigned from the per-module test sums of squares and residual variances of the cycle (no pass over the
 * cells), mixed with the prior as written there; prior_off / update_order may then be NULL (no prior: the result
 * does not depend on the order).  update_order is 0-based here too (R draws 1-based gene indices at :1702).
 * The reference's argument checks make this branch unreachable in v0.2.2 (:807-817); it is provided for
 * completeness and reproduces the source as written. */
typedef struct scrc_prior {
    const int* prior_off;      /* T + 1 */
    const int* prior_idx;      /* prior_off[T] entries */
    double baseline;           /* prior_baseline (1e-6) */
    double weight;             /* prior_weight   (0.5)  */
    const int* update_order;   /* F x T */
    int aggregate;             /* 0: per-cell votes (allocate_per_obs = TRUE); 1: aggregate branch */
} scrc_prior;
int scrc_fit_cycle_prior(scrc_ctx* ctx, int F, const double* lambda, const int* k_in, const scrc_options* opt,
                         const scrc_prior* prior, scrc_cycle_out* out);

/* Quick mode (R/scregclust.R:1435-1496, quick_mode = TRUE): from cycle two on, Step 2 of a fit only sees the targets
 * currently in a module plus the `percent` share of the noise targets whose largest absolute correlation with an
 * active regulator is highest (R's default quantile, type 7); in the last cycle (skip_allocation) the active targets
 * only.  idx_latest: F x T, R's `idx` -- the LATEST allocation of each fit, which is k_in except while an older
 * configuration of a limit cycle is re-fitted.  The skipped targets come back with the reference's values (r2 0,
 * resid_var 1e6 -- 0 in the last cycle --, zero coefficients, k_new 1).  The caller passes NULL in cycle 1. */
typedef struct scrc_quick {
    const int* idx_latest;     /* F x T, 1-based, -1 noise */
    double percent;            /* quick_mode_percent in [0, 1) */
} scrc_quick;
int scrc_fit_cycle_quick(scrc_ctx* ctx, int F, const double* lambda, const int* k_in, const scrc_options* opt,
                         int skip_allocation, const scrc_quick* quick, scrc_cycle_out* out);

/* NNLS coefficients of module k (0-based) of fit f from the last scrc_fit_cycle call:
 * n_selected x T (R/scregclust.R:1588); sel_out receives the 0-based regulator indices. */
int scrc_ctx_get_coeffs(scrc_ctx* ctx, int f, int k, double* coeffs_out, int* sel_out);
/* The same for every module of fit f in one transfer: the K coefficient matrices (n_selected_k x T,
 * column-major each) packed one after the other in coeffs_out (sum_k n_selected_k * T doubles) and
 * the regulator indices in sel_out (sum_k n_selected_k); n_selected as returned by the cycle. */
int scrc_ctx_get_coeffs_fit(scrc_ctx* ctx, int f, double* coeffs_out, int* sel_out);

/* ------------------------------------------------------------------ cancellation
 * The reference polls Rcpp::checkUserInterrupt() every ADMM iteration, NNLS outer iteration and allocated gene
 * (src/optim.cpp:307,536, src/allocation.cpp:96).  The library never calls the R API; instead the bridge (or any
 * other thread / signal handler) sets this flag, which is polled at the ADMM host polls (every second sweep)
 * and between the phases of a cycle.  The running call then returns SCRC_ERR_INTERRUPTED; in a multi-GPU job
 * the ranks agree on the interrupt at the next merge point, so a flag set on one rank stops all of them. */
int scrc_ctx_request_cancel(scrc_ctx* ctx);
int scrc_ctx_clear_cancel(scrc_ctx* ctx);

/* ------------------------------------------------------------------ multi-GPU
 * One rank per GPU.  Ranks are processes (torchrun / mpirun: rank 0 calls scrc_comm_unique_id and hands the 128
 * bytes to the others by whatever means the launcher offers) or the host threads scrc_grid_fit_multi starts.
 * A session created with a communicator is COLLECTIVE: every rank calls scrc_fit_cycle / scrc_importance /
 * scrc_grid_fit with identical arguments (and the same set of non-NULL output pointers) and every rank gets
 * the complete result.  Inside a cycle the Step-1 problems ((fit, module) pairs, R/scregclust.R:1332-1388) are
 * spread over the ranks and Step 2 (NNLS columns, residuals, votes; R/scregclust.R:1533-1615,
 * src/allocation.cpp:59-94) runs on contiguous target-gene slices; the module summaries and the per-gene
 * assignment vectors are merged with NCCL collectives over NVLink.  Results are bit-identical to one GPU.
 * NCCL is resolved at run time (libnccl.so.2); single-GPU sessions do not need it. */
typedef struct scrc_comm scrc_comm;
int scrc_comm_unique_id(char id_out[128]);
int scrc_comm_create(scrc_comm** out, const char id[128], int world, int rank, int device);
void scrc_comm_destroy(scrc_comm* comm);
/* scrc_ctx_create on comm's device; rank 0 computes Gram / eigenbasis / initial fit once and broadcasts them. */
int scrc_ctx_create_sharded(scrc_ctx** out, scrc_comm* comm, const double* z1_reg, const double* z2_reg,
                            const double* z1_target, const double* z2_target, const double* z1_target_sds,
                            int64_t n1, int64_t n2, int64_t P, int64_t T, int K);
/* Predicted ADMM iterations per (fit, module) (F x K, 0 = unknown) of the NEXT scrc_fit_cycle call; used only
 * to balance the Step-1 problems over the ranks (typically last cycle's counts); never changes a result. */
int scrc_ctx_set_cost_hint(scrc_ctx* ctx, const int* iterations, int n);
/* The partition rules themselves (host only, deterministic; every rank evaluates them identically):
 * longest-processing-time-first assignment of n problems to `world` ranks, and the contiguous slice of n
 * items (in units of `align`) owned by `rank`. */
int scrc_shard_assign(const double* cost, int n, int world, int* owner_out);
int scrc_shard_slice(int64_t n, int64_t align, int rank, int world, int64_t* begin_out, int64_t* end_out);

/* ------------------------------------------------------------------ whole-grid fit
 * The alternating loop of scregclust() for a penalization grid in ONE call (R/scregclust.R:1222-1912 for
 * allocate_per_obs = TRUE): cycle counter and forced last cycle (:1269-1284),
 * Step 1 / 2a / 2b of every penalization value that is still iterating in one batched device call per cycle,
 * noise threshold and minimum module size (:1731-1739), history match / limit cycles (:1763-1802), the re-fit
 * of every final configuration (:1307-1311) and -- compute_predictive_r2 -- the leave-one-regulator-out
 * importances (:1823-1912).  The R driver calls this once instead of ~10^3 .Call()s per grid; with a sharded
 * session (or scrc_grid_fit_multi) it uses every GPU of the box. */
typedef struct scrc_grid_params {
    int n_penalties;
    const double* penalization;             /* n_penalties */
    const int* initial_target_modules;      /* T, 1-based, -1 noise (k_start_cl, R/scregclust.R:1127-1147) */
    int n_cycles;                           /* 50    */
    double noise_threshold;                 /* 0.025 */
    int min_module_size;                    /* 0     */
    scrc_options opt;
    int compute_predictive_r2;              /* 1: importance step (the reference's default) */
    int keep_coeffs;                        /* 1: keep the NNLS coefficient matrices of the final configurations */
    int keep_diagnostics;                   /* 1: votes and NNLS iteration counts of every cycle are recorded */
    /* network prior (R/scregclust.R:584-615, src/allocation.cpp:39-56): CSR over the T targets, or NULL */
    const int* prior_off;
    const int* prior_idx;
    double prior_baseline;                  /* 1e-6 */
    double prior_weight;                    /* 0 = no prior */
    /* update_order(user, penalty index, cycle) -> the 0-based permutation R's sample() draws at
     * R/scregclust.R:1666 (NULL: identity; only matters with a prior).  Must be thread-safe and return the
     * same permutation on every rank. */
    const int* (*update_order)(void* user, int penalty, int cycle);
    void* update_order_user;
    int quick_mode;                         /* R/scregclust.R:201: 0 */
    double quick_mode_percent;              /* 0.1 */
    int allocate_per_obs;                   /* R/scregclust.R:199: 1; 0 = aggregate branch (see scrc_prior.aggregate) */
    /* Optional interrupt flag owned by the caller (NULL: none): polled before every device call of the loop; once
     * it reads non-zero the fit stops with SCRC_ERR_INTERRUPTED on every rank (same mechanism as
     * scrc_ctx_request_cancel).  The bridge sets it from the R thread while the fit runs on a worker thread. */
    const volatile int* cancel;
} scrc_grid_params;
void scrc_grid_default_params(scrc_grid_params* p);

typedef struct scrc_grid_result scrc_grid_result;
int scrc_grid_fit(scrc_ctx* ctx, const scrc_grid_params* p, scrc_grid_result** out);
/* The same on n_devices GPUs of this process: one host thread per device, each with its own session and
 * NCCL rank; inputs are read by every thread (host or device pointers).  Returns rank 0's result. */
int scrc_grid_fit_multi(const int* devices, int n_devices, const double* z1_reg, const double* z2_reg,
                        const double* z1_target, const double* z2_target, const double* z1_target_sds,
                        int64_t n1, int64_t n2, int64_t P, int64_t T, int K, const scrc_grid_params* p,
                        scrc_grid_result** out);
/* The same loop over caller-supplied compute (used by the CPU tests of the host logic, which answer the
 * callbacks from the oracle; a product caller has no reason to use it).  Callbacks return 0 or an error code. */
typedef struct scrc_grid_backend {
    void* user;
    int64_t P, T;
    int K;
    int (*fit_cycle)(void* user, int F, const double* lambda, const int* k_in, const scrc_options* opt,
                     int skip_allocation, const int* iterations_hint, const scrc_prior* prior,
                     const scrc_quick* quick, scrc_cycle_out* out);
    int (*importance)(void* user, int F, const int* k_final, const unsigned char* models, const double* signs,
                      const scrc_options* opt, double* r2_module, double* importance, double* r2_removed,
                      int64_t r2_removed_capacity);
    int (*coeffs_fit)(void* user, int f, double* coeffs_out, int* sel_out);
} scrc_grid_backend;
int scrc_grid_fit_custom(const scrc_grid_backend* backend, const scrc_grid_params* p, scrc_grid_result** out);

typedef struct scrc_grid_info {
    int n_penalties;
    int64_t P, T;
    int K;
    long long admm_sweeps;              /* batched ADMM sweeps executed by this rank */
    long long admm_problem_iterations;  /* sum of the per-problem iteration counts of the whole job */
    int fit_cycles, final_refits, device_calls;
} scrc_grid_info;
int scrc_grid_result_info(const scrc_grid_result* r, scrc_grid_info* info);
typedef struct scrc_grid_penalty_info {
    double penalization;
    int converged, n_cycles_run, final_cycle_length;
    int n_history;    /* entries of k_history (initial configuration first) */
    int n_records;    /* per-cycle records: the fit cycles in order, then one per final re-fit */
} scrc_grid_penalty_info;
int scrc_grid_result_penalty(const scrc_grid_result* r, int i, scrc_grid_penalty_info* info);
int scrc_grid_result_history(const scrc_grid_result* r, int i, int* k_history /* n_history x T */);
/* record `rec` of penalty i; votes (T x K, [t][k]) / nnls_iterations (K x T) only with keep_diagnostics */
int scrc_grid_result_record(const scrc_grid_result* r, int i, int rec, int* admm_iterations /* K */,
                            int* n_selected /* K */, unsigned char* models /* K x P */, int* votes,
                            int* nnls_iterations);
/* all records of penalty i at once: n_records x K, n_records x K, n_records x K x P (any may be NULL) */
int scrc_grid_result_records(const scrc_grid_result* r, int i, int* admm_iterations, int* n_selected,
                             unsigned char* models);
/* Final configuration m (0 <= m < final_cycle_length; m = 0 is the last entry of the history) in R's layouts
 * ([k][p] = P x K, [k][t] = T x K column-major).  Every pointer may be NULL. */
typedef struct scrc_grid_final {
    int* k_final;              /* T */
    unsigned char* models;     /* K x P */
    double* weights;           /* K x P */
    double* signs;             /* K x P, NaN where unselected */
    double* r2_test;           /* K x T */
    double* resid_var;         /* K x T  (sigmas = sqrt) */
    double* best_r2;           /* T */
    int* best_r2_idx;          /* T */
    int* n_selected;           /* K */
    double* r2_module;         /* K, NaN = NA (compute_predictive_r2) */
    double* importance;        /* K x P, NaN = NA */
} scrc_grid_final;
int scrc_grid_result_final(const scrc_grid_result* r, int i, int m, scrc_grid_final* out);
/* the K coefficient matrices (n_selected_k x T each, packed) and regulator indices, as scrc_ctx_get_coeffs_fit */
int scrc_grid_result_coeffs(const scrc_grid_result* r, int i, int m, double* coeffs_out, int* sel_out);
/* packed r2_removed blocks of configuration (i, m), as scrc_importance_ex; size in doubles first */
int64_t scrc_grid_result_r2_removed_size(const scrc_grid_result* r, int i, int m);
int scrc_grid_result_r2_removed(const scrc_grid_result* r, int i, int m, double* out);
void scrc_grid_result_free(scrc_grid_result* r);

#ifdef __cplusplus
}
#endif
#endif /* SCREGCLUST_B200_H */
