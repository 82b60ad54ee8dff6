// Fused predict -> residual -> SSE -> Gaussian log-lik -> per-cell argmax -> vote kernel
// (replaces R/scregclust.R:1590-1602,1628-1638 + src/utils.cpp:12-63 +
// src/allocation.cpp:59-94 without ever materialising the K x n2 x T array).
#pragma once
#include "dmat.cuh"
#include "gemm_tn.cuh"

namespace scrc {

constexpr int RV_BM = 64;   // cells per tile
constexpr int RV_BN = 32;   // targets per tile
constexpr int RV_STAGES = 4;   // the loads are never the limiter (profiles/): a short ring leaves shared memory for two CTAs per SM
constexpr int RV_MAX_K = 64;  // modules (+1 pseudo module) kept in shared memory

struct ResidVoteArgs {
    // stacked, 16-row padded selections for all (penalty, module) pairs
    const double* ZselT; int64_t ldzs;   // rows_total x ncells  (k contiguous)
    const double* beta; int64_t ldb;     // rows_total x T
    int64_t rows_total;
    const double* Z; int64_t ldz;        // ncells x T targets (centred)
    int ncells, T, K, L;
    const int* row_off;                  // device [L*K]
    const int* kpad;                     // device [L*K]  (0 = module without a model)
    const int* nsel = nullptr;           // device [L*K], optional: selected rows (<= kpad; the rest of a block is zero)
    const double* two_var;               // device [L][K][T]   (vote mode)
    const double* inv_two_var = nullptr; // device [L][K][T]   (vote mode): RN(1 / two_var)
    const double* half_log;              // device [L][K][T]   (vote mode)
    int* votes;                          // device [L][T][K]   (vote mode, pre-zeroed)
    int CR;                              // cell ranges per (penalty, target tile)
    // optional generalisations (used by the leave-one-regulator-out post-processing):
    int64_t rows_total_a = -1;           // rows of ZselT when they differ from the rows of `beta` (-1: same)
    const int* row_off_b = nullptr;      // device [L*K]: rows of `beta` when they differ from the rows of ZselT
    const int* col_index = nullptr;      // device: flattened per-fit lists of target columns of Z
    const int* col_off = nullptr;        // device [L]: start of fit l's list in col_index
    const int* col_cnt = nullptr;        // device [L]: number of columns of fit l (<= T = columns of beta)
    double* sq_out = nullptr;            // mode 2: K x ncells x T squared residuals (module fastest, one fit)
    int t_begin = 0, t_end = -1;         // target slice [t_begin, t_end) handled by this launch (-1: T)
};
int resid_vote_pick_cr(int ncells, int T, int L, int num_sms);
// mode 1: per-cell votes, 2: squared residuals of one fit written to sq_out (network-prior sweep)
void resid_vote(const ResidVoteArgs& a, int mode, int num_sms, cudaStream_t s);

// SSE of every (fit, module, target) from the Gram matrices of a data split instead of its cells (used for the
// training split AND the test split: pass that split's X'X, X'y and y'y):
//   ||y_t - X_sel b||^2 = y_t'y_t - 2 b'(X_sel'y_t) + b'(X_sel'X_sel) b
// with y'y = gtt[t], X'y = G_rt[sel, t], X'X = G_rr[sel, sel]  (same quantity as
// colSums(residuals_train_nnls^2), R/scregclust.R:1590-1602, in O(s^2) instead of O(n1 s) per pair).
// Modules without a model get gtt[t] exactly.  Output layout = sse[l][j][t] with K+1 module slots; with `pseudo`
// slot K of every fit is filled with gtt[t] (the "no model" sum of squares, denominator of R2).  sel_row_off
// (optional, [L*K]): rows of `sel` when they differ from the rows of `beta` (importance step: the s + 1 leave-one-out
// problems of a module share one regulator list).
void sse_train_gram(const double* G_rr, int64_t ldrr, const double* G_rt, int64_t ldrt, const double* gtt,
                    const double* beta, int64_t ldb, const int* sel /* per stacked row: regulator or -1 */,
                    const int* row_off, const int* nsel, int T, int K, int L, double* sse, cudaStream_t s, int t0 = 0,
                    int t1 = -1, const int* col_index = nullptr, const int* col_off = nullptr,
                    const int* col_cnt = nullptr, bool pseudo = false, const int* sel_row_off = nullptr);

// resid_var = SSE_train / (n1 - s_j); two_var = 2 var; inv_two_var = RN(1 / two_var); half_log = 0.5 log(2 pi var)
// (R/scregclust.R:1635-1638, allocation.cpp:59-62).  nsel: device [L*K].
void make_resid_var(const double* sse_train, const int* nsel, int n1, int T, int K, int L, double* var,
                    double* two_var, double* inv_two_var, double* half_log, cudaStream_t s, int t0 = 0, int t1 = -1);

// r2[l][t][j] = 1 - SSE_test / SS0 ; best_r2 = row max ; best_idx = first argmax (1-based)
// (R/scregclust.R:1628-1632, 1646).  k_new = first argmax of votes (1-based, allocation.cpp:88-94).
void finish_alloc(const double* sse_test, const int* votes, int T, int K, int L, double* r2, double* best_r2,
                  int* best_idx, int* k_new, cudaStream_t s, int t0 = 0, int t1 = -1);

// out[r, c] = row_src[r] >= 0 ? Zr[c, row_src[r]] : 0   (out: rows x ncells, k contiguous)
void gather_transpose(const double* Zr, int64_t ldz, int ncells, const int* row_src, int64_t rows,
                      double* out, int64_t ldo, cudaStream_t s);

}  // namespace scrc
