// Batched cooperative-lasso ADMM (replaces src/optim.cpp:63-320 for all modules
// and penalization values of a cycle at once).
//
// All problems share the design matrix (R/scregclust.R:1364 passes the same
// z1_reg_scaled / sqrt(n1) to every coop_lasso call), so the linear solve
//     beta = (X'X + rho I)^-1 (X'Y + rho zeta + mult)          (optim.cpp:136-143)
// is evaluated in the eigenbasis X'X = V diag(lam) V':
//     W = diag(1 / (lam + rho_problem)) V' R ,   beta = V W
// i.e. two FP64 tensor-core GEMMs over the concatenated columns of every active
// problem, with the per-problem rho applied in the first GEMM's epilogue.  No
// refactorisation is ever needed when rho changes (the reference refactors a
// P x P LDL^T per problem per rho change).
#pragma once
#include <atomic>
#include <vector>

#include "dmat.cuh"
#include "gemm_tn.cuh"

namespace scrc {

constexpr int ADMM_NSUM = 11;      // reductions per problem per iteration
constexpr int ADMM_SLAB = 32;      // rows per update CTA (one lane per row)
constexpr int ADMM_UPD_WARPS = 8;  // column-interleaved warps per update CTA

struct AdmmOpts {
    double rho0 = 0.2, alpha0 = 1.5, eps_corr = 0.2;
    int n_update = 2;
    int max_iter = 1000;
    double eps_rel = 1e-8, eps_abs = 1e-12;
};

// Eigen-decomposition shared by every problem of a session / drop-in call.
struct EigBasis {
    int P = 0;
    DMat V;    // P x P, columns = eigenvectors of X'X
    DMat Vt;   // V transposed
    DevBuf<double> lam;  // P eigenvalues (ascending)
};

// Device-resident batch of problems.  Columns of all problems are concatenated;
// problem p owns columns [c0[p], c1[p]).
struct AdmmBatch {
    int P = 0, C = 0, nprob = 0, nslab = 0;
    // per column / per problem descriptors
    DevBuf<int> col_prob;
    DevBuf<int> prob_c0, prob_c1;
    DevBuf<double> prob_lambda;
    DevBuf<double> prob_rho, prob_alpha;
    DevBuf<int> prob_active, prob_iters;
    DevBuf<int> n_active;      // [1]
    DevBuf<unsigned long long> col_iters;   // [1] sum over sweeps of active columns
    unsigned long long last_col_iters = 0;
    DMat w;                    // P x nprob group weights
    // state (P x C)
    DMat xty, R, W, beta, zeta, mult, beta_old, zeta_old, mult_old, mult_hat_old;
    DevBuf<double> partial;    // nprob x nslab x ADMM_NSUM
    DevBuf<int> prob_R;        // rows per update slab (32 / 16 / 8), chosen from the column count
    DevBuf<int> class_ids;     // problem ids grouped by slab height
    DevBuf<int> tiles;         // [n live 64-col tiles, n live 128-col tiles, list64..., list128...] (admm.cu)
    std::vector<int> h_cols;   // host copy of the column count of every problem (set by the caller)
    bool aborted = false;      // set by admm_solve when the cancel flag stopped it early

    void resize(int P_, int C_, int nprob_);
};

// Runs the lock-step batched ADMM until every problem has converged or hit
// max_iter.  Expects xty, w, prob_* descriptors and (optionally warm-started)
// beta/zeta to be set up; zeroes the remaining state itself when `cold`.
// Returns the number of batched sweeps executed.
// `cancel` (optional) is polled with the device counters; when set the solve stops early and b.aborted is true.
int admm_solve(AdmmBatch& b, const EigBasis& eig, const AdmmOpts& o, bool cold, int num_sms,
               cudaStream_t stream, int* h_pinned_scratch, const std::atomic<int>* cancel = nullptr);

}  // namespace scrc
