// Batched accelerated anti-lopsided NNLS (replaces src/optim.cpp:335-377, 403-549).
//
// Every (module, target column) pair is an independent small problem of size
// s = #selected regulators; the per-column result of the reference does not
// depend on the other columns.  One warp owns one column;
// the CTA's warps share the module's Q matrix in shared memory.
#pragma once
#include <vector>

#include "dmat.cuh"

namespace scrc {

// One NNLS problem group = one (penalty, module): `s` variables, all `m` columns.
struct NnlsProblem {
    int s;            // number of variables (selected regulators)
    int sel_off;      // offset into sel / sign (/ pos) arrays
    int64_t q_off;    // offset of Q (s x s, ld = s) in the Q workspace
    int64_t row_off;  // first row of this problem in the stacked grad0 / beta_hat matrices
    // --- optional generalisations (defaults reproduce the plain "all columns" problem) ---
    int ncols;        // number of columns of this problem (<= batch m); -1 = all m columns
    int xcol_off;     // local column c reads column cols[xcol_off + c] of xty / colscale; -1 = identity
    int pos_off;      // variable a is written to row row_off + pos[pos_off + a] of beta; -1 = identity
    int scale_off;    // out_scale window starts at out_scale[scale_off] ...
    int scale_len;    // ... and has this length (R's recycling modulus); 0 = use the batch default
    int pad_;
};
inline NnlsProblem make_nnls_problem(int s, int sel_off, int64_t q_off, int64_t row_off) {
    NnlsProblem p;
    p.s = s; p.sel_off = sel_off; p.q_off = q_off; p.row_off = row_off;
    p.ncols = -1; p.xcol_off = -1; p.pos_off = -1; p.scale_off = 0; p.scale_len = 0; p.pad_ = 0;
    return p;
}

struct NnlsBatch {
    int nprob = 0;
    int m = 0;                     // columns (targets)
    int64_t rows = 0;              // stacked rows (sum of 16-padded s)
    DevBuf<NnlsProblem> prob;      // device copy
    DevBuf<int> sel;               // selected variable indices (into xtx / xty rows)
    DevBuf<double> sign;           // +-1 per selected variable
    DevBuf<int> pos;               // optional output row positions (see NnlsProblem::pos_off)
    DevBuf<int> cols;              // optional column index lists (see NnlsProblem::xcol_off)
    DevBuf<double> Q;              // normalised Gram blocks
    DevBuf<double> dvec;           // 1/sqrt(diag) per stacked row
    DMat grad0;                    // rows x m : initial gradient
    DMat beta;                     // rows x m : output (rescaled; padding rows = 0)
    DevBuf<int> iters;             // nprob x m
    // host side
    std::vector<NnlsProblem> h_prob;   // host copy of `prob` (size nprob), set by the caller
    int col0 = 0, col1 = -1;           // column slice [col0, col1) to set up / solve (-1: m)
    bool use_active = false;           // true: only the problems listed in `active` are set up / solved
    std::vector<int> active;
    DevBuf<int> ids;                   // device scratch for the launch lists (kept across calls)
    std::vector<int> h_ids;            // pinned lifetime of the async upload of `ids`
};

// Builds Q, dvec and grad0 for every problem from the full Gram matrices:
//   xtx: n_var_total x n_var_total (ldx), xty: n_var_total x m (ldy),
//   colscale[m] multiplies the columns of xty (1 / z1_target_sds; nullptr = 1).
void nnls_setup(NnlsBatch& b, const double* xtx, int64_t ldx, const double* xty, int64_t ldy,
                const double* colscale, cudaStream_t s);

// Solves all problems.  out_scale (length n_out_scale, may be nullptr): the output
// element (a, t) of a problem with s variables is multiplied by
//   sign[a] * out_scale[(a + t*s) mod n_out_scale]
// which reproduces R's vector recycling in R/scregclust.R:1588.
void nnls_solve(NnlsBatch& b, double eps, int max_iter, const double* out_scale, int n_out_scale,
                int num_sms, cudaStream_t s);

}  // namespace scrc
