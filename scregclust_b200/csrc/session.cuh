// Device-resident scregclust session: data splits, Gram matrices, eigenbasis,
// initial fit and the per-cycle batched pipeline.
#pragma once
#include <vector>

#include "../../include/scregclust_b200.h"
#include <atomic>

#include "admm.cuh"
#include "alloc.cuh"
#include "comm.cuh"
#include "misc.cuh"
#include "nnls.cuh"

namespace scrc {

struct Device {
    int id = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    int* h_pinned = nullptr;   // small pinned scratch (16 ints)
    void init(int device);
    void destroy();
};

// Builds eig (V, Vt, lam) of the symmetric matrix G * scale (G: P x P device matrix).
void build_eig(const DMat& G, double scale, EigBasis& eig, Device& dev);

// out = V diag(1 / (mult * lam + shift)) V' * B     (normal-equation solves, R/utils.R:13-57)
void eig_solve(const EigBasis& eig, double mult, double shift, double cut, const DMat& B, DMat& out, DMat& tmp,
               Device& dev);

struct FitSlot {        // bookkeeping of one (fit, module) after Step 1
    int s = 0;          // selected regulators
    int row_off = 0;    // first stacked row
    int sel_off = 0;    // first entry in the stacked selection list (nnls.sel)
};

struct Ctx {
    Device dev;
    int64_t n1 = 0, n2 = 0, P = 0, T = 0;
    int K = 0;
    // Multi-GPU: with a communicator every call below is COLLECTIVE -- all ranks call it with
    // the same arguments; Step-1 problems are spread over the ranks, Step 2 runs on target slices, and the
    // outputs are merged so that every rank returns the complete result.  Not owned.
    Comm* comm = nullptr;
    int rank() const { return comm ? comm->rank : 0; }
    int world() const { return comm ? comm->world : 1; }
    // Cooperative cancellation (the reference polls Rcpp::checkUserInterrupt() in its loops,
    // optim.cpp:307,536, allocation.cpp:96): set from any thread / signal handler, polled at the ADMM
    // host polls and between the phases of a cycle; the call then fails with SCRC_ERR_INTERRUPTED.
    std::atomic<int> cancel{0};
    // predicted ADMM iterations per (fit, module) of the NEXT fit_cycle call (balances the ranks); consumed
    std::vector<int> cost_hint;
    DMat Z1r, Z2r, Z1t, Z2t;
    DevBuf<double> sds_t;        // z1_target_sds
    DevBuf<double> inv_sds_t;    // 1 / z1_target_sds
    DMat G_rr, G_rt;             // Z1r'Z1r, Z1r'Z1t
    DMat G2_rr, G2_rt;           // Z2r'Z2r, Z2r'Z2t: the test-split sums of squares come from these (alloc.cuh: sse_train_gram)
    DevBuf<double> gtt2;         // colSums(z2_target_centered^2)
    EigBasis eig;                // of G_rr / n1
    DMat beta_init;              // P x T
    DevBuf<double> res_sds;      // T
    DMat bas;                    // beta_adapt_sq, P x T
    double init_df = 0.0;

    // per-cycle state
    AdmmBatch admm;
    DevBuf<int> col_target;
    DevBuf<int> prob_fit, prob_mod;
    NnlsBatch nnls;
    DevBuf<int> row_src, row_off_d, kpad_d, nsel_d, sel_off_d, iters_fk_d;
    int* h_fk = nullptr;          // pinned: [FK admm iterations | abort flag | FK selected counts]
    size_t h_fk_cap = 0;
    DMat ZselT;
    DevBuf<double> sse_train, sse_test, var, two_var, inv_two_var, half_log, r2, best_r2;
    DevBuf<int> votes, best_idx, k_new;
    DevBuf<unsigned char> models_d;
    DevBuf<double> weights_d, signs_d;
    std::vector<FitSlot> slots;  // F*K, from the last cycle
    int last_F = 0;

    void create(int device, const double* z1r, const double* z2r, const double* z1t, const double* z2t,
                const double* sds, int64_t n1, int64_t n2, int64_t P, int64_t T, int K, Comm* comm = nullptr);
    // Raw p x n expression matrix in; split, per-stratum centring, constant-gene filter and scaling on
    // the device (stage.cu).  keep_gene[p] (may be null) marks the genes that survive the filter;
    // dims_out[4] = n1, n2, P, T after filtering.
    void create_raw(int device, const double* expr, int64_t p, int64_t n, const int* is_regulator,
                    const int* split_indices, const int* stratification, int center, int K,
                    unsigned char* keep_gene, int64_t* dims_out);
    void begin(int64_t n1, int64_t n2, int64_t P, int64_t T, int K);   // dimensions + input buffers
    void finish(cudaEvent_t inputs_ready = nullptr);                    // Gram, eigenbasis, initial fit
    void fit_cycle(int F, const double* lambda, const int* k_in, const scrc_options& opt, bool skip_alloc,
                   scrc_cycle_out* out, const scrc_prior* prior = nullptr, const scrc_quick* quick = nullptr);
    // quick mode (scregclust.R:1435-1496): |cross_corr1| (lazy), per-fit target subsets of the last cycle
    DMat abs_corr;
    DevBuf<unsigned char> q_areg;
    DevBuf<int> q_idx, q_off, q_cnt, q_inv, q_bi, q_kn, q_votes;
    DevBuf<double> q_mx, q_sds, q_r2, q_var, q_br;
    bool last_quick = false;
    DevBuf<double> sq_arr;       // K x n2 x T squared residuals, only with a network prior
    // Leave-one-regulator-out importance and module R2 (R/scregclust.R:1823-1912) for F final configurations.
    // r2_removed (optional): the per-target matrices of scregclust.R:1890-1898, packed in (fit, module) order --
    // see scrc_importance_ex in the public header.
    void importance(int F, const int* k_final, const unsigned char* models, const double* signs,
                    const scrc_options& opt, double* r2_module, double* importance_out,
                    double* r2_removed = nullptr, int64_t r2_removed_capacity = 0);
    // scratch of the post-processing and of the coefficient packing, kept across calls (cudaMalloc / cudaFree
    // synchronise the device and cost more than the kernels they serve)
    NnlsBatch imp_nb;
    DevBuf<double> imp_scale, imp_sse;
    DevBuf<int> imp_row_src, imp_arow, imp_brow, imp_kpad, imp_coff, imp_ccnt;
    std::vector<double> imp_h_sse, h_sds;
    DevBuf<int> pack_src, pack_s, pack_a;
    DevBuf<long long> pack_base;
    DevBuf<double> pack_out;
    DevBuf<double> z2t_css;      // centred sums of squares of the test targets (lazy)
    DevBuf<double> gtt;          // colSums(z1_target_centered^2)
    // All K coefficient matrices of fit f of the last (skip_allocation) cycle, packed; see scrc_ctx_get_coeffs_fit.
    void get_coeffs_fit(int f, double* coeffs_out, int* sel_out);
    void pack_coeffs_fit(int f, double* out_dev);     // the same, packed into a device buffer
    void check_cancel() const {
        if (cancel.load(std::memory_order_relaxed)) throw Error(SCRC_ERR_INTERRUPTED, "interrupted by scrc_ctx_request_cancel");
    }
    double phase_s[16] = {0};     // SCRC_DEBUG_TIMING accumulators (session.cu: kPhaseNames)
    void print_phase_times() const;
    ~Ctx() { print_phase_times(); if (h_fk) cudaFreeHost(h_fk); dev.destroy(); }
};

// O(T + nnz) host checks of caller data that index device memory in the network-prior sweep; throws SCRC_ERR_ARG.
void validate_prior_inputs(const char* who, const int* k, int64_t T, int K, const int* prior_off, const int* prior_idx,
                           const int* order);

// kmeanspp() of R/utils.R:351-375 on a device-resident n x p matrix (kmeans.cu); the caller owns R's RNG draws
void kmeanspp_device(cudaStream_t st, const double* x_dev, int64_t ldx, int64_t n, int64_t p, int K, int n_init,
                     int n_max_iter, const int* first_centers, const double* uniforms, int* cluster_out,
                     double* tot_withinss_out, int* ifault_out);

// drop-in helpers (session.cu)
void dropin_coop_lasso(Device& dev, const double* y, int64_t n, int64_t m, const double* x, int64_t p, double lambda,
                       const double* weights, const double* beta0, const AdmmOpts& o, double* beta_out, int* iters);
void dropin_nnls(Device& dev, const double* x, int64_t nobs, int64_t nvar, const double* y, int64_t m, double eps,
                 int max_iter, double* beta_out, int* iters);
void dropin_fast_cor(Device& dev, const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out);
void dropin_ols(Device& dev, const double* y, const double* x, int64_t n, int64_t p, int64_t m, double lambda,
                bool ridge, double* beta_out);
void dropin_crossprod(Device& dev, const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out);
// Sequential Gauss-Seidel allocation sweep on a device-resident array (allocation.cpp:33-97).
void alloc_aggregate_device(const double* d_sse_test, const double* d_var, int K, int64_t T, int F, double n2,
                            const int* d_k_in, const int* d_prior_off, const int* d_prior_idx, const int* d_order,
                            double baseline, double weight, double* d_mlp, int* d_k_out, cudaStream_t st);
void alloc_sequential_device(const double* d_arr, int K, int64_t n2, int64_t T, const double* d_resid_var,
                             const int* d_prior_off, const int* d_prior_idx, int* d_k /* 0-based, -2 noise, in/out */,
                             const int* d_order, double baseline, double weight, int* d_votes, cudaStream_t st);
void dropin_allocate(Device& dev, const double* sq_resid, int K, int64_t n2, int64_t T, const double* resid_var,
                     const int* prior_off, const int* prior_idx, const int* k_in, const int* order, double baseline,
                     double weight, int* k_out, int* votes_out);

}  // namespace scrc

// the opaque handle of the public header
struct scrc_ctx {
    scrc::Ctx impl;
};
struct scrc_comm {
    scrc::Comm impl;
};
extern "C" void scrc_set_last_error_(const char* msg);   // capi.cu: message returned by scrc_last_error()
