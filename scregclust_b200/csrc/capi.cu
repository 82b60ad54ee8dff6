// extern "C" boundary of libscregclust_b200 (see include/scregclust_b200.h).
#include <memory>
#include <mutex>

#include "prof.cuh"
#include "session.cuh"
#include "shard.hpp"

using namespace scrc;

namespace {
thread_local std::string g_err;
std::mutex g_dev_mutex;
std::unique_ptr<Device> g_dev;     // device used by the drop-in entry points
int g_dev_id = 0;

Device& dropin_device() {
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    if (!g_dev || g_dev->id != g_dev_id) {
        if (g_dev) g_dev->destroy();
        g_dev.reset(new Device());
        g_dev->init(g_dev_id);
    }
    SCRC_CUDA(cudaSetDevice(g_dev->id));
    return *g_dev;
}

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

template <class F>
int guarded(F&& f) {
    try {
        f();
        return SCRC_OK;
    } catch (const Error& e) {
        return fail(e.code, e.what());
    } catch (const std::bad_alloc&) {
        return fail(SCRC_ERR_INTERNAL, "host allocation failed");
    } catch (const std::exception& e) {
        return fail(SCRC_ERR_INTERNAL, e.what());
    } catch (...) {
        return fail(SCRC_ERR_INTERNAL, "unknown failure");
    }
}
}  // namespace


extern "C" {

void scrc_set_last_error_(const char* msg) { g_err = msg ? msg : ""; }

int scrc_version(void) { return 201; }
const char* scrc_last_error(void) { return g_err.c_str(); }
unsigned long long scrc_launch_count(void) { return scrc::g_launch_count.load(); }
int scrc_set_device(int device) {
    if (device < 0) return fail(SCRC_ERR_ARG, "scrc_set_device: negative device index");
    g_dev_id = device;
    return SCRC_OK;
}

void scrc_default_options(scrc_options* o) {
    if (!o) return;
    o->max_optim_iter = 10000; o->tol_coop_rel = 1e-8; o->tol_coop_abs = 1e-12; o->tol_nnls = 1e-4;
    o->rho0 = 0.2; o->alpha0 = 1.5; o->n_update = 2; o->eps_corr = 0.2;
}

int scrc_coop_lasso(const double* y, int64_t n, int64_t m, const double* x, int64_t n_x, int64_t p, double lambda,
                    const double* weights, const double* beta0, int64_t beta0_rows, int64_t beta0_cols, double rho0,
                    double alpha0, int n_update, double eps_corr, int max_iter, double eps_rel, double eps_abs,
                    double* beta_out, int* iterations_out) {
    // argument checks in the reference's order (optim.cpp:76-122)
    if (n <= 0 || m <= 0 || p <= 0 || n_x <= 0)
        return fail(SCRC_ERR_COOP_DIMS, "COOP LASSO: Matrix dimensions of y and x need to be positive.");
    if (n_x != n) return fail(SCRC_ERR_COOP_ROWS, "y and x need to have the same number of rows.");
    if (beta0 && (beta0_rows != p || beta0_cols != m))
        return fail(SCRC_ERR_COOP_BETA0, "beta_0 needs to be of size p x m");
    if (!(rho0 > 0.0)) return fail(SCRC_ERR_COOP_RHO, "rho_0 > 0 needs to hold");
    if (alpha0 <= 0.0 || alpha0 > 2.0) return fail(SCRC_ERR_COOP_ALPHA, "0 < alpha_0 < 2 needs to hold");
    if (!y || !x || !weights || !beta_out || !iterations_out || n_update <= 0)
        return fail(SCRC_ERR_ARG, "scrc_coop_lasso: null pointer or n_update <= 0");
    return guarded([&] {
        AdmmOpts o;
        o.rho0 = rho0; o.alpha0 = alpha0; o.n_update = n_update; o.eps_corr = eps_corr; o.max_iter = max_iter;
        o.eps_rel = eps_rel; o.eps_abs = eps_abs;
        dropin_coop_lasso(dropin_device(), y, n, m, x, p, lambda, weights, beta0, o, beta_out, iterations_out);
    });
}

int scrc_coef_nnls(const double* x, int64_t nobs, int64_t nvar, const double* y, int64_t nobs_y, int64_t m, double eps,
                   int max_iter, double* beta_out, int* iterations_out) {
    if (nvar <= 0 || m <= 0 || nobs <= 0 || nobs_y <= 0)
        return fail(SCRC_ERR_NNLS_DIMS, "NNLS: Matrix dimensions of y and x need to be positive.");
    if (nobs != nobs_y) return fail(SCRC_ERR_ARG, "scrc_coef_nnls: x and y need the same number of rows");
    if (!x || !y || !beta_out || !iterations_out) return fail(SCRC_ERR_ARG, "scrc_coef_nnls: null pointer");
    return guarded([&] { dropin_nnls(dropin_device(), x, nobs, nvar, y, m, eps, max_iter, beta_out, iterations_out); });
}

int scrc_allocate_into_modules(const double* sq_resid, int K, int64_t n2, int64_t T, const double* resid_var,
                               const int* prior_off, const int* prior_idx, const int* k_in, const int* update_order,
                               double prior_baseline, double prior_weight, int* k_out, int* votes_out) {
    if (!sq_resid || !resid_var || !k_in || !update_order || !k_out || K <= 0 || n2 <= 0 || T <= 0)
        return fail(SCRC_ERR_ARG, "scrc_allocate_into_modules: invalid argument");
    return guarded([&] {
        dropin_allocate(dropin_device(), sq_resid, K, n2, T, resid_var, prior_off, prior_idx, k_in, update_order,
                        prior_baseline, prior_weight, k_out, votes_out);
    });
}

int scrc_alloc_array(const double* input, int64_t n_obs, int64_t n_genes, int64_t n_cl, double* out) {
    if (!input || !out || n_obs < 0 || n_genes < 0 || n_cl < 0) return fail(SCRC_ERR_ARG, "scrc_alloc_array: invalid argument");
    const long double total = (long double)n_cl * (long double)n_obs * (long double)n_genes;
    if (total > 4503599627370496.0L)   // R_XLEN_T_MAX = 2^52 (utils.cpp:19-21)
        return fail(SCRC_ERR_ALLOC_TOO_LARGE, "alloc_array: requested allocation too large");
    for (int64_t i = 0, ub = n_obs * n_genes; i < ub; ++i)
        for (int64_t j = 0; j < n_cl; ++j) out[i * n_cl + j] = input[i];
    return SCRC_OK;
}

int scrc_reset_array(double* arr, int64_t n_cl, int64_t n_obs, int64_t n_genes, const double* input, int64_t in_rows,
                     int64_t in_cols) {
    if (!arr || !input) return fail(SCRC_ERR_ARG, "scrc_reset_array: null pointer");
    if (in_rows != n_obs || in_cols != n_genes) return fail(SCRC_ERR_RESET_DIMS, "reset_array: input has wrong dimensions");
    for (int64_t i = 0, ub = n_obs * n_genes; i < ub; ++i)
        for (int64_t j = 0; j < n_cl; ++j) arr[i * n_cl + j] = input[i];
    return SCRC_OK;
}

int scrc_fast_cor(const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out) {
    if (!x || !y || !out || n <= 0 || px <= 0 || py <= 0) return fail(SCRC_ERR_ARG, "scrc_fast_cor: invalid argument");
    return guarded([&] { dropin_fast_cor(dropin_device(), x, y, n, px, py, out); });
}
int scrc_coef_ols(const double* y, const double* x, int64_t n, int64_t p, int64_t m, double* beta_out) {
    if (!x || !y || !beta_out || n <= 0 || p <= 0 || m <= 0) return fail(SCRC_ERR_ARG, "scrc_coef_ols: invalid argument");
    return guarded([&] { dropin_ols(dropin_device(), y, x, n, p, m, 0.0, false, beta_out); });
}
int scrc_coef_ridge(const double* y, const double* x, int64_t n, int64_t p, int64_t m, double lambda, double* beta_out) {
    if (!x || !y || !beta_out || n <= 0 || p <= 0 || m <= 0) return fail(SCRC_ERR_ARG, "scrc_coef_ridge: invalid argument");
    return guarded([&] { dropin_ols(dropin_device(), y, x, n, p, m, lambda, true, beta_out); });
}
int scrc_crossprod(const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out) {
    if (!x || !y || !out || n <= 0 || px <= 0 || py <= 0) return fail(SCRC_ERR_ARG, "scrc_crossprod: invalid argument");
    return guarded([&] { dropin_crossprod(dropin_device(), x, y, n, px, py, out); });
}

__global__ void fill_pseudo_kernel(double* x, int64_t n, unsigned seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    x[i] = ((double)(h & 0xffff) - 32768.0) / 32768.0;
}

int scrc_bench_gemm(int64_t k, int64_t m, int64_t n, int symmetric, int reps, double* ms_per_launch) {
    if (k <= 0 || m <= 0 || n <= 0 || reps <= 0 || !ms_per_launch) return fail(SCRC_ERR_ARG, "scrc_bench_gemm: invalid argument");
    return guarded([&] {
        Device& dev = dropin_device();
        cudaStream_t st = dev.stream;
        AllocScope as(st);
        DMat A(k, m), B(k, symmetric ? m : n), Cm(m, symmetric ? m : n);
        fill_pseudo_kernel<<<(unsigned)ceil_div((int64_t)A.buf.n, 256), 256, 0, st>>>(A.data(), (int64_t)A.buf.n, 1u);
        fill_pseudo_kernel<<<(unsigned)ceil_div((int64_t)B.buf.n, 256), 256, 0, st>>>(B.data(), (int64_t)B.buf.n, 2u);
        DenseArgs a{A.data(), A.ld, symmetric ? A.data() : B.data(), symmetric ? A.ld : B.ld, Cm.data(), Cm.ld,
                    (int)m, (int)(symmetric ? m : n), (int)k};
        const int mode = symmetric ? EPI_SYM_LOWER : EPI_PLAIN;
        for (int i = 0; i < 2; ++i) gemm_tn_dense(mode, a, dev.num_sms, st);
        cudaEvent_t e0, e1;
        SCRC_CUDA(cudaEventCreate(&e0)); SCRC_CUDA(cudaEventCreate(&e1));
        SCRC_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) gemm_tn_dense(mode, a, dev.num_sms, st);
        SCRC_CUDA(cudaEventRecord(e1, st));
        SCRC_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SCRC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        *ms_per_launch = (double)ms / reps;
    });
}

int scrc_profile_enable(int on) {
    return guarded([&] {
        if (!on) g_prof.collect();
        g_prof.enabled = (on != 0);
    });
}
int scrc_profile_reset(void) { return guarded([&] { g_prof.reset(); }); }
int scrc_profile_get(double* ms, double* work, unsigned long long* launches) {
    return guarded([&] {
        g_prof.collect();
        for (int i = 0; i < PROF_NCLASS; ++i) {
            if (ms) ms[i] = g_prof.ms[i];
            if (work) work[i] = g_prof.work[i];
            if (launches) launches[i] = g_prof.launches[i];
        }
    });
}

namespace {
cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
}
int scrc_timer_begin(void) {
    return guarded([&] {
        if (!g_t0) { SCRC_CUDA(cudaEventCreate(&g_t0)); SCRC_CUDA(cudaEventCreate(&g_t1)); }
        SCRC_CUDA(cudaDeviceSynchronize());
        SCRC_CUDA(cudaEventRecord(g_t0, cudaStreamPerThread));
    });
}
int scrc_timer_end(double* ms) {
    if (!ms) return fail(SCRC_ERR_ARG, "scrc_timer_end: null pointer");
    return guarded([&] {
        if (!g_t0) throw Error(SCRC_ERR_ARG, "scrc_timer_end without scrc_timer_begin");
        SCRC_CUDA(cudaDeviceSynchronize());
        SCRC_CUDA(cudaEventRecord(g_t1, cudaStreamPerThread));
        SCRC_CUDA(cudaEventSynchronize(g_t1));
        float t = 0.f;
        SCRC_CUDA(cudaEventElapsedTime(&t, g_t0, g_t1));
        *ms = t;
    });
}

int scrc_ctx_create(scrc_ctx** out, int device, const double* z1_reg, const double* z2_reg, const double* z1_target,
                    const double* z2_target, const double* z1_target_sds, int64_t n1, int64_t n2, int64_t P, int64_t T,
                    int K) {
    if (!out || !z1_reg || !z2_reg || !z1_target || !z2_target || !z1_target_sds)
        return fail(SCRC_ERR_ARG, "scrc_ctx_create: null pointer");
    *out = nullptr;
    return guarded([&] {
        std::unique_ptr<scrc_ctx> c(new scrc_ctx());
        c->impl.create(device, z1_reg, z2_reg, z1_target, z2_target, z1_target_sds, n1, n2, P, T, K);
        *out = c.release();
    });
}
int scrc_ctx_create_raw(scrc_ctx** out, int device, const double* expr, int64_t p, int64_t n, const int* is_regulator,
                        const int* split_indices, const int* stratification, int center, int K,
                        unsigned char* keep_gene, int64_t* dims_out) {
    if (!out || !expr || !is_regulator || !split_indices) return fail(SCRC_ERR_ARG, "scrc_ctx_create_raw: null pointer");
    *out = nullptr;
    return guarded([&] {
        std::unique_ptr<scrc_ctx> c(new scrc_ctx());
        c->impl.create_raw(device, expr, p, n, is_regulator, split_indices, stratification, center, K, keep_gene, dims_out);
        *out = c.release();
    });
}
int scrc_ctx_get_inputs(scrc_ctx* ctx, double* z1_reg, double* z2_reg, double* z1_target, double* z2_target,
                        double* z1_target_sds) {
    if (!ctx) return fail(SCRC_ERR_ARG, "null context");
    return guarded([&] {
        Ctx& c = ctx->impl;
        SCRC_CUDA(cudaSetDevice(c.dev.id));
        cudaStream_t st = c.dev.stream;
        if (z1_reg) c.Z1r.download(z1_reg, c.n1, st);
        if (z2_reg) c.Z2r.download(z2_reg, c.n2, st);
        if (z1_target) c.Z1t.download(z1_target, c.n1, st);
        if (z2_target) c.Z2t.download(z2_target, c.n2, st);
        if (z1_target_sds) SCRC_CUDA(cudaMemcpyAsync(z1_target_sds, c.sds_t.p, c.T * 8, cudaMemcpyDeviceToHost, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
    });
}
void scrc_ctx_destroy(scrc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->impl.dev.id);
    delete ctx;
}

int scrc_ctx_get_init(scrc_ctx* ctx, double* res_sds, double* init_df, double* beta_init) {
    if (!ctx) return fail(SCRC_ERR_ARG, "null context");
    return guarded([&] {
        Ctx& c = ctx->impl;
        SCRC_CUDA(cudaSetDevice(c.dev.id));
        if (res_sds) SCRC_CUDA(cudaMemcpyAsync(res_sds, c.res_sds.p, c.T * 8, cudaMemcpyDeviceToHost, c.dev.stream));
        if (init_df) *init_df = c.init_df;
        if (beta_init) c.beta_init.download(beta_init, c.P, c.dev.stream);
        SCRC_CUDA(cudaStreamSynchronize(c.dev.stream));
    });
}
int scrc_ctx_get_gram(scrc_ctx* ctx, double* g_rr, double* g_rt) {
    if (!ctx) return fail(SCRC_ERR_ARG, "null context");
    return guarded([&] {
        Ctx& c = ctx->impl;
        SCRC_CUDA(cudaSetDevice(c.dev.id));
        if (g_rr) c.G_rr.download(g_rr, c.P, c.dev.stream);
        if (g_rt) c.G_rt.download(g_rt, c.P, c.dev.stream);
        SCRC_CUDA(cudaStreamSynchronize(c.dev.stream));
    });
}
namespace {
// cross_corr1 = fast_cor(z1_target_centered, z1_reg_scaled) on device copies (R/scregclust.R:1116), T x P
void session_cross_corr(Ctx& c, DMat& C) {
    cudaStream_t st = c.dev.stream;
    DMat X(c.n1, c.T), Y(c.n1, c.P);
    C.alloc(c.T, c.P);
    SCRC_CUDA(cudaMemcpyAsync(X.data(), c.Z1t.data(), c.Z1t.bytes(), cudaMemcpyDeviceToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(Y.data(), c.Z1r.data(), c.Z1r.bytes(), cudaMemcpyDeviceToDevice, st));
    DevBuf<double> xss(c.T), yss(c.P);
    col_center_ss(X.data(), X.ld, c.n1, c.T, xss.p, st);
    col_center_ss(Y.data(), Y.ld, c.n1, c.P, yss.p, st);
    DenseArgs a{X.data(), X.ld, Y.data(), Y.ld, C.data(), C.ld, (int)c.T, (int)c.P, (int)c.n1};
    a.rvec = xss.p; a.cvec = yss.p;
    gemm_tn_dense(EPI_COR, a, c.dev.num_sms, st);
    SCRC_CUDA(cudaStreamSynchronize(st));           // X, Y, xss, yss are released on return
}
}  // namespace

int scrc_ctx_cross_corr(scrc_ctx* ctx, double* out) {
    if (!ctx || !out) return fail(SCRC_ERR_ARG, "null pointer");
    return guarded([&] {
        Ctx& c = ctx->impl;
        SCRC_CUDA(cudaSetDevice(c.dev.id));
        cudaStream_t st = c.dev.stream;
        AllocScope as(st);
        DMat C;
        session_cross_corr(c, C);
        C.download(out, c.T, st);
        SCRC_CUDA(cudaStreamSynchronize(st));
    });
}

int scrc_kmeanspp(const double* x, int64_t n, int64_t p, int n_cluster, int n_init_clusterings, int n_max_iter,
                  const int* first_centers, const double* uniforms, int* cluster_out, double* tot_withinss_out,
                  int* ifault_out) {
    if (!x || !first_centers || !uniforms || !cluster_out || n <= 0 || p <= 0)
        return fail(SCRC_ERR_ARG, "scrc_kmeanspp: null pointer or non-positive dimension");
    return guarded([&] {
        Device& dev = dropin_device();
        cudaStream_t st = dev.stream;
        AllocScope as(st);
        DMat X(n, p);
        X.upload(x, n, st);
        kmeanspp_device(st, X.data(), X.ld, n, p, n_cluster, n_init_clusterings, n_max_iter, first_centers, uniforms,
                        cluster_out, tot_withinss_out, ifault_out);
    });
}

int scrc_ctx_kmeanspp(scrc_ctx* ctx, int n_init_clusterings, int n_max_iter, const int* first_centers,
                      const double* uniforms, int* cluster_out, double* tot_withinss_out, int* ifault_out) {
    if (!ctx || !first_centers || !uniforms || !cluster_out) return fail(SCRC_ERR_ARG, "scrc_ctx_kmeanspp: null pointer");
    return guarded([&] {
        Ctx& c = ctx->impl;
        SCRC_CUDA(cudaSetDevice(c.dev.id));
        cudaStream_t st = c.dev.stream;
        AllocScope as(st);
        DMat C;
        session_cross_corr(c, C);                   // the matrix R clusters (R/scregclust.R:1131-1136), never leaves the device
        kmeanspp_device(st, C.data(), C.ld, c.T, c.P, c.K, n_init_clusterings, n_max_iter, first_centers, uniforms,
                        cluster_out, tot_withinss_out, ifault_out);
    });
}

int scrc_fit_cycle(scrc_ctx* ctx, int F, const double* lambda, const int* k_in, const scrc_options* opt,
                   int skip_allocation, scrc_cycle_out* out) {
    if (!ctx || !lambda || !k_in) return fail(SCRC_ERR_ARG, "scrc_fit_cycle: null pointer");
    scrc_options o;
    if (opt) o = *opt; else scrc_default_options(&o);
    if (o.max_optim_iter <= 0 || o.n_update <= 0 || !(o.rho0 > 0.0) || o.alpha0 <= 0.0 || o.alpha0 > 2.0)
        return fail(SCRC_ERR_ARG, "scrc_fit_cycle: invalid options");
    return guarded([&] { ctx->impl.fit_cycle(F, lambda, k_in, o, skip_allocation != 0, out); });
}

int scrc_fit_cycle_prior(scrc_ctx* ctx, int F, const double* lambda, const int* k_in, const scrc_options* opt,
                         const scrc_prior* prior, scrc_cycle_out* out) {
    if (!ctx || !lambda || !k_in || !prior) return fail(SCRC_ERR_ARG, "scrc_fit_cycle_prior: null pointer");
    scrc_options o;
    if (opt) o = *opt; else scrc_default_options(&o);
    if (o.max_optim_iter <= 0 || o.n_update <= 0 || !(o.rho0 > 0.0) || o.alpha0 <= 0.0 || o.alpha0 > 2.0)
        return fail(SCRC_ERR_ARG, "scrc_fit_cycle_prior: invalid options");
    return guarded([&] { ctx->impl.fit_cycle(F, lambda, k_in, o, false, out, prior); });
}

int scrc_fit_cycle_quick(scrc_ctx* ctx, int F, const double* lambda, const int* k_in, const scrc_options* opt,
                         int skip_allocation, const scrc_quick* quick, scrc_cycle_out* out) {
    if (!ctx || !lambda || !k_in) return fail(SCRC_ERR_ARG, "scrc_fit_cycle_quick: null pointer");
    scrc_options o;
    if (opt) o = *opt; else scrc_default_options(&o);
    if (o.max_optim_iter <= 0 || o.n_update <= 0 || !(o.rho0 > 0.0) || o.alpha0 <= 0.0 || o.alpha0 > 2.0)
        return fail(SCRC_ERR_ARG, "scrc_fit_cycle_quick: invalid options");
    return guarded([&] { ctx->impl.fit_cycle(F, lambda, k_in, o, skip_allocation != 0, out, nullptr, quick); });
}

int scrc_importance(scrc_ctx* ctx, int F, const int* k_final, const unsigned char* models, const double* signs,
                    const scrc_options* opt, double* r2_module, double* importance) {
    if (!ctx || !k_final || !models || !signs || !r2_module || !importance)
        return fail(SCRC_ERR_ARG, "scrc_importance: null pointer");
    scrc_options o;
    if (opt) o = *opt; else scrc_default_options(&o);
    return guarded([&] { ctx->impl.importance(F, k_final, models, signs, o, r2_module, importance); });
}

int scrc_importance_ex(scrc_ctx* ctx, int F, const int* k_final, const unsigned char* models, const double* signs,
                       const scrc_options* opt, double* r2_module, double* importance, double* r2_removed,
                       int64_t r2_removed_capacity) {
    if (!ctx || !k_final || !models || !signs || !r2_module || !importance)
        return fail(SCRC_ERR_ARG, "scrc_importance_ex: null pointer");
    scrc_options o;
    if (opt) o = *opt; else scrc_default_options(&o);
    return guarded([&] {
        ctx->impl.importance(F, k_final, models, signs, o, r2_module, importance, r2_removed, r2_removed_capacity);
    });
}

int scrc_ctx_get_coeffs(scrc_ctx* ctx, int f, int k, double* coeffs_out, int* sel_out) {
    if (!ctx) return fail(SCRC_ERR_ARG, "null context");
    return guarded([&] {
        Ctx& c = ctx->impl;
        if (f < 0 || f >= c.last_F || k < 0 || k >= c.K) throw Error(SCRC_ERR_ARG, "scrc_ctx_get_coeffs: index out of range");
        SCRC_CUDA(cudaSetDevice(c.dev.id));
        AllocScope as(c.dev.stream);
        const FitSlot& sl = c.slots[(size_t)f * c.K + k];
        if (sel_out && sl.s > 0)
            SCRC_CUDA(cudaMemcpyAsync(sel_out, c.nnls.sel.p + sl.sel_off, (size_t)sl.s * 4, cudaMemcpyDeviceToHost, c.dev.stream));
        if (coeffs_out && sl.s > 0) {
            // compact on the device first: a strided device-to-host copy of T rows of s doubles is slow
            DevBuf<double> packed((size_t)sl.s * c.T);
            SCRC_CUDA(cudaMemcpy2DAsync(packed.p, (size_t)sl.s * 8, c.nnls.beta.data() + sl.row_off, c.nnls.beta.ld * 8,
                                        (size_t)sl.s * 8, c.T, cudaMemcpyDeviceToDevice, c.dev.stream));
            SCRC_CUDA(cudaMemcpyAsync(coeffs_out, packed.p, (size_t)sl.s * c.T * 8, cudaMemcpyDeviceToHost, c.dev.stream));
        }
        SCRC_CUDA(cudaStreamSynchronize(c.dev.stream));
    });
}

int scrc_ctx_get_coeffs_fit(scrc_ctx* ctx, int f, double* coeffs_out, int* sel_out) {
    if (!ctx) return fail(SCRC_ERR_ARG, "null context");
    return guarded([&] { ctx->impl.get_coeffs_fit(f, coeffs_out, sel_out); });
}

int scrc_ctx_request_cancel(scrc_ctx* ctx) {
    if (!ctx) return SCRC_ERR_ARG;          // no allocation, no locks: callable from a signal handler
    ctx->impl.cancel.store(1, std::memory_order_relaxed);
    return SCRC_OK;
}
int scrc_ctx_clear_cancel(scrc_ctx* ctx) {
    if (!ctx) return SCRC_ERR_ARG;
    ctx->impl.cancel.store(0, std::memory_order_relaxed);
    return SCRC_OK;
}

int scrc_comm_unique_id(char id_out[128]) {
    if (!id_out) return fail(SCRC_ERR_ARG, "scrc_comm_unique_id: null pointer");
    return guarded([&] { comm_unique_id(id_out); });
}
int scrc_comm_create(scrc_comm** out, const char id[128], int world, int rank, int device) {
    if (!out || !id) return fail(SCRC_ERR_ARG, "scrc_comm_create: null pointer");
    *out = nullptr;
    return guarded([&] {
        std::unique_ptr<scrc_comm> c(new scrc_comm());
        c->impl.init(id, world, rank, device);
        *out = c.release();
    });
}
void scrc_comm_destroy(scrc_comm* comm) { delete comm; }
int scrc_ctx_create_sharded(scrc_ctx** out, scrc_comm* comm, const double* z1_reg, const double* z2_reg,
                            const double* z1_target, const double* z2_target, const double* z1_target_sds, int64_t n1,
                            int64_t n2, int64_t P, int64_t T, int K) {
    if (!out || !comm || !z1_reg || !z2_reg || !z1_target || !z2_target || !z1_target_sds)
        return fail(SCRC_ERR_ARG, "scrc_ctx_create_sharded: null pointer");
    *out = nullptr;
    return guarded([&] {
        std::unique_ptr<scrc_ctx> c(new scrc_ctx());
        c->impl.create(comm->impl.device, z1_reg, z2_reg, z1_target, z2_target, z1_target_sds, n1, n2, P, T, K, &comm->impl);
        *out = c.release();
    });
}
int scrc_ctx_set_cost_hint(scrc_ctx* ctx, const int* iterations, int n) {
    if (!ctx || (n > 0 && !iterations) || n < 0) return fail(SCRC_ERR_ARG, "scrc_ctx_set_cost_hint: invalid argument");
    return guarded([&] { ctx->impl.cost_hint.assign(iterations, iterations + n); });
}
int scrc_shard_assign(const double* cost, int n, int world, int* owner_out) {
    if (n < 0 || world < 1 || (n > 0 && (!cost || !owner_out))) return fail(SCRC_ERR_ARG, "scrc_shard_assign: invalid argument");
    return guarded([&] { shard_assign(cost, n, world, owner_out); });
}
int scrc_shard_slice(int64_t n, int64_t align, int rank, int world, int64_t* begin_out, int64_t* end_out) {
    if (n < 0 || align < 1 || world < 1 || rank < 0 || rank >= world || !begin_out || !end_out)
        return fail(SCRC_ERR_ARG, "scrc_shard_slice: invalid argument");
    shard_slice(n, align, rank, world, begin_out, end_out);
    return SCRC_OK;
}

/* debug accessor (not part of the public header): stacked selection matrix of the last cycle
 * (test split), rows_total x n2 column-major, and the stacked row count */
int scrc_debug_get_zsel(scrc_ctx* ctx, double* out, long long* rows_total) {
    if (!ctx) return fail(SCRC_ERR_ARG, "null context");
    return guarded([&] {
        Ctx& c = ctx->impl;
        SCRC_CUDA(cudaSetDevice(c.dev.id));
        if (rows_total) *rows_total = c.nnls.rows;
        if (out && c.nnls.rows > 0) {
            SCRC_CUDA(cudaMemcpy2DAsync(out, (size_t)c.nnls.rows * 8, c.ZselT.data(), c.ZselT.ld * 8, (size_t)c.nnls.rows * 8,
                                        c.n2, cudaMemcpyDeviceToHost, c.dev.stream));
            SCRC_CUDA(cudaStreamSynchronize(c.dev.stream));
        }
    });
}

}  // extern "C"
