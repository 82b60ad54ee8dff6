// "Reference-faithful CPU" restatement of the scregclust native kernels -- TEST / BASELINE
// INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/scregclust_oracle.py for the parity status:
// parity unpinned except fast_cor).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the resulting oracle/_ref/liboracle_ref.so.
//
// The reference itself (src/optim.cpp, src/allocation.cpp, src/utils.cpp) needs Rcpp, RcppEigen
// and R headers, none of which exist in the build container, so it is treated as unbuildable.
// This file restates it operation for operation in plain C++17 and keeps its COST STRUCTURE:
//   * X'X is recomputed inside every coop_lasso / coef_nnls call      (optim.cpp:97, 414)
//   * the P x P matrix is re-factorised whenever pc == 0 || rho != rho_old, which after the
//     first rho change is every iteration                              (optim.cpp:136-140)
//   * the factorisation is Eigen 3.4.0's unblocked, diagonally pivoted LDL^T, single thread
//   * only general matrix products use all cores (Eigen's OpenMP GEMM; src/Makevars:1-2)
// BLAS calls go to the OpenBLAS bundled with SciPy (scipy_cblas_*), standing in for Eigen's
// own GEMM/TRSM kernels and R's BLAS.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

extern "C" {
// OpenBLAS (SciPy build, LP64, symbol prefix scipy_)
enum { RowMajor = 101, ColMajor = 102, NoTrans = 111, Trans = 112, Upper = 121, Lower = 122, NonUnit = 131, Unit = 132, Left = 141 };
void scipy_cblas_dgemm(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
void scipy_cblas_dsyrk(int, int, int, int, int, double, const double*, int, double, double*, int);
void scipy_cblas_dtrsm(int, int, int, int, int, int, int, double, const double*, int, double*, int);
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
}

namespace {
int g_threads = 1;
struct OneThread {   // Eigen only parallelises general GEMM; everything else runs on one core
    OneThread() { scipy_openblas_set_num_threads(1); }
    ~OneThread() { scipy_openblas_set_num_threads(g_threads); }
};

// src/optim.cpp:17-27
void compute_xtx(const double* x, int64_t n, int64_t p, std::vector<double>& xtx) {
    xtx.assign((size_t)p * p, 0.0);
    if (p == 0) return;
    {
        OneThread t;
        scipy_cblas_dsyrk(ColMajor, Lower, Trans, (int)p, (int)n, 1.0, x, (int)n, 0.0, xtx.data(), (int)p);
    }
    for (int64_t j = 0; j < p; ++j)
        for (int64_t i = j + 1; i < p; ++i) xtx[j + i * p] = xtx[i + j * p];
}

// Eigen 3.4.0 LDLT.h: ldlt_inplace<Lower>::unblocked + _solve_impl
struct LDLT {
    int64_t n = 0;
    std::vector<double> mat;
    std::vector<int64_t> trans;
    std::vector<double> temp;
    void compute(const std::vector<double>& a, int64_t n_) {
        n = n_; mat = a; trans.assign(n, 0); temp.assign(n, 0.0);
        double* m = mat.data();
        for (int64_t k = 0; k < n; ++k) {
            int64_t piv = k; double big = std::fabs(m[k + k * n]);
            for (int64_t i = k + 1; i < n; ++i) { const double v = std::fabs(m[i + i * n]); if (v > big) { big = v; piv = i; } }
            trans[k] = piv;
            if (piv != k) {
                for (int64_t c = 0; c < k; ++c) std::swap(m[k + c * n], m[piv + c * n]);
                for (int64_t r = piv + 1; r < n; ++r) std::swap(m[r + k * n], m[r + piv * n]);
                std::swap(m[k + k * n], m[piv + piv * n]);
                for (int64_t i = k + 1; i < piv; ++i) std::swap(m[i + k * n], m[piv + i * n]);
            }
            const int64_t rs = n - k - 1;
            if (k > 0) {
                for (int64_t c = 0; c < k; ++c) temp[c] = m[c + c * n] * m[k + c * n];
                double acc = 0.0;
                for (int64_t c = 0; c < k; ++c) acc += m[k + c * n] * temp[c];
                m[k + k * n] -= acc;
                if (rs > 0) {   // A21 -= A20 * temp  (column-major GEMV, axpy form like Eigen's col-major kernel)
                    double* a21 = m + (k + 1) + k * n;
                    for (int64_t c = 0; c < k; ++c) {
                        const double t = temp[c];
                        const double* a20 = m + (k + 1) + c * n;
                        for (int64_t r = 0; r < rs; ++r) a21[r] -= a20[r] * t;
                    }
                }
            }
            const double akk = m[k + k * n];
            if (rs > 0 && std::fabs(akk) > 0.0) {
                double* a21 = m + (k + 1) + k * n;
                for (int64_t r = 0; r < rs; ++r) a21[r] /= akk;
            }
        }
    }
    void solve(double* b, int64_t ncol) const {   // in place, b is n x ncol column-major
        const double* m = mat.data();
        for (int64_t k = 0; k < n; ++k) if (trans[k] != k) for (int64_t c = 0; c < ncol; ++c) std::swap(b[k + c * n], b[trans[k] + c * n]);
        {
            OneThread t;
            scipy_cblas_dtrsm(ColMajor, Left, Lower, NoTrans, Unit, (int)n, (int)ncol, 1.0, m, (int)n, b, (int)n);
        }
        const double tol = std::numeric_limits<double>::min();
        for (int64_t k = 0; k < n; ++k) {
            const double d = m[k + k * n];
            if (std::fabs(d) > tol) for (int64_t c = 0; c < ncol; ++c) b[k + c * n] /= d;
            else for (int64_t c = 0; c < ncol; ++c) b[k + c * n] = 0.0;
        }
        {
            OneThread t;
            scipy_cblas_dtrsm(ColMajor, Left, Lower, Trans, Unit, (int)n, (int)ncol, 1.0, m, (int)n, b, (int)n);
        }
        for (int64_t k = n - 1; k >= 0; --k) if (trans[k] != k) for (int64_t c = 0; c < ncol; ++c) std::swap(b[k + c * n], b[trans[k] + c * n]);
    }
};

inline double sqnorm(const std::vector<double>& v) { double s = 0.0; for (double x : v) s += x * x; return s; }
}  // namespace

extern "C" {

int ref_set_threads(int n) {
    g_threads = n > 0 ? n : 1;
    scipy_openblas_set_num_threads(g_threads);
    return scipy_openblas_get_num_threads();
}

// src/optim.cpp:63-320.  n_factor_out counts LDLT factorisations actually performed.
int ref_coop_lasso(const double* y, const double* x, int64_t n, int64_t m, int64_t p, double lambda, const double* w,
                   const double* beta0, double rho0, double alpha0, int n_update, double eps_corr, int max_iter,
                   double eps_rel, double eps_abs, double* beta_out, int* iters_out, int* n_factor_out) {
    if (n <= 0 || m <= 0 || p <= 0) return -1;
    if (rho0 <= 0.0) return -4;
    if (alpha0 <= 0.0 || alpha0 > 2.0) return -5;
    const size_t pm = (size_t)p * m;
    std::vector<double> xtx, xty(pm);
    compute_xtx(x, n, p, xtx);
    scipy_cblas_dgemm(ColMajor, Trans, NoTrans, (int)p, (int)m, (int)n, 1.0, x, (int)n, y, (int)n, 0.0, xty.data(), (int)p);
    std::vector<double> beta(pm, 0.0);
    if (beta0) std::memcpy(beta.data(), beta0, pm * 8);
    std::vector<double> zeta = beta, mult(pm, 0.0), beta_old = beta, zeta_old = zeta, mult_old(pm, 0.0), mult_hat_old(pm, 0.0);
    std::vector<double> zeta_new(pm), beta_relaxed(pm), mult_hat(pm), xtx_rho = xtx, sp(p), sn(p);
    double rho = rho0, rho_old = 0.0, alpha = alpha0;
    int it = 1, pc = 0, nf = 0;
    LDLT ldlt;
    while (true) {
        if (pc == 0 || rho != rho_old) {
            for (int64_t i = 0; i < p; ++i) xtx_rho[i + i * p] = xtx[i + i * p] + rho;
            ldlt.compute(xtx_rho, p);
            ++nf;
        }
        for (size_t i = 0; i < pm; ++i) beta[i] = xty[i] + rho * zeta[i] + mult[i];
        ldlt.solve(beta.data(), m);
        for (size_t i = 0; i < pm; ++i) { beta_relaxed[i] = alpha * beta[i] + (1.0 - alpha) * zeta[i]; zeta_new[i] = beta_relaxed[i] - mult[i] / rho; }
        for (int64_t i = 0; i < p; ++i) {
            double a = 0.0, b = 0.0;
            for (int64_t j = 0; j < m; ++j) { const double v = zeta_new[i + j * p]; const double pv = std::max(v, 0.0), nv = std::min(v, 0.0); a += pv * pv; b += nv * nv; }
            sp[i] = std::max(1.0 - lambda * w[i] / (rho * std::sqrt(a)), 0.0);
            sn[i] = std::max(1.0 - lambda * w[i] / (rho * std::sqrt(b)), 0.0);
        }
        for (int64_t j = 0; j < m; ++j) for (int64_t i = 0; i < p; ++i) { double& v = zeta_new[i + j * p]; v *= (v >= 0.0) ? sp[i] : sn[i]; }
        for (size_t i = 0; i < pm; ++i) mult[i] += rho * (-beta_relaxed[i] + zeta_new[i]);
        double pr = 0.0, du = 0.0, nb = 0.0, nz = 0.0, nm = 0.0;
        for (size_t i = 0; i < pm; ++i) {
            const double a = -beta[i] + zeta_new[i], b = rho * (zeta[i] - zeta_new[i]);
            pr += a * a; du += b * b; nb += beta[i] * beta[i]; nz += zeta_new[i] * zeta_new[i]; nm += mult[i] * mult[i];
        }
        if (std::sqrt(pr) <= std::fmax(eps_rel * std::fmax(std::sqrt(nb), std::sqrt(nz)), eps_abs) &&
            std::sqrt(du) <= std::fmax(eps_rel * std::sqrt(nm), eps_abs)) break;
        ++pc;
        if (pc == n_update) {
            double s5 = 0, s6 = 0, s7 = 0, s8 = 0, s9 = 0, s10 = 0;
            for (size_t i = 0; i < pm; ++i) {
                mult_hat[i] = mult[i] + rho * (-beta[i] + zeta[i]);
                const double dmh = mult_hat[i] - mult_hat_old[i], dh = beta[i] - beta_old[i], dm = mult[i] - mult_old[i], dg = zeta_old[i] - zeta[i];
                s5 += dmh * dmh; s6 += dh * dh; s7 += dh * dmh; s8 += dm * dm; s9 += dg * dg; s10 += dg * dm;
            }
            const double n_dmh = std::sqrt(s5), n_dh = std::sqrt(s6), n_dm = std::sqrt(s8), n_dg = std::sqrt(s9);
            double a = 0, a_corr = 0, b = 0, b_corr = 0;
            if (n_dmh > 0.0 && n_dh > 0.0) { const double a_sd = s5 / s7, a_mg = s7 / s6; a = (2.0 * a_mg > a_sd) ? a_mg : a_sd - a_mg / 2.0; a_corr = s7 / (n_dh * n_dmh); }
            if (n_dm > 0.0 && n_dg > 0.0) { const double b_sd = s8 / s10, b_mg = s10 / s9; b = (2.0 * b_mg > b_sd) ? b_mg : b_sd - b_mg / 2.0; b_corr = s10 / (n_dg * n_dm); }
            rho_old = rho;
            const bool ga = a_corr > eps_corr, gb = b_corr > eps_corr;
            if (ga && gb) { rho = std::sqrt(a * b); alpha = 1.0 + 2.0 / (std::sqrt(a * b) * (1.0 / a + 1.0 / b)); }
            else if (ga) { rho = a; alpha = 1.9; }
            else if (gb) { rho = b; alpha = 1.1; }
            else alpha = 1.5;
            beta_old = beta; zeta_old = zeta_new; mult_old = mult; mult_hat_old = mult_hat;
            pc = 0;
        }
        zeta = zeta_new;
        ++it;
        if (it > max_iter) break;
    }
    std::memcpy(beta_out, beta.data(), pm * 8);
    *iters_out = it;
    if (n_factor_out) *n_factor_out = nf;
    return 0;
}

// src/optim.cpp:335-377
static void greedy_cd(const double* Q, int64_t n, double* beta, double* grad, int64_t m) {
    for (int64_t t = 0; t < n; ++t) {
        int64_t empty = 0;
        for (int64_t j = 0; j < m; ++j) {
            double* b = beta + j * n; double* g = grad + j * n;
            int64_t p = 0; double mx = -1.0;
            for (int64_t i = 0; i < n; ++i) { const double sc = ((b[i] > 0.0) || (g[i] < 0.0)) ? std::fabs(g[i]) : 0.0; if (sc > mx) { mx = sc; p = i; } }
            if (mx == 0.0) { ++empty; continue; }
            const double db = std::fmax(0.0, b[p] - g[p] / Q[p + p * n]) - b[p];
            b[p] += db;
            const double* q = Q + p * n;
            for (int64_t i = 0; i < n; ++i) g[i] += db * q[i];
        }
        if (empty == m) break;
    }
}

// src/optim.cpp:403-549
int ref_coef_nnls(const double* x, int64_t nobs, int64_t n, const double* y, int64_t m_total, double eps, int max_iter,
                  double* beta_out, int* iters_out) {
    if (n <= 0 || m_total <= 0 || nobs <= 0) return -6;
    std::vector<double> xtx, d(n), Q((size_t)n * n), grad((size_t)n * m_total);
    compute_xtx(x, nobs, n, xtx);
    for (int64_t i = 0; i < n; ++i) d[i] = 1.0 / std::sqrt(xtx[i + i * n]);
    for (int64_t j = 0; j < n; ++j) for (int64_t i = 0; i < n; ++i) Q[i + j * n] = xtx[i + j * n] * (d[i] * d[j]);
    scipy_cblas_dgemm(ColMajor, Trans, NoTrans, (int)n, (int)m_total, (int)nobs, -1.0, x, (int)nobs, y, (int)nobs, 0.0, grad.data(), (int)n);
    for (int64_t j = 0; j < m_total; ++j) for (int64_t i = 0; i < n; ++i) grad[i + j * n] *= d[i];
    std::vector<double> beta_final((size_t)n * m_total, 0.0), beta((size_t)n * m_total, 0.0), grad_bar = grad;
    for (size_t i = 0; i < grad.size(); ++i) if (grad[i] > 0.0) grad_bar[i] = 0.0;
    std::vector<int64_t> remaining(m_total);
    std::iota(remaining.begin(), remaining.end(), 0);
    for (int64_t j = 0; j < m_total; ++j) iters_out[j] = max_iter;
    int64_t m = m_total;
    std::vector<double> beta_save, tmp, dbeta;
    auto neg_corr = [&](int64_t j) {
        double* b = beta.data() + j * n; double* g = grad.data() + j * n;
        for (int64_t i = 0; i < n; ++i) if (b[i] < 0.0) { const double bi = b[i]; const double* q = Q.data() + i * n; for (int64_t r = 0; r < n; ++r) g[r] -= bi * q[r]; b[i] = 0.0; }
    };
    auto ok = [](double a) { return (a == a) && std::fabs(a) >= 1e-20 && std::fabs(a) < 1e30; };
    for (int l = 0; l < max_iter; ++l) {
        beta_save.assign(beta.begin(), beta.begin() + n * m);
        tmp.assign((size_t)n * m, 0.0);
        scipy_cblas_dgemm(ColMajor, NoTrans, NoTrans, (int)n, (int)m, (int)n, 1.0, Q.data(), (int)n, grad_bar.data(), (int)n, 0.0, tmp.data(), (int)n);
        for (int64_t j = 0; j < m; ++j) {
            double num = 0, den = 0;
            for (int64_t i = 0; i < n; ++i) { const double gb = grad_bar[i + j * n]; num += gb * gb; den += gb * tmp[i + j * n]; }
            const double a = num / den;
            if (ok(a)) { for (int64_t i = 0; i < n; ++i) { beta[i + j * n] -= a * grad_bar[i + j * n]; grad[i + j * n] -= a * tmp[i + j * n]; } neg_corr(j); }
        }
        greedy_cd(Q.data(), n, beta.data(), grad.data(), m);
        dbeta.resize((size_t)n * m);
        for (size_t i = 0; i < (size_t)n * m; ++i) dbeta[i] = beta_save[i] - beta[i];
        scipy_cblas_dgemm(ColMajor, NoTrans, NoTrans, (int)n, (int)m, (int)n, 1.0, Q.data(), (int)n, dbeta.data(), (int)n, 0.0, tmp.data(), (int)n);
        for (int64_t j = 0; j < m; ++j) {
            double num = 0, den = 0;
            for (int64_t i = 0; i < n; ++i) { const double gd = grad[i + j * n] * dbeta[i + j * n]; num += gd * gd; den += dbeta[i + j * n] * tmp[i + j * n]; }
            const double a = num / den;
            if (ok(a)) { for (int64_t i = 0; i < n; ++i) { beta[i + j * n] -= a * dbeta[i + j * n]; grad[i + j * n] -= a * tmp[i + j * n]; } neg_corr(j); }
        }
        greedy_cd(Q.data(), n, beta.data(), grad.data(), m);
        int64_t kept = 0;
        std::vector<int64_t> rem2;
        for (int64_t j = 0; j < m; ++j) {
            double nrm = 0.0;
            for (int64_t i = 0; i < n; ++i) { const double g = grad[i + j * n]; const double v = (beta[i + j * n] == 0.0 && g > 0.0) ? 0.0 : g; grad_bar[i + j * n] = v; nrm += v * v; }
            if (std::sqrt(nrm) < eps) {
                std::memcpy(beta_final.data() + remaining[j] * n, beta.data() + j * n, n * 8);
                iters_out[remaining[j]] = l + 1;
            } else {
                if (kept != j) {
                    std::memcpy(beta.data() + kept * n, beta.data() + j * n, n * 8);
                    std::memcpy(grad.data() + kept * n, grad.data() + j * n, n * 8);
                    std::memcpy(grad_bar.data() + kept * n, grad_bar.data() + j * n, n * 8);
                }
                rem2.push_back(remaining[j]);
                ++kept;
            }
        }
        remaining.swap(rem2);
        m = kept;
        if (m == 0) break;
    }
    for (int64_t j = 0; j < m_total; ++j) for (int64_t i = 0; i < n; ++i) beta_out[i + j * n] = beta_final[i + j * n] * d[i];
    return 0;
}

// src/utils.cpp:12-34 / 43-63
void ref_fill_array(double* arr, const double* input, int64_t n_cl, int64_t n_obs, int64_t n_genes) {
    for (int64_t i = 0, ub = n_obs * n_genes; i < ub; ++i) for (int64_t j = 0; j < n_cl; ++j) arr[i * n_cl + j] = input[i];
}

// src/allocation.cpp:8-100
int ref_allocate_into_modules(const double* arr, int K, int64_t n_obs, int64_t T, const double* resid_var, const int* prior_off,
                              const int* prior_idx, const int* k_in, const int* order, double baseline, double weight, int* k_out) {
    std::vector<int> k(T);
    for (int64_t t = 0; t < T; ++t) k[t] = k_in[t] - 1;
    std::vector<double> tot(K, 0.0), frac(K), plp(K), sc((size_t)K * n_obs);
    for (int64_t t = 0; t < T; ++t) if (k[t] != -2) tot[k[t]] += 1;
    std::vector<int> votes(K);
    for (int64_t oi = 0; oi < T; ++oi) {
        const int64_t j = order[oi];
        const double* resid = arr + j * (int64_t)K * n_obs;
        std::fill(frac.begin(), frac.end(), 0.0);
        if (prior_off) for (int e = prior_off[j]; e < prior_off[j + 1]; ++e) if (k[prior_idx[e]] != -2) frac[k[prior_idx[e]]] += 1;
        double sum = 0.0;
        for (int q = 0; q < K; ++q) { if (tot[q] > 0) frac[q] /= tot[q]; frac[q] += baseline; sum += frac[q]; }
        for (int q = 0; q < K; ++q) plp[q] = std::log(frac[q]) - std::log(sum);
        for (int64_t c = 0; c < n_obs; ++c) for (int q = 0; q < K; ++q) {
            const double v = resid_var[j + (int64_t)q * T];
            const double ll = (-resid[q + c * K]) / (2.0 * v) - 0.5 * std::log(2.0 * M_PI * v);
            sc[q + c * K] = (1.0 - weight) * ll + weight * plp[q];
        }
        std::fill(votes.begin(), votes.end(), 0);
        for (int64_t c = 0; c < n_obs; ++c) { int bi = 0; double b = sc[c * K]; for (int q = 1; q < K; ++q) if (sc[q + c * K] > b) { b = sc[q + c * K]; bi = q; } votes[bi] += 1; }
        int best = 0;
        for (int q = 1; q < K; ++q) if (votes[q] > votes[best]) best = q;
        if (k[j] != -2) tot[k[j]] -= 1;
        tot[best] += 1;
        k[j] = best;
    }
    for (int64_t t = 0; t < T; ++t) k_out[t] = k[t] + 1;
    return 0;
}

// R-level products of the driver (R/scregclust.R:1590-1602): C = A B, all cores (R BLAS)
void ref_dgemm_nn(const double* a, const double* b, double* c, int64_t m, int64_t n, int64_t k) {
    scipy_cblas_dgemm(ColMajor, NoTrans, NoTrans, (int)m, (int)n, (int)k, 1.0, a, (int)m, b, (int)k, 0.0, c, (int)m);
}

}  // extern "C"
