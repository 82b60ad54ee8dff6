// Rcpp bridge between the scregclust R package and libscregclust_b200.so.
//
// Drop this file into src/ of scmethods/scregclust v0.2.2 IN PLACE OF src/optim.cpp, src/allocation.cpp and
// src/utils.cpp (src/jaccard.cpp stays), use bridge/Makevars as src/Makevars, and run Rcpp::compileAttributes():
// the regenerated src/RcppExports.cpp keeps the .Call symbols of the reference (src/RcppExports.cpp:104-112) --
// _scregclust_coop_lasso (13 arguments), _scregclust_coef_nnls (4), _scregclust_allocate_into_modules (7),
// _scregclust_alloc_array (2), _scregclust_reset_array (2) -- so R/RcppExports.R:4-117 and R/scregclust.R work unchanged
// (mode A).  The exports below the line "mode B" are additions: the three pure-R helpers of R/utils.R:13-57,533-541 as
// device calls and the whole-grid fit, which bridge/scregclust_b200.R calls instead of the loop of
// R/scregclust.R:1222-1912.
//
// The library never calls the R API and never throws; every entry point returns a status whose message
// (scrc_last_error) reproduces the Rcpp::stop() texts of the reference.  Interrupts: the library is polled from the R
// thread (R_ToplevelExec(R_CheckUserInterrupt)) while a worker thread runs the fit; see run_interruptible().
//
// Not compiled in the build container of this repository (no R / Rcpp there); tests/test_bridge_syntax.py type-checks
// it against include/scregclust_b200.h with the minimal Rcpp mock in bridge/mock/.
#include <Rcpp/Lightest>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "scregclust_b200.h"

static void check(int rc) {             // same R errors as Rcpp::stop() in the reference
  if (rc == SCRC_ERR_INTERRUPTED) Rcpp::checkUserInterrupt();   // raises R's interrupt condition if one is pending
  if (rc != SCRC_OK) Rcpp::stop(scrc_last_error());   // texts match optim.cpp:77,81,88,116,121,409; utils.cpp:20,52
}

// ------------------------------------------------------------------------------------------------ mode A: drop-ins

// [[Rcpp::export]]
Rcpp::List coop_lasso(Rcpp::NumericMatrix y, Rcpp::NumericMatrix x, double lambda, Rcpp::NumericVector weights,
                      Rcpp::Nullable<Rcpp::NumericMatrix> beta_0 = R_NilValue, double rho_0 = 0.2, double alpha_0 = 1.5,
                      int n_update = 2, double eps_corr = 0.2, int max_iter = 1000, double eps_rel = 1e-8,
                      double eps_abs = 1e-12, bool verbose = false) {
  const double* b0 = nullptr; int64_t b0r = 0, b0c = 0;
  Rcpp::NumericMatrix b0m;
  if (beta_0.isUsable()) { b0m = beta_0.as(); b0 = b0m.begin(); b0r = b0m.nrow(); b0c = b0m.ncol(); }
  Rcpp::NumericMatrix beta(x.ncol(), y.ncol());
  int it = 0;
  check(scrc_coop_lasso(y.begin(), y.nrow(), y.ncol(), x.begin(), x.nrow(), x.ncol(), lambda, weights.begin(),
                        b0, b0r, b0c, rho_0, alpha_0, n_update, eps_corr, max_iter, eps_rel, eps_abs,
                        beta.begin(), &it));
  if (it > max_iter) Rcpp::Rcout << "Coop-Lasso: Maximum number of iterations reached" << std::endl;  // optim.cpp:300
  (void)verbose;                        // per-iteration residual printing (optim.cpp:178-189) is not reproduced
  Rcpp::checkUserInterrupt();           // the library never calls the R API; poll here, on the R thread
  return Rcpp::List::create(Rcpp::Named("beta") = beta, Rcpp::Named("iterations") = it);
}

// [[Rcpp::export]]
Rcpp::List coef_nnls(Rcpp::NumericMatrix x, Rcpp::NumericMatrix y, double eps = 1e-12, int max_iter = 1000) {
  Rcpp::NumericMatrix beta(x.ncol(), y.ncol());
  Rcpp::IntegerVector it(y.ncol());
  check(scrc_coef_nnls(x.begin(), x.nrow(), x.ncol(), y.begin(), y.nrow(), y.ncol(), eps, max_iter,
                       beta.begin(), it.begin()));
  bool hit = false;
  for (int i = 0; i < it.size(); ++i) hit |= (it[i] == max_iter);
  if (hit) Rcpp::Rcout << "NNLS: Maximum number of iterations reached" << std::endl;                 // optim.cpp:533
  return Rcpp::List::create(Rcpp::Named("beta") = beta, Rcpp::Named("iterations") = it);
}

// list of integer vectors (prior_indicator, R/scregclust.R:584-608) -> CSR
static void prior_csr(Rcpp::List prior_indicator, int64_t T, std::vector<int>& off, std::vector<int>& idx) {
  off.assign(T + 1, 0);
  idx.clear();
  if (prior_indicator.size() != T) return;
  for (int64_t t = 0; t < T; ++t) {
    Rcpp::IntegerVector v = prior_indicator[t];
    idx.insert(idx.end(), v.begin(), v.end());
    off[t + 1] = (int)idx.size();
  }
}

// [[Rcpp::export]]
Rcpp::IntegerVector allocate_into_modules(SEXP resid_array, Rcpp::NumericMatrix resid_var, Rcpp::List prior_indicator,
                                          Rcpp::IntegerVector k_, Rcpp::IntegerVector update_order,
                                          double prior_baseline, double prior_weight) {
  int* dims = INTEGER(Rf_getAttrib(resid_array, R_DimSymbol));          // K x n_obs x n_genes (allocation.cpp:15-18)
  const int K = dims[0]; const int64_t n2 = dims[1], T = dims[2];
  std::vector<int> off, idx;
  prior_csr(prior_indicator, T, off, idx);
  const bool has_prior = prior_indicator.size() == T;
  Rcpp::IntegerVector out(T);
  check(scrc_allocate_into_modules(REAL(resid_array), K, n2, T, resid_var.begin(), has_prior ? off.data() : nullptr,
                                   idx.empty() ? nullptr : idx.data(), k_.begin(), update_order.begin(),
                                   prior_baseline, prior_weight, out.begin(), nullptr));
  Rcpp::checkUserInterrupt();
  return out;
}

// [[Rcpp::export]]
SEXP alloc_array(SEXP input, R_xlen_t n_cl) {
  const R_xlen_t n_obs = Rf_nrows(input), n_genes = Rf_ncols(input);
  if ((long double)n_cl * n_obs * n_genes > (long double)R_XLEN_T_MAX) Rcpp::stop("alloc_array: requested allocation too large");
  SEXP arr = PROTECT(Rf_allocVector(REALSXP, n_cl * n_obs * n_genes));
  check(scrc_alloc_array(REAL(input), n_obs, n_genes, n_cl, REAL(arr)));
  UNPROTECT(1);
  return arr;
}

// [[Rcpp::export]]
void reset_array(SEXP arr, SEXP input) {                                 // mutates arr in place (utils.cpp:43-63)
  int* d = INTEGER(Rf_getAttrib(arr, R_DimSymbol));
  check(scrc_reset_array(REAL(arr), d[0], d[1], d[2], REAL(input), Rf_nrows(input), Rf_ncols(input)));
}

// ------------------------------------------------------------------------------------------------ mode B: additions

// [[Rcpp::export]]
Rcpp::NumericMatrix scrc_coef_ols_(Rcpp::NumericMatrix y, Rcpp::NumericMatrix x) {                   // R/utils.R:13-37
  Rcpp::NumericMatrix beta(x.ncol(), y.ncol());
  check(scrc_coef_ols(y.begin(), x.begin(), x.nrow(), x.ncol(), y.ncol(), beta.begin()));
  return beta;
}
// [[Rcpp::export]]
Rcpp::NumericMatrix scrc_coef_ridge_(Rcpp::NumericMatrix y, Rcpp::NumericMatrix x, double lambda) {  // R/utils.R:49-57
  Rcpp::NumericMatrix beta(x.ncol(), y.ncol());
  check(scrc_coef_ridge(y.begin(), x.begin(), x.nrow(), x.ncol(), y.ncol(), lambda, beta.begin()));
  return beta;
}
// [[Rcpp::export]]
Rcpp::NumericMatrix scrc_fast_cor_(Rcpp::NumericMatrix x, Rcpp::NumericMatrix y) {                   // R/utils.R:533-541
  Rcpp::NumericMatrix out(x.ncol(), y.ncol());
  check(scrc_fast_cor(x.begin(), y.begin(), x.nrow(), x.ncol(), y.ncol(), out.begin()));
  return out;
}

// kmeanspp(x, n_cluster, n_init_clusterings, n_max_iter) of R/utils.R:351-375 on the device.  The R caller draws
// first_centers[i] <- sample(n, 1L) and uniforms[, i] <- runif(n_cluster - 1L) per initialisation (bridge/scregclust_b200.R),
// which consumes R's RNG stream exactly as kmeanspp_init() does.  uniforms: (n_cluster - 1) x n_init.
// [[Rcpp::export]]
Rcpp::IntegerVector scrc_kmeanspp_(Rcpp::NumericMatrix x, int n_cluster, Rcpp::IntegerVector first_centers,
                                   Rcpp::NumericMatrix uniforms, int n_max_iter) {
  Rcpp::IntegerVector cluster(x.nrow());
  std::vector<int> ifault(first_centers.size());
  check(scrc_kmeanspp(x.begin(), x.nrow(), x.ncol(), n_cluster, (int)first_centers.size(), n_max_iter,
                      first_centers.begin(), uniforms.begin(), cluster.begin(), nullptr, ifault.data()));
  for (size_t i = 0; i < ifault.size(); ++i) {      // the warnings of stats::kmeans
    if (ifault[i] == 2) Rcpp::Rcout << "Warning: did not converge in " << n_max_iter << " iterations" << std::endl;
    if (ifault[i] == 4) Rcpp::Rcout << "Warning: Quick-TRANSfer stage steps exceeded maximum" << std::endl;
  }
  return cluster;
}

static void poll_interrupt(void*) { R_CheckUserInterrupt(); }

// Runs `work` on a worker thread while the R thread polls for a user interrupt every 50 ms; a pending interrupt sets
// `cancel`, which the library's loop reads before every device call (scrc_grid_params::cancel).
template <class Work>
static int run_interruptible(Work&& work, volatile int* cancel) {
  std::atomic<bool> done(false);
  int rc = SCRC_OK;
  std::string err;
  std::thread worker([&] { rc = work(); if (rc != SCRC_OK) err = scrc_last_error(); done.store(true); });
  while (!done.load()) {
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    if (!*cancel && R_ToplevelExec(poll_interrupt, nullptr) == FALSE) *cancel = 1;
  }
  worker.join();
  if (*cancel) Rcpp::stop("scregclust: interrupted");
  if (rc != SCRC_OK) Rcpp::stop(err);     // (the last-error text lives on the worker thread, hence the copy)
  return rc;
}

// The whole alternating loop of scregclust() for a penalization grid on `devices` (0-based CUDA ordinals; several
// GPUs share every cycle, include/scregclust_b200.h "multi-GPU").  Inputs are the matrices of R/scregclust.R:998-1009.
// Returns list(results = <one list per penalization value>, info = <work counters>).
// [[Rcpp::export]]
Rcpp::List scrc_grid(Rcpp::NumericMatrix z1_reg, Rcpp::NumericMatrix z2_reg, Rcpp::NumericMatrix z1_target,
                     Rcpp::NumericMatrix z2_target, Rcpp::NumericVector z1_target_sds, Rcpp::NumericVector penalization,
                     int n_modules, Rcpp::IntegerVector initial_target_modules, int n_cycles, double noise_threshold,
                     int min_module_size, int max_optim_iter, double tol_coop_rel, double tol_coop_abs, double tol_nnls,
                     bool compute_predictive_r2, bool quick_mode, double quick_mode_percent, bool allocate_per_obs,
                     Rcpp::List prior_indicator,
                     double prior_baseline, double prior_weight, Rcpp::IntegerMatrix update_orders /* T x (L * cycles) or 0 x 0 */,
                     Rcpp::IntegerVector devices) {
  const int64_t n1 = z1_reg.nrow(), n2 = z2_reg.nrow(), P = z1_reg.ncol(), T = z1_target.ncol();
  const int K = n_modules, L = penalization.size();
  scrc_grid_params gp;
  scrc_grid_default_params(&gp);
  gp.n_penalties = L; gp.penalization = penalization.begin(); gp.initial_target_modules = initial_target_modules.begin();
  gp.n_cycles = n_cycles; gp.noise_threshold = noise_threshold; gp.min_module_size = min_module_size;
  gp.opt.max_optim_iter = max_optim_iter; gp.opt.tol_coop_rel = tol_coop_rel; gp.opt.tol_coop_abs = tol_coop_abs;
  gp.opt.tol_nnls = tol_nnls;
  gp.compute_predictive_r2 = compute_predictive_r2; gp.keep_coeffs = 1;
  gp.quick_mode = quick_mode; gp.quick_mode_percent = quick_mode_percent;
  gp.allocate_per_obs = allocate_per_obs;            // FALSE: the aggregate branch of R/scregclust.R:1678-1728
  std::vector<int> off, idx;
  if (prior_weight != 0.0 && prior_indicator.size() == T) {
    prior_csr(prior_indicator, T, off, idx);
    if (idx.empty()) idx.push_back(0);
    gp.prior_off = off.data(); gp.prior_idx = idx.data(); gp.prior_baseline = prior_baseline; gp.prior_weight = prior_weight;
  }
  // update_order: the permutations R's sample() would draw (R/scregclust.R:1666), pre-drawn by the R driver in the
  // order of the serial loop, column (penalty * (n_cycles + 1) + cycle - 1)
  struct Orders { const int* base; int64_t T; int per; } ord = {update_orders.begin(), T, n_cycles + 1};
  if (update_orders.ncol() > 0) {
    gp.update_order_user = &ord;
    gp.update_order = [](void* u, int penalty, int cycle) -> const int* {
      const Orders* o = static_cast<const Orders*>(u);
      return o->base + ((int64_t)penalty * o->per + (cycle - 1)) * o->T;
    };
  }
  volatile int cancel = 0;
  gp.cancel = &cancel;
  std::vector<int> dev(devices.begin(), devices.end());
  if (dev.empty()) dev.push_back(0);
  scrc_grid_result* res = nullptr;
  run_interruptible([&] {
    return scrc_grid_fit_multi(dev.data(), (int)dev.size(), z1_reg.begin(), z2_reg.begin(), z1_target.begin(),
                               z2_target.begin(), z1_target_sds.begin(), n1, n2, P, T, K, &gp, &res);
  }, &cancel);

  Rcpp::List results(L);
  for (int i = 0; i < L; ++i) {
    scrc_grid_penalty_info pi;
    check(scrc_grid_result_penalty(res, i, &pi));
    Rcpp::IntegerMatrix k_history(T, pi.n_history);                      // column h = configuration h (k_history[[h]])
    check(scrc_grid_result_history(res, i, k_history.begin()));
    Rcpp::List finals(pi.final_cycle_length);
    for (int m = 0; m < pi.final_cycle_length; ++m) {
      Rcpp::IntegerVector k_final(T), best_idx(T), nsel(K);
      Rcpp::LogicalMatrix models(P, K);
      std::vector<unsigned char> md((size_t)P * K);
      Rcpp::NumericMatrix weights(P, K), signs(P, K), r2(T, K), resid_var(T, K), importance(P, K);
      Rcpp::NumericVector best_r2(T), r2_module(K);
      scrc_grid_final gf = {k_final.begin(), md.data(), weights.begin(), signs.begin(), r2.begin(), resid_var.begin(),
                            best_r2.begin(), best_idx.begin(), nsel.begin(),
                            compute_predictive_r2 ? r2_module.begin() : nullptr,
                            compute_predictive_r2 ? importance.begin() : nullptr};
      check(scrc_grid_result_final(res, i, m, &gf));
      for (size_t q = 0; q < md.size(); ++q) models[q] = md[q];
      int tot = 0;
      for (int j = 0; j < K; ++j) tot += nsel[j];
      std::vector<double> packed((size_t)tot * T);
      std::vector<int> sel(tot > 0 ? tot : 1);
      check(scrc_grid_result_coeffs(res, i, m, packed.data(), sel.data()));
      Rcpp::List coeffs(K);                                              // coeffs_final[[m]][[j]]: s_j x T (R/scregclust.R:1606-1611)
      size_t o = 0;
      for (int j = 0; j < K; ++j) {
        if (nsel[j] == 0) continue;
        Rcpp::NumericMatrix cj(nsel[j], T);
        std::copy(packed.begin() + o, packed.begin() + o + (size_t)nsel[j] * T, cj.begin());
        coeffs[j] = cj;
        o += (size_t)nsel[j] * T;
      }
      finals[m] = Rcpp::List::create(Rcpp::Named("k") = k_final, Rcpp::Named("models") = models,
                                     Rcpp::Named("weights") = weights, Rcpp::Named("signs") = signs,
                                     Rcpp::Named("r2") = r2, Rcpp::Named("resid_var") = resid_var,
                                     Rcpp::Named("best_r2") = best_r2, Rcpp::Named("best_r2_idx") = best_idx,
                                     Rcpp::Named("coeffs") = coeffs, Rcpp::Named("r2_module") = r2_module,
                                     Rcpp::Named("importance") = importance);
    }
    results[i] = Rcpp::List::create(Rcpp::Named("penalization") = pi.penalization, Rcpp::Named("converged") = pi.converged != 0,
                                    Rcpp::Named("n_cycles") = pi.n_cycles_run,
                                    Rcpp::Named("final_cycle_length") = pi.final_cycle_length,
                                    Rcpp::Named("k_history") = k_history, Rcpp::Named("final") = finals);
  }
  scrc_grid_info info;
  check(scrc_grid_result_info(res, &info));
  scrc_grid_result_free(res);
  return Rcpp::List::create(Rcpp::Named("results") = results,
                            Rcpp::Named("info") = Rcpp::List::create(Rcpp::Named("admm_sweeps") = (double)info.admm_sweeps,
                                                                     Rcpp::Named("admm_problem_iterations") = (double)info.admm_problem_iterations,
                                                                     Rcpp::Named("fit_cycles") = info.fit_cycles,
                                                                     Rcpp::Named("final_refits") = info.final_refits));
}
