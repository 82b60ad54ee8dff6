# R side of mode B: what R/scregclust.R calls instead of its penalty loop (lines 1222-1912) once src/bridge.cpp is
# in place.  Everything before the loop (validation, split_sample, constant-gene filter, scaling: lines 203-1009)
# and everything after it (silhouette, result packaging: lines 1914-2111) stays as it is.
#
#   fit <- scrc_grid_fit_r(z1_reg_scaled, z2_reg_scaled, z1_target_centered, z2_target_centered, z1_target_sds,
#                          penalization, n_modules, k_start_cl, ...)
#
# `update_orders` reproduces R's RNG stream when a network prior is used: the permutations sample(length(k), length(k))
# of line 1666 are drawn here, before the fit, in the order the serial loop would draw them (penalty-major, cycle-minor).
# Without a prior the allocation does not depend on the order (src/allocation.cpp:39-56) and nothing is drawn.

coef_ols   <- function(y, x)         scrc_coef_ols_(y, x)             # R/utils.R:13-37
coef_ridge <- function(y, x, lambda) scrc_coef_ridge_(y, x, lambda)   # R/utils.R:49-57
fast_cor   <- function(x, y)         scrc_fast_cor_(x, y)             # R/utils.R:533-541

# kmeanspp() of R/utils.R:351-375 with the distances, the seeding bookkeeping and the Hartigan-Wong runs on the device.
# The draws below consume R's RNG stream exactly as kmeanspp_init() does (one sample(n, 1L), then one unif_rand() per
# sample.int(prob =) call), so a seeded R session gives the same initial centres with either implementation.
kmeanspp <- function(x, n_cluster, n_init_clusterings = 10L, n_max_iter = 10L) {
  n <- nrow(x)
  first <- integer(n_init_clusterings)
  u <- matrix(0, n_cluster - 1L, n_init_clusterings)
  for (i in seq_len(n_init_clusterings)) {
    first[i] <- sample(n, size = 1L)
    u[, i] <- runif(n_cluster - 1L)
  }
  scrc_kmeanspp_(x, as.integer(n_cluster), first, u, as.integer(n_max_iter))
}

scrc_grid_fit_r <- function(z1_reg_scaled, z2_reg_scaled, z1_target_centered, z2_target_centered, z1_target_sds,
                            penalization, n_modules, k_start_cl, n_cycles = 50L, noise_threshold = 0.025,
                            min_module_size = 0L, max_optim_iter = 10000L, tol_coop_rel = 1e-8, tol_coop_abs = 1e-12,
                            tol_nnls = 1e-4, compute_predictive_r2 = TRUE, quick_mode = FALSE, quick_mode_percent = 0.1,
                            allocate_per_obs = TRUE, prior_indicator_list = list(), prior_baseline = 1e-6, prior_weight = 0,
                            devices = 0L) {
  n_target <- ncol(z1_target_centered)
  update_orders <- matrix(0L, 0L, 0L)
  if (length(prior_indicator_list) > 0L && prior_weight != 0) {
    update_orders <- vapply(
      seq_len(length(penalization) * (n_cycles + 1L)),
      function(i) sample(n_target, n_target) - 1L, integer(n_target)
    )
  }
  scrc_grid(
    z1_reg_scaled, z2_reg_scaled, z1_target_centered, z2_target_centered, z1_target_sds,
    as.double(penalization), as.integer(n_modules), as.integer(k_start_cl), as.integer(n_cycles),
    noise_threshold, as.integer(min_module_size), as.integer(max_optim_iter), tol_coop_rel, tol_coop_abs, tol_nnls,
    compute_predictive_r2, quick_mode, quick_mode_percent, allocate_per_obs, prior_indicator_list, prior_baseline,
    prior_weight,
    update_orders, as.integer(devices)
  )
}
