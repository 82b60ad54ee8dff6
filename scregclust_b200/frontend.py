"""Host-side mirror of the reference's user-facing entry point ``scregclust()`` (R/scregclust.R:172-2120).

Everything that is arithmetic runs in the device session (`driver.Session`): sample split, centring,
constant-gene filter and scaling (``Session.from_expression``), the precompute, the alternating loop
and the importance post-processing (``driver.scregclust_fit``).  This module only does what the R
function does around it: argument validation with the reference's messages, bookkeeping of the genes
that survive the filter, the initial module allocation, and the packaging of the result lists.

Two things cannot be bit-compatible with an R session and are documented as such:
* when ``split_indices`` is not given, the stratified random split is drawn with NumPy's generator, not R's
  (R/scregclust.R:2283-2301 consumes R's RNG stream);
* when ``initial_target_modules`` is not given, k-means++ on the target/regulator correlations uses
  scikit-learn's Lloyd iterations on the host, not R's Hartigan-Wong ``stats::kmeans`` (R/utils.R:275-375).
Pass both explicitly (as the parity tests and the benchmark do) to take them out of the picture.
"""
from __future__ import annotations

import numpy as np

from . import driver


class ScregclustError(ValueError):
    """Raised where the reference calls ``cli::cli_abort`` during input validation."""


def _abort(msg):
    raise ScregclustError(msg)


def _validate(expression, genesymbols, is_regulator, penalization, n_modules, initial_target_modules,
              sample_assignment, split1_proportion, total_proportion, split_indices, noise_threshold, n_cycles,
              min_module_size, prior_weight):
    """The checks of R/scregclust.R:204-880 that concern arguments supported here, same wording."""
    ex = np.asarray(expression)
    if ex.ndim != 2 or not np.issubdtype(ex.dtype, np.number):
        _abort("expression needs to be a numeric matrix.")
    p, n = ex.shape
    if len(genesymbols) != p:
        _abort(f"genesymbols needs to be of length {p}.")
    isr = np.asarray(is_regulator)
    if isr.size != p or not np.all(np.isin(np.unique(isr.astype(np.int64)), (0, 1))):
        _abort("is_regulator needs to be 0/1 or TRUE/FALSE vector of length " + str(p) + ".")
    pen = np.atleast_1d(np.asarray(penalization, dtype=np.float64))
    if pen.size == 0 or not np.all(np.isfinite(pen)) or np.any(pen <= 0):
        _abort("penalization needs to be a numeric vector with positive values.")
    if not (isinstance(n_modules, (int, np.integer)) and n_modules >= 1):
        _abort("n_modules needs to be a positive integer.")
    n_target = int(np.sum(isr.astype(np.int64) == 0))
    if initial_target_modules is not None:
        itm = np.asarray(initial_target_modules)
        if itm.size != n_target or not np.all(np.isin(itm, np.arange(1, n_modules + 1))):
            _abort("initial_target_modules needs to be a vector of integers in 1, ..., n_modules with one "
                   "entry per target gene.")
    if sample_assignment is not None and len(sample_assignment) != n:
        _abort(f"sample_assignment needs to be of length {n}.")
    for name, v in (("split1_proportion", split1_proportion), ("total_proportion", total_proportion)):
        if not (0 < float(v) <= 1):
            _abort(f"{name} needs to be a number in (0, 1].")
    if split_indices is not None:
        si = np.asarray(split_indices, dtype=np.float64)
        ok = np.isnan(si) | np.isin(si, (1.0, 2.0))
        if si.size != n or not np.all(ok):
            _abort(f"split_indices needs to be a vector of length {n} containing only 1, 2 or NA.")
    if not (0 <= float(noise_threshold) <= 1):
        _abort("noise_threshold needs to be a number in [0, 1].")
    if not (isinstance(n_cycles, (int, np.integer)) and n_cycles >= 1):
        _abort("n_cycles needs to be a positive integer.")
    if not (isinstance(min_module_size, (int, np.integer)) and min_module_size >= 0):
        _abort("min_module_size needs to be a non-negative integer.")
    if not (0 <= float(prior_weight) <= 1):
        _abort("prior_weight needs to be a number between 0 and 1.")                   # R/scregclust.R:567-579
    return ex, isr.astype(np.int32), pen


def draw_split_indices(n, sample_assignment, split1_proportion, total_proportion, rng):
    """Stratified random split of R/scregclust.R:2283-2301 (NumPy generator, not R's stream)."""
    strat = np.zeros(n, dtype=np.int64) if sample_assignment is None else np.unique(np.asarray(sample_assignment),
                                                                                   return_inverse=True)[1]
    included = np.sort(np.concatenate([
        rng.choice(np.flatnonzero(strat == s), int(np.floor(np.sum(strat == s) * total_proportion)), replace=False)
        for s in np.unique(strat)]))
    sub = strat[included]
    split1 = np.sort(np.concatenate([
        rng.choice(np.flatnonzero(sub == s), int(np.floor(np.sum(sub == s) * split1_proportion)), replace=False)
        for s in np.unique(sub)]))
    split_indices = np.full(n, np.nan)
    split_indices[included] = 2
    split_indices[included[split1]] = 1
    return split_indices


def initial_modules_kmeans(cross_corr1, n_modules, n_initializations, rng):
    """Stand-in for ``kmeanspp(cross_corr1, n_modules, n_initializations, 100)`` (R/scregclust.R:1127-1147)."""
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=n_modules, init="k-means++", n_init=int(n_initializations), max_iter=100,
                random_state=int(rng.integers(0, 2 ** 31 - 1)))
    return km.fit_predict(np.ascontiguousarray(cross_corr1)).astype(np.int32) + 1


def prior_to_target_lists(prior_indicator, prior_genesymbols, genesymbols_target):
    """``prior_indicator`` (q x q, dense or scipy.sparse) + ``prior_genesymbols`` -> per target gene the sorted
    0-based indices of the target genes it is linked to (R/scregclust.R:584-608 and 1026-1047): the diagonal is
    dropped, row l lists the columns with a non-zero entry, and both sides are matched to the target symbols
    (first match, like ``match``)."""
    pg = [str(g) for g in prior_genesymbols]
    gt = [str(g) for g in genesymbols_target]
    q = len(pg)
    pi = prior_indicator.toarray() if hasattr(prior_indicator, "toarray") else np.asarray(prior_indicator)
    if pi.shape != (q, q):
        _abort("prior_genesymbols needs to have one element for each row/column of prior_indicator")
    pi = pi.copy()
    np.fill_diagonal(pi, 0)
    first = {}
    for i, g in enumerate(pg):
        first.setdefault(g, i)
    out = [[] for _ in gt]
    done = set()
    for t, g in enumerate(gt):
        if g in first and g not in done:                  # match(genes_overlap, genesymbols_target): first hit only
            done.add(g)
            linked = {pg[c] for c in np.flatnonzero(pi[first[g]] != 0)}
            out[t] = [u for u, h in enumerate(gt) if h in linked]
    return out


def silhouette_from_r2(r2, k):
    """R/scregclust.R:1973-1985: (a - b) / max(a, b) with a = own-module R2 and b = best other module (>= 0)."""
    r2 = np.asarray(r2)
    out = np.full(len(k), np.nan)
    for i, c in enumerate(k):
        if c != -1:
            others = np.delete(r2[i], c - 1)
            b = max(float(np.max(others)), 0.0) if others.size else 0.0
            a = float(r2[i, c - 1])
            with np.errstate(invalid="ignore", divide="ignore"):
                out[i] = np.float64(a - b) / np.float64(max(a, b))
    return out


def scregclust(expression, genesymbols, is_regulator, penalization, n_modules, initial_target_modules=None,
               sample_assignment=None, center=True, split1_proportion=0.5, total_proportion=1, split_indices=None,
               prior_indicator=None, prior_genesymbols=None, prior_baseline=1e-6, prior_weight=0.5, min_module_size=0,
               noise_threshold=0.025, n_cycles=50, n_initializations=50, max_optim_iter=10000, tol_coop_rel=1e-8,
               tol_coop_abs=1e-12, tol_nnls=1e-4, compute_predictive_r2=True, compute_silhouette=False,
               allocate_per_obs=True, use_kmeanspp_init=True, quick_mode=False, quick_mode_percent=0.1,
               update_orders=None, seed=None, device=0, session_factory=None):
    """``scregclust(expression, genesymbols, is_regulator, penalization, n_modules, ...)`` of the reference.

    ``expression``: p x n (genes x cells).  ``prior_indicator``: the reference's q x q indicator matrix together
    with ``prior_genesymbols`` (R/scregclust.R:52-63), or -- without ``prior_genesymbols`` -- a ready list of T lists
    of 0-based neighbour targets in target order after the constant-gene filter; ``None`` = no prior.  With a prior
    the caller supplies ``update_orders(penalty_index, cycle)`` to replay R's ``sample()`` (identity otherwise).  Returns a dict with the fields of the R object
    (R/scregclust.R:2056-2120): ``penalization``, ``results`` (one per penalization value, each with
    ``genesymbols``, ``is_regulator``, ``n_modules``, ``converged`` and ``output`` = one entry per final
    configuration), ``initial_target_modules`` and ``split_indices``.  NA is ``nan``.
    """
    ex, isr, pen = _validate(expression, genesymbols, is_regulator, penalization, n_modules, initial_target_modules,
                             sample_assignment, split1_proportion, total_proportion, split_indices, noise_threshold,
                             n_cycles, min_module_size, prior_weight)
    if not isinstance(quick_mode, (bool, np.bool_)):
        _abort("quick_mode needs to be TRUE or FALSE.")                                  # R/scregclust.R:796-806
    if quick_mode and not allocate_per_obs:
        _abort("quick_mode only available when allocate_per_obs is TRUE.")               # :809-817
    if quick_mode and not (0 <= float(quick_mode_percent) < 1):
        _abort("quick_mode_percent needs to be numeric in [0, 1).")                      # :819-831
    if not allocate_per_obs:
        # the aggregate branch is unreachable through the reference's public API
        raise NotImplementedError("allocate_per_obs = FALSE is not built in this library")
    rng = np.random.default_rng(seed)
    p, n = ex.shape
    genesymbols = np.asarray(genesymbols)
    if split_indices is None:
        split_indices = draw_split_indices(n, sample_assignment, split1_proportion, total_proportion, rng)
    split_indices = np.asarray(split_indices, dtype=np.float64)

    factory = session_factory or (lambda *a, **k: driver.Session.from_expression(*a, **k))
    ses = factory(ex, isr, split_indices, int(n_modules), stratification=sample_assignment, center=bool(center),
                  device=device)
    try:
        keep = np.asarray(ses.keep_gene, dtype=bool)            # R/scregclust.R:899-964
        gs, isr_k = genesymbols[keep], isr[keep]
        tgt_keep = keep[isr == 0]
        cross_corr1 = np.asarray(ses.cross_corr())             # T x P (R/scregclust.R:1116)
        if initial_target_modules is not None:
            k_start = np.asarray(initial_target_modules, dtype=np.int32)[tgt_keep]      # :947-949
        else:
            k_start = initial_modules_kmeans(cross_corr1, int(n_modules), n_initializations, rng)
        if prior_indicator is not None and prior_genesymbols is not None:
            prior_indicator = prior_to_target_lists(prior_indicator, prior_genesymbols, gs[isr_k == 0])
        elif prior_indicator is not None and not isinstance(prior_indicator, (list, tuple)):
            _abort("prior_genesymbols missing, but prior_indicator supplied.")          # R/scregclust.R:523-533
        use_prior = prior_indicator is not None and len(prior_indicator) > 0 and float(prior_weight) != 0.0
        fits = driver.scregclust_fit(
            penalization=pen, n_modules=int(n_modules), initial_target_modules=k_start, session=ses,
            min_module_size=int(min_module_size), noise_threshold=float(noise_threshold), n_cycles=int(n_cycles),
            max_optim_iter=int(max_optim_iter), tol_coop_rel=tol_coop_rel, tol_coop_abs=tol_coop_abs, tol_nnls=tol_nnls,
            return_coeffs=True, compute_predictive_r2=bool(compute_predictive_r2),
            prior_indicator=prior_indicator if use_prior else None, prior_baseline=prior_baseline,
            prior_weight=float(prior_weight) if use_prior else 0.0, update_orders=update_orders,
            quick_mode=bool(quick_mode), quick_mode_percent=float(quick_mode_percent))
    finally:
        ses.close()

    is_t = isr_k == 0
    results = []
    for fr in fits:
        output = []
        for m in range(fr.final_cycle_length):                  # R/scregclust.R:1992-2046
            k = fr.k_final[m]
            module = np.full(gs.size, np.nan); module[is_t] = k
            module_all = module.copy()
            ma = k.astype(np.float64); ma[k == -1] = fr.best_r2_idx[m][k == -1]
            module_all[is_t] = ma
            best_r2 = np.full(gs.size, np.nan); best_r2[is_t] = fr.best_r2[m]
            best_idx = np.full(gs.size, np.nan); best_idx[is_t] = fr.best_r2_idx[m]
            max_corr = np.zeros((cross_corr1.shape[1], int(n_modules)))
            for i in range(int(n_modules)):
                rows = cross_corr1[k == i + 1]
                max_corr[:, i] = np.median(rows, axis=0) if rows.shape[0] else np.nan   # apply(..., 2, median)
            with np.errstate(invalid="ignore"):
                reg_table = fr.models[m] * fr.weights[m] * (np.abs(max_corr) > 0.025)
            output.append({
                "reg_table": reg_table, "module": module, "module_all": module_all, "r2": fr.r2[m],
                "best_r2": best_r2, "best_r2_idx": best_idx,
                "r2_module": fr.r2_module[m] if compute_predictive_r2 else None,
                "importance": fr.importance[m] if compute_predictive_r2 else None,
                "r2_removed": fr.r2_removed[m] if compute_predictive_r2 else None,
                "r2_cross_module_per_target": fr.r2[m] if compute_silhouette else None,         # :1961
                "silhouette": silhouette_from_r2(fr.r2[m], k) if compute_silhouette else None,
                "models": fr.models[m], "signs": fr.signs[m], "weights": fr.weights[m], "coeffs": fr.coeffs[m],
                "sigmas": fr.sigmas[m]})
        results.append({"penalization": fr.penalization, "genesymbols": gs.tolist(), "is_regulator": isr_k == 1,
                        "n_modules": int(n_modules), "converged": fr.converged, "output": output})
    return {"penalization": pen, "results": results, "initial_target_modules": k_start,
            "split_indices": split_indices, "quick_mode": False}
