"""One cycle of the alternating loop on the C++ twin (oracle/ref_cpu.cpp), module by module.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ref_cpu.cpp): used by tests/ and by bench.py's
CPU leg as the checker at BASELINE.json sizes, where the NumPy oracle's Python-level LDL^T is too
slow.  Follows R/scregclust.R:1332-1388 (Step 1) and :1533-1638 (Step 2a) call by call, with the
reference's cost structure: every ``coop_lasso`` / ``coef_nnls`` call gets the cell-level matrices
and recomputes its Gram matrices (src/optim.cpp:97-99, 414-421), the linear solve is the restated
``Eigen::LDLT`` (src/optim.cpp:136-143).
"""
from __future__ import annotations

import time

import numpy as np

from . import ref_cpu
from . import scregclust_oracle as orc


def step1_module(w, pre, k, j, lam, tol_rel=1e-8, tol_abs=1e-12, max_iter=10000):
    """coop_lasso for module ``j`` (1-based) of configuration ``k`` (R/scregclust.R:1345-1386).
    Returns dict(iterations, factorizations, selected (0-based regulator indices), signs, weights, seconds)
    or None for an empty module."""
    n1 = w.z1_reg.shape[0]
    cl = np.flatnonzero(k == j)
    if cl.size == 0:
        return None
    res_sds, bas = pre["res_sds"], pre["beta_adapt_sq"]
    t0 = time.perf_counter()
    x = np.asfortranarray(w.z1_reg / np.sqrt(n1))
    y = np.asfortranarray((w.z1_target[:, cl] / res_sds[cl][None, :]) / np.sqrt(n1))
    ws = np.sqrt(cl.size) / np.sqrt(np.sqrt(np.sum(bas[:, cl], axis=1)))
    beta, it, nf = ref_cpu.coop_lasso(y, x, lam, ws, max_iter=max_iter, eps_rel=tol_rel, eps_abs=tol_abs)
    beta = beta * res_sds[cl][None, :]
    beta[np.abs(beta) < 1e-6] = 0.0
    sel = np.flatnonzero(np.sum(np.abs(beta) > 0, axis=1) > 0)
    wt = np.sum(beta, axis=1) / cl.size
    return dict(iterations=int(it), factorizations=int(nf), selected=sel, signs=np.sign(wt[sel]), weights=wt,
                seconds=time.perf_counter() - t0, columns=int(cl.size))


def step2a_module(w, sel, signs, n_selected_total=None, tol_nnls=1e-4, max_iter=10000):
    """coef_nnls on all T targets + train / test residual sums of squares for one module
    (R/scregclust.R:1571-1602).  Returns dict(beta_hat s x T, nnls_iterations T, sse_train T, sse_test T, seconds)."""
    t0 = time.perf_counter()
    y_nn = np.asfortranarray(w.z1_target / w.z1_target_sds[None, :])
    xs = np.asfortranarray(w.z1_reg[:, sel] * signs[None, :])
    b, its = ref_cpu.coef_nnls(xs, y_nn, eps=tol_nnls, max_iter=max_iter)
    bh = b * signs[:, None] * orc._recycle(w.z1_target_sds, b.shape)
    r_tr = w.z1_target - w.z1_reg[:, sel] @ bh
    r_te = w.z2_target - w.z2_reg[:, sel] @ bh
    return dict(beta_hat=bh, nnls_iterations=its, sse_train=(r_tr * r_tr).sum(axis=0), sse_test=(r_te * r_te).sum(axis=0),
                seconds=time.perf_counter() - t0)


def cycle(w, lam, k, pre=None, modules=None, tol_rel=1e-8, tol_abs=1e-12, tol_nnls=1e-4, max_iter=10000):
    """Step 1 + Step 2a of one configuration on the C++ twin, for all (or the listed, 1-based) modules.

    Returns dict with, per module j (0-based position in the K-long arrays; untouched modules keep the
    defaults of R/scregclust.R:1505-1515): admm_iterations[K], models[P, K] bool, signs[P, K] (nan where
    unselected), weights[P, K], nnls_iterations[K, T], r2[T, K], resid_var[T, K], coeffs (list), done (list)."""
    K = w.n_modules
    n1, P = w.z1_reg.shape
    T = w.z1_target.shape[1]
    if pre is None:
        pre = orc.precompute(w.z1_reg, w.z1_target)
    out = dict(admm_iterations=np.zeros(K, np.int32), models=np.zeros((P, K), bool), signs=np.full((P, K), np.nan),
               weights=np.zeros((P, K)), nnls_iterations=np.zeros((K, T), np.int32), coeffs=[None] * K, done=[],
               step1_seconds=0.0, step2a_seconds=0.0, factorizations=0)
    ss_train0 = np.sum(w.z1_target * w.z1_target, axis=0)
    ss_test0 = np.sum(w.z2_target * w.z2_target, axis=0)
    ss_train = np.repeat(ss_train0[:, None], K, axis=1)
    ss_test = np.repeat(ss_test0[:, None], K, axis=1)
    for j in (range(1, K + 1) if modules is None else modules):
        s1 = step1_module(w, pre, k, j, lam, tol_rel, tol_abs, max_iter)
        if s1 is None:
            continue
        out["done"].append(j)
        out["admm_iterations"][j - 1] = s1["iterations"]
        out["factorizations"] += s1["factorizations"]
        out["step1_seconds"] += s1["seconds"]
        out["models"][s1["selected"], j - 1] = True
        out["signs"][s1["selected"], j - 1] = s1["signs"]
        out["weights"][:, j - 1] = s1["weights"]
        if s1["selected"].size == 0:
            continue
        s2 = step2a_module(w, s1["selected"], s1["signs"], tol_nnls=tol_nnls, max_iter=max_iter)
        out["step2a_seconds"] += s2["seconds"]
        out["nnls_iterations"][j - 1] = s2["nnls_iterations"]
        out["coeffs"][j - 1] = s2["beta_hat"]
        ss_train[:, j - 1] = s2["sse_train"]
        ss_test[:, j - 1] = s2["sse_test"]
    out["r2"] = 1.0 - ss_test / ss_test0[:, None]                                  # scregclust.R:1629-1631
    out["resid_var"] = ss_train / (n1 - np.sum(out["models"], axis=0))[None, :]     # scregclust.R:1636-1638
    return out


def full_fit(w, lam, n_cycles=50, noise_threshold=0.025, min_module_size=0, pre=None, log=None):
    """The whole alternating loop of ONE penalization value on the C++ twin with the reference's cost structure:
    Step 1 + 2a per module through :func:`cycle`, the K x n2 x T array reset and strided writes
    (src/utils.cpp:43-63, R/scregclust.R:1599) and ``allocate_into_modules`` (src/allocation.cpp).  Returns
    dict(k_history, admm_iterations (per cycle), seconds, step seconds).  Used to measure the CPU baseline of a
    whole grid (profiles/) and to validate the extrapolation model of bench.py."""
    K = w.n_modules
    n2, T = w.z2_target.shape
    if pre is None:
        pre = orc.precompute(w.z1_reg, w.z1_target)
    z2sq = np.asfortranarray(w.z2_target ** 2)
    arr = np.empty((K, n2, T), order="F")
    k = np.asarray(w.k_init, dtype=np.int32).copy()
    hist = [k.copy()]
    its_hist, t_all = [], dict(step1=0.0, step2a=0.0, step2b=0.0)
    t_start = time.perf_counter()
    last_cycle, fcl, n_k = False, 1, 2
    for cyc in range(1, n_cycles + 2):
        if cyc == n_cycles + 1:
            last_cycle = True
        for m in range(1, fcl + 1):
            kk = hist[len(hist) - m]
            c = cycle(w, lam, kk, pre=pre)
            its_hist.append(c["admm_iterations"].copy())
            t_all["step1"] += c["step1_seconds"]; t_all["step2a"] += c["step2a_seconds"]
        if last_cycle:
            break
        t0 = time.perf_counter()
        ref_cpu.fill_array(arr, z2sq)                                        # reset_array
        for j in range(K):
            if c["coeffs"][j] is not None:
                sel = np.flatnonzero(c["models"][:, j])
                r = w.z2_target - w.z2_reg[:, sel] @ c["coeffs"][j]
                arr[j, :, :] = r * r                                          # sq_residuals_test[j, , ] <- ...
        idx = ref_cpu.allocate_into_modules(arr, np.asfortranarray(c["resid_var"]), None, None, kk,
                                            np.arange(T, dtype=np.int32), 1e-6, 0.0).copy()
        t_all["step2b"] += time.perf_counter() - t0
        idx[c["r2"].max(axis=1) < noise_threshold] = -1
        sizes = np.array([np.sum(idx == i) for i in range(1, K + 1)])
        idx[np.isin(idx, np.flatnonzero(sizes < min_module_size) + 1)] = -1
        k = idx.astype(np.int32)
        match = [i for i, kh in enumerate(hist, start=1) if np.array_equal(kh, k)]
        hist.append(k.copy())
        if match:
            if n_k - match[0] != 1:
                fcl = n_k - match[0]
            last_cycle = True
        n_k += 1
        if log:
            log(f"lambda={lam:.4g} cycle {cyc}: {time.perf_counter() - t_start:.1f} s so far, iterations {its_hist[-1].tolist()}")
    return dict(k_history=hist, admm_iterations=its_hist, seconds=time.perf_counter() - t_start, **t_all)
