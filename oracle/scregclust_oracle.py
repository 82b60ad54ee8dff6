"""CPU oracle for the scregclust hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Operation-for-operation NumPy restatement of the reference's fit / re-estimate /
reallocate loop.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; the product path (``scregclust_b200``) never does.

PARITY STATUS: **parity unpinned** for coop_lasso / coef_nnls /
allocate_into_modules / coef_ols / coef_ridge (the reference ships no golden
vectors for them, R + Eigen are absent in the build container so the reference
cannot be executed, and ``datasets/*.rds`` are Git-LFS pointers).  ``fast_cor``
is pinned through the reference's own test (tests/testthat/test-fast-cor.R:1-16:
equal to ``stats::cor`` == ``numpy.corrcoef`` at tolerance 1.5e-8), and the
constant-gene filter of the input staging through the reference's other test
(tests/testthat/test-constant-genes.R:1-24: T1 / R1 dropped).  Independent
cross-checks (KKT conditions, the single-column case against scikit-learn's
weighted lasso, scipy.optimize.nnls, lstsq, brute-force allocation) live in
tests/test_oracle.py.

Third-party arithmetic restated here (absent from /root/reference):
  * Eigen 3.4.0 (bundled by RcppEigen 0.3.4.x; DESCRIPTION:39 "LinkingTo: Rcpp,
    RcppEigen", no version pin): ``LDLT::compute/solve`` = diagonally pivoted
    unblocked left-looking LDL^T (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>
    ::unblocked and LDLT::_solve_impl), ``maxCoeff(&i)`` = first index of the
    maximum, ``norm()`` = sqrt(sum x^2) without scaling.
  * R's BLAS/LAPACK for crossprod / solve / %*% (R/utils.R:13-57,533-541).
Summation order inside GEMM/SYRK is implementation defined in both, so
"bit-exact" refers to discrete outputs (iteration counts, selected regulators,
module assignments); floating-point outputs are compared at 1e-8 relative.

All matrices are float64, column-major where layout matters; module ids are
1-based with -1 = noise, exactly as the R driver passes them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "compute_xtx", "EigenLDLT", "coop_lasso", "greedy_coord_descent", "coef_nnls",
    "allocate_into_modules", "alloc_array", "reset_array", "coef_ols", "coef_ridge",
    "fast_cor", "adjusted_rand_index", "precompute", "scregclust_fit", "FitResult",
    "scale_inputs",
]


# --------------------------------------------------------------------------- #
# src/optim.cpp:17-27
# --------------------------------------------------------------------------- #
def compute_xtx(x: np.ndarray) -> np.ndarray:
    """X'X via a lower rank update mirrored to the upper triangle (src/optim.cpp:17-27)."""
    p = x.shape[1]
    xtx = np.zeros((p, p))
    if p > 0:
        full = x.T @ x
        low = np.tril(full)
        xtx = low + np.tril(full, -1).T  # upper := lower'
    return xtx


# --------------------------------------------------------------------------- #
# Eigen 3.4.0 Eigen/src/Cholesky/LDLT.h (restated; see module docstring)
# --------------------------------------------------------------------------- #
class EigenLDLT:
    """Diagonally pivoted LDL^T, unblocked, lower storage (Eigen 3.4.0 LDLT.h).

    compute(): at step k pick the largest |diagonal| of the trailing block, apply
    the symmetric transposition, then the left-looking update
        A_kk -= A_k,0:k * D_0:k * A_k,0:k' ;  A_{k+1:,k} -= A_{k+1:,0:k} (D A_k,0:k')
        A_{k+1:,k} /= A_kk.
    solve(): b := P b; L^-1; D^-1 (entries with |d| <= realmin give 0); L^-T; P^-1.
    """

    def __init__(self):
        self.mat = None
        self.trans = None

    def compute(self, a: np.ndarray) -> "EigenLDLT":
        mat = np.array(a, dtype=np.float64, order="F", copy=True)
        n = mat.shape[0]
        trans = np.arange(n)
        for k in range(n):
            diag = np.abs(np.diagonal(mat)[k:])
            piv = int(np.argmax(diag)) + k
            trans[k] = piv
            if piv != k:
                # symmetric swap acting on the lower triangle only
                s = n - piv - 1
                tmp = mat[k, :k].copy(); mat[k, :k] = mat[piv, :k]; mat[piv, :k] = tmp
                tmp = mat[piv + 1:, k].copy(); mat[piv + 1:, k] = mat[piv + 1:, piv]; mat[piv + 1:, piv] = tmp
                mat[k, k], mat[piv, piv] = mat[piv, piv], mat[k, k]
                for i in range(k + 1, piv):
                    mat[i, k], mat[piv, i] = mat[piv, i], mat[i, k]
                del s
            rs = n - k - 1
            if k > 0:
                temp = np.diagonal(mat)[:k] * mat[k, :k]
                mat[k, k] -= mat[k, :k] @ temp
                if rs > 0:
                    mat[k + 1:, k] -= mat[k + 1:, :k] @ temp
            akk = mat[k, k]
            if rs > 0 and abs(akk) > 0.0:
                mat[k + 1:, k] /= akk
        self.mat = mat
        self.trans = trans
        return self

    def solve(self, b: np.ndarray) -> np.ndarray:
        mat, trans = self.mat, self.trans
        n = mat.shape[0]
        x = np.array(b, dtype=np.float64, order="F", copy=True)
        if x.ndim == 1:
            x = x[:, None]
        for k in range(n):  # dst = P b
            t = trans[k]
            if t != k:
                x[[k, t], :] = x[[t, k], :]
        for k in range(n):  # L^-1 (unit lower)
            if k > 0:
                x[k, :] -= mat[k, :k] @ x[:k, :]
        d = np.diagonal(mat)
        tol = np.finfo(np.float64).tiny
        for k in range(n):
            if abs(d[k]) > tol:
                x[k, :] /= d[k]
            else:
                x[k, :] = 0.0
        for k in range(n - 1, -1, -1):  # L^-T
            if k < n - 1:
                x[k, :] -= mat[k + 1:, k] @ x[k + 1:, :]
        for k in range(n - 1, -1, -1):  # P^-1
            t = trans[k]
            if t != k:
                x[[k, t], :] = x[[t, k], :]
        return x


def _fnorm(a: np.ndarray) -> float:
    """Eigen ``.norm()``: sqrt of the plain sum of squares."""
    return math.sqrt(float(np.sum(a * a)))


# --------------------------------------------------------------------------- #
# src/optim.cpp:63-320
# --------------------------------------------------------------------------- #
def coop_lasso(y, x, lam, weights, beta_0=None, rho_0=0.2, alpha_0=1.5, n_update=2,
               eps_corr=0.2, max_iter=1000, eps_rel=1e-8, eps_abs=1e-12,
               xtx=None, xty=None, trace=None, fast_solver=True):
    """Adaptive relaxed ADMM for the cooperative lasso (src/optim.cpp:63-320).

    Returns ``(beta, iterations)`` with ``beta`` the un-thresholded ADMM variable
    (optim.cpp:315) and ``iterations == max_iter + 1`` on overrun (optim.cpp:295-305).
    ``xtx`` / ``xty`` may be supplied to skip the per-call recomputation
    (optim.cpp:97-99) -- same values, used by the loop-level oracle for speed.
    ``fast_solver`` replaces the restated Eigen LDLT by LAPACK Cholesky
    (identical to rounding on the SPD system; the restated LDLT is exercised in
    tests/test_oracle.py).
    """
    if xtx is None or xty is None:
        n, m = y.shape
        p = x.shape[1]
        if n <= 0 or m <= 0 or p <= 0 or x.shape[0] <= 0:
            raise ValueError("COOP LASSO: Matrix dimensions of y and x need to be positive.")
        if x.shape[0] != n:
            raise ValueError("y and x need to have the same number of rows.")
        xtx = compute_xtx(x)
        xty = x.T @ y
    p, m = xty.shape
    weights = np.asarray(weights, dtype=np.float64)

    beta = np.zeros((p, m))
    if beta_0 is not None:
        if beta_0.shape != (p, m):
            raise ValueError("beta_0 needs to be of size p x m")
        beta = np.array(beta_0, dtype=np.float64)
    zeta = beta.copy()
    mult = np.zeros((p, m))
    beta_old = beta.copy(); zeta_old = zeta.copy()
    mult_old = mult.copy(); mult_hat_old = mult.copy()

    if rho_0 <= 0.0:
        raise ValueError("rho_0 > 0 needs to hold")
    rho = float(rho_0)
    rho_old = 0.0
    if alpha_0 <= 0.0 or alpha_0 > 2.0:
        raise ValueError("0 < alpha_0 < 2 needs to hold")
    alpha = float(alpha_0)

    it = 1
    pc = 0
    factor = None
    factor_rho = None
    n_factor = 0
    idx = np.arange(p)

    while True:
        if pc == 0 or rho != rho_old:
            # refactorising the same matrix is idempotent -> only redo on a new rho
            if factor is None or factor_rho != rho:
                a = xtx.copy()
                a[idx, idx] = xtx[idx, idx] + rho
                if fast_solver:
                    import scipy.linalg as sla
                    factor = sla.cho_factor(a, lower=True, check_finite=False)
                else:
                    factor = EigenLDLT().compute(a)
                factor_rho = rho
            n_factor += 1

        rhs = xty + rho * zeta + mult
        if fast_solver:
            import scipy.linalg as sla
            beta = sla.cho_solve(factor, rhs, check_finite=False)
        else:
            beta = factor.solve(rhs)

        beta_relaxed = alpha * beta + (1.0 - alpha) * zeta
        zeta_new = beta_relaxed - mult / rho

        with np.errstate(divide="ignore", invalid="ignore"):
            npos = np.sqrt(np.sum(np.maximum(zeta_new, 0.0) ** 2, axis=1))
            nneg = np.sqrt(np.sum(np.minimum(zeta_new, 0.0) ** 2, axis=1))
            shrink_pos = np.maximum(1.0 - lam * weights / (rho * npos), 0.0)
            shrink_neg = np.maximum(1.0 - lam * weights / (rho * nneg), 0.0)
        zeta_new = np.where(zeta_new >= 0.0, zeta_new * shrink_pos[:, None],
                            zeta_new * shrink_neg[:, None])

        mult = mult + rho * (-beta_relaxed + zeta_new)

        primal = -beta + zeta_new
        dual = rho * (zeta - zeta_new)
        n_primal = _fnorm(primal); n_dual = _fnorm(dual)
        n_beta = _fnorm(beta); n_zeta = _fnorm(zeta_new); n_mult = _fnorm(mult)
        if trace is not None:
            trace.append(dict(it=it, rho=rho, alpha=alpha, primal=n_primal, dual=n_dual,
                              nbeta=n_beta, nzeta=n_zeta, nmult=n_mult))
        if (n_primal <= max(eps_rel * max(n_beta, n_zeta), eps_abs)
                and n_dual <= max(eps_rel * n_mult, eps_abs)):
            break

        pc += 1
        if pc == n_update:
            mult_hat = mult + rho * (-beta + zeta)
            d_mh = mult_hat - mult_hat_old
            d_h = beta - beta_old
            d_m = mult - mult_old
            d_g = zeta_old - zeta
            n_dmh = _fnorm(d_mh); n_dh = _fnorm(d_h); n_dm = _fnorm(d_m); n_dg = _fnorm(d_g)
            # IEEE semantics (inf / nan instead of Python's ZeroDivisionError), as in C++
            f8 = np.float64
            with np.errstate(all="ignore"):
                a_ = 0.0; a_corr = 0.0
                if n_dmh > 0.0 and n_dh > 0.0:
                    hm = f8(np.sum(d_h * d_mh))
                    a_sd = f8(np.sum(d_mh * d_mh)) / hm
                    a_mg = hm / f8(np.sum(d_h * d_h))
                    a_ = float(a_mg if 2.0 * a_mg > a_sd else a_sd - a_mg / 2.0)
                    a_corr = float(hm / (f8(n_dh) * f8(n_dmh)))
                b_ = 0.0; b_corr = 0.0
                if n_dm > 0.0 and n_dg > 0.0:
                    gm = f8(np.sum(d_g * d_m))
                    b_sd = f8(np.sum(d_m * d_m)) / gm
                    b_mg = gm / f8(np.sum(d_g * d_g))
                    b_ = float(b_mg if 2.0 * b_mg > b_sd else b_sd - b_mg / 2.0)
                    b_corr = float(gm / (f8(n_dg) * f8(n_dm)))
            rho_old = rho
            with np.errstate(all="ignore"):
                ga = a_corr > eps_corr; gb = b_corr > eps_corr
                if ga and gb:
                    rho = float(np.sqrt(f8(a_) * f8(b_)))
                    alpha = float(1.0 + 2.0 / (np.sqrt(f8(a_) * f8(b_)) * (1.0 / f8(a_) + 1.0 / f8(b_))))
                elif ga and not gb:
                    rho = a_; alpha = 1.9
                elif (not ga) and gb:
                    rho = b_; alpha = 1.1
                else:
                    alpha = 1.5
            beta_old = beta.copy(); zeta_old = zeta_new.copy()
            mult_old = mult.copy(); mult_hat_old = mult_hat
            pc = 0

        zeta = zeta_new
        it += 1
        if it > max_iter:
            break
    return np.asfortranarray(beta), it


# --------------------------------------------------------------------------- #
# src/optim.cpp:335-377
# --------------------------------------------------------------------------- #
def greedy_coord_descent(Q, beta, grad):
    """In-place greedy coordinate descent (src/optim.cpp:335-377), all columns."""
    n, m = beta.shape
    cols = np.arange(m)
    qdiag = np.diagonal(Q)
    for _ in range(n):
        score = ((beta > 0.0) | (grad < 0.0)).astype(np.float64) * np.abs(grad)
        p = np.argmax(score, axis=0)           # first index of the max (Eigen maxCoeff)
        max_val = score[p, cols]
        act = max_val != 0.0
        if not np.any(act):
            break
        pa = p[act]; ca = cols[act]
        b = beta[pa, ca]
        dbeta = np.maximum(0.0, b - grad[pa, ca] / qdiag[pa]) - b
        beta[pa, ca] = b + dbeta
        grad[:, ca] += dbeta[None, :] * Q[:, pa]


def _nnls_neg_correction(Q, beta, grad, colmask):
    n = beta.shape[0]
    for i in range(n):
        neg = colmask & (beta[i, :] < 0.0)
        if np.any(neg):
            grad[:, neg] -= beta[i, neg][None, :] * Q[:, [i]]
            beta[i, neg] = 0.0


# --------------------------------------------------------------------------- #
# src/optim.cpp:403-549
# --------------------------------------------------------------------------- #
def coef_nnls(x=None, y=None, eps=1e-12, max_iter=1000, xtx=None, xty=None):
    """Accelerated anti-lopsided NNLS with matrix RHS (src/optim.cpp:403-549).

    Returns ``(beta[n x m], iterations[m])``; never-converged columns keep
    ``beta == 0`` and ``iterations == max_iter`` (optim.cpp:423,431-432,502).
    ``xtx`` / ``xty`` may replace ``x`` / ``y`` (same values as optim.cpp:414,421).
    """
    if xtx is None:
        if x.shape[1] <= 0 or y.shape[1] <= 0 or x.shape[0] <= 0 or y.shape[0] <= 0:
            raise ValueError("NNLS: Matrix dimensions of y and x need to be positive.")
        xtx = compute_xtx(x)
        xty = x.T @ y
    n = xtx.shape[0]
    m_total = xty.shape[1]
    with np.errstate(divide="ignore", invalid="ignore"):
        d = 1.0 / np.sqrt(np.diagonal(xtx))
    Q = xtx * np.outer(d, d)
    grad = (-xty) * d[:, None]
    beta_final = np.zeros((n, m_total))
    beta = np.zeros((n, m_total))
    grad_bar = grad * (1.0 - ((beta == 0.0) & (grad > 0.0)))
    iterations = np.full(m_total, max_iter, dtype=np.int32)
    remaining = np.arange(m_total)

    for l in range(max_iter):
        m = beta.shape[1]
        allcols = np.ones(m, dtype=bool)
        beta_save = beta.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            q_gb = Q @ grad_bar
            alpha1 = np.sum(grad_bar * grad_bar, axis=0) / np.sum(grad_bar * q_gb, axis=0)
        ok = (alpha1 == alpha1) & (np.abs(alpha1) >= 1e-20) & (np.abs(alpha1) < 1e30)
        if np.any(ok):
            beta[:, ok] -= alpha1[ok][None, :] * grad_bar[:, ok]
            grad[:, ok] -= alpha1[ok][None, :] * q_gb[:, ok]
            _nnls_neg_correction(Q, beta, grad, ok)
        greedy_coord_descent(Q, beta, grad)

        dbeta = beta_save - beta
        with np.errstate(divide="ignore", invalid="ignore"):
            q_db = Q @ dbeta
            alpha2 = np.sum((grad * dbeta) ** 2, axis=0) / np.sum(dbeta * q_db, axis=0)
        ok = (alpha2 == alpha2) & (np.abs(alpha2) >= 1e-20) & (np.abs(alpha2) < 1e30)
        if np.any(ok):
            beta[:, ok] -= alpha2[ok][None, :] * dbeta[:, ok]
            grad[:, ok] -= alpha2[ok][None, :] * q_db[:, ok]
            _nnls_neg_correction(Q, beta, grad, ok)
        greedy_coord_descent(Q, beta, grad)

        grad_bar = grad * (1.0 - ((beta == 0.0) & (grad > 0.0)))
        norms = np.sqrt(np.sum(grad_bar * grad_bar, axis=0))
        conv = norms < eps
        if np.any(conv):
            beta_final[:, remaining[conv]] = beta[:, conv]
            iterations[remaining[conv]] = l + 1
        keep = ~conv
        remaining = remaining[keep]
        if remaining.size == 0:
            break
        beta = np.ascontiguousarray(beta[:, keep])
        grad = np.ascontiguousarray(grad[:, keep])
        grad_bar = np.ascontiguousarray(grad_bar[:, keep])
        del allcols
    beta_final *= d[:, None]
    return np.asfortranarray(beta_final), iterations


# --------------------------------------------------------------------------- #
# src/utils.cpp:12-63
# --------------------------------------------------------------------------- #
def alloc_array(inp: np.ndarray, n_cl: int) -> np.ndarray:
    """K x n_obs x n_genes broadcast of ``inp`` along the first axis (src/utils.cpp:12-34)."""
    n_obs, n_genes = inp.shape
    arr = np.empty((n_cl, n_obs, n_genes), order="F")
    arr[:] = inp[None, :, :]
    return arr


def reset_array(arr: np.ndarray, inp: np.ndarray) -> None:
    """In-place refill (src/utils.cpp:43-63)."""
    if inp.shape != arr.shape[1:]:
        raise ValueError("reset_array: input has wrong dimensions")
    arr[:] = inp[None, :, :]


# --------------------------------------------------------------------------- #
# src/allocation.cpp:8-100
# --------------------------------------------------------------------------- #
# Relative score gap below which a cell's argmax over two structurally distinct modules counts as
# decided by rounding noise.  Calibrated on what was measured, not on convenience: predictions of the
# CUDA path agree with this oracle to 3e-15 relative and the Gram-form training SSE (hence sigma^2, which
# scales the whole score) to 1e-14, so score differences between implementations stay below ~1e-13;
# 1e-12 leaves one order of magnitude and is three orders tighter than round 1's 1e-9.
AMBIGUITY_REL = 1e-12


def allocate_into_modules(resid_array, resid_var, prior_indicator, k_, update_order,
                          prior_baseline, prior_weight, return_votes=False, ambiguity=None,
                          distinct=None):
    """Per-cell vote reallocation (src/allocation.cpp:8-100).

    ``resid_array`` is K x n2 x T of *squared* residuals, ``resid_var`` T x K,
    ``k_`` 1-based (-1 noise), ``update_order`` / ``prior_indicator`` 0-based.

    ``ambiguity`` (optional int array of length T, filled in place) receives, per gene, the
    number of cells whose two best scores -- over structurally distinct modules, ``distinct`` being
    a boolean mask that keeps one representative of every group of modules with identical
    regulator sets and signs -- agree to ``AMBIGUITY_REL`` relative, i.e. cells whose vote is decided by
    rounding noise.  (Modules with different regulator sets can predict a gene identically when
    its coefficients live on the shared regulators; whether such a tie is exact or off by one ulp
    depends on the summation order of the BLAS / tensor-core kernel.  Structurally identical
    modules tie bit-exactly in every implementation and go to the first index.)  Diagnostics only.
    """
    n_modules, n_obs, n_genes = resid_array.shape
    k = np.asarray(k_, dtype=np.int64) - 1
    module_totals = np.zeros(n_modules)
    for idx in k:
        if idx != -2:
            module_totals[idx] += 1
    votes_all = np.zeros((n_genes, n_modules), dtype=np.int32) if return_votes else None
    for j in update_order:
        resid = resid_array[:, :, j]                       # K x n_obs
        prior_frac = np.zeros(n_modules)
        pri = prior_indicator[j] if len(prior_indicator) > 0 else ()
        for idx in pri:
            if k[idx] != -2:
                prior_frac[k[idx]] += 1
        nz = module_totals > 0
        prior_frac[nz] /= module_totals[nz]
        prior_frac += prior_baseline
        prior_log_prob = np.log(prior_frac) - math.log(float(np.sum(prior_frac)))
        var = resid_var[j, :]
        ll = (-resid) / (2.0 * var)[:, None] - (0.5 * np.log(2.0 * math.pi * var))[:, None]
        score = (1.0 - prior_weight) * ll + (prior_weight * prior_log_prob)[:, None]
        best_per_cell = np.argmax(score, axis=0)            # first max
        if ambiguity is not None:
            sc = score if distinct is None else score[np.asarray(distinct, dtype=bool)]
            nd = sc.shape[0]
            if nd > 1:
                part = np.partition(sc, nd - 2, axis=0)
                top1 = part[nd - 1]; top2 = part[nd - 2]
                ambiguity[j] = int(np.sum((top1 - top2) <= AMBIGUITY_REL * np.maximum(1.0, np.abs(top1))))
        votes = np.bincount(best_per_cell, minlength=n_modules).astype(np.int32)
        best = int(np.argmax(votes))                        # first max
        if return_votes:
            votes_all[j] = votes
        if k[j] != -2:
            module_totals[k[j]] -= 1
        module_totals[best] += 1
        k[j] = best
    out = (k + 1).astype(np.int32)
    return (out, votes_all) if return_votes else out


def allocate_aggregate(k_, sum_squares_test, resid_var, n2, prior_indicator, update_order,
                       prior_baseline, prior_weight, ambiguity=None):
    """Aggregate reallocation, the ``allocate_per_obs = FALSE`` branch of Step 2b (R/scregclust.R:1678-1728).

    ``sum_squares_test`` / ``resid_var`` are T x K, ``k_`` 1-based (-1 noise), ``update_order`` 0-based here (R draws
    1-based gene indices, ``sample(length(idx), length(idx))``), ``prior_indicator`` 0-based as everywhere.
    Restated as written, including what looks unintended: the scores are built from
    ``n2 log(2 pi s2) + SS / s2`` (minus twice a log-likelihood) but normalised like log-probabilities, mixed with
    the prior's log-probabilities and minimised (``which.min``), and ``module_totals`` is computed once from the
    incoming configuration and never updated while genes move.  The branch cannot be reached through the
    public API of v0.2.2 (the argument check at R/scregclust.R:807-817 rejects ``allocate_per_obs = FALSE``).
    ``ambiguity`` (optional int array of length T, filled in place): 1 where the two smallest scores of a gene agree to
    ``AMBIGUITY_REL`` relative -- ``log`` / ``exp`` differ in the last place between math libraries.
    """
    idx = np.asarray(k_, dtype=np.int64).copy()
    n_genes, n_modules = resid_var.shape
    module_totals = np.array([np.sum(idx == i) for i in range(1, n_modules + 1)], dtype=np.float64)   # :1682
    with np.errstate(all="ignore"):
        mll = float(n2) * np.log((2.0 * math.pi) * resid_var) + sum_squares_test / resid_var          # :1685-1688
        d = mll - np.max(mll, axis=1)[:, None]                                                        # :1691-1694
        rs = np.zeros(n_genes, dtype=np.longdouble)            # rowSums accumulates in long double, column by column
        for j in range(n_modules):
            rs = rs + np.exp(d[:, j])
        mlp = d - np.log(rs.astype(np.float64))[:, None]                                               # :1695-1698
    for gene in update_order:
        prior_frac = np.zeros(n_modules)
        pri = prior_indicator[gene] if len(prior_indicator) > 0 else ()
        for i in pri:                                                                                  # :1706-1714
            if 1 <= idx[i] <= n_modules:
                prior_frac[idx[i] - 1] += 1.0
        nz = module_totals > 0
        prior_frac[nz] = prior_frac[nz] / module_totals[nz]                                            # :1716-1718
        prior_frac = prior_frac + prior_baseline
        tot = np.longdouble(0.0)                               # sum() accumulates in long double
        for v in prior_frac:
            tot = tot + v
        with np.errstate(all="ignore"):
            prior_log_prob = np.log(prior_frac) - math.log(float(tot))                                 # :1720
            total = (1.0 - prior_weight) * mlp[gene] + prior_weight * prior_log_prob                   # :1722-1725
        ok = ~np.isnan(total)
        if not ok.any():
            raise ValueError("which.min of an all-NaN score vector (R: replacement has length zero)")
        cand = np.where(ok, total, np.inf)
        best = int(np.argmin(cand))                            # which.min: first minimum, NaN skipped
        if ambiguity is not None and n_modules > 1:
            two = np.partition(cand, 1)[:2]
            ambiguity[gene] = int((two[1] - two[0]) <= AMBIGUITY_REL * max(1.0, abs(two[0])))
        idx[gene] = best + 1                                                                           # :1727
    return idx.astype(np.int32)


# --------------------------------------------------------------------------- #
# R/utils.R:13-57, 533-541, 631-656
# --------------------------------------------------------------------------- #
def coef_ols(y, x):
    """OLS via normal equations, pseudo-inverse when n < p (R/utils.R:13-37)."""
    n, p = x.shape
    xtx = x.T @ x
    xty = x.T @ y
    if n < p:
        u, d, vt = np.linalg.svd(xtx)
        keep = d > np.finfo(np.float64).eps * p * d.max()
        dinv = np.where(keep, 1.0 / np.where(keep, d, 1.0), 0.0)
        return (vt.T * dinv) @ u.T @ xty
    return np.linalg.solve(xtx, xty)


def coef_ridge(y, x, lam):
    """Ridge via normal equations (R/utils.R:49-57)."""
    p = x.shape[1]
    return np.linalg.solve(x.T @ x + lam * np.eye(p), x.T @ y)


def fast_cor(x, y):
    """Column-centred Pearson cross-correlation clamped to [-1, 1] (R/utils.R:533-541)."""
    xv = x - x.mean(axis=0)
    yv = y - y.mean(axis=0)
    xvss = np.sum(xv * xv, axis=0)
    yvss = np.sum(yv * yv, axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        res = (xv.T @ yv) / np.sqrt(np.outer(xvss, yvss))
    return np.maximum(np.minimum(res, 1.0), -1.0)


def adjusted_rand_index(k1, k2) -> float:
    """Hubert-Arabie adjusted Rand index (R/utils.R:631-656)."""
    k1 = np.asarray(k1); k2 = np.asarray(k2)
    n = k1.size
    _, i1 = np.unique(k1, return_inverse=True)
    _, i2 = np.unique(k2, return_inverse=True)
    ct = np.zeros((i1.max() + 1, i2.max() + 1), dtype=np.int64)
    np.add.at(ct, (i1, i2), 1)
    ch2 = lambda v: v * (v - 1) / 2.0
    sum_as = ch2(ct.sum(axis=1)).sum(); sum_bs = ch2(ct.sum(axis=0)).sum()
    sum_ns = ch2(ct).sum(); denom = ch2(n)
    return float((sum_ns - sum_as * sum_bs / denom) / (0.5 * (sum_as + sum_bs) - sum_as * sum_bs / denom))


# --------------------------------------------------------------------------- #
# Driver: R/scregclust.R:996-1113 (pre-compute)
# --------------------------------------------------------------------------- #
def _sd(a):
    return np.std(a, axis=0, ddof=1)


def split_sample(z, stratification, is_regulator, split_indices, center=True):
    """split_sample with explicit ``split_indices`` (R/scregclust.R:2302-2325).

    ``z``: p x n expression (genes x cells); ``split_indices``: 1 / 2 / NaN (= NA) per cell.
    Quirk kept on purpose: with ``center`` the reference rebuilds ``z_`` with ``cbind`` over the
    strata (:2309-2316), which re-orders its columns stratum by stratum, and then still selects the
    first split by the positions computed on the original order (:2320-2323).
    Returns z1_reg, z2_reg, z1_target, z2_target (cells x genes)."""
    z = np.asarray(z, dtype=np.float64)
    si = np.asarray(split_indices, dtype=np.float64)
    is_included = np.flatnonzero(~np.isnan(si))                 # which(!is.na(split_indices))
    is_split1 = np.flatnonzero(si[is_included] == 1)            # positions inside is_included
    z_ = z[:, is_included]
    if center:
        strat = (np.zeros(is_included.size, dtype=np.int64) if stratification is None
                 else np.asarray(stratification)[is_included])
        blocks = []
        for lev in np.unique(strat):                            # split(): sorted factor levels
            s_ = np.flatnonzero(strat == lev)
            idx = np.intersect1d(s_, is_split1)
            blocks.append(z_[:, s_] - z_[:, idx].mean(axis=1, keepdims=True))
        z_ = np.concatenate(blocks, axis=1)
    in1 = np.zeros(z_.shape[1], dtype=bool)
    in1[is_split1] = True
    isr = np.asarray(is_regulator)
    reg, tgt = isr == 1, isr == 0
    return (np.asfortranarray(z_[:, in1][reg].T), np.asfortranarray(z_[:, ~in1][reg].T),
            np.asfortranarray(z_[:, in1][tgt].T), np.asfortranarray(z_[:, ~in1][tgt].T))


def constant_genes(z1, z2):
    """Columns with ``sd == 0`` in either split (R/scregclust.R:900-907).  R's ``sd`` of a column of
    identical values is exactly 0 (long-double mean with a refinement pass), which a plain float64
    two-pass formula does not guarantee -- so the test is stated as "all values coincide"."""
    return (np.ptp(z1, axis=0) == 0) | (np.ptp(z2, axis=0) == 0)


def stage_inputs(z, stratification, is_regulator, split_indices, center=True):
    """Raw expression -> the five scaled inputs of the loop + per-gene keep mask
    (R/scregclust.R:880-964 + 996-1009)."""
    z1r, z2r, z1t, z2t = split_sample(z, stratification, is_regulator, split_indices, center)
    creg, ctgt = constant_genes(z1r, z2r), constant_genes(z1t, z2t)
    isr = np.asarray(is_regulator)
    keep = np.zeros(isr.size, dtype=bool)
    keep[np.flatnonzero(isr == 1)[~creg]] = True
    keep[np.flatnonzero(isr == 0)[~ctgt]] = True
    out = scale_inputs(z1r[:, ~creg], z2r[:, ~creg], z1t[:, ~ctgt], z2t[:, ~ctgt])
    return out + (keep,)


def scale_inputs(z1_reg, z2_reg, z1_target, z2_target):
    """Scaling with split-1 statistics (R/scregclust.R:998-1009)."""
    m1t = z1_target.mean(axis=0)
    z1_target_centered = z1_target - m1t
    z1_target_sds = _sd(z1_target)
    m1r = z1_reg.mean(axis=0); s1r = _sd(z1_reg)
    z1_reg_scaled = (z1_reg - m1r) / s1r
    z2_target_centered = z2_target - m1t
    z2_reg_scaled = (z2_reg - m1r) / s1r
    f = np.asfortranarray
    return f(z1_reg_scaled), f(z2_reg_scaled), f(z1_target_centered), f(z2_target_centered), z1_target_sds


def precompute(z1_reg_scaled, z1_target_centered):
    """beta_init, res_sds, beta_adapt_sq (R/scregclust.R:1049-1113)."""
    n1, n_reg = z1_reg_scaled.shape
    if n1 >= n_reg:
        beta_init = coef_ols(z1_target_centered, z1_reg_scaled)
        init_df = float(n_reg)
        lam_min = None
    else:
        ztz = z1_reg_scaled.T @ z1_reg_scaled
        es = np.sort(np.linalg.eigvalsh(ztz))[::-1]       # R eigen(): decreasing
        lam_min = es[n1 - 2] + 1e-4                        # es[n1 - 1] 1-based
        init_df = float(np.sum(es / (es + lam_min)))
        while init_df >= n1:
            lam_min = 2 * lam_min
            init_df = float(np.sum(es / (es + lam_min)))
        beta_init = coef_ridge(z1_target_centered, z1_reg_scaled, lam_min)
    res = z1_target_centered - z1_reg_scaled @ beta_init
    res_sds = np.sqrt(np.sum(res * res, axis=0) / (n1 - init_df))
    beta_adapt_sq = (beta_init / res_sds[None, :]) ** 2
    return dict(beta_init=beta_init, init_df=init_df, res_sds=res_sds,
                beta_adapt_sq=beta_adapt_sq, lambda_minimal=lam_min)


@dataclass
class FitResult:
    """Per-penalty outcome of the alternating loop (subset of R/scregclust.R:2002-2088)."""
    penalization: float
    converged: bool
    n_cycles_run: int
    k_history: list
    final_cycle_length: int
    k_final: list = field(default_factory=list)            # one per final configuration
    models: list = field(default_factory=list)             # P x K bool
    signs: list = field(default_factory=list)              # P x K (nan where unselected)
    weights: list = field(default_factory=list)            # P x K
    coeffs: list = field(default_factory=list)             # list over modules of s_j x T (or None)
    sigmas: list = field(default_factory=list)             # T x K
    r2: list = field(default_factory=list)                 # T x K
    best_r2: list = field(default_factory=list)            # T
    best_r2_idx: list = field(default_factory=list)        # T (1-based)
    admm_iterations: list = field(default_factory=list)    # per cycle: length-K int (0 = skipped)
    nnls_iterations: list = field(default_factory=list)    # per cycle: K x T int (0 = skipped)
    votes: list = field(default_factory=list)              # per cycle: T x K int (None in the aggregate branch)
    allocations: list = field(default_factory=list)        # per cycle: T int, Step 2b's result before the noise / size rules
    ambiguity: list = field(default_factory=list)          # per cycle: T int (rounding-tied cells)
    models_history: list = field(default_factory=list)
    resid_var_history: list = field(default_factory=list)  # per cycle T x K
    r2_history: list = field(default_factory=list)         # per cycle T x K
    coeffs_history: list = field(default_factory=list)     # per cycle list of s_j x T
    r2_module: list = field(default_factory=list)
    importance: list = field(default_factory=list)
    r2_removed: list = field(default_factory=list)         # per final configuration: list over modules


def _recycle(v, shape):
    """R recycling of a length-T vector over an s x T matrix (column-major flat index).

    ``coef_nnls(...)$beta * signs_cl * z1_target_sds_tmp`` (R/scregclust.R:1580-1588)
    multiplies element (a, t) by ``sds[(a + t*s) mod T]`` -- NOT ``sds[t]`` -- because R
    recycles the vector down the columns of the s x T matrix.  Observable in every
    output downstream of Step 2a, hence reproduced here and on the GPU.
    """
    s_, t_ = shape
    flat = np.arange(s_)[:, None] + np.arange(t_)[None, :] * s_
    return np.asarray(v)[flat % len(v)]


def _step1(pre, z1r_x, G_scaled, xty_scaled_all, k, lam, n_modules, tol_rel, tol_abs,
           max_iter, admm_it):
    """Step 1 for one configuration (R/scregclust.R:1313-1388)."""
    n_reg = G_scaled.shape[0]
    models = np.zeros((n_reg, n_modules), dtype=bool)
    weights = np.zeros((n_reg, n_modules))
    signs = np.full((n_reg, n_modules), np.nan)
    res_sds = pre["res_sds"]; bas = pre["beta_adapt_sq"]
    for j in range(1, n_modules + 1):
        cl = np.flatnonzero(k == j)
        m_j = cl.size
        if m_j == 0:
            continue
        ws = math.sqrt(m_j) / np.sqrt(np.sqrt(np.sum(bas[:, cl], axis=1)))
        # y = (Y_cl diag(1/sd)) / sqrt(n1), x = Z1r / sqrt(n1)  (scregclust.R:1353-1364)
        beta, its = coop_lasso(None, None, lam, ws, eps_rel=tol_rel, eps_abs=tol_abs,
                               max_iter=max_iter, xtx=G_scaled, xty=xty_scaled_all[:, cl])
        admm_it[j - 1] = its
        beta = beta * res_sds[cl][None, :]
        beta[np.abs(beta) < 1e-6] = 0.0
        models[:, j - 1] = np.sum(np.abs(beta) > 0, axis=1) > 0
        weights[:, j - 1] = np.sum(beta, axis=1) / m_j
        sel = models[:, j - 1]
        signs[sel, j - 1] = np.sign(weights[sel, j - 1])
    return models, weights, signs


def scregclust_fit(z1_reg_scaled, z2_reg_scaled, z1_target_centered, z2_target_centered,
                   z1_target_sds, penalization, n_modules, initial_target_modules,
                   prior_indicator=(), prior_baseline=1e-6, prior_weight=0.0,
                   min_module_size=0, noise_threshold=0.025, n_cycles=50,
                   max_optim_iter=10000, tol_coop_rel=1e-8, tol_coop_abs=1e-12,
                   tol_nnls=1e-4, update_orders=None, compute_predictive_r2=False,
                   materialize_array=True, pre=None, keep_history=False, quick_mode=False,
                   quick_mode_percent=0.1, quick_first_cycle_idx=None, allocate_per_obs=True):
    """The alternating loop of ``scregclust()`` (R/scregclust.R:1222-1803) for
    ``allocate_per_obs = TRUE`` (default) or the aggregate branch (``allocate_per_obs=False``, scregclust.R:1678-1728,
    unreachable through the reference's argument checks but part of its source).

    ``update_orders`` -- optional callable ``(penalty_index, cycle) -> permutation``
    standing in for R's ``sample()`` (scregclust.R:1666); default = identity, which
    is result-neutral when ``prior_weight == 0`` (allocation.cpp:39-56).
    ``materialize_array=False`` skips the K x n2 x T array (scregclust.R:1177) and
    streams gene by gene with identical arithmetic.
    ``quick_mode`` (scregclust.R:1435-1496): from cycle two on Step 2 only sees the targets currently in
    a module plus the ``quick_mode_percent`` noise targets best correlated with an active regulator.
    ``quick_first_cycle_idx`` (test hook): treat the first cycle of this call as a later cycle of a longer run whose
    latest allocation (R's ``idx``) is the given vector -- lets a test replay single cycles of a quick-mode trajectory.
    """
    z1r, z2r = z1_reg_scaled, z2_reg_scaled
    z1t, z2t = z1_target_centered, z2_target_centered
    n1, n_reg = z1r.shape
    n2 = z2r.shape[0]
    n_target = z1t.shape[1]
    if pre is None:
        pre = precompute(z1r, z1t)
    res_sds = pre["res_sds"]
    # Quantities the reference recomputes inside every coop_lasso / coef_nnls call
    # (optim.cpp:97-99,414-421); values identical up to summation order.
    x = z1r / math.sqrt(n1)
    G_scaled = compute_xtx(x)
    y_all = (z1t / res_sds[None, :]) / math.sqrt(n1)
    xty_scaled_all = x.T @ y_all
    G_rr = compute_xtx(z1r)
    y_nnls = z1t / z1_target_sds[None, :]
    G_rt_nnls = z1r.T @ y_nnls
    z2sq = z2t * z2t
    ss_train0 = np.sum(z1t * z1t, axis=0)
    ss_test0 = np.sum(z2sq, axis=0)

    if not allocate_per_obs:
        if quick_mode:
            raise ValueError("quick_mode only available when allocate_per_obs is TRUE.")          # scregclust.R:807-817
        materialize_array = False                                                                 # scregclust.R:1177,1521
    sq_arr = alloc_array(z2sq, n_modules) if materialize_array else None
    abs_corr = np.abs(fast_cor(z1t, z1r)) if quick_mode else None          # scregclust.R:1116-1122
    results = []
    k_start = np.asarray(initial_target_modules, dtype=np.int32)

    for li, lam in enumerate(penalization):
        r2_test = np.zeros((n_target, n_modules))
        resid_var = np.full((n_target, n_modules), 1e6)
        last_cycle = False
        converged = False
        k = k_start.copy()
        n_k = 2
        k_history = [k.copy()]
        best_r2 = {}
        fr = FitResult(penalization=float(lam), converged=False, n_cycles_run=0,
                       k_history=k_history, final_cycle_length=1)
        final_cycle_length = 1
        idx_latest = None                                   # R's `idx`: the latest allocation (scregclust.R:1731-1739)
        if quick_first_cycle_idx is not None:
            idx_latest = np.asarray(quick_first_cycle_idx, dtype=np.int32)
        for cycle in range(1, n_cycles + 2):
            if cycle == n_cycles + 1:
                last_cycle = True
            for m in range(1, final_cycle_length + 1):
                k = k_history[len(k_history) - m]
                admm_it = np.zeros(n_modules, dtype=np.int32)
                models, weights, signs = _step1(pre, x, G_scaled, xty_scaled_all, k, lam,
                                                n_modules, tol_coop_rel, tol_coop_abs,
                                                max_optim_iter, admm_it)
                fr.admm_iterations.append(admm_it)
                fr.models_history.append(models.copy())
                # ---- Step 1.5 (scregclust.R:1435-1496): quick-mode target subset ----
                if quick_mode and (cycle != 1 or quick_first_cycle_idx is not None) and np.sum(models) != 0:
                    filt = quick_mode_targets(abs_corr, models, idx_latest, quick_mode_percent, last_cycle)
                else:
                    filt = np.arange(n_target)
                nt = filt.size
                z1t_f, z2t_f, sds_f = z1t[:, filt], z2t[:, filt], z1_target_sds[filt]
                # ---- Step 2a (scregclust.R:1505-1647) ----
                ss_train = np.repeat(np.sum(z1t_f * z1t_f, axis=0)[:, None], n_modules, axis=1)
                ss_test = np.repeat(np.sum(z2t_f * z2t_f, axis=0)[:, None], n_modules, axis=1)
                if not last_cycle and materialize_array:
                    reset_array(sq_arr, z2sq)
                nn_it = np.zeros((n_modules, n_target), dtype=np.int32)
                betas = [None] * n_modules
                for j in range(n_modules):
                    reg_cl = np.flatnonzero(models[:, j])
                    if reg_cl.size == 0:
                        continue
                    sg = signs[reg_cl, j]
                    xtx_j = G_rr[np.ix_(reg_cl, reg_cl)] * np.outer(sg, sg)
                    xty_j = G_rt_nnls[np.ix_(reg_cl, filt)] * sg[:, None]
                    b_nn, its = coef_nnls(eps=tol_nnls, max_iter=max_optim_iter, xtx=xtx_j, xty=xty_j)
                    nn_it[j, filt] = its
                    beta_hat = b_nn * sg[:, None] * _recycle(sds_f, b_nn.shape)      # modulus = #filtered targets
                    full = np.zeros((reg_cl.size, n_target)); full[:, filt] = beta_hat   # scregclust.R:1606-1610
                    betas[j] = full
                    r_train = z1t_f - z1r[:, reg_cl] @ beta_hat
                    r_test = z2t_f - z2r[:, reg_cl] @ beta_hat
                    if not last_cycle and materialize_array:
                        sq_arr[j][:, filt] = r_test * r_test
                    ss_test[:, j] = np.sum(r_test * r_test, axis=0)
                    ss_train[:, j] = np.sum(r_train * r_train, axis=0)
                fr.nnls_iterations.append(nn_it)
                r2_test = np.zeros((n_target, n_modules))                              # scregclust.R:1628-1631
                r2_test[filt] = 1.0 - ss_test / np.sum(z2t_f * z2t_f, axis=0)[:, None]
                best_r2[cycle] = np.max(r2_test, axis=1)
                resid_var = np.full((n_target, n_modules), 1e6)                        # scregclust.R:1635-1638
                resid_var[filt] = ss_train / (n1 - np.sum(models, axis=0))[None, :]
                if last_cycle and nt < n_target:
                    rest = np.setdiff1d(np.arange(n_target), filt)
                    resid_var[rest] = 0.0                                              # scregclust.R:1642
                if keep_history:
                    fr.resid_var_history.append(resid_var.copy()); fr.r2_history.append(r2_test.copy())
                    fr.coeffs_history.append(betas)
                if last_cycle:
                    fr.k_final.append(k.copy()); fr.models.append(models); fr.signs.append(signs)
                    fr.weights.append(weights); fr.coeffs.append(betas)
                    fr.sigmas.append(np.sqrt(resid_var)); fr.r2.append(r2_test)
                    fr.best_r2.append(best_r2[cycle])
                    fr.best_r2_idx.append(np.argmax(r2_test, axis=1) + 1)
            fr.n_cycles_run = cycle
            if last_cycle:
                break
            # ---- Step 2b (scregclust.R:1665-1739) ----
            order = (update_orders(li, cycle) if update_orders is not None
                     else np.arange(n_target))
            amb = np.zeros(n_target, dtype=np.int64)
            seen, distinct = set(), np.zeros(n_modules, dtype=bool)
            for j in range(n_modules):                      # one representative per identical model
                reg_cl = np.flatnonzero(models[:, j])
                key = (tuple(reg_cl), tuple(signs[reg_cl, j]))
                if key not in seen:
                    seen.add(key); distinct[j] = True
            if not allocate_per_obs:
                idx = allocate_aggregate(k, ss_test, resid_var, n2, prior_indicator, order, prior_baseline,
                                         prior_weight, ambiguity=amb)
                votes = None
            elif materialize_array:
                idx, votes = allocate_into_modules(sq_arr, resid_var, prior_indicator, k, order,
                                                   prior_baseline, prior_weight, return_votes=True,
                                                   ambiguity=amb, distinct=distinct)
            else:
                idx, votes = _allocate_streaming(z2t, z2r, models, betas, resid_var,
                                                 prior_indicator, k, order, prior_baseline,
                                                 prior_weight, amb, distinct)
            fr.votes.append(votes)
            fr.ambiguity.append(amb)
            fr.allocations.append(np.asarray(idx, dtype=np.int32).copy())
            idx = idx.copy()
            idx[best_r2[cycle] < noise_threshold] = -1
            module_size = np.array([np.sum(idx == i) for i in range(1, n_modules + 1)])
            small = np.flatnonzero(module_size < min_module_size) + 1
            idx[np.isin(idx, small)] = -1
            k = idx.astype(np.int32)
            idx_latest = k.copy()
            matching = [i for i, kh in enumerate(k_history, start=1) if np.array_equal(kh, k)]
            k_history.append(k.copy())
            if matching:
                converged = True
                mc = matching[0]
                if n_k - mc != 1:
                    final_cycle_length = n_k - mc
                if len(matching) > 1:   # R would error on a length>1 condition; cannot happen
                    pass                # (history entries before the first repeat are distinct)
                last_cycle = True
            n_k += 1
        fr.converged = converged
        fr.final_cycle_length = final_cycle_length
        if compute_predictive_r2:
            _importance(fr, z1r, z2r, z1t, z2t, z1_target_sds, G_rr, G_rt_nnls, n_modules,
                        tol_nnls, max_optim_iter)
        results.append(fr)
    return results


def r_quantile7(x, prob):
    """R's default quantile (type 7), with its back-compatible arithmetic (stats:::quantile.default):
    index = 1 + (n-1) p; qs = x[lo]; where index > lo and x[hi] != qs: qs = (1-h) qs + h x[hi]."""
    x = np.sort(np.asarray(x, dtype=np.float64))
    n = x.size
    if n == 0:
        return float("nan")
    index = 1.0 + max(n - 1, 0) * prob
    lo = int(math.floor(index)); hi = int(math.ceil(index))
    qs = x[lo - 1]
    if index > lo and x[hi - 1] != qs:
        h = index - lo
        qs = (1.0 - h) * qs + h * x[hi - 1]
    return float(qs)


def quick_mode_targets(abs_corr, models, idx, percent, last_cycle):
    """Step 1.5 (scregclust.R:1444-1477): targets currently in a module (``idx != -1``, increasing) followed by the
    noise targets whose largest absolute correlation with an ACTIVE regulator reaches the (1 - percent) quantile of
    those maxima (increasing).  ``idx`` is R's ``idx``: the latest allocation, also while an older configuration of a
    limit cycle is re-fitted in the last cycle."""
    active_reg = np.flatnonzero(models.sum(axis=1) >= 1)
    active_t = np.flatnonzero(idx != -1)
    if last_cycle:
        return active_t
    noise_t = np.setdiff1d(np.arange(abs_corr.shape[0]), active_t)
    if noise_t.size == 0:
        return active_t
    mx = abs_corr[np.ix_(noise_t, active_reg)].max(axis=1)
    thr = r_quantile7(mx, 1.0 - percent)
    return np.concatenate([active_t, noise_t[mx >= thr]])


def _allocate_streaming(z2t, z2r, models, betas, resid_var, prior_indicator, k_, order,
                        prior_baseline, prior_weight, ambiguity=None, distinct=None):
    """allocate_into_modules without the K x n2 x T array: rebuilds each gene's
    K x n2 slab on the fly (same arithmetic as scregclust.R:1593-1599 + allocation.cpp)."""
    n2, n_target = z2t.shape
    n_modules = models.shape[1]
    preds = []
    for j in range(n_modules):
        reg_cl = np.flatnonzero(models[:, j])
        preds.append(None if reg_cl.size == 0 else z2r[:, reg_cl] @ betas[j])

    class _Lazy:
        shape = (n_modules, n2, n_target)

        def __getitem__(self, key):
            _, _, j = key
            out = np.empty((n_modules, n2))
            for c in range(n_modules):
                r = z2t[:, j] if preds[c] is None else z2t[:, j] - preds[c][:, j]
                out[c] = r * r
            return out
    return allocate_into_modules(_Lazy(), resid_var, prior_indicator, k_, order,
                                 prior_baseline, prior_weight, return_votes=True, ambiguity=ambiguity,
                                 distinct=distinct)


def _importance(fr, z1r, z2r, z1t, z2t, z1_target_sds, G_rr, G_rt_nnls, n_modules,
                tol_nnls, max_iter):
    """Leave-one-regulator-out importance and module R2 (R/scregclust.R:1823-1912)."""
    n_reg = z1r.shape[1]
    for m in range(fr.final_cycle_length):
        k = fr.k_final[m]
        r2_module = np.full(n_modules, np.nan)
        importance = np.full((n_reg, n_modules), np.nan)
        removed = [None] * n_modules
        for j in range(n_modules):
            reg_cl = np.flatnonzero(fr.models[m][:, j])
            target_cl = np.flatnonzero(k == j + 1)
            if target_cl.size > 0 and reg_cl.size > 1:
                ssq = np.full((target_cl.size, reg_cl.size + 1), np.nan)
                for r in range(reg_cl.size + 1):
                    reg_mod = reg_cl if r == 0 else np.delete(reg_cl, r - 1)
                    sg = fr.signs[m][reg_mod, j]
                    xtx_j = G_rr[np.ix_(reg_mod, reg_mod)] * np.outer(sg, sg)
                    xty_j = G_rt_nnls[np.ix_(reg_mod, target_cl)] * sg[:, None]
                    b, _ = coef_nnls(eps=tol_nnls, max_iter=max_iter, xtx=xtx_j, xty=xty_j)
                    bh = b * sg[:, None] * _recycle(z1_target_sds[target_cl], b.shape)
                    rt = z2t[:, target_cl] - z2r[:, reg_mod] @ bh
                    ssq[:, r] = np.sum(rt * rt, axis=0)
                zc = z2t[:, target_cl] - z2t[:, target_cl].mean(axis=0)
                r2_removed = 1.0 - ssq.sum(axis=0) / np.sum(zc * zc)
                # scregclust.R:1890-1896: 1 - t(t(ssq) / colSums(zc^2)) -- the length-m vector is recycled
                # down the columns of the (s+1) x m matrix t(ssq), not matched to its columns
                tss = ssq.T
                removed[j] = 1.0 - (tss / _recycle(np.sum(zc * zc, axis=0), tss.shape)).T
                r2_module[j] = r2_removed[0]
                importance[reg_cl, j] = 1.0 - r2_removed[1:] / r2_module[j]
        fr.r2_module.append(r2_module)
        fr.importance.append(importance)
        fr.r2_removed.append(removed)
