"""Host-side mirror of the reference's operator interface (same names, argument
meaning and error behaviour as R/RcppExports.R:4-117 and R/utils.R:13-57,533-541),
routed through the C ABI of libscregclust_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import ScrcError, check, dptr, f64, iptr


def coop_lasso(y, x, lambda_, weights, beta_0=None, rho_0=0.2, alpha_0=1.5, n_update=2, eps_corr=0.2,
               max_iter=1000, eps_rel=1e-8, eps_abs=1e-12, verbose=False):
    """``coop_lasso()`` of R/RcppExports.R:66-68 -> ``{"beta": p x m, "iterations": int}``."""
    lib = _capi.load()
    y = f64(np.atleast_2d(y)); x = f64(np.atleast_2d(x))
    n, m = y.shape
    nx, p = x.shape
    w = np.ascontiguousarray(weights, dtype=np.float64)
    b0 = None if beta_0 is None else f64(np.atleast_2d(beta_0))
    if w.size != p and n > 0 and m > 0 and p > 0 and nx == n:
        raise ScrcError(-20, "weights needs to have one entry per column of x")
    beta = np.zeros((max(p, 0), max(m, 0)), order="F")
    its = C.c_int(0)
    check(lib.scrc_coop_lasso(dptr(y), n, m, dptr(x), nx, p, float(lambda_), dptr(w), dptr(b0),
                              0 if b0 is None else b0.shape[0], 0 if b0 is None else b0.shape[1],
                              float(rho_0), float(alpha_0), int(n_update), float(eps_corr), int(max_iter),
                              float(eps_rel), float(eps_abs), dptr(beta), C.byref(its)))
    if its.value > max_iter:
        # optim.cpp:300 prints this notice on Rcout
        if verbose:
            print("Coop-Lasso: Maximum number of iterations reached")
    return {"beta": beta, "iterations": int(its.value)}


def coef_nnls(x, y, eps=1e-12, max_iter=1000):
    """``coef_nnls()`` of R/RcppExports.R:93-95 -> ``{"beta": nvar x m, "iterations": int[m]}``."""
    lib = _capi.load()
    x = f64(np.atleast_2d(x)); y = f64(np.atleast_2d(y))
    nobs, nvar = x.shape
    nobs_y, m = y.shape
    beta = np.zeros((max(nvar, 0), max(m, 0)), order="F")
    its = np.zeros(max(m, 1), dtype=np.int32)
    check(lib.scrc_coef_nnls(dptr(x), nobs, nvar, dptr(y), nobs_y, m, float(eps), int(max_iter), dptr(beta), iptr(its)))
    return {"beta": beta, "iterations": its[:m]}


def alloc_array(input_, n_cl):
    """``alloc_array()`` of R/RcppExports.R:104-106: K x n_obs x n_genes broadcast."""
    lib = _capi.load()
    inp = f64(input_)
    n_obs, n_genes = inp.shape
    out = np.empty((int(n_cl), n_obs, n_genes), order="F")
    check(lib.scrc_alloc_array(dptr(inp), n_obs, n_genes, int(n_cl), dptr(out)))
    return out


def reset_array(arr, input_):
    """``reset_array()`` of R/RcppExports.R:115-117: refills ``arr`` in place."""
    lib = _capi.load()
    if not (isinstance(arr, np.ndarray) and arr.flags.f_contiguous and arr.dtype == np.float64 and arr.ndim == 3):
        raise ScrcError(-20, "reset_array: arr must be a 3-d column-major float64 array")
    inp = f64(input_)
    check(lib.scrc_reset_array(dptr(arr), arr.shape[0], arr.shape[1], arr.shape[2], dptr(inp), inp.shape[0], inp.shape[1]))


def _csr(prior_indicator, n):
    if prior_indicator is None or len(prior_indicator) == 0:
        return None, None
    off = np.zeros(n + 1, dtype=np.int32)
    for i in range(n):
        off[i + 1] = off[i] + len(prior_indicator[i])
    idx = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
    for i in range(n):
        idx[off[i]:off[i + 1]] = np.asarray(prior_indicator[i], dtype=np.int32)
    return off, idx


def allocate_into_modules(resid_array, resid_var, prior_indicator, k_, update_order, prior_baseline, prior_weight,
                          return_votes=False):
    """``allocate_into_modules()`` of R/RcppExports.R:4-6 -> integer[T] (1-based)."""
    lib = _capi.load()
    arr = f64(resid_array)
    K, n2, T = arr.shape
    rv = f64(resid_var)
    k_in = np.ascontiguousarray(k_, dtype=np.int32)
    order = np.ascontiguousarray(update_order, dtype=np.int32)
    off, idx = _csr(prior_indicator, T)
    k_out = np.zeros(T, dtype=np.int32)
    votes = np.zeros((T, K), dtype=np.int32) if return_votes else None
    check(lib.scrc_allocate_into_modules(dptr(arr), K, n2, T, dptr(rv), iptr(off), iptr(idx), iptr(k_in), iptr(order),
                                         float(prior_baseline), float(prior_weight), iptr(k_out), iptr(votes)))
    return (k_out, votes) if return_votes else k_out


def fast_cor(x, y):
    """``fast_cor()`` of R/utils.R:533-541."""
    lib = _capi.load()
    x = f64(np.atleast_2d(x)); y = f64(np.atleast_2d(y))
    if x.shape[0] != y.shape[0]:
        raise ScrcError(-20, "fast_cor: x and y need the same number of rows")
    out = np.zeros((x.shape[1], y.shape[1]), order="F")
    check(lib.scrc_fast_cor(dptr(x), dptr(y), x.shape[0], x.shape[1], y.shape[1], dptr(out)))
    return out


def coef_ols(y, x):
    """``coef_ols()`` of R/utils.R:13-37."""
    lib = _capi.load()
    y = f64(np.atleast_2d(y)); x = f64(np.atleast_2d(x))
    out = np.zeros((x.shape[1], y.shape[1]), order="F")
    check(lib.scrc_coef_ols(dptr(y), dptr(x), x.shape[0], x.shape[1], y.shape[1], dptr(out)))
    return out


def coef_ridge(y, x, lambda_):
    """``coef_ridge()`` of R/utils.R:49-57."""
    lib = _capi.load()
    y = f64(np.atleast_2d(y)); x = f64(np.atleast_2d(x))
    out = np.zeros((x.shape[1], y.shape[1]), order="F")
    check(lib.scrc_coef_ridge(dptr(y), dptr(x), x.shape[0], x.shape[1], y.shape[1], float(lambda_), dptr(out)))
    return out


def kmeanspp(x, n_cluster, first_centers, uniforms, n_max_iter=10, return_details=False):
    """``kmeanspp(x, n_cluster, n_init_clusterings, n_max_iter)`` of R/utils.R:351-375 -> cluster labels (1-based).

    R's RNG stays with the caller: ``first_centers[i]`` is ``sample(n, 1L)`` (1-based) and ``uniforms[i, :]`` the
    ``n_cluster - 1`` ``unif_rand()`` draws of initialisation ``i``; the number of initialisations is their length."""
    lib = _capi.load()
    x = f64(np.atleast_2d(x))
    n, p = x.shape
    fc = np.ascontiguousarray(first_centers, dtype=np.int32)
    u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(fc.size, max(int(n_cluster) - 1, 0))
    cl = np.zeros(n, dtype=np.int32)
    tot = np.zeros(fc.size)
    fault = np.zeros(fc.size, dtype=np.int32)
    check(lib.scrc_kmeanspp(dptr(x), n, p, int(n_cluster), int(fc.size), int(n_max_iter), iptr(fc), dptr(u), iptr(cl),
                            dptr(tot), iptr(fault)))
    return (cl, tot, fault) if return_details else cl


def crossprod(x, y=None):
    """R's ``crossprod(x, y)`` / ``compute_xtx`` (src/optim.cpp:17-27) on the FP64 tensor pipe."""
    lib = _capi.load()
    x = f64(np.atleast_2d(x))
    yy = x if y is None else f64(np.atleast_2d(y))
    out = np.zeros((x.shape[1], yy.shape[1]), order="F")
    check(lib.scrc_crossprod(dptr(x), dptr(yy), x.shape[0], x.shape[1], yy.shape[1], dptr(out)))
    return out


def launch_count() -> int:
    return int(_capi.load().scrc_launch_count())


def set_device(device: int) -> None:
    check(_capi.load().scrc_set_device(int(device)))
