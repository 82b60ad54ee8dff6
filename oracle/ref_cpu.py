"""ctypes loader for oracle/_ref/liboracle_ref.so (C++ "reference-faithful CPU" twin).

TEST / BASELINE INFRASTRUCTURE ONLY -- see oracle/ref_cpu.cpp.  Used to cross-check the
NumPy oracle and as the timed CPU baseline in bench.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_ref.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ref_set_threads.restype = C.c_int
        lib.ref_set_threads.argtypes = [C.c_int]
        lib.ref_coop_lasso.restype = C.c_int
        lib.ref_coop_lasso.argtypes = [_dp, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_double, _dp, _dp, C.c_double,
                                       C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, _dp, _ip, _ip]
        lib.ref_coef_nnls.restype = C.c_int
        lib.ref_coef_nnls.argtypes = [_dp, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_double, C.c_int, _dp, _ip]
        lib.ref_fill_array.restype = None
        lib.ref_fill_array.argtypes = [_dp, _dp, C.c_int64, C.c_int64, C.c_int64]
        lib.ref_allocate_into_modules.restype = C.c_int
        lib.ref_allocate_into_modules.argtypes = [_dp, C.c_int, C.c_int64, C.c_int64, _dp, _ip, _ip, _ip, _ip,
                                                  C.c_double, C.c_double, _ip]
        lib.ref_dgemm_nn.restype = None
        lib.ref_dgemm_nn.argtypes = [_dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int64]
        _lib = lib
    return _lib


def _f(a):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F_CONTIGUOUS", "ALIGNED"])


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def set_threads(n: int) -> int:
    return load().ref_set_threads(int(n))


def coop_lasso(y, x, lam, weights, beta_0=None, rho_0=0.2, alpha_0=1.5, n_update=2, eps_corr=0.2, max_iter=1000,
               eps_rel=1e-8, eps_abs=1e-12):
    y = _f(y); x = _f(x); w = np.ascontiguousarray(weights, dtype=np.float64)
    n, m = y.shape; p = x.shape[1]
    b0 = None if beta_0 is None else _f(beta_0)
    beta = np.zeros((p, m), order="F")
    it = C.c_int(0); nf = C.c_int(0)
    rc = load().ref_coop_lasso(_d(y), _d(x), n, m, p, float(lam), _d(w), _d(b0), rho_0, alpha_0, n_update, eps_corr,
                               max_iter, eps_rel, eps_abs, _d(beta), C.byref(it), C.byref(nf))
    if rc != 0:
        raise ValueError(f"ref_coop_lasso failed with {rc}")
    return beta, it.value, nf.value


def coef_nnls(x, y, eps=1e-12, max_iter=1000):
    x = _f(x); y = _f(y)
    nobs, n = x.shape; m = y.shape[1]
    beta = np.zeros((n, m), order="F"); its = np.zeros(m, dtype=np.int32)
    rc = load().ref_coef_nnls(_d(x), nobs, n, _d(y), m, float(eps), int(max_iter), _d(beta), _i(its))
    if rc != 0:
        raise ValueError(f"ref_coef_nnls failed with {rc}")
    return beta, its


def fill_array(arr, inp):
    load().ref_fill_array(_d(arr), _d(_f(inp)), arr.shape[0], arr.shape[1], arr.shape[2])


def allocate_into_modules(arr, resid_var, prior_off, prior_idx, k_in, order, baseline, weight):
    arr = _f(arr); rv = _f(resid_var)
    K, n_obs, T = arr.shape
    k_in = np.ascontiguousarray(k_in, dtype=np.int32); order = np.ascontiguousarray(order, dtype=np.int32)
    out = np.zeros(T, dtype=np.int32)
    load().ref_allocate_into_modules(_d(arr), K, n_obs, T, _d(rv), _i(prior_off), _i(prior_idx), _i(k_in), _i(order),
                                     float(baseline), float(weight), _i(out))
    return out
