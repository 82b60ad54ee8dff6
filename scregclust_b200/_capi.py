"""ctypes binding of libscregclust_b200.so (include/scregclust_b200.h).

This is the Python stand-in for the thin Rcpp bridge: it passes
the same column-major FP64 buffers and 1-based module ids R would pass.  There
is no fallback: if the shared library is missing or no CUDA device is usable,
every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libscregclust_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)


class ScrcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[scrc {code}] {msg}")
        self.code = code
        self.message = msg


class Options(C.Structure):
    _fields_ = [("max_optim_iter", C.c_int), ("tol_coop_rel", C.c_double), ("tol_coop_abs", C.c_double),
                ("tol_nnls", C.c_double), ("rho0", C.c_double), ("alpha0", C.c_double),
                ("n_update", C.c_int), ("eps_corr", C.c_double)]


class CycleOut(C.Structure):
    _fields_ = [("k_new", c_int_p), ("votes", c_int_p), ("models", c_ubyte_p), ("weights", c_double_p),
                ("signs", c_double_p), ("r2_test", c_double_p), ("resid_var", c_double_p),
                ("best_r2", c_double_p), ("best_r2_idx", c_int_p), ("admm_iterations", c_int_p),
                ("nnls_iterations", c_int_p), ("n_selected", c_int_p), ("admm_sweeps", C.POINTER(C.c_longlong))]


class Prior(C.Structure):
    _fields_ = [("prior_off", c_int_p), ("prior_idx", c_int_p), ("baseline", C.c_double), ("weight", C.c_double),
                ("update_order", c_int_p), ("aggregate", C.c_int)]


class Quick(C.Structure):
    _fields_ = [("idx_latest", c_int_p), ("percent", C.c_double)]


UPDATE_ORDER_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_int)   # returns the address of T ints


class GridParams(C.Structure):
    _fields_ = [("n_penalties", C.c_int), ("penalization", c_double_p), ("initial_target_modules", c_int_p),
                ("n_cycles", C.c_int), ("noise_threshold", C.c_double), ("min_module_size", C.c_int),
                ("opt", Options), ("compute_predictive_r2", C.c_int), ("keep_coeffs", C.c_int),
                ("keep_diagnostics", C.c_int), ("prior_off", c_int_p), ("prior_idx", c_int_p),
                ("prior_baseline", C.c_double), ("prior_weight", C.c_double), ("update_order", UPDATE_ORDER_FN),
                ("update_order_user", C.c_void_p), ("quick_mode", C.c_int), ("quick_mode_percent", C.c_double),
                ("allocate_per_obs", C.c_int), ("cancel", C.POINTER(C.c_int))]


class GridInfo(C.Structure):
    _fields_ = [("n_penalties", C.c_int), ("P", C.c_int64), ("T", C.c_int64), ("K", C.c_int),
                ("admm_sweeps", C.c_longlong), ("admm_problem_iterations", C.c_longlong), ("fit_cycles", C.c_int),
                ("final_refits", C.c_int), ("device_calls", C.c_int)]


class GridPenaltyInfo(C.Structure):
    _fields_ = [("penalization", C.c_double), ("converged", C.c_int), ("n_cycles_run", C.c_int),
                ("final_cycle_length", C.c_int), ("n_history", C.c_int), ("n_records", C.c_int)]


class GridFinal(C.Structure):
    _fields_ = [("k_final", c_int_p), ("models", c_ubyte_p), ("weights", c_double_p), ("signs", c_double_p),
                ("r2_test", c_double_p), ("resid_var", c_double_p), ("best_r2", c_double_p), ("best_r2_idx", c_int_p),
                ("n_selected", c_int_p), ("r2_module", c_double_p), ("importance", c_double_p)]


BACKEND_FIT_CYCLE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_double_p, c_int_p, C.POINTER(Options), C.c_int, c_int_p,
                                   C.POINTER(Prior), C.POINTER(Quick), C.POINTER(CycleOut))
BACKEND_IMPORTANCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_int_p, c_ubyte_p, c_double_p, C.POINTER(Options),
                                    c_double_p, c_double_p, c_double_p, C.c_int64)
BACKEND_COEFFS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_double_p, c_int_p)


class GridBackend(C.Structure):
    _fields_ = [("user", C.c_void_p), ("P", C.c_int64), ("T", C.c_int64), ("K", C.c_int),
                ("fit_cycle", BACKEND_FIT_CYCLE_FN), ("importance", BACKEND_IMPORTANCE_FN),
                ("coeffs_fit", BACKEND_COEFFS_FN)]


# name -> (restype, argtypes); also the list of symbols the header declares
SIGNATURES = {
    "scrc_version": (C.c_int, []),
    "scrc_last_error": (C.c_char_p, []),
    "scrc_launch_count": (C.c_ulonglong, []),
    "scrc_set_device": (C.c_int, [C.c_int]),
    "scrc_default_options": (None, [C.POINTER(Options)]),
    "scrc_coop_lasso": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_double_p, C.c_int64, C.c_int64, C.c_double,
                                  c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int,
                                  C.c_double, C.c_int, C.c_double, C.c_double, c_double_p, c_int_p]),
    "scrc_coef_nnls": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_double_p, C.c_int64, C.c_int64, C.c_double,
                                 C.c_int, c_double_p, c_int_p]),
    "scrc_allocate_into_modules": (C.c_int, [c_double_p, C.c_int, C.c_int64, C.c_int64, c_double_p, c_int_p, c_int_p,
                                             c_int_p, c_int_p, C.c_double, C.c_double, c_int_p, c_int_p]),
    "scrc_alloc_array": (C.c_int, [c_double_p, C.c_int64, C.c_int64, C.c_int64, c_double_p]),
    "scrc_reset_array": (C.c_int, [c_double_p, C.c_int64, C.c_int64, C.c_int64, c_double_p, C.c_int64, C.c_int64]),
    "scrc_fast_cor": (C.c_int, [c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int64, c_double_p]),
    "scrc_coef_ols": (C.c_int, [c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int64, c_double_p]),
    "scrc_coef_ridge": (C.c_int, [c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_double_p]),
    "scrc_kmeanspp": (C.c_int, [c_double_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, c_int_p, c_double_p, c_int_p,
                                c_double_p, c_int_p]),
    "scrc_ctx_kmeanspp": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p, c_double_p, c_int_p, c_double_p, c_int_p]),
    "scrc_crossprod": (C.c_int, [c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int64, c_double_p]),
    "scrc_bench_gemm": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, c_double_p]),
    "scrc_profile_enable": (C.c_int, [C.c_int]),
    "scrc_profile_reset": (C.c_int, []),
    "scrc_profile_get": (C.c_int, [c_double_p, c_double_p, C.POINTER(C.c_ulonglong)]),
    "scrc_timer_begin": (C.c_int, []),
    "scrc_timer_end": (C.c_int, [c_double_p]),
    "scrc_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                                  c_double_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "scrc_ctx_create_raw": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_double_p, C.c_int64, C.c_int64, c_int_p, c_int_p,
                                      c_int_p, C.c_int, C.c_int, c_ubyte_p, C.POINTER(C.c_int64)]),
    "scrc_ctx_get_inputs": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "scrc_ctx_destroy": (None, [C.c_void_p]),
    "scrc_ctx_get_init": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "scrc_ctx_cross_corr": (C.c_int, [C.c_void_p, c_double_p]),
    "scrc_ctx_get_gram": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "scrc_fit_cycle": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p, C.POINTER(Options), C.c_int,
                                 C.POINTER(CycleOut)]),
    "scrc_fit_cycle_prior": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p, C.POINTER(Options), C.POINTER(Prior),
                                       C.POINTER(CycleOut)]),
    "scrc_fit_cycle_quick": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p, C.POINTER(Options), C.c_int,
                                       C.POINTER(Quick), C.POINTER(CycleOut)]),
    "scrc_importance": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_ubyte_p, c_double_p, C.POINTER(Options), c_double_p,
                                  c_double_p]),
    "scrc_importance_ex": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_ubyte_p, c_double_p, C.POINTER(Options), c_double_p,
                                     c_double_p, c_double_p, C.c_int64]),
    "scrc_ctx_get_coeffs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_int_p]),
    "scrc_ctx_get_coeffs_fit": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p]),
    "scrc_ctx_request_cancel": (C.c_int, [C.c_void_p]),
    "scrc_ctx_clear_cancel": (C.c_int, [C.c_void_p]),
    "scrc_comm_unique_id": (C.c_int, [C.c_char_p]),
    "scrc_comm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "scrc_comm_destroy": (None, [C.c_void_p]),
    "scrc_ctx_create_sharded": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, c_double_p, c_double_p, c_double_p,
                                          c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "scrc_ctx_set_cost_hint": (C.c_int, [C.c_void_p, c_int_p, C.c_int]),
    "scrc_shard_assign": (C.c_int, [c_double_p, C.c_int, C.c_int, c_int_p]),
    "scrc_shard_slice": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "scrc_grid_default_params": (None, [C.POINTER(GridParams)]),
    "scrc_grid_fit": (C.c_int, [C.c_void_p, C.POINTER(GridParams), C.POINTER(C.c_void_p)]),
    "scrc_grid_fit_multi": (C.c_int, [c_int_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(GridParams),
                                      C.POINTER(C.c_void_p)]),
    "scrc_grid_fit_custom": (C.c_int, [C.POINTER(GridBackend), C.POINTER(GridParams), C.POINTER(C.c_void_p)]),
    "scrc_grid_result_info": (C.c_int, [C.c_void_p, C.POINTER(GridInfo)]),
    "scrc_grid_result_penalty": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GridPenaltyInfo)]),
    "scrc_grid_result_history": (C.c_int, [C.c_void_p, C.c_int, c_int_p]),
    "scrc_grid_result_record": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p, c_int_p, c_ubyte_p, c_int_p, c_int_p]),
    "scrc_grid_result_records": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p, c_ubyte_p]),
    "scrc_grid_result_final": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(GridFinal)]),
    "scrc_grid_result_coeffs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_int_p]),
    "scrc_grid_result_r2_removed_size": (C.c_int64, [C.c_void_p, C.c_int, C.c_int]),
    "scrc_grid_result_r2_removed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "scrc_grid_result_free": (None, [C.c_void_p]),
}

_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScrcError(-1000, f"{LIB_PATH} is missing: build it with `make -C scregclust_b200/csrc` "
                               "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise ScrcError(code, load().scrc_last_error().decode("utf-8", "replace"))


def f64(a, order="F"):
    """Column-major float64 view/copy, the layout R hands to .Call."""
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F_CONTIGUOUS", "ALIGNED"]) if order == "F" \
        else np.ascontiguousarray(a, dtype=np.float64)


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None
