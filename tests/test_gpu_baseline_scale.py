"""Parity at the BASELINE.json sizes: the CUDA path against the C++ twin of the reference
(oracle/ref_cpu.cpp through oracle/twin_cycle.py -- restated Eigen::LDLT per ADMM iteration,
src/optim.cpp:136-143, Gram matrices recomputed per call), where the NumPy oracle is too slow.

What is compared, per (penalization value, module) of a first cycle from the initial configuration:
ADMM iteration counts (bit-exact; this is the validation of the eigenbasis solve against LDL^T at
P = 500 and P = 1600 that SURVEY.md section 7 asks for), selected-regulator sets and signs (bit-exact),
NNLS iteration counts of every (module, target) (bit-exact), R2 (absolute 1e-8: the quantity lives in
[.., 1]) and residual variances / coefficients within 1e-8 relative ELEMENT-WISE (coefficients smaller than
1e-3 of the largest one of their matrix are held to 1e-8 of that floor instead: they are differences of
O(1) terms and carry the absolute error of those terms).
"""
import json
import os

import numpy as np
import pytest

from oracle import ref_cpu
from oracle import scregclust_oracle as orc
from oracle import twin_cycle
from scregclust_b200.synthetic import make_workload

pytestmark = pytest.mark.gpu


def elem_rel_err(a, b, floor_frac=1e-3):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    floor = floor_frac * max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def _record(entry):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as fh:
            fh.write(json.dumps(entry) + "\n")
    except OSError:
        pass


def _compare_cycle(tag, w, lams, modules=None):
    from scregclust_b200.driver import Session
    ref_cpu.set_threads(os.cpu_count() or 1)
    K, T = w.n_modules, w.z1_target.shape[1]
    pre = orc.precompute(w.z1_reg, w.z1_target)
    with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, K) as s:
        res = s.fit_cycle(lams, np.stack([w.k_init] * len(lams)), want_nnls_iters=True, skip_allocation=True)
        s._last_nsel = res["n_selected"]
        worst = dict(r2_abs=0.0, resid_var_rel=0.0, coeff_rel=0.0, weights_rel=0.0)
        for f, lam in enumerate(lams):
            tw = twin_cycle.cycle(w, lam, w.k_init, pre=pre, modules=modules)
            done = np.array(tw["done"]) - 1
            t = f"{tag} lambda={lam:.4g}"
            assert np.array_equal(res["admm_iterations"][f][done], tw["admm_iterations"][done]), \
                f"{t}: ADMM iterations gpu={res['admm_iterations'][f][done]} twin={tw['admm_iterations'][done]}"
            assert np.array_equal(res["models"][f].T.astype(bool)[:, done], tw["models"][:, done]), f"{t}: selected sets differ"
            sg, st = res["signs"][f].T[:, done], tw["signs"][:, done]
            assert np.array_equal(np.isnan(sg), np.isnan(st)) and np.array_equal(np.nan_to_num(sg), np.nan_to_num(st)), f"{t}: signs"
            assert np.array_equal(res["nnls_iterations"][f][done], tw["nnls_iterations"][done]), f"{t}: NNLS iterations differ"
            worst["weights_rel"] = max(worst["weights_rel"], elem_rel_err(res["weights"][f].T[:, done], tw["weights"][:, done]))
            worst["r2_abs"] = max(worst["r2_abs"], float(np.max(np.abs(res["r2_test"][f].T[:, done] - tw["r2"][:, done]))))
            worst["resid_var_rel"] = max(worst["resid_var_rel"],
                                         float(np.max(np.abs(res["resid_var"][f].T[:, done] / tw["resid_var"][:, done] - 1.0))))
            cos, _ = s.coeffs_fit(f)
            for j in done:
                assert (cos[j] is None) == (tw["coeffs"][j] is None), t
                if cos[j] is not None:
                    assert cos[j].shape == tw["coeffs"][j].shape, t
                    worst["coeff_rel"] = max(worst["coeff_rel"], elem_rel_err(cos[j], tw["coeffs"][j]))
            print(f"{t}: ADMM iterations {tw['admm_iterations'][done]} ({tw['factorizations']} LDL^T factorisations on the twin), "
                  f"selected {tw['models'][:, done].sum(axis=0)}")
    print(f"{tag}: worst element-wise errors {worst}")
    _record(dict(kind="baseline_scale_cycle", case=tag, penalties=[float(x) for x in lams],
                 modules="all" if modules is None else list(modules), **worst))
    assert worst["r2_abs"] < 1e-8 and worst["resid_var_rel"] < 1e-8 and worst["coeff_rel"] < 1e-8 and worst["weights_rel"] < 1e-8, worst


def test_config2_first_cycle_matches_cpp_twin():
    """BASELINE configs[1] (5k cells x 2k targets x 500 regulators, 10 modules): smallest, middle and largest
    penalization value of the grid, every module."""
    w = make_workload("config2", seed=0)
    _compare_cycle("config2", w, [float(w.penalization[i]) for i in (0, 4, 9)])


def test_config2_correlated_regulators_match_cpp_twin():
    """Same shape with AR(1) = 0.9 regulators: cond(X'X) ~ 360, the hard case for eigenbasis-vs-LDL^T counts."""
    w = make_workload("config2", seed=1, ar1=0.9)
    _compare_cycle("config2/ar1=0.9", w, [float(w.penalization[i]) for i in (1, 8)])


def test_p1600_eigenbasis_matches_ldlt_iteration_counts():
    """P = 1600 regulators (the humanTFs size of BASELINE configs[2], [3]) on fewer cells / targets so that the
    twin's unblocked LDL^T (1.4 GFLOP per ADMM iteration, one core) finishes in about a minute."""
    w = make_workload(shape=(4000, 600, 1600, 3), penalization=[0.1, 0.3], seed=2)
    _compare_cycle("P1600", w, [0.1, 0.3])
