"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle.

Tolerances: discrete outputs (iteration counts, module ids, votes, selected
regulators) bit-exact; floating point 1e-8 relative (BASELINE.json north_star),
Gram / correlation 1e-12 relative.
"""
import numpy as np
import pytest

from conftest import rel_err
from oracle import scregclust_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from scregclust_b200 import api as a
    return a


@pytest.mark.parametrize("n,px,py", [(64, 8, 8), (200, 50, 10), (1000, 130, 257), (2500, 500, 300), (333, 77, 1),
                                     (17, 3, 5), (4096, 256, 1024)])
def test_crossprod(api, rng, n, px, py):
    x = rng.standard_normal((n, px)); y = rng.standard_normal((n, py))
    got = api.crossprod(x, y)
    assert rel_err(got, x.T @ y) < 1e-12
    g = api.crossprod(x)
    assert rel_err(g, x.T @ x) < 1e-12
    assert np.array_equal(g, g.T)          # mirrored like optim.cpp:22-23


def test_crossprod_linearity(api, rng):
    # size-independent property at a BASELINE-scale contraction length
    n, p = 10000, 192
    x = rng.standard_normal((n, p)); y1 = rng.standard_normal((n, 64)); y2 = rng.standard_normal((n, 64))
    a = api.crossprod(x, y1 + 2.0 * y2)
    b = api.crossprod(x, y1) + 2.0 * api.crossprod(x, y2)
    assert rel_err(a, b) < 1e-12


@pytest.mark.parametrize("n,px,py", [(200, 50, 10), (1001, 33, 129)])
def test_fast_cor(api, rng, n, px, py):
    # tests/testthat/test-fast-cor.R: fast_cor == stats::cor
    zt = rng.standard_normal((n, px)) + 3.0; zr = rng.standard_normal((n, py)) - 1.0
    ref = np.corrcoef(zt.T, zr.T)[:px, px:]
    got = api.fast_cor(zt, zr)
    assert rel_err(got, ref) < 1e-12
    assert rel_err(got, orc.fast_cor(zt, zr)) < 1e-12
    assert np.all(np.abs(got) <= 1.0)


def test_fast_cor_self_is_one(api, rng):
    z = rng.standard_normal((500, 20))
    got = api.fast_cor(z, z)
    assert np.allclose(np.diag(got), 1.0, atol=1e-14)
    assert np.all(got <= 1.0)


def test_coef_ols_ridge(api, rng):
    n, p, m = 600, 40, 25
    x = rng.standard_normal((n, p)); y = x @ rng.standard_normal((p, m)) + rng.standard_normal((n, m))
    assert rel_err(api.coef_ols(y, x), orc.coef_ols(y, x)) < 1e-9
    assert rel_err(api.coef_ridge(y, x, 3.5), orc.coef_ridge(y, x, 3.5)) < 1e-9
    # n < p: Moore-Penrose branch (R/utils.R:20-30)
    x2 = rng.standard_normal((30, 50)); y2 = rng.standard_normal((30, 4))
    assert rel_err(api.coef_ols(y2, x2), orc.coef_ols(y2, x2)) < 1e-8


def _coop_case(rng, n, p, m, lam):
    x = rng.standard_normal((n, p)) / np.sqrt(n)
    b = np.zeros((p, m)); b[: max(2, p // 5)] = rng.standard_normal((max(2, p // 5), 1)) + 0.2 * rng.standard_normal((max(2, p // 5), m))
    y = (x * np.sqrt(n)) @ b / np.sqrt(n) + rng.standard_normal((n, m)) / np.sqrt(n)
    w = np.full(p, np.sqrt(m)) * (0.5 + rng.random(p))
    return x, y, w, lam


@pytest.mark.parametrize("n,p,m,lam", [(200, 10, 5, 0.1), (1000, 100, 80, 0.2), (2500, 300, 200, 0.05), (300, 20, 1, 0.3)])
def test_coop_lasso(api, rng, n, p, m, lam):
    x, y, w, lam = _coop_case(rng, n, p, m, lam)
    ref_b, ref_it = orc.coop_lasso(y, x, lam, w, max_iter=10000)
    got = api.coop_lasso(y, x, lam, w, max_iter=10000)
    assert got["iterations"] == ref_it
    assert rel_err(got["beta"], ref_b) < 1e-8
    # selected set through the driver's 1e-6 threshold (R/scregclust.R:1382)
    assert np.array_equal(np.abs(got["beta"]) >= 1e-6, np.abs(ref_b) >= 1e-6)


def test_coop_lasso_warm_start_and_maxiter(api, rng):
    x, y, w, lam = _coop_case(rng, 400, 30, 12, 0.15)
    ref_b, ref_it = orc.coop_lasso(y, x, lam, w, max_iter=10000)
    b0 = ref_b + 1e-3 * rng.standard_normal(ref_b.shape)
    r2_b, r2_it = orc.coop_lasso(y, x, lam, w, beta_0=b0, max_iter=10000)
    got = api.coop_lasso(y, x, lam, w, beta_0=b0, max_iter=10000)
    assert got["iterations"] == r2_it and rel_err(got["beta"], r2_b) < 1e-8
    r3_b, r3_it = orc.coop_lasso(y, x, lam, w, max_iter=5)
    got = api.coop_lasso(y, x, lam, w, max_iter=5)
    assert r3_it == 6 and got["iterations"] == 6          # optim.cpp:295-305
    assert rel_err(got["beta"], r3_b) < 1e-8


def test_coop_lasso_errors(api):
    from scregclust_b200._capi import ScrcError
    y = np.zeros((4, 2)); x = np.zeros((5, 3)); w = np.ones(3)
    with pytest.raises(ScrcError, match="same number of rows"):
        api.coop_lasso(y, x, 0.1, w)
    x = np.ones((4, 3))
    with pytest.raises(ScrcError, match="rho_0 > 0"):
        api.coop_lasso(y, x, 0.1, w, rho_0=0.0)
    with pytest.raises(ScrcError, match="alpha_0"):
        api.coop_lasso(y, x, 0.1, w, alpha_0=2.5)
    with pytest.raises(ScrcError, match="beta_0 needs"):
        api.coop_lasso(y, x, 0.1, w, beta_0=np.zeros((2, 2)))


@pytest.mark.parametrize("n,s,m", [(300, 1, 7), (500, 4, 33), (800, 13, 100), (1200, 64, 150), (900, 130, 40)])
def test_coef_nnls(api, rng, n, s, m):
    x = rng.standard_normal((n, s))
    b = np.maximum(rng.standard_normal((s, m)), 0.0)
    y = x @ b + 0.5 * rng.standard_normal((n, m))
    for eps in (1e-4, 1e-10):
        ref_b, ref_it = orc.coef_nnls(x, y, eps=eps, max_iter=10000)
        got = api.coef_nnls(x, y, eps=eps, max_iter=10000)
        assert np.array_equal(got["iterations"], ref_it)
        assert rel_err(got["beta"], ref_b) < 1e-8
        assert np.all(got["beta"] >= 0.0)
        assert np.array_equal(got["beta"] == 0.0, ref_b == 0.0)


def test_coef_nnls_unconverged_gives_zero(api, rng):
    x = rng.standard_normal((200, 12)); x[:, 1] = x[:, 0] + 1e-3 * rng.standard_normal(200)
    y = x @ np.abs(rng.standard_normal((12, 9))) + rng.standard_normal((200, 9))
    ref_b, ref_it = orc.coef_nnls(x, y, eps=1e-13, max_iter=1)
    got = api.coef_nnls(x, y, eps=1e-13, max_iter=1)
    assert np.array_equal(got["iterations"], ref_it)
    assert np.array_equal(got["beta"] == 0.0, ref_b == 0.0)      # optim.cpp:423,502
    assert rel_err(got["beta"], ref_b) < 1e-8 or np.all(ref_b == 0)


@pytest.mark.parametrize("K,n2,T,weight", [(3, 50, 20, 0.0), (5, 333, 64, 0.0), (10, 1000, 40, 0.0), (4, 120, 30, 0.5)])
def test_allocate_into_modules(api, rng, K, n2, T, weight):
    z2 = rng.standard_normal((n2, T))
    arr = orc.alloc_array(z2 * z2, K)
    got_arr = api.alloc_array(z2 * z2, K)
    assert np.array_equal(arr, got_arr)
    arr *= rng.random((K, 1, T)) + 0.5
    arr[1] = arr[0]                                  # exact tie between module 1 and 2 -> first wins
    resid_var = np.asfortranarray(rng.random((T, K)) + 0.5)
    resid_var[:, 1] = resid_var[:, 0]
    k = rng.integers(1, K + 1, size=T).astype(np.int32); k[::7] = -1
    order = rng.permutation(T).astype(np.int32)
    prior = [list(rng.choice(T, size=rng.integers(0, 5), replace=False)) for _ in range(T)] if weight > 0 else []
    ref, ref_votes = orc.allocate_into_modules(arr, resid_var, prior, k, order, 1e-6, weight, return_votes=True)
    got, votes = api.allocate_into_modules(arr, resid_var, prior, k, order, 1e-6, weight, return_votes=True)
    assert np.array_equal(votes, ref_votes)
    assert np.array_equal(got, ref)
    if weight == 0.0:
        assert not np.any(got == 2)
    api.reset_array(got_arr, z2 * z2)
    assert np.array_equal(got_arr, orc.alloc_array(z2 * z2, K))


def test_allocate_into_modules_rejects_indices_that_would_leave_the_arrays(api, rng):
    """The reference documents these as assumptions (allocation.cpp:4-6); the ABI checks them because the sequential
    prior sweep indexes device memory with them."""
    from scregclust_b200._capi import ScrcError
    K, n2, T = 3, 20, 6
    arr = np.asfortranarray(rng.random((K, n2, T))); var = np.asfortranarray(rng.random((T, K)) + 0.5)
    k = rng.integers(1, K + 1, T).astype(np.int32)
    prior = [[(t + 1) % T] for t in range(T)]
    good = np.arange(T, dtype=np.int32)
    api.allocate_into_modules(arr, var, prior, k, good, 1e-6, 0.5)
    for bad_k, bad_order, bad_prior in [(np.where(np.arange(T) == 2, K + 1, k).astype(np.int32), good, prior),
                                        (k, np.array([0, 1, 1, 3, 4, 5], np.int32), prior),
                                        (k, good, [[T + 3]] + prior[1:])]:
        with pytest.raises(ScrcError) as e:
            api.allocate_into_modules(arr, var, bad_prior, bad_k, bad_order, 1e-6, 0.5)
        assert e.value.code == -20


def test_plain_c_client_runs_on_the_device(tmp_path):
    """tests/c/abi_smoke.c --gpu: coop_lasso and coef_nnls through the C ABI from a C program."""
    import subprocess
    from test_capi_cpu import build_c_client
    exe = build_c_client(tmp_path)
    out = subprocess.run([exe, "--gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "C ABI smoke: ok" in out.stdout, out.stdout + out.stderr


def test_kernels_match_golden_fixtures(api):
    """The committed golden vectors (tests/golden/kernels_seed7.npz, written by tests/golden/make_golden.py and
    checked against the oracle in tests/test_oracle.py) through the C ABI."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kernels_seed7.npz"))
    got = api.coop_lasso(g["coop_y"], g["coop_x"], float(g["coop_lambda"]), g["coop_w"], max_iter=10000)
    assert got["iterations"] == int(g["coop_iterations"])
    assert rel_err(got["beta"], g["coop_beta"]) < 1e-8
    got = api.coef_nnls(g["nnls_x"], g["nnls_y"], eps=1e-4, max_iter=10000)
    assert np.array_equal(got["iterations"], g["nnls_iterations"])
    assert rel_err(got["beta"], g["nnls_beta"]) < 1e-8
    k = api.allocate_into_modules(np.asfortranarray(g["alloc_arr"]), np.asfortranarray(g["alloc_var"]), [],
                                  g["alloc_k"], g["alloc_order"], 1e-6, 0.0)
    assert np.array_equal(k, g["alloc_out"])


@pytest.mark.gpu
def test_gram_form_sse_with_large_selections():
    """Weak penalization (or the ridge-initialised n1 < P regime, R/scregclust.R:1089-1100) keeps hundreds of regulators
    per module: the Gram-form sums of squares then run through the tiled path (64 x 64 blocks of the Gram matrix).
    Checked against the explicit residuals of the coefficients the same cycle returns: resid_var = ||z1 - X1 b||^2 /
    (n1 - s), r2_test = 1 - ||z2 - X2 b||^2 / ||z2||^2 (R/scregclust.R:1590-1638)."""
    from scregclust_b200.driver import Session
    from scregclust_b200.synthetic import make_workload
    w = make_workload(seed=4, shape=(800, 96, 220, 3), penalization=[0.002, 0.01])
    n1 = w.z1_reg.shape[0]
    L = len(w.penalization)
    with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules) as ses:
        res = ses.fit_cycle(w.penalization, np.stack([w.k_init] * L), skip_allocation=True)
        assert res["n_selected"].max() > 72                      # the tiled path is the one exercised
        ss2 = (w.z2_target ** 2).sum(axis=0)
        for f in range(L):
            cos, sels = ses.coeffs_fit(f)
            for j, (b, sel) in enumerate(zip(cos, sels)):
                if b is None:
                    continue
                sse1 = ((w.z1_target - w.z1_reg[:, sel] @ b) ** 2).sum(axis=0)
                sse2 = ((w.z2_target - w.z2_reg[:, sel] @ b) ** 2).sum(axis=0)
                assert np.allclose(res["resid_var"][f, j], sse1 / (n1 - sel.size), rtol=1e-8, atol=0)
                assert np.allclose(res["r2_test"][f, j], 1.0 - sse2 / ss2, rtol=0, atol=1e-8)
