"""CPU tests of the oracle (the checker itself): golden vectors, the one numeric pin the
reference's own tests provide (fast_cor == stats::cor, tests/testthat/test-fast-cor.R) and
independent cross-checks (KKT conditions, scipy NNLS, lstsq, brute-force allocation, the
restated Eigen LDLT, the C++ twin)."""
import os

import numpy as np
import pytest
import scipy.optimize

from conftest import rel_err
from oracle import scregclust_oracle as orc
from scregclust_b200.synthetic import make_workload

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fast_cor_matches_stats_cor(rng):
    # mirrors tests/testthat/test-fast-cor.R:1-16 (pt=50, pr=10, n=200)
    zt = rng.standard_normal((200, 50)); zr = rng.standard_normal((200, 10))
    ref = np.corrcoef(zt.T, zr.T)[:50, 50:]
    assert np.allclose(orc.fast_cor(zt, zr), ref, rtol=0, atol=1.5e-8)
    assert rel_err(orc.fast_cor(zt, zr), ref) < 1e-13


def test_constant_gene_smoke(rng):
    # spirit of tests/testthat/test-constant-genes.R: 8 genes x 100 cells, lambda 0.1, K=2 runs through
    expr = np.vstack([rng.standard_normal((5, 100)), rng.standard_normal((1, 100))])
    zr = expr[5:].T; zt = expr[:5].T
    from scregclust_b200.synthetic import scale_split
    z1r, z2r, z1t, z2t, sds = scale_split(zr, zt)
    res = orc.scregclust_fit(z1r, z2r, z1t, z2t, sds, [0.1], 2, np.array([1, 2, 1, 2, 1]))
    assert len(res) == 1 and res[0].k_final[0].shape == (5,)


def test_ldlt_restatement_solves(rng):
    for n in (1, 5, 40):
        a = rng.standard_normal((n + 5, n)); g = a.T @ a + 0.2 * np.eye(n)
        b = rng.standard_normal((n, 3))
        x = orc.EigenLDLT().compute(g).solve(b)
        assert rel_err(x, np.linalg.solve(g, b)) < 1e-10
    # semi-definite input: pivoting keeps it finite
    g = np.zeros((4, 4)); g[0, 0] = 2.0; g[2, 2] = 1.0
    x = orc.EigenLDLT().compute(g).solve(np.ones((4, 1)))
    assert np.all(np.isfinite(x))


def _coop_problem(rng, n=300, p=25, m=8, lam=0.15):
    x = rng.standard_normal((n, p)) / np.sqrt(n)
    b = np.zeros((p, m)); b[:5] = rng.standard_normal((5, 1)) + 0.3 * rng.standard_normal((5, m))
    y = x @ b + rng.standard_normal((n, m)) / np.sqrt(n)
    w = np.sqrt(m) * (0.5 + rng.random(p))
    return x, y, w, lam


def test_coop_lasso_kkt(rng):
    x, y, w, lam = _coop_problem(rng)
    beta, it = orc.coop_lasso(y, x, lam, w, max_iter=10000)
    assert 1 < it < 10000
    b = np.where(np.abs(beta) < 1e-6, 0.0, beta)
    grad = x.T @ (y - x @ b)                                   # = subgradient of the penalty
    for i in range(x.shape[1]):
        for sign in (1.0, -1.0):
            part = np.maximum(sign * b[i], 0.0)
            gpart = np.maximum(sign * grad[i], 0.0)
            if np.linalg.norm(part) > 0:                      # active group: gradient aligned, norm lam*w
                act = part > 0
                assert abs(np.linalg.norm(sign * grad[i][act]) - lam * w[i]) < 1e-5
            else:                                             # inactive group: inside the dual ball
                assert np.linalg.norm(gpart) <= lam * w[i] + 1e-5


def test_coop_lasso_ldlt_vs_cholesky_same_iterations(rng):
    x, y, w, lam = _coop_problem(rng)
    b1, i1 = orc.coop_lasso(y, x, lam, w, max_iter=10000, fast_solver=True)
    b2, i2 = orc.coop_lasso(y, x, lam, w, max_iter=10000, fast_solver=False)
    assert i1 == i2 and rel_err(b1, b2) < 1e-10


def test_coop_lasso_quirks(rng):
    x, y, w, lam = _coop_problem(rng)
    _, it = orc.coop_lasso(y, x, lam, w, max_iter=3)
    assert it == 4                                            # optim.cpp:295-305
    with pytest.raises(ValueError, match="rho_0"):
        orc.coop_lasso(y, x, lam, w, rho_0=0.0)
    with pytest.raises(ValueError, match="alpha_0"):
        orc.coop_lasso(y, x, lam, w, alpha_0=2.1)
    with pytest.raises(ValueError, match="same number of rows"):
        orc.coop_lasso(y[:-1], x, lam, w)
    # infinite weight (a regulator with zero initial coefficients) is shrunk to exactly zero in zeta
    w2 = w.copy(); w2[0] = np.inf
    beta, _ = orc.coop_lasso(y, x, lam, w2, max_iter=10000)
    assert abs(beta[0]).max() < 1e-6


def test_nnls_against_scipy(rng):
    x = rng.standard_normal((400, 9))
    y = x @ np.maximum(rng.standard_normal((9, 12)), 0) + 0.3 * rng.standard_normal((400, 12))
    beta, its = orc.coef_nnls(x, y, eps=1e-12, max_iter=10000)
    assert np.all(its < 10000) and np.all(beta >= 0)
    for j in range(12):
        ref, _ = scipy.optimize.nnls(x, y[:, j])
        assert np.allclose(beta[:, j], ref, atol=1e-8)
    grad = x.T @ (x @ beta - y)
    assert np.all(grad[beta == 0] > -1e-7)                    # complementary slackness


def test_nnls_unconverged_columns_are_zero():
    rng = np.random.default_rng(7)          # own stream: every column must need more than two outer iterations
    x = rng.standard_normal((100, 6))
    y = x @ (rng.random((6, 4)) + 0.5) + 0.1 * rng.standard_normal((100, 4))
    beta, its = orc.coef_nnls(x, y, eps=1e-300, max_iter=2)
    assert np.all(its == 2) and np.all(beta == 0)             # optim.cpp:423,502


def _brute_alloc(arr, resid_var, k_in):
    K, n, T = arr.shape
    out = np.zeros(T, dtype=np.int32)
    for t in range(T):
        votes = np.zeros(K, dtype=int)
        for c in range(n):
            ll = [-arr[q, c, t] / (2 * resid_var[t, q]) - 0.5 * np.log(2 * np.pi * resid_var[t, q]) for q in range(K)]
            votes[int(np.argmax(ll))] += 1
        out[t] = int(np.argmax(votes)) + 1
    return out


def test_allocation_bruteforce_and_ties(rng):
    K, n, T = 4, 60, 15
    arr = np.asfortranarray(rng.random((K, n, T)))
    arr[2] = arr[1]                                           # exact tie -> first index wins
    rv = np.asfortranarray(rng.random((T, K)) + 0.5); rv[:, 2] = rv[:, 1]
    k = rng.integers(1, K + 1, T); k[0] = -1
    got = orc.allocate_into_modules(arr, rv, [], k, rng.permutation(T), 1e-6, 0.0)
    assert np.array_equal(got, _brute_alloc(arr, rv, k))
    assert not np.any(got == 3)


def test_allocation_prior_is_order_dependent_only_with_weight(rng):
    K, n, T = 3, 40, 25
    arr = np.asfortranarray(rng.random((K, n, T))); rv = np.asfortranarray(rng.random((T, K)) + 0.5)
    k = rng.integers(1, K + 1, T)
    prior = [list(rng.choice(T, 3, replace=False)) for _ in range(T)]
    a = orc.allocate_into_modules(arr, rv, prior, k, np.arange(T), 1e-6, 0.0)
    b = orc.allocate_into_modules(arr, rv, prior, k, np.arange(T)[::-1], 1e-6, 0.0)
    assert np.array_equal(a, b)                               # weight 0 => genes independent


def test_alloc_and_reset_array(rng):
    z = rng.random((7, 5))
    arr = orc.alloc_array(z, 3)
    assert arr.shape == (3, 7, 5) and arr.flags.f_contiguous
    assert all(np.array_equal(arr[q], z) for q in range(3))
    arr[1] = 0
    orc.reset_array(arr, z)
    assert np.array_equal(arr[1], z)
    with pytest.raises(ValueError, match="wrong dimensions"):
        orc.reset_array(arr, z[:, :4])


def test_ols_ridge(rng):
    x = rng.standard_normal((80, 10)); y = rng.standard_normal((80, 3))
    assert rel_err(orc.coef_ols(y, x), np.linalg.lstsq(x, y, rcond=None)[0]) < 1e-10
    x2 = rng.standard_normal((8, 12)); y2 = rng.standard_normal((8, 2))
    assert rel_err(orc.coef_ols(y2, x2), np.linalg.pinv(x2.T @ x2) @ x2.T @ y2) < 1e-8
    lam = 2.0
    aug_x = np.vstack([x, np.sqrt(lam) * np.eye(10)]); aug_y = np.vstack([y, np.zeros((10, 3))])
    assert rel_err(orc.coef_ridge(y, x, lam), np.linalg.lstsq(aug_x, aug_y, rcond=None)[0]) < 1e-10


def test_adjusted_rand_index():
    a = np.array([1, 1, 2, 2, 3, 3, -1])
    assert orc.adjusted_rand_index(a, a) == 1.0
    assert orc.adjusted_rand_index(a, np.array([2, 2, 3, 3, 1, 1, 7])) == 1.0
    assert orc.adjusted_rand_index(a, np.array([1, 2, 1, 2, 1, 2, 1])) < 0.2


def test_recycling_quirk_of_the_driver():
    # R: matrix(1, 2, 3) * c(10, 20, 30)  ->  [[10, 30, 20], [20, 10, 30]]
    got = orc._recycle(np.array([10.0, 20.0, 30.0]), (2, 3))
    assert np.array_equal(got, np.array([[10.0, 30.0, 20.0], [20.0, 10.0, 30.0]]))


def test_loop_golden_trajectory():
    # seed 3: no vote of this trajectory is decided by rounding noise (oracle ambiguity == 0),
    # so the fixture is host independent
    g = np.load(os.path.join(GOLD, "loop_tiny_seed3.npz"))
    w = make_workload("tiny", seed=3)
    res = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, compute_predictive_r2=True)
    for i, r in enumerate(res):
        assert sum(int(a.sum()) for a in r.ambiguity) == 0
        assert np.array_equal(np.stack(r.k_history), g[f"k_history_{i}"])
        assert np.array_equal(np.stack(r.admm_iterations), g[f"admm_iterations_{i}"])
        assert np.array_equal(r.models[0], g[f"models_{i}"])
        assert rel_err(r.r2[0], g[f"r2_{i}"]) < 1e-9
        assert rel_err(r.sigmas[0], g[f"sigmas_{i}"]) < 1e-9
        assert np.allclose(r.importance[0], g[f"importance_{i}"], rtol=1e-7, atol=1e-9, equal_nan=True)


def test_kernel_golden_vectors():
    g = np.load(os.path.join(GOLD, "kernels_seed7.npz"))
    b, it = orc.coop_lasso(g["coop_y"], g["coop_x"], float(g["coop_lambda"]), g["coop_w"], max_iter=10000)
    assert it == int(g["coop_iterations"]) and rel_err(b, g["coop_beta"]) < 1e-9
    b, its = orc.coef_nnls(g["nnls_x"], g["nnls_y"], eps=1e-4, max_iter=10000)
    assert np.array_equal(its, g["nnls_iterations"]) and rel_err(b, g["nnls_beta"]) < 1e-9
    k = orc.allocate_into_modules(np.asfortranarray(g["alloc_arr"]), np.asfortranarray(g["alloc_var"]), [],
                                  g["alloc_k"], g["alloc_order"], 1e-6, 0.0)
    assert np.array_equal(k, g["alloc_out"])


def test_cpp_twin_matches_numpy_oracle(rng):
    ref_cpu = pytest.importorskip("oracle.ref_cpu")
    if not os.path.exists(ref_cpu.LIB_PATH):
        pytest.skip("oracle/_ref/liboracle_ref.so not built (run __graft_entry__.build())")
    x, y, w, lam = _coop_problem(rng, n=400, p=40, m=10)
    b1, i1 = orc.coop_lasso(y, x, lam, w, max_iter=10000)
    b2, i2, nf = ref_cpu.coop_lasso(y, x, lam, w, max_iter=10000)
    assert i1 == i2 and rel_err(b2, b1) < 1e-9 and nf >= i2 // 2
    xs = rng.standard_normal((300, 7)); ys = xs @ np.abs(rng.standard_normal((7, 20))) + rng.standard_normal((300, 20))
    n1, t1 = orc.coef_nnls(xs, ys, eps=1e-4, max_iter=10000)
    n2, t2 = ref_cpu.coef_nnls(xs, ys, eps=1e-4, max_iter=10000)
    assert np.array_equal(t1, t2) and rel_err(n2, n1) < 1e-9
    K, n, T = 4, 50, 12
    arr = np.asfortranarray(rng.random((K, n, T))); rv = np.asfortranarray(rng.random((T, K)) + 0.5)
    k = rng.integers(1, K + 1, T).astype(np.int32); order = rng.permutation(T).astype(np.int32)
    assert np.array_equal(ref_cpu.allocate_into_modules(arr, rv, None, None, k, order, 1e-6, 0.0),
                          orc.allocate_into_modules(arr, rv, [], k, order, 1e-6, 0.0))


def _raw_expression(seed, p=60, n=90, n_strata=3, sort_strata=False):
    r = np.random.default_rng(seed)
    z = np.round(r.gamma(1.5, 1.0, size=(p, n)), 3)
    is_reg = (r.random(p) < 0.3).astype(np.int32)
    is_reg[:2] = 1; is_reg[2:4] = 0
    strat = r.integers(0, n_strata, size=n) * 7 + 3          # arbitrary integer labels
    if sort_strata:
        strat = np.sort(strat)
    split = np.where(r.random(n) < 0.5, 1.0, 2.0)
    split[r.random(n) < 0.1] = np.nan                        # unused cells
    z[5, :] = 0.0                                            # constant everywhere
    z[9, split == 2] = 4.0                                   # constant in the second split only
    return z, is_reg, strat, split


def test_split_sample_single_stratum_is_split1_row_centring():
    z, is_reg, _, split = _raw_expression(1)
    z1r, z2r, z1t, z2t = orc.split_sample(z, None, is_reg, split, center=True)
    inc = ~np.isnan(split)
    mu = z[:, split == 1].mean(axis=1, keepdims=True)
    assert np.allclose(z1r, (z[:, split == 1] - mu)[is_reg == 1].T, rtol=0, atol=1e-14)
    assert np.allclose(z2t, (z[:, inc & (split == 2)] - mu)[is_reg == 0].T, rtol=0, atol=1e-14)
    # no centring: plain selection
    a1r, _, _, a2t = orc.split_sample(z, None, is_reg, split, center=False)
    assert np.array_equal(a1r, z[:, split == 1][is_reg == 1].T)
    assert np.array_equal(a2t, z[:, inc & (split == 2)][is_reg == 0].T)


def test_split_sample_reorders_columns_by_stratum_when_centring():
    """R/scregclust.R:2309-2323: cbind over strata permutes the cells, the split positions stay."""
    z, is_reg, strat, split = _raw_expression(2)
    z1r, z2r, z1t, z2t = orc.split_sample(z, strat, is_reg, split, center=True)
    inc = np.flatnonzero(~np.isnan(split))
    pos1 = np.flatnonzero(split[inc] == 1)
    perm = np.argsort(strat[inc], kind="stable")
    # centred values cell by cell
    cen = np.empty((z.shape[0], inc.size))
    for lev in np.unique(strat[inc]):
        s_ = np.flatnonzero(strat[inc] == lev)
        idx = np.intersect1d(s_, pos1)
        cen[:, s_] = z[:, inc[s_]] - z[:, inc[idx]].mean(axis=1, keepdims=True)
    expect = cen[:, perm][:, pos1][is_reg == 1].T
    assert np.allclose(z1r, expect, rtol=0, atol=1e-14)
    assert z1r.shape[0] + z2r.shape[0] == inc.size
    # with strata already sorted the permutation is the identity
    zs, ir, st, sp = _raw_expression(2, sort_strata=True)
    b1r, *_ = orc.split_sample(zs, st, ir, sp, center=True)
    incs = np.flatnonzero(~np.isnan(sp))
    assert np.array_equal(np.argsort(st[incs], kind="stable"), np.arange(incs.size))
    assert b1r.shape == (int((sp == 1).sum()), int((ir == 1).sum()))


def test_stage_inputs_filters_constant_genes_and_scales():
    z, is_reg, strat, split = _raw_expression(3)
    # per-stratum centring makes gene 9 vary across strata again; gene 5 (all zero) stays constant
    *_, keep_c = orc.stage_inputs(z, strat, is_reg, split, center=True)
    assert not keep_c[5] and keep_c[9] and keep_c.sum() == z.shape[0] - 1
    z1r, z2r, z1t, z2t, sds, keep = orc.stage_inputs(z, strat, is_reg, split, center=False)
    assert not keep[5] and not keep[9] and keep.sum() == z.shape[0] - 2
    assert z1r.shape[1] == int((is_reg[keep] == 1).sum()) and z1t.shape[1] == int((is_reg[keep] == 0).sum())
    assert np.allclose(z1r.mean(axis=0), 0, atol=1e-13) and np.allclose(z1r.std(axis=0, ddof=1), 1, atol=1e-12)
    assert np.allclose(z1t.mean(axis=0), 0, atol=1e-13) and np.allclose(z1t.std(axis=0, ddof=1), sds, rtol=1e-12)
    assert abs(z2r.mean()) > 1e-6                               # the test split keeps split-1 statistics


def _constant_genes_case():
    """The fixture of the reference's tests/testthat/test-constant-genes.R:1-10."""
    r = np.random.default_rng(0)
    expression = np.vstack([np.ones((1, 100)), r.standard_normal((5, 100)), np.full((1, 100), 0.5),
                            r.standard_normal((1, 100))])
    genesymbols = np.array(["T1", "T2", "T3", "T4", "T5", "T6", "R1", "R2"])
    is_regulator = np.array([0, 0, 0, 0, 0, 0, 1, 1], dtype=np.int32)
    split = np.where(r.permutation(100) < 50, 1.0, 2.0)           # split_indices drawn by the caller
    return expression, genesymbols, is_regulator, split


@pytest.mark.parametrize("center", [True, False])
def test_constant_genes_are_discarded_correctly(center):
    """tests/testthat/test-constant-genes.R:12-23: T1 and R1 are constant and must be dropped."""
    expression, genesymbols, is_regulator, split = _constant_genes_case()
    *_, keep = orc.stage_inputs(expression, None, is_regulator, split, center=center)
    assert genesymbols[keep].tolist() == ["T2", "T3", "T4", "T5", "T6", "R2"]
    assert (is_regulator[keep] == 1).tolist() == [False, False, False, False, False, True]


def test_r2_removed_recycles_the_column_sums_like_r():
    """scregclust.R:1890-1896: `t(ssq) / v` with an (s+1) x m matrix and a length-m vector divides the
    element with column-major flat index i by v[i mod m]."""
    w = make_workload("tiny", seed=3)
    r = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization[:1],
                           w.n_modules, w.k_init, compute_predictive_r2=True)[0]
    k = r.k_final[0]
    checked = 0
    for j, blk in enumerate(r.r2_removed[0]):
        if blk is None:
            continue
        tc = np.flatnonzero(k == j + 1)
        zc = w.z2_target[:, tc] - w.z2_target[:, tc].mean(axis=0)
        css = (zc * zc).sum(axis=0)
        m, s1 = blk.shape
        assert m == tc.size and s1 == int(r.models[0][:, j].sum()) + 1
        # recover ssq from column 0 of a correctly aligned entry: flat index of (r=0, c) is c*(s+1)
        for c in range(m):
            ssq_c0 = (1.0 - blk[c, 0]) * css[(c * s1) % m]
            assert ssq_c0 > 0
        # module R2 is the properly summed version of column 0
        ssq0 = np.array([(1.0 - blk[c, 0]) * css[(c * s1) % m] for c in range(m)])
        assert abs((1.0 - ssq0.sum() / css.sum()) - r.r2_module[0][j]) < 1e-12
        checked += 1
    assert checked > 0


def test_coop_lasso_single_column_is_the_weighted_lasso(rng):
    """Independent anchor: for one response column the cooperative-lasso penalty
    lambda * w_i * (||beta_i^+|| + ||beta_i^-||) is lambda * w_i * |beta_i|, i.e. a weighted lasso, whose optimum
    scikit-learn's coordinate descent finds by a completely different route.  With the objective of
    src/optim.cpp:134-176 (1/2 ||y - x beta||^2 + penalty) and columns rescaled by 1 / w_i:
    Lasso(alpha = lambda / n) on x / w returns w * beta."""
    from sklearn.linear_model import Lasso
    n, p = 400, 30
    x = rng.standard_normal((n, p)) / np.sqrt(n)
    beta_true = np.zeros(p); beta_true[:6] = rng.standard_normal(6)
    y = x @ beta_true + 0.3 * rng.standard_normal(n) / np.sqrt(n)
    w = 0.5 + rng.random(p)
    lam = 0.05
    beta, it = orc.coop_lasso(y[:, None], x, lam, w, max_iter=100000, eps_rel=1e-12, eps_abs=1e-14)
    sk = Lasso(alpha=lam / n, fit_intercept=False, tol=1e-14, max_iter=1000000).fit(x / w[None, :], y)
    ref = sk.coef_ / w
    assert it < 100000
    assert np.max(np.abs(beta[:, 0] - ref)) < 1e-7 * max(1.0, np.max(np.abs(ref)))
    assert np.array_equal(np.abs(beta[:, 0]) > 1e-6, np.abs(ref) > 1e-6) and 0 < np.sum(np.abs(ref) > 1e-6) < p


def test_aggregate_allocation_branch():
    """R/scregclust.R:1678-1728 (allocate_per_obs = FALSE).  Without a prior the normalisation is monotone per gene, so the
    result is the arg-min of n2 log(2 pi s2) + SS / s2; with a prior the sweep depends on the order, module totals
    stay those of the incoming configuration, and noise genes (-1) are reassigned like any other gene."""
    rng = np.random.default_rng(11)
    T, K, n2 = 60, 5, 200
    ss = rng.random((T, K)) * 80 + 20; rv = rng.random((T, K)) * 0.5 + 0.1
    k = rng.integers(1, K + 1, T).astype(np.int32); k[:5] = -1
    plain = orc.allocate_aggregate(k, ss, rv, n2, (), np.arange(T), 1e-6, 0.0)
    assert np.array_equal(plain, np.argmin(n2 * np.log(2 * np.pi * rv) + ss / rv, axis=1) + 1)
    assert np.array_equal(plain, orc.allocate_aggregate(k, ss, rv, n2, (), np.arange(T)[::-1], 1e-6, 0.0))
    # exact tie -> which.min takes the first module
    ss2 = ss.copy(); rv2 = rv.copy(); ss2[:, 3] = ss2[:, 1]; rv2[:, 3] = rv2[:, 1]
    tied = orc.allocate_aggregate(k, ss2, rv2, n2, (), np.arange(T), 1e-6, 0.0)
    assert not np.any(tied == 4)
    # hand-worked prior step: one gene, two modules with equal model scores, all neighbours in module 2 ->
    # the prior's log-probability is larger for module 2, and which.min therefore picks module 1
    ss3 = np.full((3, 2), 10.0); rv3 = np.full((3, 2), 1.0)
    out = orc.allocate_aggregate(np.array([1, 2, 2]), ss3, rv3, 50, [[1, 2], [], []], np.array([0]), 1e-6, 0.5)
    assert out.tolist() == [1, 2, 2]
    out = orc.allocate_aggregate(np.array([2, 1, 1]), ss3, rv3, 50, [[1, 2], [], []], np.array([0]), 1e-6, 0.5)
    assert out.tolist() == [2, 1, 1]
    # order dependence through the neighbours' current modules
    prior = [sorted(rng.choice(T, size=3, replace=False).tolist()) for _ in range(T)]
    a = orc.allocate_aggregate(k, ss, rv, n2, prior, np.arange(T), 1e-6, 0.9)
    b = orc.allocate_aggregate(k, ss, rv, n2, prior, np.arange(T)[::-1], 1e-6, 0.9)
    assert np.all((a >= 1) & (a <= K)) and not np.array_equal(a, b)


def test_aggregate_golden_vectors():
    """tests/golden/aggregate_seed9.npz (tests/golden/make_golden.py aggregate): the aggregate branch, gene by gene and through the loop."""
    g = np.load(os.path.join(GOLD, "aggregate_seed9.npz"))
    prior = [list(r) for r in g["prior"]]
    T = g["k"].size
    assert np.array_equal(orc.allocate_aggregate(g["k"], g["ss"], g["rv"], int(g["n2"]), (), np.arange(T), 1e-6, 0.0), g["out_plain"])
    assert np.array_equal(orc.allocate_aggregate(g["k"], g["ss"], g["rv"], int(g["n2"]), prior, g["order"], 1e-6, 0.5), g["out_prior"])
    assert not np.array_equal(g["out_plain"], g["out_prior"])
    w = make_workload("tiny", seed=3)
    res = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, allocate_per_obs=False, min_module_size=2)
    for i, r in enumerate(res):
        assert np.array_equal(np.stack(r.k_history), g[f"k_history_{i}"])
