"""Input staging on the device against the oracle's restatement of
split_sample + constant-gene filter + scaling (R/scregclust.R:2302-2325, 899-964, 996-1009).

Floating point: the staged matrices are sums / differences / quotients of the same numbers in a
different summation order than NumPy's, tolerance 1e-12 relative to the column scale; the keep
mask, the split sizes and the cell order (including the reference's stratum re-ordering quirk)
are compared exactly."""
import numpy as np
import pytest

from oracle import scregclust_oracle as orc
from test_oracle import _raw_expression

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-12):
    scale = max(1.0, float(np.max(np.abs(b))))
    return float(np.max(np.abs(a - b))) <= tol * scale


@pytest.mark.parametrize("seed,n_strata,center,p,n", [(3, 3, True, 60, 90), (4, 1, True, 45, 70), (5, 3, False, 60, 90),
                                                      (6, 4, True, 333, 517), (7, 2, True, 130, 1100)])
def test_staged_inputs_match_oracle(seed, n_strata, center, p, n):
    from scregclust_b200.driver import Session
    z, is_reg, strat, split = _raw_expression(seed, p=p, n=n, n_strata=n_strata)
    ref = orc.stage_inputs(z, strat if n_strata > 1 else None, is_reg, split, center=center)
    with Session.from_expression(z, is_reg, split, n_modules=3, stratification=strat if n_strata > 1 else None,
                                 center=center) as ses:
        got = ses.inputs()
        assert np.array_equal(ses.keep_gene, ref[5])
        assert (ses.n1, ses.n2) == (ref[0].shape[0], ref[1].shape[0])
        assert (ses.P, ses.T) == (ref[0].shape[1], ref[2].shape[1])
        for a, b, nm in zip(got, ref[:5], ["z1_reg", "z2_reg", "z1_target", "z2_target", "sds"]):
            assert a.shape == b.shape, nm
            assert _close(a, b), f"{nm}: max abs diff {np.max(np.abs(a - b)):.3e}"
        # the rest of the session is built from the staged matrices exactly as from uploaded ones
        with Session(*ref[:5], n_modules=3) as ses2:
            i1, i2 = ses.initial_fit(), ses2.initial_fit()
            assert i1["init_df"] == i2["init_df"]
            assert np.allclose(i1["res_sds"], i2["res_sds"], rtol=1e-9, atol=0)


def test_staged_session_runs_the_loop_like_an_uploaded_one():
    """Same assignments whether the driver scales on the host or the session stages the raw matrix."""
    from scregclust_b200.driver import Session, scregclust_fit
    from scregclust_b200.synthetic import CONFIGS, make_raw, make_workload
    w = make_workload("tiny", seed=3)
    n_cells, T, P, K, _ = CONFIGS["tiny"]
    zr, zt, _ = make_raw(n_cells, T, P, K, seed=3)
    # genes x cells, cells shuffled: the split indicator, not the position, decides the data split
    perm = np.random.default_rng(11).permutation(n_cells)
    raw = np.asfortranarray(np.concatenate([zr, zt], axis=1).T[:, perm])
    split = np.where(perm < n_cells // 2, 1.0, 2.0)
    is_reg = np.r_[np.ones(P, dtype=np.int32), np.zeros(T, dtype=np.int32)]
    ref = orc.stage_inputs(raw, None, is_reg, split, center=False)
    with Session.from_expression(raw, is_reg, split, n_modules=w.n_modules, center=False) as ses, \
            Session(*ref[:5], n_modules=w.n_modules) as ses2:
        a = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init, session=ses)
        b = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init, session=ses2)
        for ra, rb in zip(a, b):
            assert ra.n_cycles_run == rb.n_cycles_run
            assert all(np.array_equal(x, y) for x, y in zip(ra.k_final, rb.k_final))


def test_staging_rejects_bad_inputs():
    from scregclust_b200 import _capi
    from scregclust_b200.driver import Session
    z, is_reg, strat, split = _raw_expression(8)
    with pytest.raises(_capi.ScrcError):
        Session.from_expression(z, is_reg, np.full(split.shape, 2.0), n_modules=3)       # empty first split
    strat2 = strat.copy(); strat2[:] = 1; strat2[0] = 99; split2 = split.copy(); split2[0] = 2.0
    with pytest.raises(_capi.ScrcError):
        Session.from_expression(z, is_reg, split2, n_modules=3, stratification=strat2)   # stratum without split-1 cells


@pytest.mark.parametrize("center", [True, False])
def test_constant_genes_are_discarded_correctly(center):
    """The reference's own test of this step (tests/testthat/test-constant-genes.R:1-24), through the C ABI:
    after staging, T1 (all 1) and R1 (all 0.5) are gone and a fit on the remaining genes runs."""
    from test_oracle import _constant_genes_case
    from scregclust_b200.driver import Session, scregclust_fit
    expression, genesymbols, is_regulator, split = _constant_genes_case()
    with Session.from_expression(expression, is_regulator, split, n_modules=2, center=center) as ses:
        assert genesymbols[ses.keep_gene].tolist() == ["T2", "T3", "T4", "T5", "T6", "R2"]
        assert (is_regulator[ses.keep_gene] == 1).tolist() == [False, False, False, False, False, True]
        assert (ses.P, ses.T) == (1, 5)
        res = scregclust_fit(penalization=[0.1], n_modules=2, initial_target_modules=np.array([1, 2, 1, 2, 1], dtype=np.int32),
                             session=ses)
        assert res[0].k_final[0].shape == (5,)


def test_reference_constant_genes_test_through_the_frontend():
    """tests/testthat/test-constant-genes.R:1-24 with the reference's own call shape:
    ``scregclust(expression, genesymbols, is_regulator, 0.1, 2)`` -> genesymbols / is_regulator of the result."""
    from test_oracle import _constant_genes_case
    from scregclust_b200.frontend import scregclust
    expression, genesymbols, is_regulator, split = _constant_genes_case()
    fit = scregclust(expression, genesymbols, is_regulator, 0.1, 2, split_indices=split, seed=1, n_initializations=2)
    assert fit["results"][0]["genesymbols"] == ["T2", "T3", "T4", "T5", "T6", "R2"]
    assert fit["results"][0]["is_regulator"].tolist() == [False, False, False, False, False, True]
    out = fit["results"][0]["output"][0]
    assert out["module"].shape == (6,) and np.isnan(out["module"][5])          # NA for the regulator
