"""The user-facing mirror of ``scregclust()`` (scregclust_b200/frontend.py) without a GPU: validation messages,
bookkeeping of the constant-gene filter, packaging -- with the device session replaced by an oracle-backed one.
The first test is the reference's own tests/testthat/test-constant-genes.R, line for line."""
import numpy as np
import pytest

from oracle import scregclust_oracle as orc
from test_driver_host_logic import OracleSession
from test_oracle import _constant_genes_case, _raw_expression


class _W:      # the slice of synthetic.Workload that OracleSession reads
    pass


class OracleExpressionSession(OracleSession):
    """OracleSession built from a raw expression matrix, like driver.Session.from_expression."""

    def __init__(self, expression, is_regulator, split_indices, n_modules, stratification=None, center=True, device=0):
        z1r, z2r, z1t, z2t, sds, keep = orc.stage_inputs(expression, stratification, is_regulator, split_indices, center)
        w = _W()
        w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules = z1r, z2r, z1t, z2t, sds, n_modules
        super().__init__(w)
        self.keep_gene = keep

    def cross_corr(self):
        return orc.fast_cor(self.w.z1_target, self.w.z1_reg)

    def importance(self, k_final, models, signs, options=None, want_r2_removed=False):
        w = self.w
        F = len(k_final)
        r2m = np.full((F, self.K), np.nan); imp = np.full((F, self.K, self.P), np.nan); rem = []
        G_rr = w.z1_reg.T @ w.z1_reg
        G_rt = w.z1_reg.T @ (w.z1_target / w.z1_target_sds[None, :])
        for f in range(F):
            fr = orc.FitResult(penalization=0.0, converged=True, n_cycles_run=0, k_history=[], final_cycle_length=1)
            fr.k_final = [np.asarray(k_final[f])]
            fr.models = [np.asarray(models[f]).T.astype(bool)]
            fr.signs = [np.asarray(signs[f]).T]
            orc._importance(fr, w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, G_rr, G_rt, self.K,
                            1e-4, 10000)
            r2m[f] = fr.r2_module[0]; imp[f] = fr.importance[0].T; rem.append(fr.r2_removed[0])
        return (r2m, imp, rem) if want_r2_removed else (r2m, imp)


def test_constant_genes_are_discarded_correctly_through_the_frontend():
    """tests/testthat/test-constant-genes.R:1-24."""
    from scregclust_b200.frontend import scregclust
    expression, genesymbols, is_regulator, split = _constant_genes_case()
    fit = scregclust(expression, genesymbols, is_regulator, 0.1, 2, split_indices=split, seed=1,
                     n_initializations=2, session_factory=OracleExpressionSession)
    assert fit["results"][0]["genesymbols"] == ["T2", "T3", "T4", "T5", "T6", "R2"]
    assert fit["results"][0]["is_regulator"].tolist() == [False, False, False, False, False, True]


def test_frontend_packaging_matches_the_oracle_pipeline():
    from scregclust_b200.frontend import scregclust
    z, is_reg, strat, split = _raw_expression(21, p=70, n=160, n_strata=2)
    K = 3
    staged = orc.stage_inputs(z, strat, is_reg, split, center=True)
    keep = staged[5]
    n_t = int((is_reg[keep] == 0).sum())
    k0_all = np.random.default_rng(5).integers(1, K + 1, size=int((is_reg == 0).sum())).astype(np.int32)
    k0 = k0_all[keep[is_reg == 0]]
    assert k0.size == n_t
    ref = orc.scregclust_fit(*staged[:5], [0.05, 0.2], K, k0, compute_predictive_r2=True)
    fit = scregclust(z, [f"g{i}" for i in range(z.shape[0])], is_reg, [0.05, 0.2], K, initial_target_modules=k0_all,
                     sample_assignment=strat, split_indices=split, session_factory=OracleExpressionSession)
    assert np.array_equal(fit["initial_target_modules"], k0)
    is_t = is_reg[keep] == 0
    for res, r in zip(fit["results"], ref):
        assert res["converged"] == r.converged and len(res["output"]) == r.final_cycle_length
        assert len(res["genesymbols"]) == int(keep.sum())
        out = res["output"][0]
        assert np.array_equal(out["module"][is_t], r.k_final[0]) and np.all(np.isnan(out["module"][~is_t]))
        noise = r.k_final[0] == -1
        assert np.array_equal(out["module_all"][is_t][noise], r.best_r2_idx[0][noise])
        assert np.array_equal(out["module_all"][is_t][~noise], r.k_final[0][~noise])
        assert np.allclose(out["best_r2"][is_t], r.best_r2[0], rtol=0, atol=1e-12)
        assert np.allclose(out["r2_module"], r.r2_module[0], equal_nan=True)
        assert np.allclose(out["importance"], r.importance[0], equal_nan=True)
        # reg_table = models * weights * (|median correlation with the module's targets| > 0.025)   (:2025-2036)
        cc = orc.fast_cor(staged[2], staged[0])
        for j in range(K):
            rows = cc[r.k_final[0] == j + 1]
            if rows.shape[0]:
                expect = r.models[0][:, j] * r.weights[0][:, j] * (np.abs(np.median(rows, axis=0)) > 0.025)
                assert np.allclose(out["reg_table"][:, j], expect, rtol=0, atol=1e-12)


def test_frontend_draws_a_stratified_split_and_initial_modules_when_not_given():
    from scregclust_b200.frontend import draw_split_indices, scregclust
    z, is_reg, strat, _ = _raw_expression(22, p=50, n=200, n_strata=2)
    si = draw_split_indices(200, strat, 0.5, 0.8, np.random.default_rng(0))
    for s in np.unique(strat):
        n_s = int((strat == s).sum())
        inc = ~np.isnan(si[strat == s])
        assert inc.sum() == int(np.floor(n_s * 0.8))
        assert (si[strat == s] == 1).sum() == int(np.floor(inc.sum() * 0.5))
    fit = scregclust(z, [f"g{i}" for i in range(50)], is_reg, 0.2, 2, sample_assignment=strat, seed=3,
                     n_initializations=2, n_cycles=3, compute_predictive_r2=False, session_factory=OracleExpressionSession)
    assert set(np.unique(fit["initial_target_modules"])) <= {1, 2}
    assert np.isin(fit["split_indices"][~np.isnan(fit["split_indices"])], (1.0, 2.0)).all()


@pytest.mark.parametrize("kwargs,needle", [
    (dict(genesymbols=["a"]), "genesymbols needs to be of length"),
    (dict(is_regulator=np.array([0, 1, 2, 0, 0, 0, 1, 1])), "is_regulator needs to be 0/1"),
    (dict(penalization=-1.0), "penalization"),
    (dict(n_modules=0), "n_modules"),
    (dict(split_indices=np.arange(100.0)), "split_indices"),
    (dict(noise_threshold=2.0), "noise_threshold"),
    (dict(initial_target_modules=np.array([1, 2, 3, 1, 1, 1])), "initial_target_modules"),
])
def test_frontend_argument_errors(kwargs, needle):
    from scregclust_b200.frontend import ScregclustError, scregclust
    expression, genesymbols, is_regulator, split = _constant_genes_case()
    args = dict(expression=expression, genesymbols=genesymbols, is_regulator=is_regulator, penalization=0.1,
                n_modules=2, split_indices=split, session_factory=OracleExpressionSession)
    args.update(kwargs)
    with pytest.raises(ScregclustError, match=needle):
        scregclust(**args)


def test_prior_matrix_is_mapped_to_target_neighbour_lists():
    """R/scregclust.R:584-608,1026-1047 on a small example worked out by hand."""
    from scregclust_b200.frontend import prior_to_target_lists
    prior_genes = ["A", "B", "X", "C", "D"]                      # X is not a target gene
    pi = np.zeros((5, 5), dtype=int)
    pi[0, 1] = pi[1, 0] = 1      # A - B
    pi[0, 2] = pi[2, 0] = 1      # A - X  (dropped: X is no target)
    pi[3, 3] = 1                 # C - C  (diagonal: dropped)
    pi[4, 0] = 1                 # D -> A (directed entry: only D's list gets A)
    targets = ["B", "A", "E", "D", "C"]                           # E is not in the prior
    got = prior_to_target_lists(pi, prior_genes, targets)
    assert got == [[1], [0], [], [1], []]
    import scipy.sparse
    assert prior_to_target_lists(scipy.sparse.csr_matrix(pi), prior_genes, targets) == got


def test_silhouette_follows_the_reference_formula():
    from scregclust_b200.frontend import silhouette_from_r2
    r2 = np.array([[0.5, 0.2, -0.1], [0.1, -0.3, -0.2], [0.0, 0.4, 0.6], [0.3, 0.3, 0.3]])
    k = np.array([1, 2, -1, 3])
    s = silhouette_from_r2(r2, k)
    assert np.isclose(s[0], (0.5 - 0.2) / 0.5)
    assert np.isclose(s[1], (-0.3 - 0.1) / 0.1)                  # b = max(0.1, -0.2) = 0.1
    assert np.isnan(s[2])
    assert np.isclose(s[3], 0.0)


def test_frontend_accepts_the_prior_matrix_form():
    from scregclust_b200.frontend import ScregclustError, scregclust
    z, is_reg, strat, split = _raw_expression(23, p=40, n=120, n_strata=1)
    genes = [f"g{i}" for i in range(40)]
    tgt = [g for g, r in zip(genes, is_reg) if r == 0]
    q = len(tgt)
    pi = (np.random.default_rng(1).random((q, q)) < 0.1).astype(int)
    common = dict(expression=z, genesymbols=genes, is_regulator=is_reg, penalization=0.2, n_modules=2,
                  split_indices=split, n_cycles=2, compute_predictive_r2=False, compute_silhouette=True, seed=0,
                  n_initializations=2, session_factory=OracleExpressionSession)
    with pytest.raises(ScregclustError, match="prior_genesymbols missing"):
        scregclust(prior_indicator=pi, **common)
    # weight 0 keeps the sweep order-independent (the oracle-backed session has no prior path)
    fit = scregclust(prior_indicator=pi, prior_genesymbols=tgt, prior_weight=0.0, **common)
    out = fit["results"][0]["output"][0]
    assert out["silhouette"].shape == (len(fit["initial_target_modules"]),)
    assert out["r2_cross_module_per_target"] is out["r2"]


def test_unsupported_modes_fail_loudly():
    from scregclust_b200.frontend import ScregclustError, scregclust
    expression, genesymbols, is_regulator, split = _constant_genes_case()
    args = dict(expression=expression, genesymbols=genesymbols, is_regulator=is_regulator, penalization=0.1,
                n_modules=2, split_indices=split, session_factory=OracleExpressionSession)
    with pytest.raises(NotImplementedError):
        scregclust(allocate_per_obs=False, **args)
    with pytest.raises(ScregclustError, match="quick_mode only available"):
        scregclust(quick_mode=True, allocate_per_obs=False, **args)
    with pytest.raises(ScregclustError, match="quick_mode_percent"):
        scregclust(quick_mode=True, quick_mode_percent=1.5, **args)
