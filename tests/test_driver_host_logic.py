"""Host logic of the session driver (scregclust_b200/driver.py) without a GPU.

The driver's control flow -- noise threshold, minimum module size, history match / limit cycles,
forced last cycle, deferral and batching of the final re-fits, dynamic claiming with a bounded
number of penalization values in flight -- is exercised against the oracle's monolithic loop by
substituting the device session with one that answers each batched cycle from the oracle
(tests may import oracle/, the product path never does)."""
import numpy as np
import pytest

from oracle import scregclust_oracle as orc
from scregclust_b200.synthetic import make_workload


class OracleSession:
    """Implements the slice of driver.Session that scregclust_fit() uses."""

    def __init__(self, w):
        self.w = w
        self.K = w.n_modules
        self.T = w.z1_target.shape[1]
        self.P = w.z1_reg.shape[1]
        self.pre = orc.precompute(w.z1_reg, w.z1_target)
        self.calls = []          # (n_fits, skip_allocation) per batched call
        self.quick_calls = []    # idx_latest handed to each call (None: no quick-mode subset)
        self._last = None

    def _one(self, lam, k, n_cycles, quick=None, aggregate=False):
        w = self.w
        kw = {}
        if quick is not None:
            kw = dict(quick_mode=True, quick_mode_percent=quick[1], quick_first_cycle_idx=quick[0])
        if aggregate:
            kw["allocate_per_obs"] = False
        return orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, [lam], self.K, k,
                                  n_cycles=n_cycles, pre=self.pre, keep_history=True, materialize_array=False, **kw)[0]

    def fit_cycle(self, lambdas, ks, options=None, skip_allocation=False, want_votes=False, want_nnls_iters=False,
                  prior=None, light=False, quick=None, aggregate=False):
        F, K, T, P = len(lambdas), self.K, self.T, self.P
        ks = np.asarray(ks).reshape(F, T)
        self.calls.append((F, bool(skip_allocation)))
        self.aggregate_calls = getattr(self, "aggregate_calls", 0) + (1 if aggregate else 0)
        self.quick_calls.append(None if quick is None else np.asarray(quick[0]).copy())
        res = {"k_new": np.zeros((F, T), np.int32), "models": np.zeros((F, K, P), np.uint8),
               "best_r2": np.zeros((F, T)), "admm_iterations": np.zeros((F, K), np.int32),
               "n_selected": np.zeros((F, K), np.int32), "votes": None, "nnls_iterations": None,
               "signs": np.zeros((F, K, P)), "weights": np.zeros((F, K, P)), "r2_test": np.zeros((F, K, T)),
               "resid_var": np.zeros((F, K, T)), "best_r2_idx": np.zeros((F, T), np.int32)}
        self._last = []
        for f in range(F):
            if skip_allocation:                     # the forced last cycle of the reference: re-fit only
                r = self._one(lambdas[f], ks[f], 0, None if quick is None else (quick[0][f], quick[1]))
                res["models"][f] = r.models[0].T
                res["signs"][f] = r.signs[0].T; res["weights"][f] = r.weights[0].T
                res["r2_test"][f] = r.r2[0].T; res["resid_var"][f] = (r.sigmas[0] ** 2).T
                res["best_r2"][f] = r.best_r2[0]; res["best_r2_idx"][f] = r.best_r2_idx[0]
                res["admm_iterations"][f] = r.admm_iterations[0]
                res["n_selected"][f] = r.models[0].sum(axis=0)
                self._last.append(r.coeffs[0])
            else:
                r = self._one(lambdas[f], ks[f], 1, None if quick is None else (quick[0][f], quick[1]), aggregate)
                res["models"][f] = r.models_history[0].T
                res["k_new"][f] = r.allocations[0] if aggregate else np.argmax(r.votes[0], axis=1) + 1
                res["best_r2"][f] = r.r2_history[0].max(axis=1)
                res["admm_iterations"][f] = r.admm_iterations[0]
                res["n_selected"][f] = r.models_history[0].sum(axis=0)
                self._last.append(None)
        return res

    def coeffs_fit(self, f):
        return self._last[f], None

    def close(self):
        pass


def _same(got, ref):
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g.n_cycles_run == r.n_cycles_run and g.converged == r.converged
        assert g.final_cycle_length == r.final_cycle_length
        assert len(g.k_history) == len(r.k_history)
        assert all(np.array_equal(a, b) for a, b in zip(g.k_history, r.k_history))
        for m in range(r.final_cycle_length):
            assert np.array_equal(g.k_final[m], r.k_final[m])
            assert np.array_equal(g.models[m], r.models[m])
            assert np.allclose(g.r2[m], r.r2[m], rtol=0, atol=1e-12)
            assert np.allclose(g.sigmas[m], r.sigmas[m], rtol=1e-12, atol=0)
            for cg, cr in zip(g.coeffs[m], r.coeffs[m]):
                assert (cg is None) == (cr is None)
                if cg is not None:
                    assert np.array_equal(cg, cr)


@pytest.fixture(scope="module")
def case():
    w = make_workload("tiny", seed=3)
    kw = dict(min_module_size=2, noise_threshold=0.025)
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, **kw)
    return w, kw, ref


def test_lockstep_driver_reproduces_the_oracle_loop(case):
    from scregclust_b200.driver import scregclust_fit
    w, kw, ref = case
    ses = OracleSession(w)
    got = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=ses, **kw)
    _same(got, ref)
    # the final re-fits of all penalization values arrive in ONE batched call, after the last cycle
    assert [c for c in ses.calls if c[1]] == [(sum(r.final_cycle_length for r in ref), True)]
    assert ses.calls[-1][1] is True


def test_forced_last_cycle_after_n_cycles(case):
    from scregclust_b200.driver import scregclust_fit
    w, kw, _ = case
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, n_cycles=2, **kw)
    got = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=OracleSession(w), n_cycles=2, **kw)
    _same(got, ref)
    assert all(g.n_cycles_run <= 3 for g in got)


def test_zero_cycles_refits_the_initial_configuration(case):
    """n_cycles = 0: cycle 1 is already the forced last cycle (R/scregclust.R:1269-1284) -- one re-fit of k_start, no allocation."""
    from scregclust_b200.driver import scregclust_fit
    w, kw, _ = case
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, n_cycles=0, **kw)
    ses = OracleSession(w)
    got = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=ses, n_cycles=0, **kw)
    _same(got, ref)
    assert ses.calls == [(len(w.penalization), True)] and all(not g.converged and g.n_cycles_run == 1 for g in got)


def test_aggregate_branch_through_the_loop(case):
    """allocate_per_obs = FALSE (R/scregclust.R:1678-1728): the loop hands the flag to every fit cycle (not to the final
    re-fit) and reproduces the oracle's trajectory; quick mode is refused with the reference's message."""
    from scregclust_b200.driver import scregclust_fit
    from scregclust_b200._capi import ScrcError
    w, kw, per_obs = case
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, allocate_per_obs=False, **kw)
    ses = OracleSession(w)
    got = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=ses, allocate_per_obs=False, **kw)
    _same(got, ref)
    assert ses.aggregate_calls == sum(1 for c in ses.calls if not c[1])
    assert any(not np.array_equal(a.k_final[0], b.k_final[0]) for a, b in zip(ref, per_obs))   # it is a different rule
    with pytest.raises(ScrcError, match="quick_mode only available when allocate_per_obs is TRUE"):
        scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                       session=OracleSession(w), allocate_per_obs=False, quick_mode=True, **kw)


def test_quick_mode_bookkeeping_of_the_loop(case):
    """quick mode (R/scregclust.R:1435-1496): no subset in cycle 1, afterwards R's `idx` (the latest allocation) goes
    with every call -- also with the final re-fit -- and the trajectory equals the oracle's quick-mode run."""
    from scregclust_b200.driver import scregclust_fit
    w, kw, _ = case
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, quick_mode=True, quick_mode_percent=0.3, **kw)
    ses = OracleSession(w)
    got = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=ses, quick_mode=True, quick_mode_percent=0.3, **kw)
    _same(got, ref)
    assert ses.quick_calls[0] is None and all(q is not None for q in ses.quick_calls[1:])
    assert any(np.any(k == -1) for r in ref for k in r.k_history)        # the case does put targets into the noise module


class ScriptedSession:
    """Answers the cycles from a script of configurations: exercises the history / limit-cycle logic of the
    library's loop (csrc/grid.cu) without any numerics."""

    def __init__(self, T, K, P, script):
        self.T, self.K, self.P, self.script = T, K, P, script
        self.step = 0
        self.final_ks = []

    def fit_cycle(self, lambdas, ks, options=None, skip_allocation=False, want_votes=False, want_nnls_iters=False,
                  prior=None, light=False):
        F, K, T, P = len(lambdas), self.K, self.T, self.P
        ks = np.asarray(ks).reshape(F, T)
        res = {"k_new": np.zeros((F, T), np.int32), "models": np.zeros((F, K, P), np.uint8),
               "best_r2": np.ones((F, T)), "admm_iterations": np.full((F, K), 7, np.int32),
               "n_selected": np.zeros((F, K), np.int32), "votes": None, "nnls_iterations": None,
               "signs": np.full((F, K, P), np.nan), "weights": np.zeros((F, K, P)), "r2_test": np.zeros((F, K, T)),
               "resid_var": np.ones((F, K, T)), "best_r2_idx": np.ones((F, T), np.int32)}
        if skip_allocation:
            self.final_ks = [k.copy() for k in ks]
            # mark every final configuration so that the order of the re-fits can be checked
            for f in range(F):
                res["r2_test"][f] = float(ks[f][0])
        else:
            for f in range(F):
                res["k_new"][f] = self.script[min(self.step, len(self.script) - 1)]
            self.step += 1
        return res

    def coeffs_fit(self, f):
        return [None] * self.K, None

    def close(self):
        pass


def test_limit_cycle_yields_one_final_configuration_per_member():
    """k_start -> A -> B -> C -> B: the new configuration equals history entry 3 while n_k = 5, so the limit cycle
    has length 2 and the re-fit runs on the last two configurations, last one first (R/scregclust.R:1763-1800,
    1307-1311)."""
    from scregclust_b200.driver import scregclust_fit
    T, K, P = 6, 2, 3
    A = np.array([1, 1, 2, 2, 1, 2], np.int32); B = np.array([2, 1, 2, 1, 1, 2], np.int32); Cc = np.array([2, 2, 2, 1, 1, 1], np.int32)
    ses = ScriptedSession(T, K, P, [A, B, Cc, B])
    got = scregclust_fit(penalization=[0.1], n_modules=K, initial_target_modules=np.array([1, 2, 1, 2, 1, 2], np.int32),
                         session=ses, noise_threshold=0.0, return_coeffs=False)[0]
    assert got.converged and got.final_cycle_length == 2 and got.n_cycles_run == 5
    assert [k.tolist() for k in got.k_history] == [[1, 2, 1, 2, 1, 2], A.tolist(), B.tolist(), Cc.tolist(), B.tolist()]
    assert [k.tolist() for k in got.k_final] == [B.tolist(), Cc.tolist()]
    assert got.r2[0][0, 0] == float(B[0]) and got.r2[1][0, 0] == float(Cc[0])
    assert len(got.admm_iterations) == 4 + 2            # four fit cycles, two final re-fits


def test_fixed_point_and_noise_rules():
    """A repeats immediately: converged with a single final configuration; genes below the noise threshold go to -1
    and modules smaller than min_module_size are dissolved before the history comparison (R/scregclust.R:1731-1739)."""
    from scregclust_b200.driver import scregclust_fit
    T, K, P = 6, 3, 2
    A = np.array([1, 1, 1, 2, 2, 3], np.int32)
    ses = ScriptedSession(T, K, P, [A, A])
    got = scregclust_fit(penalization=[0.2], n_modules=K, initial_target_modules=np.ones(T, np.int32), session=ses,
                         noise_threshold=0.0, min_module_size=2, return_coeffs=False)[0]
    want = np.array([1, 1, 1, 2, 2, -1], np.int32)      # module 3 has one member < min_module_size
    assert got.converged and got.final_cycle_length == 1 and got.n_cycles_run == 3
    assert np.array_equal(got.k_history[1], want) and np.array_equal(got.k_final[0], want)


def test_cancel_and_errors_surface_as_exceptions():
    from scregclust_b200.driver import scregclust_fit

    class Failing(ScriptedSession):
        def fit_cycle(self, *a, **k):
            raise RuntimeError("backend failure")
    with pytest.raises(RuntimeError, match="backend failure"):
        scregclust_fit(penalization=[0.1], n_modules=2, initial_target_modules=np.ones(4, np.int32),
                       session=Failing(4, 2, 2, []), return_coeffs=False)
