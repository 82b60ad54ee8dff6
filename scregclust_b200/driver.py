"""Host driver of the alternating fit / reallocate loop on top of the session API.

Mirrors the control flow of ``scregclust()`` (R/scregclust.R:1222-1803) for
``allocate_per_obs = TRUE``, ``quick_mode = FALSE`` and no network prior: module
history, noise threshold, minimum module size, convergence / limit-cycle
detection and the final re-fit of every configuration of a limit cycle.  All
numerics run in ``libscregclust_b200.so``; every penalization value that is still
iterating is advanced by the same batched device call (the fits are independent,
R/scregclust.R:1223,1245).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import CycleOut, Options, Prior, check, dptr, f64, iptr


@dataclass
class FitResult:
    """Per-penalty outcome (subset of the list built in R/scregclust.R:2002-2088)."""
    penalization: float
    converged: bool = False
    n_cycles_run: int = 0
    k_history: list = field(default_factory=list)
    final_cycle_length: int = 1
    k_final: list = field(default_factory=list)
    models: list = field(default_factory=list)        # P x K bool
    signs: list = field(default_factory=list)         # P x K (nan where unselected)
    weights: list = field(default_factory=list)       # P x K
    coeffs: list = field(default_factory=list)        # per module: s_j x T or None
    sigmas: list = field(default_factory=list)        # T x K
    r2: list = field(default_factory=list)            # T x K
    best_r2: list = field(default_factory=list)
    best_r2_idx: list = field(default_factory=list)
    admm_iterations: list = field(default_factory=list)
    nnls_iterations: list = field(default_factory=list)
    votes: list = field(default_factory=list)
    models_history: list = field(default_factory=list)
    r2_module: list = field(default_factory=list)     # per final configuration: K (nan = NA)
    importance: list = field(default_factory=list)    # per final configuration: P x K (nan = NA)
    r2_removed: list = field(default_factory=list)    # per final configuration: list over modules of m_j x (s_j+1)


class Session:
    """Owns a ``scrc_ctx``: data splits, Gram matrices, eigenbasis and initial fit on one GPU."""

    def __init__(self, z1_reg_scaled, z2_reg_scaled, z1_target_centered, z2_target_centered, z1_target_sds,
                 n_modules, device=0):
        self._lib = _capi.load()
        z1r, z2r = f64(z1_reg_scaled), f64(z2_reg_scaled)
        z1t, z2t = f64(z1_target_centered), f64(z2_target_centered)
        sds = np.ascontiguousarray(z1_target_sds, dtype=np.float64)
        n1, P = z1r.shape
        n2, T = z2t.shape
        if z2r.shape != (n2, P) or z1t.shape != (n1, T) or sds.size != T:
            raise ValueError("inconsistent split dimensions")
        self._create(device, z1r.ctypes.data, z2r.ctypes.data, z1t.ctypes.data, z2t.ctypes.data, sds.ctypes.data,
                     n1, n2, P, T, n_modules)
        self.h2d_bytes = z1r.nbytes + z2r.nbytes + z1t.nbytes + z2t.nbytes + sds.nbytes

    @classmethod
    def from_pointers(cls, z1_reg, z2_reg, z1_target, z2_target, z1_target_sds, n1, n2, P, T, n_modules, device=0):
        """Create a session from raw addresses of column-major FP64 buffers.  The buffers may live in
        host memory or already in device memory (e.g. ``tensor.data_ptr()`` of a CUDA tensor); the
        library copies them into its own padded layout either way."""
        self = cls.__new__(cls)
        self._lib = _capi.load()
        self._create(device, int(z1_reg), int(z2_reg), int(z1_target), int(z2_target), int(z1_target_sds),
                     int(n1), int(n2), int(P), int(T), n_modules)
        self.h2d_bytes = 0
        return self

    @classmethod
    def from_expression(cls, expression, is_regulator, split_indices, n_modules, stratification=None, center=True,
                        device=0):
        """Session from the raw ``p x n`` expression matrix (genes x cells): sample split, per-stratum
        centring, constant-gene filter and scaling run on the device (R/scregclust.R:2280-2326,
        899-964, 996-1009).  ``split_indices``: 1 / 2 per cell, anything else (NaN / NA) = unused.
        The surviving genes are in ``self.keep_gene``; ``is_regulator`` refers to the rows of ``expression``."""
        self = cls.__new__(cls)
        self._lib = _capi.load()
        ex = f64(expression)
        p, n = ex.shape
        isr = np.ascontiguousarray(is_regulator, dtype=np.int32)
        si = np.asarray(split_indices, dtype=np.float64)
        si = np.ascontiguousarray(np.where(np.isfinite(si), si, 0.0).astype(np.int32))
        if isr.size != p or si.size != n:
            raise ValueError("is_regulator / split_indices do not match the expression matrix")
        st = None
        if stratification is not None:
            _, st = np.unique(np.asarray(stratification), return_inverse=True)   # sorted levels, like split()
            st = np.ascontiguousarray(st, dtype=np.int32)
            if st.size != n:
                raise ValueError("stratification does not match the expression matrix")
        keep = np.zeros(p, dtype=np.uint8)
        dims = (C.c_int64 * 4)()
        self._h = C.c_void_p()
        check(self._lib.scrc_ctx_create_raw(C.byref(self._h), int(device), dptr(ex), p, n, iptr(isr), iptr(si), iptr(st),
                                            1 if center else 0, int(n_modules),
                                            keep.ctypes.data_as(_capi.c_ubyte_p), dims))
        self.n1, self.n2, self.P, self.T, self.K = int(dims[0]), int(dims[1]), int(dims[2]), int(dims[3]), int(n_modules)
        self.keep_gene = keep.astype(bool)
        self.h2d_bytes = ex.nbytes + isr.nbytes + si.nbytes + (st.nbytes if st is not None else 0)
        self.d2h_bytes = keep.nbytes
        self.admm_sweeps = 0
        self.admm_problem_iterations = 0
        return self

    def inputs(self):
        """The staged matrices as R would hold them: z1_reg_scaled, z2_reg_scaled, z1_target_centered,
        z2_target_centered, z1_target_sds."""
        z1r = np.zeros((self.n1, self.P), order="F"); z2r = np.zeros((self.n2, self.P), order="F")
        z1t = np.zeros((self.n1, self.T), order="F"); z2t = np.zeros((self.n2, self.T), order="F")
        sds = np.zeros(self.T)
        check(self._lib.scrc_ctx_get_inputs(self._h, dptr(z1r), dptr(z2r), dptr(z1t), dptr(z2t), dptr(sds)))
        return z1r, z2r, z1t, z2t, sds

    def _create(self, device, p1r, p2r, p1t, p2t, psd, n1, n2, P, T, n_modules):
        self.n1, self.n2, self.P, self.T, self.K = int(n1), int(n2), int(P), int(T), int(n_modules)
        self._h = C.c_void_p()
        cast = lambda a: C.cast(C.c_void_p(a), _capi.c_double_p)
        check(self._lib.scrc_ctx_create(C.byref(self._h), int(device), cast(p1r), cast(p2r), cast(p1t), cast(p2t),
                                        cast(psd), self.n1, self.n2, self.P, self.T, self.K))
        self.d2h_bytes = 0
        self.admm_sweeps = 0
        self.admm_problem_iterations = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.scrc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ accessors
    def initial_fit(self):
        res_sds = np.zeros(self.T)
        df = C.c_double(0.0)
        beta = np.zeros((self.P, self.T), order="F")
        check(self._lib.scrc_ctx_get_init(self._h, dptr(res_sds), C.byref(df), dptr(beta)))
        return {"res_sds": res_sds, "init_df": df.value, "beta_init": beta}

    def gram(self):
        g_rr = np.zeros((self.P, self.P), order="F")
        g_rt = np.zeros((self.P, self.T), order="F")
        check(self._lib.scrc_ctx_get_gram(self._h, dptr(g_rr), dptr(g_rt)))
        return g_rr, g_rt

    def cross_corr(self):
        out = np.zeros((self.T, self.P), order="F")
        check(self._lib.scrc_ctx_cross_corr(self._h, dptr(out)))
        return out

    # ------------------------------------------------------------------ one batched cycle
    def fit_cycle(self, lambdas, ks, options=None, skip_allocation=False, want_votes=False, want_nnls_iters=False,
                  prior=None, light=False):
        """Step 1 + 2a (+ 2b) for ``F`` (lambda, k) pairs; returns a dict of arrays in R's layouts.

        ``prior``: optional ``(prior_indicator, prior_baseline, prior_weight, update_orders)`` with
        ``prior_indicator`` a list of T lists of 0-based neighbour targets and ``update_orders`` an
        ``F x T`` array of 0-based permutations (R/scregclust.R:1666; src/allocation.cpp:39-56)."""
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        F = lam.size
        k = np.ascontiguousarray(ks, dtype=np.int32).reshape(F, self.T)
        opt = options if options is not None else default_options()
        K, P, T = self.K, self.P, self.T
        res = {
            "k_new": None if skip_allocation else np.zeros((F, T), dtype=np.int32),
            "votes": np.zeros((F, T, K), dtype=np.int32) if (want_votes and not skip_allocation) else None,
            "models": np.zeros((F, K, P), dtype=np.uint8),
            # ``light``: an intermediate cycle of the loop only needs k_new, best_r2 and models back
            "weights": None if light else np.zeros((F, K, P)),
            "signs": None if light else np.zeros((F, K, P)),
            "r2_test": None if light else np.zeros((F, K, T)),
            "resid_var": None if light else np.zeros((F, K, T)),
            "best_r2": np.zeros((F, T)),
            "best_r2_idx": None if light else np.zeros((F, T), dtype=np.int32),
            "admm_iterations": np.zeros((F, K), dtype=np.int32),
            "nnls_iterations": np.zeros((F, K, T), dtype=np.int32) if want_nnls_iters else None,
            "n_selected": np.zeros((F, K), dtype=np.int32),
        }
        sweeps = C.c_longlong(0)
        out = CycleOut()
        out.k_new = iptr(res["k_new"]); out.votes = iptr(res["votes"])
        out.models = res["models"].ctypes.data_as(_capi.c_ubyte_p)
        out.weights = dptr(res["weights"]); out.signs = dptr(res["signs"])
        out.r2_test = dptr(res["r2_test"]); out.resid_var = dptr(res["resid_var"])
        out.best_r2 = dptr(res["best_r2"]); out.best_r2_idx = iptr(res["best_r2_idx"])
        out.admm_iterations = iptr(res["admm_iterations"]); out.nnls_iterations = iptr(res["nnls_iterations"])
        out.n_selected = iptr(res["n_selected"])
        out.admm_sweeps = C.pointer(sweeps)
        if prior is not None and not skip_allocation and float(prior[2]) != 0.0:
            from .api import _csr
            off, idx = _csr(prior[0], T)
            if off is None:
                off = np.zeros(T + 1, dtype=np.int32); idx = np.zeros(1, dtype=np.int32)
            orders = np.ascontiguousarray(prior[3], dtype=np.int32).reshape(F, T)
            pr = Prior(iptr(off), iptr(idx), float(prior[1]), float(prior[2]), iptr(orders))
            check(self._lib.scrc_fit_cycle_prior(self._h, F, dptr(lam), iptr(k), C.byref(opt), C.byref(pr), C.byref(out)))
        else:
            check(self._lib.scrc_fit_cycle(self._h, F, dptr(lam), iptr(k), C.byref(opt), 1 if skip_allocation else 0,
                                           C.byref(out)))
        self.h2d_bytes += lam.nbytes + k.nbytes
        self.d2h_bytes += sum(v.nbytes for v in res.values() if v is not None)
        self.admm_sweeps += int(sweeps.value)
        self.admm_problem_iterations += int(res["admm_iterations"].sum())
        return res

    def importance(self, k_final, models, signs, options=None, want_r2_removed=False):
        """Leave-one-regulator-out importance (R/scregclust.R:1823-1912) for F final configurations.

        ``k_final``: F x T, ``models``: F x K x P (bool/uint8), ``signs``: F x K x P (nan where unselected).
        Returns ``(r2_module[F, K], importance[F, K, P])`` with nan where the reference leaves NA and, with
        ``want_r2_removed``, a third item: per fit a list over modules of the ``m_j x (s_j + 1)`` matrices
        ``r2_removed`` (``None`` where the reference leaves the list entry empty)."""
        k = np.ascontiguousarray(k_final, dtype=np.int32).reshape(-1, self.T)
        F = k.shape[0]
        md = np.ascontiguousarray(models, dtype=np.uint8).reshape(F, self.K, self.P)
        sg = np.ascontiguousarray(signs, dtype=np.float64).reshape(F, self.K, self.P)
        opt = options if options is not None else default_options()
        r2m = np.zeros((F, self.K)); imp = np.zeros((F, self.K, self.P))
        self.h2d_bytes += k.nbytes + md.nbytes + sg.nbytes
        self.d2h_bytes += r2m.nbytes + imp.nbytes
        if not want_r2_removed:
            check(self._lib.scrc_importance(self._h, F, iptr(k), md.ctypes.data_as(_capi.c_ubyte_p), dptr(sg),
                                            C.byref(opt), dptr(r2m), dptr(imp)))
            return r2m, imp
        shapes = []                                   # (f, j, m_j, s_j) of every block, in the library's order
        for f in range(F):
            for j in range(self.K):
                m_j, s_j = int(np.count_nonzero(k[f] == j + 1)), int(md[f, j].sum())
                if m_j > 0 and s_j > 1:
                    shapes.append((f, j, m_j, s_j))
        cap = sum(m_j * (s_j + 1) for _, _, m_j, s_j in shapes)
        buf = np.zeros(max(cap, 1))
        check(self._lib.scrc_importance_ex(self._h, F, iptr(k), md.ctypes.data_as(_capi.c_ubyte_p), dptr(sg),
                                           C.byref(opt), dptr(r2m), dptr(imp), dptr(buf), cap))
        self.d2h_bytes += cap * 8
        removed = [[None] * self.K for _ in range(F)]
        o = 0
        for f, j, m_j, s_j in shapes:
            removed[f][j] = buf[o:o + m_j * (s_j + 1)].reshape((m_j, s_j + 1), order="F")
            o += m_j * (s_j + 1)
        return r2m, imp, removed

    def coeffs(self, f, module):
        """NNLS coefficients (s x T) and 0-based regulator indices of ``module`` (0-based) of fit ``f``."""
        s = self._last_nsel[f, module]
        if s == 0:
            return None, np.zeros(0, dtype=np.int32)
        co = np.zeros((s, self.T), order="F")
        sel = np.zeros(s, dtype=np.int32)
        check(self._lib.scrc_ctx_get_coeffs(self._h, int(f), int(module), dptr(co), iptr(sel)))
        self.d2h_bytes += co.nbytes + sel.nbytes
        return co, sel


    def coeffs_fit(self, f):
        """All modules of fit ``f`` in one transfer: list of K coefficient matrices (s_k x T, ``None``
        where the module has no model) -- views into one packed buffer -- and the list of index vectors."""
        ns = [int(v) for v in self._last_nsel[f]]
        tot = sum(ns)
        buf = np.zeros(max(tot, 1) * self.T)
        sel = np.zeros(max(tot, 1), dtype=np.int32)
        check(self._lib.scrc_ctx_get_coeffs_fit(self._h, int(f), dptr(buf), iptr(sel)))
        self.d2h_bytes += tot * self.T * 8 + tot * 4
        cos, sels, o, r = [], [], 0, 0
        for s_ in ns:
            cos.append(buf[o:o + s_ * self.T].reshape((s_, self.T), order="F") if s_ else None)
            sels.append(sel[r:r + s_].copy())
            o += s_ * self.T; r += s_
        return cos, sels


def default_options(**kw) -> Options:
    o = Options()
    _capi.load().scrc_default_options(C.byref(o))
    for k_, v in kw.items():
        setattr(o, k_, v)
    return o


@dataclass
class _PenState:
    k: np.ndarray
    k_history: list
    n_k: int = 2
    last_cycle: bool = False
    final_cycle_length: int = 1
    converged: bool = False
    cycle: int = 0
    done: bool = False


def _unpack(fr, res, f, K, keep_diag):
    fr.admm_iterations.append(res["admm_iterations"][f].copy())
    fr.models_history.append(res["models"][f].T.astype(bool))
    if keep_diag and res["nnls_iterations"] is not None:
        fr.nnls_iterations.append(res["nnls_iterations"][f].copy())


def scregclust_fit(z1_reg_scaled=None, z2_reg_scaled=None, z1_target_centered=None, z2_target_centered=None,
                   z1_target_sds=None, penalization=(), n_modules=1, initial_target_modules=None,
                   min_module_size=0, noise_threshold=0.025, n_cycles=50, max_optim_iter=10000,
                   tol_coop_rel=1e-8, tol_coop_abs=1e-12, tol_nnls=1e-4, device=0, session=None,
                   keep_diagnostics=False, return_coeffs=True, claim=None, concurrency=None,
                   compute_predictive_r2=False, prior_indicator=None, prior_baseline=1e-6, prior_weight=0.0,
                   update_orders=None):
    """The alternating loop of ``scregclust()`` (R/scregclust.R:1222-1803) for all penalization values.

    Returns one :class:`FitResult` per penalization value, in order.  ``session`` may be passed to
    reuse the device-resident precompute (R/scregclust.R:996-1183) across calls.

    Network prior (R/scregclust.R:584-615): ``prior_indicator`` is a list of T lists of 0-based
    neighbour targets, ``update_orders(penalty_index, cycle)`` returns the permutation R's
    ``sample()`` would draw (R/scregclust.R:1666; identity when omitted).

    Multi-GPU work sharing: the fits of different penalization values are independent
    (R/scregclust.R:1223,1245), so ``claim`` may be a callable returning the index of the next
    penalization value this process should fit (or ``None`` when the grid is exhausted) --
    e.g. an atomic counter shared by all ranks -- and ``concurrency`` bounds how many claimed
    values are advanced together by one batched device call.  Values never claimed here come
    back as ``None``.  A fit's result does not depend on what it is batched with.
    """
    own = session is None
    ses = session or Session(z1_reg_scaled, z2_reg_scaled, z1_target_centered, z2_target_centered, z1_target_sds,
                             n_modules, device=device)
    try:
        K, T = ses.K, ses.T
        opt = default_options(max_optim_iter=int(max_optim_iter), tol_coop_rel=float(tol_coop_rel),
                              tol_coop_abs=float(tol_coop_abs), tol_nnls=float(tol_nnls))
        pens = [float(p) for p in penalization]
        k_start = np.asarray(initial_target_modules, dtype=np.int32)
        if k_start.size != T:
            raise ValueError("initial_target_modules needs one entry per target gene")
        if claim is None:
            queue = iter(range(len(pens)))
            claim = lambda: next(queue, None)
        concurrency = len(pens) if concurrency is None else max(1, int(concurrency))
        states, results = {}, [None] * len(pens)
        active, pending_final = [], []

        while True:
            while len(active) < concurrency:
                i = claim()
                if i is None:
                    break
                states[i] = _PenState(k=k_start.copy(), k_history=[k_start.copy()])
                results[i] = FitResult(penalization=pens[i])
                results[i].k_history = states[i].k_history
                active.append(i)
            if not active:
                break
            # R/scregclust.R:1269-1284: cycle counter, forced last cycle after n_cycles
            for i in active:
                s = states[i]
                s.cycle += 1
                if s.cycle == n_cycles + 1:
                    s.last_cycle = True
            run = [i for i in active if not states[i].last_cycle]
            fin = [i for i in active if states[i].last_cycle]

            if run:
                prior = None
                if prior_indicator is not None and len(prior_indicator) > 0 and prior_weight != 0.0:
                    orders = np.stack([np.asarray(update_orders(i, states[i].cycle), dtype=np.int32)
                                       if update_orders is not None else np.arange(T, dtype=np.int32) for i in run])
                    prior = (prior_indicator, prior_baseline, prior_weight, orders)
                res = ses.fit_cycle([pens[i] for i in run], np.stack([states[i].k for i in run]), opt,
                                    skip_allocation=False, want_votes=keep_diagnostics,
                                    want_nnls_iters=keep_diagnostics, prior=prior, light=True)
                for f, i in enumerate(run):
                    s, fr = states[i], results[i]
                    _unpack(fr, res, f, K, keep_diagnostics)
                    if keep_diagnostics:
                        fr.votes.append(res["votes"][f].copy())
                    idx = res["k_new"][f].copy()
                    # noise module and minimum module size (R/scregclust.R:1731-1739)
                    idx[res["best_r2"][f] < noise_threshold] = -1
                    sizes = np.bincount(idx[idx > 0], minlength=K + 1)[1:]
                    small = np.flatnonzero(sizes < min_module_size) + 1
                    if small.size:
                        idx[np.isin(idx, small)] = -1
                    k = idx.astype(np.int32)
                    # convergence / limit cycle (R/scregclust.R:1763-1802)
                    matching = [j for j, kh in enumerate(s.k_history, start=1) if np.array_equal(kh, k)]
                    s.k_history.append(k.copy())
                    if matching:
                        s.converged = True
                        if s.n_k - matching[0] != 1:
                            s.final_cycle_length = s.n_k - matching[0]
                        s.last_cycle = True
                    s.n_k += 1
                    s.k = k
                    fr.n_cycles_run = s.cycle

            # Penalization values entering their last cycle only need the re-fit of their final
            # configuration(s) (R/scregclust.R:1307-1311, Step 2b skipped) and the post-processing.
            # Nothing depends on those results, so they are deferred and batched into as few device
            # calls as possible instead of one small call whenever a value happens to finish.
            pending_final.extend(fin)
            active = [i for i in active if i not in fin]

        def finalize(batch):
            lam_f, k_f, owner = [], [], []
            for i in batch:
                s = states[i]
                for m in range(1, s.final_cycle_length + 1):     # R/scregclust.R:1307-1311
                    lam_f.append(pens[i]); k_f.append(s.k_history[len(s.k_history) - m]); owner.append(i)
            res = ses.fit_cycle(lam_f, np.stack(k_f), opt, skip_allocation=True,
                                want_nnls_iters=keep_diagnostics)
            ses._last_nsel = res["n_selected"]
            for f, i in enumerate(owner):
                s, fr = states[i], results[i]
                _unpack(fr, res, f, K, keep_diagnostics)
                fr.k_final.append(np.asarray(k_f[f]).copy())
                fr.models.append(res["models"][f].T.astype(bool))
                fr.signs.append(res["signs"][f].T.copy())
                fr.weights.append(res["weights"][f].T.copy())
                fr.sigmas.append(np.sqrt(res["resid_var"][f].T))
                fr.r2.append(res["r2_test"][f].T.copy())
                fr.best_r2.append(res["best_r2"][f].copy())
                fr.best_r2_idx.append(res["best_r2_idx"][f].copy())
                if return_coeffs:
                    fr.coeffs.append(ses.coeffs_fit(f)[0])
            if compute_predictive_r2:                             # R/scregclust.R:1823-1912
                r2m, imp, rem = ses.importance(np.stack(k_f), res["models"], res["signs"], opt, want_r2_removed=True)
                for f, i in enumerate(owner):
                    results[i].r2_module.append(r2m[f].copy())
                    results[i].importance.append(imp[f].T.copy())
                    results[i].r2_removed.append(rem[f])
            for i in batch:
                s, fr = states[i], results[i]
                s.done = True
                fr.converged = s.converged
                fr.final_cycle_length = s.final_cycle_length
                fr.n_cycles_run = s.cycle

        # at most ~2x the regular batch width per call (bounds the state matrices of the re-fit)
        batch, width = [], 0
        for i in pending_final:
            w_i = states[i].final_cycle_length
            if batch and width + w_i > 2 * concurrency:
                finalize(batch)
                batch, width = [], 0
            batch.append(i); width += w_i
        if batch:
            finalize(batch)
        return results
    finally:
        if own:
            ses.close()
