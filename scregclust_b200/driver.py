"""Python caller of the session and whole-grid entry points of ``libscregclust_b200.so``.

The alternating loop of ``scregclust()`` (R/scregclust.R:1222-1912: module history, noise
threshold, minimum module size, convergence / limit-cycle detection, the re-fit of every final
configuration and the importance step) runs inside the library (``scrc_grid_fit``,
csrc/grid.cu); :func:`scregclust_fit` marshals its arguments and unpacks the result handle into
:class:`FitResult` objects.  :class:`Session` exposes the per-cycle entry points the parity tests
drive directly, and :class:`Comm` the NCCL communicator of a multi-GPU job.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import CycleOut, Options, Prior, Quick, check, dptr, f64, iptr


class Comm:
    """One rank of a multi-GPU job (one rank per GPU; NCCL over NVLink).  ``unique_id`` is the
    128-byte token rank 0 obtains from :meth:`new_unique_id` and hands to the other ranks."""

    def __init__(self, unique_id: bytes, world: int, rank: int, device: int):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.world, self.rank, self.device = int(world), int(rank), int(device)
        check(self._lib.scrc_comm_create(C.byref(self._h), C.c_char_p(bytes(unique_id)), self.world, self.rank, self.device))

    @staticmethod
    def new_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_capi.load().scrc_comm_unique_id(buf))
        return buf.raw

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.scrc_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
                 n_modules, device=0, comm=None):
        self._lib = _capi.load()
        self._comm = comm
        z1r, z2r = f64(z1_reg_scaled), f64(z2_reg_scaled)
        z1t, z2t = f64(z1_target_centered), f64(z2_target_centered)
        sds = np.ascontiguousarray(z1_target_sds, dtype=np.float64)
        n1, P = z1r.shape
        n2, T = z2t.shape
        if z2r.shape != (n2, P) or z1t.shape != (n1, T) or sds.size != T:
            raise ValueError("inconsistent split dimensions")
        self._create(device, z1r.ctypes.data, z2r.ctypes.data, z1t.ctypes.data, z2t.ctypes.data, sds.ctypes.data,
                     n1, n2, P, T, n_modules, comm)
        self.h2d_bytes = z1r.nbytes + z2r.nbytes + z1t.nbytes + z2t.nbytes + sds.nbytes

    @classmethod
    def from_pointers(cls, z1_reg, z2_reg, z1_target, z2_target, z1_target_sds, n1, n2, P, T, n_modules, device=0,
                      comm=None):
        """Create a session from raw addresses of column-major FP64 buffers.  The buffers may live in
        host memory or already in device memory (e.g. ``tensor.data_ptr()`` of a CUDA tensor); the
        library copies them into its own padded layout either way."""
        self = cls.__new__(cls)
        self._lib = _capi.load()
        self._comm = comm
        self._create(device, int(z1_reg), int(z2_reg), int(z1_target), int(z2_target), int(z1_target_sds),
                     int(n1), int(n2), int(P), int(T), n_modules, comm)
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
        self._comm = None
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

    def _create(self, device, p1r, p2r, p1t, p2t, psd, n1, n2, P, T, n_modules, comm=None):
        self.n1, self.n2, self.P, self.T, self.K = int(n1), int(n2), int(P), int(T), int(n_modules)
        self._h = C.c_void_p()
        cast = lambda a: C.cast(C.c_void_p(a), _capi.c_double_p)
        if comm is not None:      # collective: rank 0 builds the precompute and broadcasts it
            check(self._lib.scrc_ctx_create_sharded(C.byref(self._h), comm._h, cast(p1r), cast(p2r), cast(p1t), cast(p2t),
                                                    cast(psd), self.n1, self.n2, self.P, self.T, self.K))
        else:
            check(self._lib.scrc_ctx_create(C.byref(self._h), int(device), cast(p1r), cast(p2r), cast(p1t), cast(p2t),
                                            cast(psd), self.n1, self.n2, self.P, self.T, self.K))
        self.d2h_bytes = 0
        self.admm_sweeps = 0
        self.admm_problem_iterations = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.scrc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def request_cancel(self):
        """Cooperative interrupt (the reference polls Rcpp::checkUserInterrupt()): the running / next
        device call of this session fails with SCRC_ERR_INTERRUPTED.  Safe from another thread."""
        check(self._lib.scrc_ctx_request_cancel(self._h))

    def clear_cancel(self):
        check(self._lib.scrc_ctx_clear_cancel(self._h))

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

    def kmeanspp(self, first_centers, uniforms, n_max_iter=100, return_details=False):
        """``k_start_cl <- kmeanspp(cross_corr1, n_modules, n_initializations, 100L)`` (R/scregclust.R:1131-1136) on the
        session's own cross_corr1 (it never leaves the device).  ``first_centers`` / ``uniforms``: R's draws, see
        :func:`scregclust_b200.api.kmeanspp`."""
        fc = np.ascontiguousarray(first_centers, dtype=np.int32)
        u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(fc.size, self.K - 1)
        cl = np.zeros(self.T, dtype=np.int32); tot = np.zeros(fc.size); fault = np.zeros(fc.size, dtype=np.int32)
        check(self._lib.scrc_ctx_kmeanspp(self._h, int(fc.size), int(n_max_iter), iptr(fc), dptr(u), iptr(cl), dptr(tot),
                                          iptr(fault)))
        return (cl, tot, fault) if return_details else cl

    def cross_corr(self):
        out = np.zeros((self.T, self.P), order="F")
        check(self._lib.scrc_ctx_cross_corr(self._h, dptr(out)))
        return out

    # ------------------------------------------------------------------ one batched cycle
    def fit_cycle(self, lambdas, ks, options=None, skip_allocation=False, want_votes=False, want_nnls_iters=False,
                  prior=None, light=False, quick=None, aggregate=False):
        """Step 1 + 2a (+ 2b) for ``F`` (lambda, k) pairs; returns a dict of arrays in R's layouts.

        ``aggregate``: Step 2b by the ``allocate_per_obs = FALSE`` branch (R/scregclust.R:1678-1728) -- from the cycle's
        test sums of squares and variances, no votes; ``prior`` is optional then.

        ``quick``: optional ``(idx_latest[F, T], quick_mode_percent)`` -- quick mode (R/scregclust.R:1435-1496).

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
        if aggregate and quick is not None:
            raise ValueError("quick_mode only available when allocate_per_obs is TRUE.")
        if prior is not None and not skip_allocation and float(prior[2]) != 0.0:
            from .api import _csr
            off, idx = _csr(prior[0], T)
            if off is None:
                off = np.zeros(T + 1, dtype=np.int32); idx = np.zeros(1, dtype=np.int32)
            orders = np.ascontiguousarray(prior[3], dtype=np.int32).reshape(F, T)
            pr = Prior(iptr(off), iptr(idx), float(prior[1]), float(prior[2]), iptr(orders), 1 if aggregate else 0)
            check(self._lib.scrc_fit_cycle_prior(self._h, F, dptr(lam), iptr(k), C.byref(opt), C.byref(pr), C.byref(out)))
        elif aggregate and not skip_allocation:
            pr = Prior(None, None, float(prior[1]) if prior is not None else 1e-6, 0.0, None, 1)
            check(self._lib.scrc_fit_cycle_prior(self._h, F, dptr(lam), iptr(k), C.byref(opt), C.byref(pr), C.byref(out)))
        elif quick is not None:
            idx = np.ascontiguousarray(quick[0], dtype=np.int32).reshape(F, self.T)
            qk = Quick(iptr(idx), float(quick[1]))
            check(self._lib.scrc_fit_cycle_quick(self._h, F, dptr(lam), iptr(k), C.byref(opt), 1 if skip_allocation else 0,
                                                 C.byref(qk), C.byref(out)))
        else:
            check(self._lib.scrc_fit_cycle(self._h, F, dptr(lam), iptr(k), C.byref(opt), 1 if skip_allocation else 0,
                                           C.byref(out)))
        self.h2d_bytes += lam.nbytes + k.nbytes
        self.d2h_bytes += sum(v.nbytes for v in res.values() if v is not None)
        self.admm_sweeps += int(sweeps.value)
        self.admm_problem_iterations += int(res["admm_iterations"].sum())
        self._last_nsel = res["n_selected"]                 # coeffs_fit() unpacks the last cycle's coefficients with it
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


def _csr_lists(off, idx, T):
    return [list(idx[off[t]:off[t + 1]]) for t in range(T)]


class _PyBackend:
    """Adapts a duck-typed session (``fit_cycle`` / ``importance`` / ``coeffs_fit`` returning the
    dictionaries of :class:`Session`) to the callback table of ``scrc_grid_fit_custom``.  Used by the
    CPU tests, which answer the cycles from the oracle."""

    def __init__(self, ses, keep_diag):
        self.ses, self.keep_diag = ses, keep_diag
        self.error = None
        K, P, T = ses.K, ses.P, ses.T
        self._keep = []

        def fill(ptr, arr, dtype):
            if ptr and arr is not None:
                a = np.ascontiguousarray(arr, dtype=dtype)
                C.memmove(ptr, a.ctypes.data, a.nbytes)

        def fit_cycle(user, F, lam, k_in, opt, skip, hint, prior, quick, out):
            try:
                lams = np.ctypeslib.as_array(lam, shape=(F,)).copy()
                ks = np.ctypeslib.as_array(k_in, shape=(F, T)).copy()
                o = out.contents
                pr = None
                kw = {}
                if prior:
                    pc = prior.contents
                    if pc.prior_off and pc.update_order:
                        off = np.ctypeslib.as_array(pc.prior_off, shape=(T + 1,)).copy()
                        idx = np.ctypeslib.as_array(pc.prior_idx, shape=(max(int(off[T]), 1),)).copy()
                        orders = np.ctypeslib.as_array(pc.update_order, shape=(F, T)).copy()
                        pr = (_csr_lists(off, idx, T), pc.baseline, pc.weight, orders)
                    if pc.aggregate:
                        kw["aggregate"] = True
                if quick:
                    kw["quick"] = (np.ctypeslib.as_array(quick.contents.idx_latest, shape=(F, T)).copy(), quick.contents.percent)
                if getattr(ses, "accepts_hint", False) and hint:
                    kw["hint"] = np.ctypeslib.as_array(hint, shape=(F, K)).copy()
                res = ses.fit_cycle(lams, ks, opt.contents, skip_allocation=bool(skip), want_votes=bool(o.votes),
                                    want_nnls_iters=bool(o.nnls_iterations), prior=pr, light=False, **kw)
                if not skip:
                    fill(o.k_new, res["k_new"], np.int32); fill(o.votes, res.get("votes"), np.int32)
                fill(o.models, res["models"], np.uint8); fill(o.weights, res.get("weights"), np.float64)
                fill(o.signs, res.get("signs"), np.float64); fill(o.r2_test, res.get("r2_test"), np.float64)
                fill(o.resid_var, res.get("resid_var"), np.float64); fill(o.best_r2, res["best_r2"], np.float64)
                fill(o.best_r2_idx, res.get("best_r2_idx"), np.int32); fill(o.admm_iterations, res["admm_iterations"], np.int32)
                fill(o.nnls_iterations, res.get("nnls_iterations"), np.int32); fill(o.n_selected, res["n_selected"], np.int32)
                if o.admm_sweeps:
                    o.admm_sweeps[0] = int(res.get("admm_sweeps", 0))
                return 0
            except Exception as exc:                     # never unwind through the C frames
                self.error = exc
                return -50

        def importance(user, F, k_final, models, signs, opt, r2m, imp, rem, cap):
            try:
                kf = np.ctypeslib.as_array(k_final, shape=(F, T)).copy()
                md = np.ctypeslib.as_array(models, shape=(F, K, P)).copy()
                sg = np.ctypeslib.as_array(signs, shape=(F, K, P)).copy()
                r2, im, removed = ses.importance(kf, md, sg, opt.contents, want_r2_removed=True)
                fill(r2m, r2, np.float64); fill(imp, im, np.float64)
                flat = [np.asarray(b, dtype=np.float64).ravel(order="F") for f in range(F) for b in removed[f] if b is not None]
                if flat:
                    buf = np.concatenate(flat)
                    if buf.size > cap:
                        raise ValueError("r2_removed does not fit the buffer")
                    fill(rem, buf, np.float64)
                return 0
            except Exception as exc:
                self.error = exc
                return -50

        def coeffs_fit(user, f, coeffs_out, sel_out):
            try:
                cos, sels = ses.coeffs_fit(f)
                flat = [np.asarray(c, dtype=np.float64).ravel(order="F") for c in cos if c is not None]
                if flat:
                    fill(coeffs_out, np.concatenate(flat), np.float64)
                if sels is not None:
                    sl = [np.asarray(x, dtype=np.int32) for x in sels if x is not None and len(x)]
                    if sl:
                        fill(sel_out, np.concatenate(sl), np.int32)
                return 0
            except Exception as exc:
                self.error = exc
                return -50

        self.table = _capi.GridBackend(None, P, T, K, _capi.BACKEND_FIT_CYCLE_FN(fit_cycle),
                                       _capi.BACKEND_IMPORTANCE_FN(importance), _capi.BACKEND_COEFFS_FN(coeffs_fit))


def _unpack_grid(lib, handle, K, P, T, keep_diag, return_coeffs, compute_r2):
    """scrc_grid_result -> list of FitResult (R's layouts: P x K, T x K)."""
    info = _capi.GridInfo()
    check(lib.scrc_grid_result_info(handle, C.byref(info)))
    out = []
    d2h = 0
    for i in range(info.n_penalties):
        pi = _capi.GridPenaltyInfo()
        check(lib.scrc_grid_result_penalty(handle, i, C.byref(pi)))
        fr = FitResult(penalization=pi.penalization, converged=bool(pi.converged), n_cycles_run=pi.n_cycles_run,
                       final_cycle_length=pi.final_cycle_length)
        # (every buffer below is filled completely by its accessor: np.empty, and transposed views instead of copies)
        hist = np.empty((pi.n_history, T), dtype=np.int32)
        check(lib.scrc_grid_result_history(handle, i, iptr(hist)))
        fr.k_history = [hist[h].copy() for h in range(pi.n_history)]
        n_fit_records = pi.n_records - pi.final_cycle_length
        ai = np.empty((pi.n_records, K), dtype=np.int32)
        md = np.empty((pi.n_records, K, P), dtype=np.uint8)
        check(lib.scrc_grid_result_records(handle, i, iptr(ai), None, md.ctypes.data_as(_capi.c_ubyte_p)))
        mdb = md.view(bool).transpose(0, 2, 1)               # views: P x K per record (the bytes are 0 / 1)
        fr.admm_iterations = [ai[rec] for rec in range(pi.n_records)]
        fr.models_history = [mdb[rec] for rec in range(pi.n_records)]
        if keep_diag:
            for rec in range(pi.n_records):
                vt = np.zeros((T, K), dtype=np.int32) if rec < n_fit_records else None
                ni = np.zeros((K, T), dtype=np.int32)
                check(lib.scrc_grid_result_record(handle, i, rec, None, None, None, iptr(vt), iptr(ni)))
                fr.nnls_iterations.append(ni)
                if vt is not None:
                    fr.votes.append(vt)
        for m in range(pi.final_cycle_length):
            kf = np.empty(T, dtype=np.int32); md = np.empty((K, P), dtype=np.uint8)
            wt = np.empty((K, P)); sg = np.empty((K, P)); r2 = np.empty((K, T)); rv = np.empty((K, T))
            br = np.empty(T); bi = np.empty(T, dtype=np.int32); ns = np.empty(K, dtype=np.int32)
            r2m = np.empty(K) if compute_r2 else None
            imp = np.empty((K, P)) if compute_r2 else None
            gf = _capi.GridFinal(iptr(kf), md.ctypes.data_as(_capi.c_ubyte_p), dptr(wt), dptr(sg), dptr(r2), dptr(rv),
                                 dptr(br), iptr(bi), iptr(ns), dptr(r2m), dptr(imp))
            check(lib.scrc_grid_result_final(handle, i, m, C.byref(gf)))
            d2h += sum(a.nbytes for a in (kf, md, wt, sg, r2, rv, br, bi, ns))
            fr.k_final.append(kf); fr.models.append(md.view(bool).T); fr.signs.append(sg.T)
            fr.weights.append(wt.T); fr.sigmas.append(np.sqrt(rv).T); fr.r2.append(r2.T)
            fr.best_r2.append(br); fr.best_r2_idx.append(bi)
            if return_coeffs:
                tot = int(ns.sum())
                buf = np.zeros(max(tot, 1) * T); sel = np.zeros(max(tot, 1), dtype=np.int32)
                check(lib.scrc_grid_result_coeffs(handle, i, m, dptr(buf), iptr(sel)))
                d2h += tot * T * 8 + tot * 4
                cos, o = [], 0
                for s_ in (int(v) for v in ns):
                    cos.append(buf[o:o + s_ * T].reshape((s_, T), order="F") if s_ else None)
                    o += s_ * T
                fr.coeffs.append(cos)
            if compute_r2:
                fr.r2_module.append(r2m); fr.importance.append(imp.T)
                n = int(lib.scrc_grid_result_r2_removed_size(handle, i, m))
                buf = np.zeros(max(n, 1))
                check(lib.scrc_grid_result_r2_removed(handle, i, m, dptr(buf)))
                d2h += r2m.nbytes + imp.nbytes + n * 8
                removed, o = [None] * K, 0
                for j in range(K):
                    m_j, s_j = int(np.count_nonzero(kf == j + 1)), int(ns[j])
                    if m_j > 0 and s_j > 1:
                        removed[j] = buf[o:o + m_j * (s_j + 1)].reshape((m_j, s_j + 1), order="F")
                        o += m_j * (s_j + 1)
                fr.r2_removed.append(removed)
        out.append(fr)
    return out, info, d2h


def scregclust_fit(z1_reg_scaled=None, z2_reg_scaled=None, z1_target_centered=None, z2_target_centered=None,
                   z1_target_sds=None, penalization=(), n_modules=1, initial_target_modules=None,
                   min_module_size=0, noise_threshold=0.025, n_cycles=50, max_optim_iter=10000,
                   tol_coop_rel=1e-8, tol_coop_abs=1e-12, tol_nnls=1e-4, device=0, session=None,
                   keep_diagnostics=False, return_coeffs=True, compute_predictive_r2=False, prior_indicator=None,
                   prior_baseline=1e-6, prior_weight=0.0, update_orders=None, comm=None, quick_mode=False,
                   quick_mode_percent=0.1, allocate_per_obs=True):
    """The alternating loop of ``scregclust()`` (R/scregclust.R:1222-1912) for all penalization values,
    one ``scrc_grid_fit`` call.  Returns one :class:`FitResult` per penalization value, in order.

    ``session`` may be passed to reuse the device-resident precompute (R/scregclust.R:996-1183) across
    calls; any object with the ``fit_cycle`` / ``importance`` / ``coeffs_fit`` methods of
    :class:`Session` is accepted and then driven through ``scrc_grid_fit_custom`` (the CPU tests pass an
    oracle-backed one).  Network prior (R/scregclust.R:584-615): ``prior_indicator`` is a list of T lists of
    0-based neighbour targets, ``update_orders(penalty_index, cycle)`` returns the permutation R's
    ``sample()`` would draw (R/scregclust.R:1666; identity when omitted).  ``allocate_per_obs=False`` selects the
    aggregate branch of Step 2b (R/scregclust.R:1678-1728; unreachable through the reference's own argument checks).

    Multi-GPU: with ``comm`` (or a session created with one) the call is collective -- every rank calls
    it with the same arguments and every rank gets the complete result.
    """
    lib = _capi.load()
    own = session is None
    ses = session or Session(z1_reg_scaled, z2_reg_scaled, z1_target_centered, z2_target_centered, z1_target_sds,
                             n_modules, device=device, comm=comm)
    try:
        K, T, P = ses.K, ses.T, ses.P
        pens = np.ascontiguousarray([float(p) for p in penalization], dtype=np.float64)
        k_start = np.ascontiguousarray(initial_target_modules, dtype=np.int32)
        if k_start.size != T:
            raise ValueError("initial_target_modules needs one entry per target gene")
        gp = _capi.GridParams()
        lib.scrc_grid_default_params(C.byref(gp))
        gp.n_penalties = pens.size; gp.penalization = dptr(pens); gp.initial_target_modules = iptr(k_start)
        gp.n_cycles = int(n_cycles); gp.noise_threshold = float(noise_threshold); gp.min_module_size = int(min_module_size)
        gp.opt = default_options(max_optim_iter=int(max_optim_iter), tol_coop_rel=float(tol_coop_rel),
                                 tol_coop_abs=float(tol_coop_abs), tol_nnls=float(tol_nnls))
        gp.compute_predictive_r2 = 1 if compute_predictive_r2 else 0
        gp.keep_coeffs = 1 if return_coeffs else 0
        gp.keep_diagnostics = 1 if keep_diagnostics else 0
        gp.quick_mode = 1 if quick_mode else 0
        gp.quick_mode_percent = float(quick_mode_percent)
        gp.allocate_per_obs = 1 if allocate_per_obs else 0
        keep = []
        if prior_indicator is not None and len(prior_indicator) > 0 and float(prior_weight) != 0.0:
            from .api import _csr
            off, idx = _csr(prior_indicator, T)
            if off is None:
                off = np.zeros(T + 1, dtype=np.int32); idx = np.zeros(1, dtype=np.int32)
            gp.prior_off = iptr(off); gp.prior_idx = iptr(idx)
            gp.prior_baseline = float(prior_baseline); gp.prior_weight = float(prior_weight)
            keep += [off, idx]
            if update_orders is not None:
                def _order(user, penalty, cycle):
                    o = np.ascontiguousarray(update_orders(int(penalty), int(cycle)), dtype=np.int32)
                    keep.append(o)
                    return o.ctypes.data
                gp.update_order = _capi.UPDATE_ORDER_FN(_order)
        handle = C.c_void_p()
        if isinstance(ses, Session):
            check(lib.scrc_grid_fit(ses._h, C.byref(gp), C.byref(handle)))
        else:
            be = _PyBackend(ses, keep_diagnostics)
            rc = lib.scrc_grid_fit_custom(C.byref(be.table), C.byref(gp), C.byref(handle))
            if be.error is not None:
                raise be.error
            check(rc)
        try:
            results, info, d2h = _unpack_grid(lib, handle, K, P, T, keep_diagnostics, return_coeffs, compute_predictive_r2)
        finally:
            lib.scrc_grid_result_free(handle)
        if isinstance(ses, Session):
            ses.h2d_bytes += pens.nbytes + k_start.nbytes
            ses.d2h_bytes += d2h
            ses.admm_sweeps += int(info.admm_sweeps)
            ses.admm_problem_iterations += int(info.admm_problem_iterations)
        ses.last_grid_info = {"fit_cycles": int(info.fit_cycles), "final_refits": int(info.final_refits),
                              "device_calls": int(info.device_calls), "admm_sweeps": int(info.admm_sweeps),
                              "admm_problem_iterations": int(info.admm_problem_iterations)}
        return results
    finally:
        if own:
            ses.close()
