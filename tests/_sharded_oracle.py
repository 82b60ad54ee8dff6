"""CPU emulation of one rank of the multi-GPU cycle of csrc/session.cu (Ctx::fit_cycle), answered from the
oracle's primitives: same partition rules (scrc_shard_assign / scrc_shard_slice from the library), same
ownership protocol (every merged entry is written by exactly one rank, zero elsewhere) and the same two
merge points, with the exchange supplied by the caller (gloo all-reduce in tests/test_distributed.py).

TEST INFRASTRUCTURE: imports oracle/; the product path never does.
"""
import ctypes as C
import math

import numpy as np

from oracle import scregclust_oracle as orc
from scregclust_b200 import _capi


def shard_assign(cost, world):
    lib = _capi.load()
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    owner = np.zeros(max(cost.size, 1), dtype=np.int32)
    _capi.check(lib.scrc_shard_assign(_capi.dptr(cost), cost.size, world, _capi.iptr(owner)))
    return owner[:cost.size]


def shard_slice(n, align, rank, world):
    lib = _capi.load()
    a, b = C.c_int64(0), C.c_int64(0)
    _capi.check(lib.scrc_shard_slice(n, align, rank, world, C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


class ShardedOracleSession:
    """Duck-types driver.Session for scregclust_fit(); `allreduce(array)` sums an array over the ranks in place."""
    accepts_hint = True

    def __init__(self, w, rank=0, world=1, allreduce=None):
        self.w, self.rank, self.world = w, rank, world
        self.allreduce = allreduce if allreduce is not None else (lambda a: a)
        self.K = w.n_modules
        self.n1, self.P = w.z1_reg.shape
        self.n2, self.T = w.z2_target.shape
        self.pre = orc.precompute(w.z1_reg, w.z1_target)
        n1 = self.n1
        self.x = w.z1_reg / math.sqrt(n1)
        self.G_scaled = orc.compute_xtx(self.x)
        self.xty_scaled_all = self.x.T @ ((w.z1_target / self.pre["res_sds"][None, :]) / math.sqrt(n1))
        self.G_rr = orc.compute_xtx(w.z1_reg)
        self.G_rt_nnls = w.z1_reg.T @ (w.z1_target / w.z1_target_sds[None, :])
        self.calls = []
        self.problems_solved = 0
        self._coeffs = None

    def fit_cycle(self, lambdas, ks, options=None, skip_allocation=False, want_votes=False, want_nnls_iters=False,
                  prior=None, light=False, hint=None):
        w, K, T, P, n1 = self.w, self.K, self.T, self.P, self.n1
        F = len(lambdas)
        ks = np.asarray(ks).reshape(F, T)
        self.calls.append((F, bool(skip_allocation)))
        tol_rel = options.tol_coop_rel if options is not None else 1e-8
        tol_abs = options.tol_coop_abs if options is not None else 1e-12
        tol_nnls = options.tol_nnls if options is not None else 1e-4
        max_iter = options.max_optim_iter if options is not None else 10000
        # ---- plan of Step 1 (Ctx::fit_cycle): non-empty (fit, module) pairs in order, cost = columns x predicted iterations
        cnt = np.stack([np.bincount(ks[f][(ks[f] >= 1) & (ks[f] <= K)], minlength=K + 1)[1:] for f in range(F)])
        g_fk = np.flatnonzero(cnt.ravel() > 0)
        h = np.asarray(hint).ravel() if hint is not None else np.zeros(F * K, dtype=np.int64)
        cost = cnt.ravel()[g_fk] * np.where(h[g_fk] > 0, h[g_fk], 25)
        owner = shard_assign(cost, self.world)
        mine = np.zeros(F * K, dtype=bool)
        mine[g_fk[owner == self.rank]] = True
        mine = mine.reshape(F, K)
        # ---- Step 1 on the owned problems; everything else stays zero
        models = np.zeros((F, K, P), dtype=np.uint8); weights = np.zeros((F, K, P)); iters = np.zeros((F, K), dtype=np.int32)
        for f in range(F):
            kmask = np.where((ks[f] >= 1) & mine[f][np.clip(ks[f], 1, K) - 1], ks[f], -1)
            it = np.zeros(K, dtype=np.int32)
            md, wt, _ = orc._step1(self.pre, self.x, self.G_scaled, self.xty_scaled_all, kmask, lambdas[f], K, tol_rel,
                                   tol_abs, max_iter, it)
            models[f] = md.T; weights[f] = wt.T; iters[f] = it
            self.problems_solved += int(mine[f].sum())
        # ---- merge 1
        self.allreduce(models); self.allreduce(weights); self.allreduce(iters)
        signs = np.where(models > 0, np.sign(weights), np.nan)
        nsel = models.sum(axis=2).astype(np.int32)
        # ---- Step 2 on this rank's target slice
        t0, t1 = shard_slice(T, 32, self.rank, self.world)
        Ts = t1 - t0
        r2 = np.zeros((F, K, T)); rv = np.zeros((F, K, T)); best_r2 = np.zeros((F, T))
        best_idx = np.zeros((F, T), dtype=np.int32); k_new = np.zeros((F, T), dtype=np.int32)
        votes = np.zeros((F, T, K), dtype=np.int32); nn_it = np.zeros((F, K, T), dtype=np.int32)
        coeffs = [[None] * K for _ in range(F)]
        z2sq = w.z2_target[:, t0:t1] ** 2
        for f in range(F):
            ss_tr = np.repeat(np.sum(w.z1_target[:, t0:t1] ** 2, axis=0)[:, None], K, axis=1)
            ss_te = np.repeat(np.sum(z2sq, axis=0)[:, None], K, axis=1)
            arr = np.repeat(z2sq[None, :, :], K, axis=0)
            for j in range(K):
                reg = np.flatnonzero(models[f, j])
                if reg.size == 0:
                    continue
                sg = signs[f, j, reg]
                xtx_j = self.G_rr[np.ix_(reg, reg)] * np.outer(sg, sg)
                xty_j = self.G_rt_nnls[reg, t0:t1] * sg[:, None]
                if Ts > 0:
                    b, its = orc.coef_nnls(eps=tol_nnls, max_iter=max_iter, xtx=xtx_j, xty=xty_j)
                else:
                    b, its = np.zeros((reg.size, 0)), np.zeros(0, dtype=np.int32)
                nn_it[f, j, t0:t1] = its
                bh = b * sg[:, None] * orc._recycle(w.z1_target_sds, (reg.size, T))[:, t0:t1]
                full = np.zeros((reg.size, T)); full[:, t0:t1] = bh
                coeffs[f][j] = full
                r_tr = w.z1_target[:, t0:t1] - w.z1_reg[:, reg] @ bh
                r_te = w.z2_target[:, t0:t1] - w.z2_reg[:, reg] @ bh
                arr[j] = r_te * r_te
                ss_te[:, j] = np.sum(r_te * r_te, axis=0); ss_tr[:, j] = np.sum(r_tr * r_tr, axis=0)
            r2s = 1.0 - ss_te / np.sum(z2sq, axis=0)[:, None]
            rvs = ss_tr / (n1 - nsel[f])[None, :]
            r2[f, :, t0:t1] = r2s.T; rv[f, :, t0:t1] = rvs.T
            if Ts > 0:
                best_r2[f, t0:t1] = r2s.max(axis=1); best_idx[f, t0:t1] = np.argmax(r2s, axis=1) + 1
            if not skip_allocation and Ts > 0:
                idx, vt = orc.allocate_into_modules(arr, rvs, (), ks[f][t0:t1], np.arange(Ts), 1e-6, 0.0, return_votes=True)
                k_new[f, t0:t1] = idx; votes[f, t0:t1] = vt
        # ---- merge 2 (every slice entry has one owner)
        for a in (r2, rv, best_r2, best_idx, k_new, votes, nn_it):
            self.allreduce(a)
        if skip_allocation:
            for f in range(F):
                for j in range(K):
                    if coeffs[f][j] is not None:
                        self.allreduce(coeffs[f][j])
        self._coeffs = coeffs
        return {"k_new": None if skip_allocation else k_new, "votes": votes if want_votes else None, "models": models,
                "weights": weights, "signs": signs, "r2_test": r2, "resid_var": rv, "best_r2": best_r2,
                "best_r2_idx": best_idx, "admm_iterations": iters, "nnls_iterations": nn_it if want_nnls_iters else None,
                "n_selected": nsel}

    def coeffs_fit(self, f):
        return self._coeffs[f], None

    def close(self):
        pass
