"""Stage-by-stage comparison of one cycle (GPU session vs oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import scregclust_oracle as orc
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session

name = sys.argv[1] if len(sys.argv) > 1 else "small"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ar1 = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
w = make_workload(name, seed=seed, ar1=ar1)
K, T = w.n_modules, w.z1_target.shape[1]
ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization, K, w.k_init,
                         min_module_size=2, n_cycles=1, keep_history=True)
with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, K) as s:
    L = len(w.penalization)
    res = s.fit_cycle(w.penalization, np.stack([w.k_init] * L), want_votes=True, want_nnls_iters=True)
    s._last_nsel = res["n_selected"]
    for f, r in enumerate(ref):
        print("=== lambda", r.penalization)
        print(" admm it gpu", res["admm_iterations"][f], "ref", r.admm_iterations[0])
        mg = res["models"][f].T.astype(bool); mr = r.models_history[0]
        print(" models equal", np.array_equal(mg, mr), "nsel gpu", mg.sum(0), "ref", mr.sum(0))
        ng = res["nnls_iterations"][f]; nr = r.nnls_iterations[0]
        print(" nnls iters equal", np.array_equal(ng, nr), "ndiff", int((ng != nr).sum()))
        if not np.array_equal(ng, nr):
            jj, tt = np.nonzero(ng != nr)
            print("   first diffs (module, target, gpu, ref):", [(int(a), int(b), int(ng[a, b]), int(nr[a, b])) for a, b in list(zip(jj, tt))[:8]])
        vg = res["votes"][f]; vr = r.votes[0]
        print(" votes equal", np.array_equal(vg, vr), "genes differing", int((vg != vr).any(axis=1).sum()))
        rvg = res["resid_var"][f].T; rvr = r.resid_var_history[0]
        r2g = res["r2_test"][f].T; r2r = r.r2_history[0]
        print(" resid_var max rel err", float(np.max(np.abs(rvg - rvr) / np.abs(rvr))), " r2 max abs err", float(np.max(np.abs(r2g - r2r))))
        rel = np.abs(rvg - rvr) / np.abs(rvr)
        tt, jj = np.unravel_index(np.argsort(rel, axis=None)[::-1][:4], rel.shape)
        for t_, j_ in zip(tt, jj):
            cg, sel = s.coeffs(f, int(j_)); cr = r.coeffs_history[0][int(j_)]
            print("  worst var: gene", int(t_), "module", int(j_), "gpu", rvg[t_, j_], "ref", rvr[t_, j_], "nsel", res["n_selected"][f][j_],
                  "nnls it", ng[j_, t_], nr[j_, t_])
            if cg is not None and cg.shape[0] > 16:
                n1 = w.z1_target.shape[0]
                for nm, keep in (("all", slice(None)), ("first16", slice(0, 16)), ("after16", slice(16, None))):
                    rr = w.z1_target[:, t_] - w.z1_reg[:, sel[keep]] @ cr[keep, t_]
                    print("     hypothesis", nm, float(rr @ rr) / (n1 - cg.shape[0]))
            if cg is not None:
                print("     coeff gpu", np.array2string(cg[:, t_], precision=6), "\n     coeff ref", np.array2string(cr[:, t_], precision=6))
        for j in range(K):
            cg, sel = s.coeffs(f, j); cr = r.coeffs_history[0][j]
            if cg is None or cr is None:
                print("  module", j, "coeffs None", cg is None, cr is None); continue
            e = np.abs(cg - cr); print("  module", j, "coeff max abs err", float(e.max()), "at target", int(np.argmax(e.max(axis=0))))
        bad = np.flatnonzero((vg != vr).any(axis=1))
        for t in bad[:5]:
            print("  gene", int(t), "votes gpu", vg[t], "ref", vr[t], "var gpu", rvg[t], "ref", rvr[t], "r2 gpu", r2g[t], "ref", r2r[t])
        # residual variances and r2 are only kept for the final cycle in the oracle; recompute from oracle pieces
        print(" k_new equal", np.array_equal(res["k_new"][f], ref[f].k_history[1]) if len(ref[f].k_history) > 1 else None)
