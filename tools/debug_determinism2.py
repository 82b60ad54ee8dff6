import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session
name = sys.argv[1] if len(sys.argv) > 1 else "config2"
w = make_workload(name, seed=0)
K = w.n_modules
ses = Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, K)
L = len(w.penalization)
outs, coefs = [], []
for rep in range(3):
    res = ses.fit_cycle(w.penalization, np.stack([w.k_init] * L), want_votes=True, want_nnls_iters=True)
    ses._last_nsel = res["n_selected"]
    outs.append(res)
    coefs.append([[ses.coeffs(f, j)[0] for j in range(K)] for f in range(L)])
for rep in (1, 2):
    for key in ("resid_var", "r2_test", "votes"):
        a = outs[rep][key].astype(np.float64); b = outs[0][key].astype(np.float64)
        d = np.abs(a - b)
        nz = np.argwhere(d > 0)
        print("rep", rep, key, "max abs diff", d.max(), "max rel", (d / np.maximum(np.abs(b), 1e-300)).max(), "n differing", len(nz),
              "fits", sorted(set(nz[:, 0]))[:10], "modules", sorted(set(nz[:, 1]))[:12] if key != "votes" else "", flush=True)
        if key == "resid_var" and len(nz):
            tt = sorted(set(nz[:, 2]))
            print("     target range of differing entries:", tt[0], tt[-1], "count distinct", len(tt), "tiles", sorted(set(t // 64 for t in tt))[:40])
    cd = max((np.abs(coefs[rep][f][j] - coefs[0][f][j]).max() if coefs[0][f][j] is not None else 0.0) for f in range(L) for j in range(K))
    print("rep", rep, "max coefficient diff", cd)
print("n_selected", outs[0]["n_selected"].tolist())
ses.close()
