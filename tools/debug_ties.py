"""Debug: are modules with identical models bitwise identical on the GPU?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import scregclust_oracle as orc
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session

name, seed, ar1 = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
w = make_workload(name, seed=seed, ar1=ar1)
K = w.n_modules
ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization, K, w.k_init,
                         min_module_size=2, n_cycles=1, keep_history=True)
with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, K) as s:
    L = len(w.penalization)
    res = s.fit_cycle(w.penalization, np.stack([w.k_init] * L), want_votes=True, want_nnls_iters=True)
    s._last_nsel = res["n_selected"]
    for f, r in enumerate(ref):
        mods = r.models_history[0]
        for a in range(K):
            for b in range(a + 1, K):
                if mods[:, a].sum() > 0 and np.array_equal(mods[:, a], mods[:, b]):
                    ca, _ = s.coeffs(f, a); cb, _ = s.coeffs(f, b)
                    sg_a = res["signs"][f][a]; sg_b = res["signs"][f][b]
                    oa = r.coeffs_history[0][a]; ob = r.coeffs_history[0][b]
                    print(f"fit {f} modules {a},{b} same set (s={int(mods[:, a].sum())}): "
                          f"gpu signs equal {np.array_equal(np.nan_to_num(sg_a), np.nan_to_num(sg_b))}, "
                          f"gpu coeffs bitwise equal {np.array_equal(ca, cb)} (max diff {np.abs(ca - cb).max():.3e}), "
                          f"oracle coeffs bitwise equal {np.array_equal(oa, ob)} (max diff {np.abs(oa - ob).max():.3e}), "
                          f"gpu var equal {np.array_equal(res['resid_var'][f][a], res['resid_var'][f][b])}, "
                          f"gpu r2 equal {np.array_equal(res['r2_test'][f][a], res['r2_test'][f][b])}, "
                          f"weights equal gpu {np.array_equal(res['weights'][f][a], res['weights'][f][b])}")
                    if not np.array_equal(ca, cb):
                        t = int(np.argmax(np.abs(ca - cb).max(axis=0)))
                        print("    target", t, "nnls iters", res["nnls_iterations"][f][a][t], res["nnls_iterations"][f][b][t])
                        print("    gpu a", ca[:, t], "\n    gpu b", cb[:, t])
