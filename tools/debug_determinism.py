import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session, scregclust_fit
name = sys.argv[1] if len(sys.argv) > 1 else "config2"
w = make_workload(name, seed=0)
def h(a): return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
inits = []
for sidx in range(2):
    ses = Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules)
    init = ses.initial_fit(); g = ses.gram()
    print("session", sidx, "gram", h(g[0]), h(g[1]), "beta_init", h(init["beta_init"]), "res_sds", h(init["res_sds"]), flush=True)
    L = len(w.penalization)
    for rep in range(2):
        res = ses.fit_cycle(w.penalization, np.stack([w.k_init] * L), want_votes=True, want_nnls_iters=True)
        print("   cycle-1 rep", rep, {k: h(v) for k, v in res.items() if v is not None and k in ("admm_iterations", "weights", "models", "nnls_iterations", "resid_var", "r2_test", "votes", "k_new")}, flush=True)
    for rep in range(2):
        out = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init, session=ses, return_coeffs=False)
        print("   fit rep", rep, [r.n_cycles_run for r in out], h(np.concatenate([np.concatenate(r.k_history) for r in out])), flush=True)
    ses.close()
