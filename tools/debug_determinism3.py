import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session
name = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = make_workload(name, seed=0)
K = w.n_modules
bad = 0
with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, K) as ses:
    for sel in ([0, 2, 5, 9], list(range(len(w.penalization))), [1], [3, 4]):
        sel = [i for i in sel if i < len(w.penalization)]
        pens = w.penalization[sel]
        outs = [ses.fit_cycle(pens, np.stack([w.k_init] * len(pens)), want_votes=True, want_nnls_iters=True) for _ in range(reps)]
        for rep in range(1, reps):
            diffs = {k: int((np.nan_to_num(v.astype(np.float64)) != np.nan_to_num(outs[rep][k].astype(np.float64))).sum()) for k, v in outs[0].items() if v is not None}
            diffs = {k: v for k, v in diffs.items() if v}
            if diffs:
                bad += 1
                print("composition", sel, "rep", rep, "DIFFERS:", diffs, flush=True)
print("determinism check:", "FAILED" if bad else "ok", "(", bad, "differing repetitions )")
