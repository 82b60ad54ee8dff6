"""Times the device k-means++ initialisation (scrc_ctx_kmeanspp) on the cross_corr1 matrix of a workload with the
reference's settings (n_initializations = 50, iter.max = 100; R/scregclust.R:1131-1136)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
n_init = int(sys.argv[2]) if len(sys.argv) > 2 else 50
w = make_workload(name, seed=0)
T = w.z1_target.shape[1]
r = np.random.default_rng(1)
firsts = r.integers(1, T + 1, n_init); us = r.random((n_init, w.n_modules - 1))
with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules) as s:
    for rep in range(2):
        t0 = time.perf_counter()
        cl, tot, fault = s.kmeanspp(firsts, us, n_max_iter=100, return_details=True)
        dt = time.perf_counter() - t0
        print(f"{name}: kmeanspp on {T} x {w.z1_reg.shape[1]} cross_corr1, K={w.n_modules}, {n_init} initialisations: {dt:.2f} s; "
              f"tot.withinss min {tot.min():.6g} max {tot.max():.6g}; ifault {np.bincount(fault, minlength=5).tolist()}; "
              f"cluster sizes {np.bincount(cl)[1:].tolist()}", flush=True)
