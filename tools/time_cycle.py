"""Times ONE batched cycle (all penalization values from the initial configuration) and prints the kernel classes."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200 import _capi
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
w = make_workload(name, seed=0)
lib = _capi.load()
ses = Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules)
ks = np.stack([w.k_init] * len(w.penalization))
for rep in range(2):
    _capi.check(lib.scrc_profile_reset()); _capi.check(lib.scrc_profile_enable(1))
    t0 = time.perf_counter(); res = ses.fit_cycle(w.penalization, ks, light=True); t1 = time.perf_counter()
    pm = (C.c_double * 6)(); pw = (C.c_double * 6)(); pl = (C.c_ulonglong * 6)()
    _capi.check(lib.scrc_profile_get(pm, pw, pl)); _capi.check(lib.scrc_profile_enable(0))
    names = ["gemm_dense", "gemm_admm", "admm_elem", "nnls", "resid_vote", "other"]
    print(f"rep {rep}: cycle {t1 - t0:.3f} s", {n: round(pm[i], 1) for i, n in enumerate(names)},
          "selected/module mean %.1f" % res["n_selected"].mean(), "k_new checksum", int(res["k_new"].sum()), flush=True)
ses.close()
