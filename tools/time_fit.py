"""Times one full penalty-grid fit (host API) for a named workload; prints per-phase kernel-class times."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200 import _capi
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session, scregclust_fit

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t0 = time.time(); w = make_workload(name, seed=0); print("workload", name, "generated in %.1f s" % (time.time() - t0), flush=True)
lib = _capi.load()
for rep in range(reps):
    _capi.check(lib.scrc_profile_reset()); _capi.check(lib.scrc_profile_enable(1))
    t0 = time.time()
    ses = Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules)
    t1 = time.time()
    acc = {"fit_cycle": 0.0, "calls": 0, "importance": 0.0, "coeffs": 0.0}
    _fc, _imp, _cf = ses.fit_cycle, ses.importance, ses.coeffs_fit
    def fc(*a, **k):
        t = time.perf_counter(); r = _fc(*a, **k); acc["fit_cycle"] += time.perf_counter() - t; acc["calls"] += 1; return r
    def im(*a, **k):
        t = time.perf_counter(); r = _imp(*a, **k); acc["importance"] += time.perf_counter() - t; return r
    def cf(*a, **k):
        t = time.perf_counter(); r = _cf(*a, **k); acc["coeffs"] += time.perf_counter() - t; return r
    ses.fit_cycle, ses.importance, ses.coeffs_fit = fc, im, cf
    res = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init, session=ses,
                         return_coeffs="--coeffs" in sys.argv, compute_predictive_r2="--r2" in sys.argv)
    t2 = time.time()
    print(f"    host split: {acc['calls']} fit_cycle calls {acc['fit_cycle']:.3f} s, importance {acc['importance']:.3f} s, "
          f"coeffs {acc['coeffs']:.3f} s, python remainder {t2 - t1 - acc['fit_cycle'] - acc['importance'] - acc['coeffs']:.3f} s", flush=True)
    if "--importance" in sys.argv:
        ks = np.stack([r.k_final[0] for r in res]); md = np.stack([r.models[0].T for r in res]); sg = np.stack([r.signs[0].T for r in res])
        ti = time.time(); r2m, imp = ses.importance(ks, md, sg); ti = time.time() - ti
        print(f"    importance for {len(res)} final configurations: {ti:.3f} s; r2_module[0][:5] = {np.round(r2m[0][:5], 4)}", flush=True)
    pm = (C.c_double * 6)(); pw = (C.c_double * 6)(); pl = (C.c_ulonglong * 6)()
    _capi.check(lib.scrc_profile_get(pm, pw, pl)); _capi.check(lib.scrc_profile_enable(0))
    print(f"rep {rep}: create {t1 - t0:.3f} s, loop {t2 - t1:.3f} s, cycles/penalty {[r.n_cycles_run for r in res]}, "
          f"sweeps {ses.admm_sweeps}, problem-iterations {ses.admm_problem_iterations}", flush=True)
    names = ["gemm_dense", "gemm_admm", "admm_elem", "nnls", "resid_vote", "other"]
    print("   ", {n: (round(pm[i], 1), f"{pw[i] / max(pm[i], 1e-9) / 1e9:.1f} G/s") for i, n in enumerate(names)}, flush=True)
    print("    selected regs per module (first penalty, final):", res[0].models[0].sum(axis=0))
    ses.close()
