"""Fully measured CPU run of a whole penalty grid on the C++ twin (oracle/ref_cpu.cpp, reference cost structure):
the number bench.py's bounded CPU sample is extrapolated to.  Writes profiles/r02_cpu_full_<workload>.json.

    python tools/cpu_full_grid.py config2 [threads]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_cpu, twin_cycle
from oracle import scregclust_oracle as orc
from scregclust_b200.synthetic import make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
threads = ref_cpu.set_threads(int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1))
w = make_workload(name, seed=0)
t0 = time.perf_counter()
pre = orc.precompute(w.z1_reg, w.z1_target)
t_pre = time.perf_counter() - t0
out = dict(workload=name, seed=0, threads=threads, precompute_seconds=t_pre, penalties=[])
total = t_pre
for lam in w.penalization:
    r = twin_cycle.full_fit(w, float(lam), pre=pre, log=lambda s: print(s, flush=True))
    total += r["seconds"]
    out["penalties"].append(dict(penalization=float(lam), seconds=r["seconds"], step1=r["step1"], step2a=r["step2a"],
                                 step2b=r["step2b"], cycles=len(r["k_history"]),
                                 problem_iterations=int(sum(int(a.sum()) for a in r["admm_iterations"])),
                                 assignment_checksum=int(np.clip(r["k_history"][-1], 0, None).sum())))
    print(f"lambda={lam:.4g}: {r['seconds']:.1f} s, {len(r['k_history'])} configurations", flush=True)
out["seconds_per_grid"] = total
out["note"] = "no importance step (compute_predictive_r2) in this run; Step 1 + 2a + 2b of every cycle and the final re-fits"
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"r02_cpu_full_{name}.json"), "w"), indent=1)
print("total", total)
