"""Session-creation phases over repeated sessions (SCRC_DEBUG_TIMING): is the eigen-decomposition time stable?

    SCRC_DEBUG_TIMING=1 python tools/eig_probe.py [P] [repeats]      (optionally with CUDA_MODULE_LOADING=EAGER)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200.driver import Session

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1600
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rng = np.random.default_rng(0)
n, T, K = 4000, 512, 4
z1r = np.asfortranarray(rng.standard_normal((n, P))); z2r = np.asfortranarray(rng.standard_normal((n, P)))
z1t = np.asfortranarray(rng.standard_normal((n, T))); z2t = np.asfortranarray(rng.standard_normal((n, T)))
sds = z1t.std(0, ddof=1)
for r in range(reps):
    t0 = time.perf_counter()
    ses = Session(z1r, z2r, z1t, z2t, sds, K)
    t1 = time.perf_counter()
    ses.close()
    print(f"session {r}: create {1e3 * (t1 - t0):.1f} ms", file=sys.stderr, flush=True)
