"""SURVEY.md section 8(d) "config 5" micro-benchmark.

coop_lasso as the session executes it (Gram + eigenbasis once, then one batched Step 1) over
P in {256..4096} x n in {1k..200k} x batch in {1,4,16,64} modules of m = 128 target columns, and
fast_cor over the same n with (px = 4096, py = P).  Prints one JSON object per point and writes the
list to gpurun_out/microbench.json.  Points whose host matrices would exceed --max-doubles are
listed as skipped, not silently dropped.

    python tools/microbench.py [--quick] [--max-doubles 6e8]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from scregclust_b200 import _capi, api
from scregclust_b200.driver import Session

CLASSES = ["gemm_dense", "gemm_admm", "admm_elem", "nnls", "resid_vote", "other"]


def profile(lib):
    pm = (C.c_double * 6)(); pw = (C.c_double * 6)(); pl = (C.c_ulonglong * 6)()
    _capi.check(lib.scrc_profile_get(pm, pw, pl))
    return {n: (pm[i], pw[i], int(pl[i])) for i, n in enumerate(CLASSES)}


def planted(rng, n, P, batch, m):
    """Regulators N(0,1); every module: 8 planted regulators shared by its m targets (SURVEY 8d generator)."""
    T = batch * m
    zr = rng.standard_normal((n, P))
    B = np.zeros((P, T))
    for j in range(batch):
        regs = rng.choice(P, size=min(8, P), replace=False)
        B[np.ix_(regs, np.arange(j * m, (j + 1) * m))] = (0.4 * rng.standard_normal((regs.size, 1))
                                                         + 0.1 * rng.standard_normal((regs.size, m)))
    zt = zr @ B + rng.standard_normal((n, T))
    return zr, zt


def scale_like_driver(zr, zt):
    """R/scregclust.R:996-1009: regulators scaled to unit sd, targets centred (sd kept aside)."""
    zr = (zr - zr.mean(0)) / zr.std(0, ddof=1)
    zt = zt - zt.mean(0)
    return np.asfortranarray(zr), np.asfortranarray(zt), zt.std(0, ddof=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--max-doubles", type=float, default=6e8)
    ap.add_argument("--lam", type=float, default=0.1)
    ap.add_argument("--out", default="gpurun_out/microbench.json")
    ap.add_argument("--ns", default="", help="comma-separated cell counts (default 1000,10000,50000,200000)")
    ap.add_argument("--Ps", default="", help="comma-separated regulator counts (default 256..4096)")
    ap.add_argument("--batches", default="", help="comma-separated module counts (default 1,4,16,64)")
    ap.add_argument("--no-fast-cor", action="store_true")
    a = ap.parse_args()
    Ps = [256, 512, 1024, 2048, 4096]
    ns = [1000, 10000, 50000, 200000]
    batches = [1, 4, 16, 64]
    if a.quick:
        Ps, ns, batches = [256, 1024], [1000, 10000], [1, 16]
    if a.ns:
        ns = [int(v) for v in a.ns.split(",")]
    if a.Ps:
        Ps = [int(v) for v in a.Ps.split(",")]
    if a.batches:
        batches = [int(v) for v in a.batches.split(",")]
    m, n2 = 128, 256
    lib = _capi.load()
    rows = []
    rng = np.random.default_rng(5)
    for P in Ps:
        for n in ns:
            for batch in batches:
                T = batch * m
                if (n + n2) * (P + T) > a.max_doubles:
                    rows.append({"kind": "coop_lasso", "P": P, "n": n, "batch": batch, "skipped": "host matrices too large"})
                    print(json.dumps(rows[-1]), flush=True)
                    continue
                zr, zt = planted(rng, n + n2, P, batch, m)
                zr, zt, _ = scale_like_driver(zr, zt)
                z1r, z2r = np.asfortranarray(zr[:n]), np.asfortranarray(zr[n:])
                z1t, z2t = np.asfortranarray(zt[:n]), np.asfortranarray(zt[n:])
                sds = z1t.std(0, ddof=1)
                k = np.repeat(np.arange(1, batch + 1, dtype=np.int32), m)
                t0 = time.perf_counter()
                ses = Session(z1r, z2r, z1t, z2t, sds, batch)
                t_create = time.perf_counter() - t0
                best = None
                for rep in range(2):        # second repetition = warm
                    _capi.check(lib.scrc_profile_reset()); _capi.check(lib.scrc_profile_enable(1))
                    s0, i0 = ses.admm_sweeps, ses.admm_problem_iterations
                    t0 = time.perf_counter()
                    res = ses.fit_cycle([a.lam], k, skip_allocation=True)
                    wall = time.perf_counter() - t0
                    pr = profile(lib); _capi.check(lib.scrc_profile_enable(0))
                    best = (wall, pr, ses.admm_sweeps - s0, ses.admm_problem_iterations - i0, res)
                wall, pr, sweeps, piters, res = best
                admm_ms = pr["gemm_admm"][0] + pr["admm_elem"][0]
                row = {"kind": "coop_lasso", "P": P, "n": n, "batch": batch, "m": m, "lambda": a.lam,
                       "session_create_s": round(t_create, 4), "step1_wall_s": round(wall, 5),
                       "admm_phase_ms": round(admm_ms, 3), "sweeps": int(sweeps), "problem_iterations": int(piters),
                       "admm_problem_iters_per_s": round(piters / max(admm_ms, 1e-9) * 1e3, 1),
                       "sweeps_per_s": round(sweeps / max(admm_ms, 1e-9) * 1e3, 1),
                       "gemm_admm_tflops": round(pr["gemm_admm"][1] / max(pr["gemm_admm"][0], 1e-9) / 1e9, 2),
                       "selected_per_module_mean": float(res["n_selected"].mean()),
                       "class_ms": {c: round(v[0], 3) for c, v in pr.items()}}
                rows.append(row)
                print(json.dumps(row), flush=True)
                ses.close()
                del zr, zt, z1r, z2r, z1t, z2t
    # fast_cor(n, px = 4096, py = P) through the drop-in entry point (host buffers, copies included)
    px = 4096 if not a.quick else 512
    for P in ([] if a.no_fast_cor else Ps):
        for n in ns:
            if n * (px + P) > a.max_doubles:
                rows.append({"kind": "fast_cor", "n": n, "px": px, "py": P, "skipped": "host matrices too large"})
                print(json.dumps(rows[-1]), flush=True)
                continue
            x = np.asfortranarray(rng.standard_normal((n, px))); y = np.asfortranarray(rng.standard_normal((n, P)))
            api.fast_cor(x[:64], y[:64])
            t0 = time.perf_counter(); api.fast_cor(x, y); wall = time.perf_counter() - t0
            row = {"kind": "fast_cor", "n": n, "px": px, "py": P, "wall_s": round(wall, 5),
                   "tflops_incl_copies": round(2.0 * n * px * P / wall / 1e12, 2),
                   "host_bytes": 8 * (n * (px + P) + px * P)}
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
