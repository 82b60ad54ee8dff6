#!/usr/bin/env python
"""bench.py -- scregclust wall time per penalty grid on N B200s, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config3|...] [--impl reference]

A *step* is one complete penalization-grid fit of the alternating loop (R/scregclust.R:996-1912:
Gram / eigenbasis / initial fit, then Step 1 + 2a + 2b cycles until every penalty has converged,
its final configurations are re-fitted and their leave-one-regulator-out importances computed,
i.e. the default compute_predictive_r2 = TRUE) on one seeded synthetic workload of a BASELINE.json
shape.  With N > 1 (torchrun, one rank per GPU) every cycle of the grid is spread over the ranks inside
the library -- Step-1 problems by (penalization value, module), Step 2 by target-gene slice, module
summaries and assignment vectors merged with NCCL all-reduces (csrc/session.cu) -- "scaling": "strong":
the grid is fixed, the wall time per grid falls with N.

Printed JSON (rank 0, one line):
  value        device-resident step time [s] (inputs already in HBM when the clock starts)
  e2e          the same step through the public host API: host (pinned) buffers -> C ABI -> results
               back on the host, copies inside the timed region
  roofline     dominant kernel class (the two eigenbasis GEMMs of every ADMM sweep): algorithmic
               flops / CUDA-event time on the launching stream, against cuBLAS DGEMM measured in
               this run (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline reference-faithful C++ twin (oracle/_ref/liboracle_ref.so) timed on a bounded sample
               of the same workload and extrapolated to the grid on the exact per-problem iteration
               counts of the GPU run (N=1, rank 0 only; "extrapolated": true)
  parity_spot  the modules the CPU leg solved, compared with the GPU's results for the same modules
  --impl reference prints only that CPU arm.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scregclust_wall_time_per_penalty_grid"
CYCLES_TABLE = os.path.join(ROOT, "profiles", "bench_cycles.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SCRC_BENCH_WORKLOAD", "config3"))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU-baseline sample work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0,
                    help="threads of the CPU legs (0 = all host cores; 1 = what stock R + reference BLAS give outside Eigen's GEMMs)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: samples taken before it (warm-up) are not reported."""
        self.lines = []

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU arm: reference-faithful C++ twin on a bounded sample
# --------------------------------------------------------------------------------------------
def cpu_sample(w, budget_s, lam, totals, gpu_first_cycle=None, n_threads=0):
    """Times the first cycle (Step 1 + 2a + 2b) of one penalization value with the cost structure of the
    reference (oracle/ref_cpu.cpp through oracle/twin_cycle.py) on as many modules / genes as fit the
    budget, and extrapolates to the whole grid on the work counters of the GPU run:

      Step 1   seconds per ADMM problem-iteration (measured) x problem-iterations of the grid (exact: the GPU's
               per-problem iteration counts are bit-equal to the twin's, tests/test_gpu_baseline_scale.py)
      Step 2a  seconds per (module with a model) x such modules over all cycles of the grid
      Step 2b  seconds per gene of reset + allocation x T x fit cycles
      importance  seconds per module x modules of the final configurations

    `totals`: dict(problem_iterations, modules_with_model, fit_cycles, final_modules) of the grid.
    `gpu_first_cycle` (optional): dict(admm_iterations[K], models[P, K]) of the GPU for the same penalization
    value and configuration -> parity_spot.  Returns (seconds per grid, description, threads, details)."""
    from oracle import ref_cpu, twin_cycle
    from oracle import scregclust_oracle as orc

    threads = ref_cpu.set_threads(n_threads if n_threads > 0 else (os.cpu_count() or 1))
    n1, P = w.z1_reg.shape
    n2, T = w.z2_target.shape
    K = w.n_modules
    t_pre0 = time.perf_counter()
    pre = orc.precompute(w.z1_reg, w.z1_target)               # R BLAS path of R/scregclust.R:1049-1113
    t_pre = time.perf_counter() - t_pre0
    k = w.k_init
    # ---- Step 1 + 2a, module by module until the budget is used
    t1 = t2 = 0.0
    its = cols = 0
    done, done2 = [], 0
    spot = dict(modules=[], iterations_equal=True, sets_equal=True, max_rel=0.0)
    sample = {}
    for j in range(1, K + 1):
        if not np.any(k == j):
            continue
        if (t1 + t2) > budget_s * 0.8 and len(done) >= 1:
            break
        c = twin_cycle.cycle(w, lam, k, pre=pre, modules=[j])
        t1 += c["step1_seconds"]; t2 += c["step2a_seconds"]
        its += int(c["admm_iterations"][j - 1]); cols += int(np.count_nonzero(k == j))
        done.append(j); done2 += int(c["coeffs"][j - 1] is not None)
        sample[j] = c
        if gpu_first_cycle is not None:
            spot["modules"].append(j)
            spot["iterations_equal"] &= bool(gpu_first_cycle["admm_iterations"][j - 1] == c["admm_iterations"][j - 1])
            spot["sets_equal"] &= bool(np.array_equal(gpu_first_cycle["models"][:, j - 1], c["models"][:, j - 1]))
    sec_per_iter = t1 / max(its, 1)
    sec_per_model = t2 / max(done2, 1)
    # ---- array reset, the per-module strided writes and the allocation on a gene subset (the full K x n2 x T array is
    # 8 GB at config 3): exactly the operations of one cycle of oracle/twin_cycle.full_fit
    g = int(max(8, min(T, 2.5e8 // (K * n2 * 8))))
    z2sq = np.asfortranarray(w.z2_target[:, :g] ** 2)
    arr = np.empty((K, n2, g), order="F")
    nonempty = int(sum(1 for j in range(1, K + 1) if np.any(k == j)))
    t0 = time.perf_counter()
    ref_cpu.fill_array(arr, z2sq)                             # reset_array (src/utils.cpp:43-63)
    for j in range(nonempty):
        arr[j, :, :] = z2sq * 0.9                             # sq_residuals_test[j, , ] <- ... (scregclust.R:1599)
    t_fill = time.perf_counter() - t0
    rv = np.asfortranarray(np.ones((g, K)))
    t0 = time.perf_counter()
    ref_cpu.allocate_into_modules(arr, rv, None, None, np.ones(g, dtype=np.int32), np.arange(g, dtype=np.int32), 1e-6, 0.0)
    t_alloc = time.perf_counter() - t0
    sec_per_gene_cycle = (t_fill + t_alloc) / g
    # ---- post-processing: leave-one-regulator-out NNLS importance (R/scregclust.R:1823-1912) on one module
    t_imp, imp_mods = 0.0, 0
    for j, c in sample.items():
        reg = np.flatnonzero(c["models"][:, j - 1])
        cl = np.flatnonzero(k == j)
        if reg.size < 2 or cl.size == 0:
            continue
        sg = c["signs"][reg, j - 1]
        t0 = time.perf_counter()
        y_cl = np.asfortranarray(w.z1_target[:, cl] / w.z1_target_sds[cl][None, :])
        for r in range(reg.size + 1):
            keep = np.ones(reg.size, dtype=bool)
            if r > 0:
                keep[r - 1] = False
            xs = np.asfortranarray(w.z1_reg[:, reg[keep]] * sg[keep][None, :])
            b, _ = ref_cpu.coef_nnls(xs, y_cl, eps=1e-4, max_iter=10000)
            bh = b * sg[keep][:, None] * orc._recycle(w.z1_target_sds[cl], b.shape)
            rt = w.z2_target[:, cl] - w.z2_reg[:, reg[keep]] @ bh
            _ = (rt * rt).sum(axis=0)
        t_imp += time.perf_counter() - t0
        imp_mods += 1
        break
    sec_per_final_module = t_imp / max(imp_mods, 1)
    parts = dict(precompute=t_pre, step1=sec_per_iter * totals["problem_iterations"],
                 step2a=sec_per_model * totals["modules_with_model"],
                 step2b=sec_per_gene_cycle * T * totals["fit_cycles"],
                 importance=sec_per_final_module * totals["final_modules"])
    total = float(sum(parts.values()))
    measured = t_pre + t1 + t2 + t_fill + t_alloc + t_imp
    desc = (f"reference-faithful C++ twin (oracle/ref_cpu.cpp: Gram per call, LDL^T per iteration, K x n2 x T array), "
            f"lambda={lam:.3g}, first cycle from the initial configuration: Step 1 + 2a on modules {done} of {nonempty} "
            f"({its} ADMM iterations over {cols} columns, {t1:.1f} s + {t2:.1f} s), reset+allocate on {g}/{T} genes "
            f"({t_fill + t_alloc:.2f} s), importance refits of {imp_mods} module ({t_imp:.1f} s), precompute {t_pre:.1f} s; "
            f"extrapolated on the grid's work counters: {totals['problem_iterations']} ADMM problem-iterations, "
            f"{totals['modules_with_model']} module refits, {totals['fit_cycles']} allocation sweeps, "
            f"{totals['final_modules']} final modules")
    details = dict(extrapolated=True, measured_seconds=measured, scale_factor=total / max(measured, 1e-9),
                   seconds_per_problem_iteration=sec_per_iter, parts=parts,
                   model_check="profiles/r02_cpu_full_config2.json: the same model against a fully measured config-2 grid")
    if gpu_first_cycle is not None:
        details["parity_spot"] = spot
    return total, desc, threads, details


def load_totals(name, seed, n_pen, K):
    """Work counters of the grid (and the GPU's first-cycle results of the sampled penalization value) from the
    last GPU run of this seeded workload -- written by the GPU arm, read by `--impl reference`."""
    try:
        e = json.load(open(CYCLES_TABLE))[f"{name}/seed{seed}"]
        return e["totals"], e.get("first_cycle"), "profiles/bench_cycles.json"
    except Exception:
        return ({"problem_iterations": 25 * K * 7 * n_pen, "modules_with_model": K * 7 * n_pen, "fit_cycles": 6 * n_pen,
                 "final_modules": K * n_pen}, None,
                "assumed 6 cycles and 25 ADMM iterations per problem (no table for this workload)")


def grid_totals(res, K):
    """Exact work counters of a finished grid from its per-cycle records."""
    tot = dict(problem_iterations=0, modules_with_model=0, fit_cycles=0, final_modules=0)
    for r in res:
        tot["fit_cycles"] += r.n_cycles_run - 1
        for it, md in zip(r.admm_iterations, r.models_history):
            tot["problem_iterations"] += int(np.sum(it))
            tot["modules_with_model"] += int(np.count_nonzero(md.sum(axis=0)))
        for kf in r.k_final:
            tot["final_modules"] += int(len(set(kf[kf > 0].tolist())))
    return tot


# --------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    args = parse()
    # Native libraries (NCCL's version banner, cuSOLVER notices) write to file descriptor 1 directly; keep the
    # original stdout for the JSON line only and send everything else to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    from scregclust_b200.synthetic import CONFIGS, make_workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_cells, n_targets, n_reg, n_modules, pens = CONFIGS[args.workload]
    i_med = len(pens) // 2                                     # the penalization value the CPU leg samples
    config = {"workload": f"{args.workload}: synthetic {n_cells} cells x {n_targets} targets x {n_reg} regulators, "
                          f"{n_modules} modules, {len(pens)}-value penalty grid (seed {args.seed})",
              "n1": n_cells // 2, "n2": n_cells - n_cells // 2, "targets": n_targets, "regulators": n_reg,
              "modules": n_modules, "penalties": [float(p) for p in pens],
              "baseline_config": {"config1": "BASELINE.json configs[0] (synthetic stand-in of the same shape)",
                                  "config2": "BASELINE.json configs[1]", "config3": "BASELINE.json configs[2]",
                                  "config4": "BASELINE.json configs[3]"}.get(args.workload, "test-sized workload"),
              "parallelism": (f"every cycle of the whole grid spread over {world} GPUs inside the library: Step-1 problems by "
                              f"(penalization value, module), Step 2 by target-gene slice, NCCL all-reduce merges"
                              if world > 1 else "whole penalty grid advanced in lock-step on 1 GPU"),
              "l2": "working set (FP64 state of all problems) exceeds the 126 MB L2; no flush between steps"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        w = make_workload(args.workload, seed=args.seed)
        totals, first, src = load_totals(args.workload, args.seed, len(pens), n_modules)
        gfc = None
        if first is not None:
            md = np.zeros((n_reg, n_modules), dtype=bool)
            for j, sel in enumerate(first["selected"]):
                md[sel, j] = True
            gfc = {"admm_iterations": np.array(first["admm_iterations"]), "models": md}
        vals, desc, threads, details = [], "", 1, {}
        for _ in range(max(1, min(args.steps, 2))):            # each step = one bounded sample
            v, desc, threads, details = cpu_sample(w, args.cpu_budget, float(pens[i_med]), totals, gfc, args.cpu_threads)
            vals.append(v)
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "s", "cores": threads, "kind": "port",
                                 "sample": desc + f" [work counters: {src}]", **details},
                "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist

    from scregclust_b200 import _capi, api, parallel
    from scregclust_b200.driver import Session, scregclust_fit

    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _capi.load()
    api.set_device(local_rank)
    comm = parallel.comm_from_torch(local_rank) if world > 1 else None    # the library's own NCCL communicator

    w = make_workload(args.workload, seed=args.seed)
    all_pens = [float(p) for p in pens]
    T = n_targets
    host = [np.asfortranarray(a) for a in (w.z1_reg, w.z2_reg, w.z1_target, w.z2_target)]
    pinned = []
    for a in host:                                             # pinned host buffers for the e2e leg
        t = torch.empty(a.shape[::-1], dtype=torch.float64).pin_memory()   # (cols, rows) C-order == column-major
        t.numpy()[...] = a.T
        pinned.append(t)
    sds_pinned = torch.from_numpy(np.ascontiguousarray(w.z1_target_sds)).pin_memory()
    dev = [t.cuda(non_blocking=True) for t in pinned]           # device-resident copies for `value`
    sds_dev = sds_pinned.cuda()
    torch.cuda.synchronize()
    n1, P = w.z1_reg.shape
    n2 = w.z2_reg.shape[0]

    dbg = os.environ.get("SCRC_BENCH_DEBUG") is not None

    def run_step(resident: bool):
        t_a = time.perf_counter()
        src = dev if resident else pinned
        sd = sds_dev if resident else sds_pinned
        ses = Session.from_pointers(src[0].data_ptr(), src[1].data_ptr(), src[2].data_ptr(), src[3].data_ptr(),
                                    sd.data_ptr(), n1, n2, P, T, n_modules, device=local_rank, comm=comm)
        if not resident:
            ses.h2d_bytes = sum(t.numel() * 8 for t in pinned) + sds_pinned.numel() * 8
        try:
            t_b = time.perf_counter()
            # one call: the whole alternating loop of the grid runs inside the library (scrc_grid_fit); with
            # N > 1 it is collective -- every rank passes the same arguments and gets the complete result
            res = scregclust_fit(penalization=all_pens, n_modules=n_modules, initial_target_modules=w.k_init,
                                 session=ses, return_coeffs=not resident, compute_predictive_r2=True)
            gi = ses.last_grid_info
            stats = dict(h2d=ses.h2d_bytes, d2h=ses.d2h_bytes, sweeps=gi["admm_sweeps"],
                         admm_iters=gi["admm_problem_iterations"], fit_cycles=gi["fit_cycles"],
                         last_cycles=gi["final_refits"], device_calls=gi["device_calls"])
            t_c = time.perf_counter()
        finally:
            ses.close()
        allk = np.stack([r.k_final[0] for r in res])
        checksum = int(np.clip(allk, 0, None).sum())
        if dbg:
            print(f"[bench rank {rank} resident={resident}] create {t_b - t_a:.3f} s, fit {t_c - t_b:.3f} s, "
                  f"close {time.perf_counter() - t_c:.3f} s", file=sys.stderr, flush=True)
        return res, stats, checksum

    def timed(resident: bool, steps: int):
        ms = C.c_double(0.0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _capi.check(lib.scrc_timer_begin())
        t0 = time.perf_counter()
        out = None
        for _ in range(steps):
            out = run_step(resident)
        torch.cuda.synchronize()
        _capi.check(lib.scrc_timer_end(C.byref(ms)))
        wall = time.perf_counter() - t0
        t = torch.tensor([ms.value, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()) / steps, float(t[1].item()) / steps, out

    # cuBLAS DGEMM on this GPU = the FP64 roofline denominator
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        a @ b
    e1.record(); torch.cuda.synchronize()
    dgemm_tf = 3 * 2.0 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b

    # nvidia-smi takes a few hundred milliseconds to attach to the device, during which it holds up the launches of
    # the process it watches: start it before the warm-up steps so that none of that lands in a timed step
    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(max(args.warmup, 0)):
        run_step(True)
    sampler.mark()
    # pass 1 -- `value`: K steps, nothing but the library's own launches on the stream
    launches0 = api.launch_count()
    ms_dev, ms_wall, (res, stats, checksum) = timed(True, args.steps)
    launches = api.launch_count() - launches0
    # pass 2 -- the same K steps with a CUDA-event pair around every kernel group on the launching stream: per-class
    # kernel times for the roofline figures.  The events cost a few microseconds per group, which matters for the
    # launch-bound small workloads, so this pass has its own clock (`profiled_pass`) and `value` does not carry it.
    _capi.check(lib.scrc_profile_reset()); _capi.check(lib.scrc_profile_enable(1))
    prof_steps = max(1, min(args.steps, 3))      # enough launches for the per-class averages; keeps long runs short
    ms_prof, _, _ = timed(True, prof_steps)
    pm = (C.c_double * 6)(); pw = (C.c_double * 6)(); pl = (C.c_ulonglong * 6)()
    _capi.check(lib.scrc_profile_get(pm, pw, pl)); _capi.check(lib.scrc_profile_enable(0))
    clocks = sampler.stop()
    run_step(False)                                  # untimed: first touch of the host-side result buffers
    e2e_dev, e2e_wall, (_, e2e_stats, _) = timed(False, max(1, min(args.steps, 2)))

    if rank == 0:
        names = ["gemm_dense", "gemm_admm", "admm_elementwise", "nnls", "resid_vote", "other"]
        prof = {n: {"ms_per_step": pm[i] / prof_steps, "work_per_step": pw[i] / prof_steps, "timed_groups": int(pl[i])}
                for i, n in enumerate(names)}
        adm = prof["gemm_admm"]
        achieved = adm["work_per_step"] / (adm["ms_per_step"] * 1e-3) / 1e12 if adm["ms_per_step"] > 0 else 0.0
        kernel_ms = sum(v["ms_per_step"] for v in prof.values())
        traffic = None
        try:   # DRAM bytes of one (all-columns-active) launch of the dominant kernel from the committed ncu capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload]
        except Exception:
            pass
        hbm = None
        try:
            hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": ms_dev * 1e-3, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_dev * 1e-3, "unit": "s", "h2d_bytes_per_step": int(e2e_stats["h2d"]),
                    "d2h_bytes_per_step": int(e2e_stats["d2h"]), "host_wall_s": e2e_wall * 1e-3},
            "gpu_launches": int(launches),
            "admm_iters_per_s": stats["admm_iters"] / (ms_dev * 1e-3),
            "admm": {"problem_iterations": int(stats["admm_iters"]), "batched_sweeps_rank0": int(stats["sweeps"]),
                     "fit_cycles": int(stats["fit_cycles"]), "final_refits": int(stats["last_cycles"]),
                     "device_calls_per_step": int(stats["device_calls"])},
            "roofline": {"kernel": "gemm_tn_kernel<AdmmGemmPolicy> (two eigenbasis GEMMs per ADMM sweep)",
                         "bound": "tensor", "achieved": achieved, "peak": dgemm_tf, "unit": "TFLOP/s",
                         "frac": achieved / dgemm_tf if dgemm_tf > 0 else None,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_note": (f"ncu dram bytes of a launch with every column active; algorithmic "
                                          f"{traffic['algorithmic_bytes_per_launch']} B for that launch") if traffic else None,
                         "share_of_step": adm["ms_per_step"] / ms_prof if ms_prof > 0 else None,
                         "launches_per_step": adm["timed_groups"] / prof_steps,
                         "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 figure"},
            "kernel_classes_rank0": prof, "kernel_share_of_step": kernel_ms / ms_prof if ms_prof > 0 else None,
            "profiled_pass": {"ms_per_step": ms_prof, "steps": prof_steps,
                              "note": "min(K, 3) steps repeated with CUDA events around every kernel group; kernel_classes_rank0, "
                                      "roofline.achieved and the shares are measured in this pass and relate to its own time"},
            "clocks": clocks, "host_wall_ms_per_step": ms_wall, "assignment_checksum": checksum,
        }
        sec = []
        if hbm:
            el = prof["admm_elementwise"]
            if el["ms_per_step"] > 0:
                gbs = el["work_per_step"] / (el["ms_per_step"] * 1e-3) / 1e9
                sec.append({"kernel": "admm_rhs_kernel / admm_update_kernel / admm_finalize_kernel (ADMM elementwise sweep)",
                            "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                            "share_of_step": el["ms_per_step"] / ms_prof,
                            "bytes_note": "algorithmic bytes per active column: 56 P on plain sweeps, 120 P on rho-adaptation "
                                          "sweeps (DESIGN.md section 4; counted per sweep on the device)",
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy-measured)"})
            rv = prof["resid_vote"]
            if rv["ms_per_step"] > 0:
                gbs = rv["work_per_step"] / (rv["ms_per_step"] * 1e-3) / 1e9
                sec.append({"kernel": "resid_vote_kernel (+ sse_train_gram_kernel): fused predict / residual / SSE / vote",
                            "bound": "latency (DMMA chain + FP64 score chain per warp); HBM figure for reference",
                            "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                            "share_of_step": rv["ms_per_step"] / ms_prof,
                            "bytes_note": "algorithmic: 8 n2 T per fit + the gathered regulators / coefficients per tile",
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy-measured)"})
        nn = prof["nnls"]
        if nn["ms_per_step"] > 0:
            sec.append({"kernel": "nnls_kernel (one warp per (module, target) column)", "bound": "latency / issue",
                        "share_of_step": nn["ms_per_step"] / ms_prof, "ms_per_step": nn["ms_per_step"]})
        line["roofline_secondary"] = sec
        os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
        totals = grid_totals(res, n_modules)
        line["work_counters"] = totals
        r_med = res[i_med]
        first = {"lambda": float(pens[i_med]), "admm_iterations": [int(v) for v in r_med.admm_iterations[0]],
                 "selected": [np.flatnonzero(r_med.models_history[0][:, j]).tolist() for j in range(n_modules)]}
        if world == 1:
            try:
                tab = json.load(open(CYCLES_TABLE)) if os.path.exists(CYCLES_TABLE) else {}
                tab[f"{args.workload}/seed{args.seed}"] = {"totals": totals, "first_cycle": first,
                                                           "cycles_per_penalty": [int(r.n_cycles_run) for r in res]}
                json.dump(tab, open(CYCLES_TABLE, "w"), indent=1, sort_keys=True)
                side = os.path.join(ROOT, "gpurun_out")          # travels back from the GPU box
                if os.path.isdir(side):
                    json.dump(tab, open(os.path.join(side, "bench_cycles.json"), "w"), indent=1, sort_keys=True)
            except Exception:
                pass
            if not args.no_cpu_baseline:
                try:
                    gfc = {"admm_iterations": r_med.admm_iterations[0], "models": r_med.models_history[0]}
                    v, desc, threads, details = cpu_sample(w, args.cpu_budget, float(pens[i_med]), totals, gfc, args.cpu_threads)
                    line["parity_spot"] = details.pop("parity_spot", None)
                    line["cpu_baseline"] = {"value": v, "unit": "s", "cores": threads, "kind": "port", "sample": desc, **details}
                except Exception as exc:   # the checker library is optional for the GPU arm
                    line["cpu_baseline"] = {"value": None, "unit": "s", "cores": 0, "kind": "port",
                                            "sample": f"unavailable: {exc}"}
        emit(line)
    if world > 1:
        dist.barrier()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
