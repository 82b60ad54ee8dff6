#!/usr/bin/env python
"""bench.py -- scregclust wall time per penalty grid on N B200s, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config3|...] [--impl reference]

A *step* is one complete penalization-grid fit of the alternating loop (R/scregclust.R:996-1912:
Gram / eigenbasis / initial fit, then Step 1 + 2a + 2b cycles until every penalty has converged,
its final configurations are re-fitted and their leave-one-regulator-out importances computed,
i.e. the default compute_predictive_r2 = TRUE) on one seeded synthetic workload of a BASELINE.json
shape.  With N > 1 (torchrun, one rank per GPU) the penalty grid is sharded round-robin over the
ranks -- the fits are independent (R/scregclust.R:1223,1245) -- and the per-penalty module
assignments are all-gathered over NCCL at the end of the step ("scaling": "strong": the grid is
fixed, the wall time per grid falls with N).

Printed JSON (rank 0, one line):
  value        device-resident step time [s] (inputs already in HBM when the clock starts)
  e2e          the same step through the public host API: host (pinned) buffers -> C ABI -> results
               back on the host, copies inside the timed region
  roofline     dominant kernel class (the two eigenbasis GEMMs of every ADMM sweep): algorithmic
               flops / CUDA-event time on the launching stream, against cuBLAS DGEMM measured in
               this run (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline reference-faithful C++ twin (oracle/_ref/liboracle_ref.so) timed on a bounded sample
               of the same workload and extrapolated to the grid (N=1, rank 0 only)
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
def cpu_sample(w, budget_s, total_fit_cycles, total_last_cycles, lam):
    """Times one cycle (Step 1 + 2a + 2b) of one penalization value with the cost structure of the
    reference (oracle/ref_cpu.cpp) on as many modules / genes as fit the budget, and extrapolates
    to the whole grid.  Returns (estimated seconds per grid, description, threads)."""
    from oracle import ref_cpu
    from oracle import scregclust_oracle as orc

    threads = ref_cpu.set_threads(os.cpu_count() or 1)
    n1, P = w.z1_reg.shape
    n2, T = w.z2_target.shape
    K = w.n_modules
    t_pre0 = time.perf_counter()
    pre = orc.precompute(w.z1_reg, w.z1_target)               # R BLAS path of R/scregclust.R:1049-1113
    t_pre = time.perf_counter() - t_pre0
    res_sds, bas = pre["res_sds"], pre["beta_adapt_sq"]
    x = np.asfortranarray(w.z1_reg / np.sqrt(n1))
    k = w.k_init
    # ---- Step 1: coop_lasso per module (Gram recomputed, LDL^T per iteration) ----
    t1, done1, models, signs, admm_it = 0.0, 0, {}, {}, 0
    for j in range(1, K + 1):
        cl = np.flatnonzero(k == j)
        if cl.size == 0:
            continue
        if t1 > budget_s * 0.6 and done1 >= 1:
            break
        t0 = time.perf_counter()
        y = np.asfortranarray((w.z1_target[:, cl] / res_sds[cl][None, :]) / np.sqrt(n1))
        ws = np.sqrt(cl.size) / np.sqrt(np.sqrt(np.sum(bas[:, cl], axis=1)))
        beta, it, nf = ref_cpu.coop_lasso(y, x, lam, ws, max_iter=10000)
        beta = beta * res_sds[cl][None, :]
        beta[np.abs(beta) < 1e-6] = 0.0
        t1 += time.perf_counter() - t0
        models[j] = np.flatnonzero(np.sum(np.abs(beta) > 0, axis=1) > 0)
        signs[j] = np.sign(beta.sum(axis=1))[models[j]]
        admm_it += it
        done1 += 1
    nonempty = int(sum(1 for j in range(1, K + 1) if np.any(k == j)))
    step1_est = t1 * nonempty / max(done1, 1)
    # ---- Step 2a for the modules fitted above: NNLS on all T targets + residual GEMMs ----
    t2, done2 = 0.0, 0
    y_nn = np.asfortranarray(w.z1_target / w.z1_target_sds[None, :])
    for j, reg in models.items():
        if reg.size == 0:
            continue
        if t2 > budget_s * 0.25 and done2 >= 1:
            break
        t0 = time.perf_counter()
        xs = np.asfortranarray(w.z1_reg[:, reg] * signs[j][None, :])
        b, _ = ref_cpu.coef_nnls(xs, y_nn, eps=1e-4, max_iter=10000)
        bh = b * signs[j][:, None] * orc._recycle(w.z1_target_sds, b.shape)
        r_tr = w.z1_target - w.z1_reg[:, reg] @ bh
        r_te = w.z2_target - w.z2_reg[:, reg] @ bh
        _ = (r_tr * r_tr).sum(axis=0), (r_te * r_te).sum(axis=0)
        t2 += time.perf_counter() - t0
        done2 += 1
    step2a_est = t2 * nonempty / max(done2, 1)
    # ---- array reset + allocation on a gene subset (the full K x n2 x T array is 8 GB at config 3) ----
    g = int(max(8, min(T, 2.5e8 // (K * n2 * 8))))
    z2sq = np.asfortranarray(w.z2_target[:, :g] ** 2)
    arr = np.empty((K, n2, g), order="F")
    t0 = time.perf_counter()
    ref_cpu.fill_array(arr, z2sq)                             # reset_array (src/utils.cpp:43-63)
    arr[0, :, :] = z2sq * 0.9                                 # one strided module write (scregclust.R:1599)
    t_fill = time.perf_counter() - t0
    rv = np.asfortranarray(np.ones((g, K)))
    t0 = time.perf_counter()
    ref_cpu.allocate_into_modules(arr, rv, None, None, np.ones(g, dtype=np.int32), np.arange(g, dtype=np.int32), 1e-6, 0.0)
    t_alloc = time.perf_counter() - t0
    step2b_est = (t_fill * (1 + (nonempty - 1) * 0.5) + t_alloc) * T / g
    # ---- post-processing: leave-one-regulator-out NNLS importance (R/scregclust.R:1823-1912) on one module ----
    t_imp, imp_mods = 0.0, 0
    for j, reg in models.items():
        cl = np.flatnonzero(k == j)
        if reg.size < 2 or cl.size == 0:
            continue
        t0 = time.perf_counter()
        y_cl = np.asfortranarray(w.z1_target[:, cl] / w.z1_target_sds[cl][None, :])
        for r in range(reg.size + 1):
            keep = np.ones(reg.size, dtype=bool)
            if r > 0:
                keep[r - 1] = False
            xs = np.asfortranarray(w.z1_reg[:, reg[keep]] * signs[j][keep][None, :])
            b, _ = ref_cpu.coef_nnls(xs, y_cl, eps=1e-4, max_iter=10000)
            bh = b * signs[j][keep][:, None] * orc._recycle(w.z1_target_sds[cl], b.shape)
            rt = w.z2_target[:, cl] - w.z2_reg[:, reg[keep]] @ bh
            _ = (rt * rt).sum(axis=0)
        t_imp += time.perf_counter() - t0
        imp_mods += 1
        break
    imp_est = t_imp * nonempty / max(imp_mods, 1)
    cycle_est = step1_est + step2a_est + step2b_est
    last_est = step1_est + step2a_est + imp_est
    total = t_pre + cycle_est * total_fit_cycles + last_est * total_last_cycles
    desc = (f"reference-faithful C++ twin (oracle/ref_cpu.cpp: Gram per call, LDL^T per iteration, K x n2 x T array), "
            f"lambda={lam:.3g}, first cycle from the initial configuration: Step 1 on {done1}/{nonempty} modules "
            f"({admm_it} ADMM iterations, {t1:.1f} s), Step 2a on {done2} modules ({t2:.1f} s), reset+allocate on {g}/{T} genes "
            f"({t_fill + t_alloc:.2f} s), importance refits of {imp_mods} module ({t_imp:.1f} s), precompute {t_pre:.1f} s; extrapolated to {total_fit_cycles} fit cycles + "
            f"{total_last_cycles} final re-fits of the grid")
    return total, desc, threads


def load_cycles(name, seed, n_pen):
    """(fit cycles, final re-fit cycles) summed over the grid, from the last GPU run of this workload."""
    try:
        tab = json.load(open(CYCLES_TABLE))
        e = tab[f"{name}/seed{seed}"]
        return int(e["fit_cycles"]), int(e["last_cycles"]), "profiles/bench_cycles.json"
    except Exception:
        return 6 * n_pen, n_pen, "assumed 6 cycles per penalization value (no cycle table for this workload)"


# --------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def in_flight(n_pens, world):
    """Penalization values a rank advances together when the grid is shared dynamically."""
    if os.environ.get("SCRC_BENCH_CONCURRENCY"):
        return max(1, int(os.environ["SCRC_BENCH_CONCURRENCY"]))
    return max(1, n_pens // (2 * world))


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
    config = {"workload": f"{args.workload}: synthetic {n_cells} cells x {n_targets} targets x {n_reg} regulators, "
                          f"{n_modules} modules, {len(pens)}-value penalty grid (seed {args.seed})",
              "n1": n_cells // 2, "n2": n_cells - n_cells // 2, "targets": n_targets, "regulators": n_reg,
              "modules": n_modules, "penalties": [float(p) for p in pens],
              "baseline_config": {"config1": "BASELINE.json configs[0] (synthetic stand-in of the same shape)",
                                  "config2": "BASELINE.json configs[1]", "config3": "BASELINE.json configs[2]",
                                  "config4": "BASELINE.json configs[3]"}.get(args.workload, "test-sized workload"),
              "parallelism": (f"penalization values claimed dynamically from a shared counter by {world} GPUs, "
                               f"{in_flight(len(pens), world)} in flight per GPU"
                              if world > 1 else "whole penalty grid advanced in lock-step on 1 GPU"),
              "l2": "working set (FP64 state of all problems) exceeds the 126 MB L2; no flush between steps"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        w = make_workload(args.workload, seed=args.seed)
        fit_c, last_c, src = load_cycles(args.workload, args.seed, len(pens))
        vals = []
        desc, threads = "", 1
        for _ in range(max(1, min(args.steps, 2))):            # each step = one bounded sample
            v, desc, threads = cpu_sample(w, args.cpu_budget, fit_c, last_c, float(np.median(pens)))
            vals.append(v)
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "s", "cores": threads, "kind": "port",
                                 "sample": desc + f" [cycle counts: {src}]"},
                "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist

    from scregclust_b200 import _capi, api
    from scregclust_b200.driver import Session, scregclust_fit

    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _capi.load()
    api.set_device(local_rank)

    w = make_workload(args.workload, seed=args.seed)
    from scregclust_b200 import parallel
    all_pens = [float(p) for p in pens]
    store = None
    if world > 1:   # shared counter for dynamic claiming of penalization values (host side)
        store = dist.TCPStore("127.0.0.1", int(os.environ.get("MASTER_PORT", "29500")) + 17, world, rank == 0,
                              wait_for_workers=True)
    counter = parallel.GridCounter(len(all_pens), store)
    step_no = [0]
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
        if resident:
            ses = Session.from_pointers(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(),
                                        sds_dev.data_ptr(), n1, n2, P, T, n_modules, device=local_rank)
        else:
            ses = Session.from_pointers(pinned[0].data_ptr(), pinned[1].data_ptr(), pinned[2].data_ptr(),
                                        pinned[3].data_ptr(), sds_pinned.data_ptr(), n1, n2, P, T, n_modules,
                                        device=local_rank)
            ses.h2d_bytes = sum(t.numel() * 8 for t in pinned) + sds_pinned.numel() * 8
        try:
            t_b = time.perf_counter()
            step_no[0] += 1
            counter.reset(f"scrc_next_{step_no[0]}")
            # one GPU: the whole grid advances in lock-step; several GPUs: each rank keeps
            # L / (2 N) penalization values in flight and claims the next one from the shared counter
            # (most expensive first) whenever one of them finishes -- the longest fit (about twice
            # the average number of cycles) then still fits inside a rank's share of the work
            allres = scregclust_fit(penalization=all_pens, n_modules=n_modules, initial_target_modules=w.k_init,
                                    session=ses, return_coeffs=not resident, compute_predictive_r2=True,
                                    claim=counter if world > 1 else None,
                                    concurrency=in_flight(len(all_pens), world) if world > 1 else None)
            res = [r for r in allres if r is not None]
            local_k = {i: r.k_final[0] for i, r in enumerate(allres) if r is not None}
            stats = dict(h2d=ses.h2d_bytes, d2h=ses.d2h_bytes, sweeps=ses.admm_sweeps,
                         admm_iters=ses.admm_problem_iterations,
                         fit_cycles=sum(r.n_cycles_run - 1 for r in res),
                         last_cycles=sum(r.final_cycle_length for r in res))
            t_c = time.perf_counter()
        finally:
            ses.close()
        t_d = time.perf_counter()
        # exchange step: every rank learns every penalty's final module assignment (int32 x T)
        allk = parallel.merge_assignments(local_k, len(pens), T, world, device="cuda")
        assert allk.min() >= -1, "a penalization value was not fitted by any rank"
        checksum = int(np.clip(allk, 0, None).sum())
        if dbg:
            print(f"[bench step {step_no[0]} resident={resident}] create {t_b - t_a:.3f} s, fit {t_c - t_b:.3f} s, "
                  f"close {t_d - t_c:.3f} s, merge {time.perf_counter() - t_d:.3f} s", file=sys.stderr, flush=True)
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

    for _ in range(max(args.warmup, 0)):
        run_step(True)
    sampler = ClockSampler(local_rank); sampler.start()
    _capi.check(lib.scrc_profile_reset()); _capi.check(lib.scrc_profile_enable(1))
    launches0 = api.launch_count()
    ms_dev, ms_wall, (res, stats, checksum) = timed(True, args.steps)
    launches = api.launch_count() - launches0
    pm = (C.c_double * 6)(); pw = (C.c_double * 6)(); pl = (C.c_ulonglong * 6)()
    _capi.check(lib.scrc_profile_get(pm, pw, pl)); _capi.check(lib.scrc_profile_enable(0))
    clocks = sampler.stop()
    run_step(False)                                  # untimed: first touch of the host-side result buffers
    e2e_dev, e2e_wall, (_, e2e_stats, _) = timed(False, max(1, min(args.steps, 2)))

    if rank == 0:
        names = ["gemm_dense", "gemm_admm", "admm_elementwise", "nnls", "resid_vote", "other"]
        prof = {n: {"ms_per_step": pm[i] / args.steps, "work_per_step": pw[i] / args.steps, "timed_groups": int(pl[i])}
                for i, n in enumerate(names)}
        adm = prof["gemm_admm"]
        achieved = adm["work_per_step"] / (adm["ms_per_step"] * 1e-3) / 1e12 if adm["ms_per_step"] > 0 else 0.0
        sweeps = max(stats["sweeps"], 1)
        traffic = None
        try:   # DRAM bytes of one (all-columns-active) launch of the dominant kernel from the committed ncu capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": ms_dev * 1e-3, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_dev * 1e-3, "unit": "s", "h2d_bytes_per_step": int(e2e_stats["h2d"]),
                    "d2h_bytes_per_step": int(e2e_stats["d2h"]), "host_wall_s": e2e_wall * 1e-3},
            "gpu_launches": int(launches),
            "admm_iters_per_s": stats["admm_iters"] / (prof["gemm_admm"]["ms_per_step"] + prof["admm_elementwise"]["ms_per_step"] + 1e-9) * 1e3,
            "admm": {"problem_iterations_rank0": int(stats["admm_iters"]), "batched_sweeps_rank0": int(stats["sweeps"]),
                     "fit_cycles_rank0": int(stats["fit_cycles"]), "final_refits_rank0": int(stats["last_cycles"])},
            "roofline": {"kernel": "gemm_tn_kernel<AdmmGemmPolicy> (two eigenbasis GEMMs per ADMM sweep)",
                         "bound": "tensor", "achieved": achieved, "peak": dgemm_tf, "unit": "TFLOP/s",
                         "frac": achieved / dgemm_tf if dgemm_tf > 0 else None,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_note": (f"ncu dram bytes of a launch with every column active; algorithmic "
                                          f"{traffic['algorithmic_bytes_per_launch']} B for that launch") if traffic else None,
                         "share_of_step": adm["ms_per_step"] / ms_dev if ms_dev > 0 else None,
                         "launches_per_step": 2 * sweeps / max(args.steps, 1),
                         "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 figure"},
            "kernel_classes_rank0": prof, "clocks": clocks, "host_wall_ms_per_step": ms_wall,
            "assignment_checksum": checksum,
        }
        try:   # second-largest HBM-bound class against the driver-measured copy bandwidth
            el = prof["admm_elementwise"]
            hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            gbs = el["work_per_step"] / (el["ms_per_step"] * 1e-3) / 1e9
            line["roofline_secondary"] = [{
                "kernel": "admm_rhs_kernel / admm_update_kernel (ADMM elementwise sweep)", "bound": "hbm",
                "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                "share_of_step": el["ms_per_step"] / ms_dev,
                "bytes_note": "algorithmic 80 B per active matrix element and sweep (DESIGN.md section 4)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy-measured)"}]
        except Exception:
            pass
        os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
        if world == 1:
            try:
                tab = json.load(open(CYCLES_TABLE)) if os.path.exists(CYCLES_TABLE) else {}
                tab[f"{args.workload}/seed{args.seed}"] = {"fit_cycles": int(stats["fit_cycles"]), "last_cycles": int(stats["last_cycles"]),
                                                           "cycles_per_penalty": [int(r.n_cycles_run) for r in res]}
                json.dump(tab, open(CYCLES_TABLE, "w"), indent=1, sort_keys=True)
            except Exception:
                pass
            if not args.no_cpu_baseline:
                try:
                    v, desc, threads = cpu_sample(w, args.cpu_budget, stats["fit_cycles"], stats["last_cycles"],
                                                  float(np.median(pens)))
                    line["cpu_baseline"] = {"value": v, "unit": "s", "cores": threads, "kind": "port", "sample": desc}
                except Exception as exc:   # the checker library is optional for the GPU arm
                    line["cpu_baseline"] = {"value": None, "unit": "s", "cores": 0, "kind": "port",
                                            "sample": f"unavailable: {exc}"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
