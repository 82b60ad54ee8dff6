"""Calibrates the device-side tile-shape rule of the ADMM GEMMs (csrc/admm.cu: small_cost): runs the config-3 shape with
3 penalization values (about the per-GPU column count of an 8-GPU run) and with all 20, for several relative costs of a
64 x 64 tile, and prints the ADMM GEMM kernel time of each."""
import sys, os, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import ctypes as C, numpy as np
    from scregclust_b200 import _capi
    from scregclust_b200.synthetic import make_workload
    from scregclust_b200.driver import Session, scregclust_fit
    w = make_workload("config3", seed=0)
    pens = w.penalization if sys.argv[2] == "all" else w.penalization[[0, 8, 19]]
    lib = _capi.load()
    for rep in range(2):
        _capi.check(lib.scrc_profile_reset()); _capi.check(lib.scrc_profile_enable(1))
        with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules) as s:
            res = scregclust_fit(penalization=pens, n_modules=w.n_modules, initial_target_modules=w.k_init, session=s, return_coeffs=False)
        pm = (C.c_double * 6)(); pw = (C.c_double * 6)(); pl = (C.c_ulonglong * 6)()
        _capi.check(lib.scrc_profile_get(pm, pw, pl)); _capi.check(lib.scrc_profile_enable(0))
    print(json.dumps({"gemm_admm_ms": pm[1], "tflops": pw[1] / pm[1] / 1e9, "elem_ms": pm[2],
                      "checksum": int(sum(int(np.clip(r.k_final[0], 0, None).sum()) for r in res))}))
else:
    for which in ("three", "all"):
        for cost in ("0.20", "0.22", "0.235", "0.25", "0.267"):
            env = dict(os.environ, SCRC_SMALL_TILE_COST=cost)
            out = subprocess.run([sys.executable, __file__, "child", which], env=env, capture_output=True, text=True)
            print(which, "small_cost", cost, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
