"""Debug: check ZselT gather and run-to-run determinism of one fit_cycle."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200.synthetic import make_workload
from scregclust_b200.driver import Session
from scregclust_b200 import _capi

w = make_workload(sys.argv[1] if len(sys.argv) > 1 else "small", seed=int(sys.argv[2]) if len(sys.argv) > 2 else 2, ar1=float(sys.argv[3]) if len(sys.argv) > 3 else 0.5)
K = w.n_modules
lib = _capi.load()
lib.scrc_debug_get_zsel.restype = C.c_int
lib.scrc_debug_get_zsel.argtypes = [C.c_void_p, _capi.c_double_p, C.POINTER(C.c_longlong)]
with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, K) as s:
    L = len(w.penalization)
    outs = []
    for rep in range(3):
        res = s.fit_cycle(w.penalization, np.stack([w.k_init] * L), want_votes=True, want_nnls_iters=True)
        outs.append(res)
    for rep in (1, 2):
        for key in ("resid_var", "r2_test", "votes"):
            d = np.abs(outs[rep][key].astype(np.float64) - outs[0][key].astype(np.float64))
            print("rep", rep, key, "max abs diff vs rep0", d.max(), "n differing", int((d > 0).sum()))
    res = outs[0]
    s._last_nsel = res["n_selected"]
    rows = C.c_longlong(0)
    _capi.check(lib.scrc_debug_get_zsel(s._h, None, C.byref(rows)))
    zs = np.zeros((rows.value, s.n2), order="F")
    _capi.check(lib.scrc_debug_get_zsel(s._h, _capi.dptr(zs), C.byref(rows)))
    ro = 0
    n1 = s.n1
    for f in range(L):
        for j in range(K):
            co, sel = s.coeffs(f, j)
            sj = len(sel)
            if sj == 0: continue
            kp = (sj + 15) // 16 * 16
            blk = zs[ro:ro + kp]
            ok = np.array_equal(blk[:sj], w.z2_reg[:, sel].T) and np.all(blk[sj:] == 0)
            # recompute SSE test / train with numpy from the GPU's own coefficients
            rt = w.z2_target - w.z2_reg[:, sel] @ co
            sse_t = (rt * rt).sum(axis=0)
            r2_np = 1 - sse_t / (w.z2_target ** 2).sum(axis=0)
            rtr = w.z1_target - w.z1_reg[:, sel] @ co
            var_np = (rtr * rtr).sum(axis=0) / (n1 - sj)
            e_r2 = np.abs(res["r2_test"][f][j] - r2_np).max()
            e_var = (np.abs(res["resid_var"][f][j] - var_np) / var_np).max()
            print(f"fit {f} module {j} s={sj} ro={ro} gather_ok={ok} r2 err {e_r2:.2e} var relerr {e_var:.2e}")
            ro += kp
