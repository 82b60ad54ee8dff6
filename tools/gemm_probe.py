"""GEMM probe: our TN DMMA kernel vs cuBLAS DGEMM (torch.matmul float64) on the same box."""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scregclust_b200 import _capi

lib = _capi.load()
out = {}
def ours(k, m, n, sym=0, reps=5):
    ms = C.c_double(0)
    _capi.check(lib.scrc_bench_gemm(k, m, n, sym, reps, C.byref(ms)))
    fl = (k * m * (m + 1) if sym else 2.0 * k * m * n)
    return ms.value, fl / ms.value / 1e9
def cublas(k, m, n, reps=5):
    a = torch.randn(m, k, device="cuda", dtype=torch.float64); b = torch.randn(k, n, device="cuda", dtype=torch.float64)
    for _ in range(2): (a @ b)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): c = a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * k * m * n / ms / 1e9
for (k, m, n) in [(8192, 8192, 8192), (10000, 1600, 5000), (1600, 1600, 100000), (2500, 500, 2000), (500, 500, 20000)]:
    o = ours(k, m, n); c = cublas(k, m, n)
    out[f"{k}x{m}x{n}"] = {"ours_ms": o[0], "ours_tflops": o[1], "cublas_ms": c[0], "cublas_tflops": c[1]}
    print(k, m, n, "ours %.3f ms %.2f TF | cublas %.3f ms %.2f TF" % (o[0], o[1], c[0], c[1]), flush=True)
o = ours(10000, 1600, 1600, sym=1)
print("sym 10000x1600: %.3f ms %.2f TF (algorithmic n*p*(p+1))" % o)
out["sym_10000x1600"] = {"ours_ms": o[0], "ours_tflops": o[1]}
json.dump(out, open("gpurun_out/gemm_probe.json", "w"), indent=1)
