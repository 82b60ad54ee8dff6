"""Generates the golden fixtures in this directory from the NumPy oracle (seeded).

The reference ships no golden vectors for the hot path and cannot be executed in the build
container (no R / Rcpp / Eigen), so these vectors pin the *oracle* (regression protection and
a fixed target for the GPU parity tests); they are not outputs of the reference binary.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import scregclust_oracle as orc  # noqa: E402
from scregclust_b200.synthetic import make_workload  # noqa: E402


def loop_fixture():
    w = make_workload("tiny", seed=3)
    res = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, compute_predictive_r2=True)
    out = {}
    for i, r in enumerate(res):
        out[f"k_history_{i}"] = np.stack(r.k_history)
        out[f"admm_iterations_{i}"] = np.stack(r.admm_iterations)
        out[f"models_{i}"] = r.models[0]
        out[f"r2_{i}"] = r.r2[0]
        out[f"sigmas_{i}"] = r.sigmas[0]
        out[f"importance_{i}"] = r.importance[0]
    np.savez_compressed(os.path.join(HERE, "loop_tiny_seed3.npz"), **out)


def kernel_fixture():
    rng = np.random.default_rng(7)
    n, p, m = 300, 30, 12
    x = rng.standard_normal((n, p)) / np.sqrt(n)
    b = np.zeros((p, m)); b[:6] = rng.standard_normal((6, 1)) + 0.3 * rng.standard_normal((6, m))
    y = x @ b + rng.standard_normal((n, m)) / np.sqrt(n)
    w = np.sqrt(m) * (0.5 + rng.random(p))
    beta, it = orc.coop_lasso(y, x, 0.12, w, max_iter=10000)
    xs = rng.standard_normal((250, 9)); ys = xs @ np.abs(rng.standard_normal((9, 16))) + rng.standard_normal((250, 16))
    nb, nit = orc.coef_nnls(xs, ys, eps=1e-4, max_iter=10000)
    K, n2, T = 4, 80, 20
    arr = np.asfortranarray(rng.random((K, n2, T))); arr[3] = arr[2]
    var = np.asfortranarray(rng.random((T, K)) + 0.5); var[:, 3] = var[:, 2]
    k = rng.integers(1, K + 1, T).astype(np.int32); k[3] = -1
    order = rng.permutation(T).astype(np.int32)
    kout = orc.allocate_into_modules(arr, var, [], k, order, 1e-6, 0.0)
    np.savez_compressed(os.path.join(HERE, "kernels_seed7.npz"), coop_x=x, coop_y=y, coop_w=w, coop_lambda=0.12,
                        coop_beta=beta, coop_iterations=it, nnls_x=xs, nnls_y=ys, nnls_beta=nb, nnls_iterations=nit,
                        alloc_arr=arr, alloc_var=var, alloc_k=k, alloc_order=order, alloc_out=kout)


def aggregate_fixture():
    """allocate_per_obs = FALSE (R/scregclust.R:1678-1728): one reassignment with a prior, one without, and the loop."""
    rng = np.random.default_rng(9)
    T, K, n2 = 50, 4, 120
    ss = rng.random((T, K)) * 60 + 30; rv = rng.random((T, K)) * 0.6 + 0.2
    k = rng.integers(1, K + 1, T).astype(np.int32); k[:4] = -1
    prior = [sorted(rng.choice(T, size=3, replace=False).tolist()) for _ in range(T)]
    order = rng.permutation(T).astype(np.int32)
    plain = orc.allocate_aggregate(k, ss, rv, n2, (), np.arange(T), 1e-6, 0.0)
    withp = orc.allocate_aggregate(k, ss, rv, n2, prior, order, 1e-6, 0.5)
    w = make_workload("tiny", seed=3)
    res = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, allocate_per_obs=False, min_module_size=2)
    out = dict(ss=ss, rv=rv, k=k, n2=n2, prior=np.array(prior, dtype=np.int32), order=order, out_plain=plain,
               out_prior=withp)
    for i, r in enumerate(res):
        out[f"k_history_{i}"] = np.stack(r.k_history)
    np.savez_compressed(os.path.join(HERE, "aggregate_seed9.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "aggregate":     # added in round 2; leaves the older files untouched
        aggregate_fixture()
        sys.exit(0)
    loop_fixture()
    kernel_fixture()
    aggregate_fixture()
    print("golden fixtures written to", HERE)
