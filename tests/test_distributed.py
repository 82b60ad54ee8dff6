"""Multi-rank paths on the CPU: the partition rules exported by the library, and a world_size-2 gloo run of
the library's whole-grid loop (scrc_grid_fit_custom, csrc/grid.cu) on both ranks, each rank answering its cycles
from the sharded oracle session (tests/_sharded_oracle.py) that follows the ownership / merge protocol of the
CUDA session (csrc/session.cu) with gloo all-reduces in the place of NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_assign_is_a_balanced_partition():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _sharded_oracle import shard_assign
    rng = np.random.default_rng(5)
    for n, world in [(1, 1), (7, 2), (400, 8), (3, 8), (40, 4)]:
        cost = rng.integers(1, 1000, n).astype(np.float64)
        owner = shard_assign(cost, world)
        assert owner.min() >= 0 and owner.max() < world
        loads = np.bincount(owner, weights=cost, minlength=world)
        # longest-processing-time-first: no rank exceeds the mean by more than the largest single problem
        assert loads.max() <= cost.sum() / world + cost.max() + 1e-9
        assert np.array_equal(owner, shard_assign(cost, world))         # deterministic
    # ties: equal costs are dealt round-robin in index order
    assert shard_assign(np.ones(6), 3).tolist() == [0, 1, 2, 0, 1, 2]


def test_shard_slice_tiles_the_range():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _sharded_oracle import shard_slice
    for n, align, world in [(5000, 32, 8), (120, 32, 2), (31, 32, 4), (0, 32, 3), (10000, 32, 8), (65, 1, 64)]:
        edges = [shard_slice(n, align, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (a0, a1), (b0, b1) in zip(edges[:-1], edges[1:]):
            assert a1 == b0 and a0 <= a1
            assert a1 % align == 0 or a1 == n
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) < 2 * align      # one unit, plus the ragged last unit


def _worker_spmd(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from _sharded_oracle import ShardedOracleSession
    from scregclust_b200.driver import scregclust_fit
    from scregclust_b200.synthetic import make_workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(a):
        t = torch.from_numpy(a)            # shares memory: the reduction lands in the caller's array
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return a

    w = make_workload("tiny", seed=3, penalization=[0.1, 0.15, 0.2, 0.3])
    ses = ShardedOracleSession(w, rank, world, allreduce)
    res = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=ses, min_module_size=2, keep_diagnostics=True)
    np.savez(os.path.join(out, f"spmd{rank}.npz"), solved=ses.problems_solved,
             hist=np.array([len(r.k_history) for r in res]),
             **{f"k{i}": np.stack(r.k_history) for i, r in enumerate(res)},
             **{f"r2_{i}": r.r2[0] for i, r in enumerate(res)},
             **{f"sig_{i}": r.sigmas[0] for i, r in enumerate(res)},
             **{f"it_{i}": np.stack(r.admm_iterations) for i, r in enumerate(res)},
             **{f"co_{i}": np.concatenate([c.ravel() for c in r.coeffs[0] if c is not None] or [np.zeros(1)])
                for i, r in enumerate(res)})
    dist.barrier()
    dist.destroy_process_group()


def test_grid_loop_sharded_over_two_ranks_matches_the_oracle(tmp_path):
    import torch.multiprocessing as mp
    from oracle import scregclust_oracle as orc
    from scregclust_b200.synthetic import make_workload
    port = _free_port()
    mp.spawn(_worker_spmd, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "spmd0.npz"); b = np.load(tmp_path / "spmd1.npz")
    w = make_workload("tiny", seed=3, penalization=[0.1, 0.15, 0.2, 0.3])
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, min_module_size=2)
    # both ranks solved Step-1 problems, and between them every problem exactly once
    total = sum(int(np.count_nonzero(np.stack(r.admm_iterations))) for r in ref)
    assert int(a["solved"]) > 0 and int(b["solved"]) > 0 and int(a["solved"]) + int(b["solved"]) == total
    for i, r in enumerate(ref):
        for z in (a, b):                                             # every rank holds the complete result
            assert np.array_equal(z[f"k{i}"], np.stack(r.k_history))
            assert np.array_equal(z[f"it_{i}"], np.stack(r.admm_iterations))
            assert np.allclose(z[f"r2_{i}"], r.r2[0], rtol=0, atol=1e-12)
            assert np.allclose(z[f"sig_{i}"], r.sigmas[0], rtol=1e-12, atol=0)
            co = np.concatenate([c.ravel() for c in r.coeffs[0] if c is not None] or [np.zeros(1)])
            assert np.allclose(z[f"co_{i}"], co, rtol=1e-12, atol=1e-14)
        assert np.array_equal(a[f"r2_{i}"], b[f"r2_{i}"])            # and the ranks agree bit for bit


def test_sharded_session_single_rank_is_the_plain_loop():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _sharded_oracle import ShardedOracleSession
    from oracle import scregclust_oracle as orc
    from scregclust_b200.driver import scregclust_fit
    from scregclust_b200.synthetic import make_workload
    w = make_workload("tiny", seed=3)
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init)
    got = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=ShardedOracleSession(w))
    for g, r in zip(got, ref):
        assert all(np.array_equal(x, y) for x, y in zip(g.k_history, r.k_history))
        assert np.array_equal(g.models[0], r.models[0]) and np.allclose(g.r2[0], r.r2[0], atol=1e-12)
