"""world_size-2 gloo test of the penalty-grid sharding and the assignment all-gather (CPU)."""
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


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import scregclust_oracle as orc
    from scregclust_b200 import parallel
    from scregclust_b200.synthetic import make_workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = make_workload("tiny", seed=3, penalization=[0.1, 0.15, 0.2])
    idx = parallel.shard_penalties(len(w.penalization), rank, world)
    res = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds,
                             [w.penalization[i] for i in idx], w.n_modules, w.k_init)
    allk = parallel.gather_assignments([r.k_final[0] for r in res], len(w.penalization), w.z1_target.shape[1], rank, world)
    np.save(os.path.join(out, f"rank{rank}.npy"), allk)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import scregclust_oracle as orc
    from scregclust_b200 import parallel
    from scregclust_b200.synthetic import make_workload
    assert parallel.shard_penalties(5, 0, 2) == [0, 2, 4] and parallel.shard_penalties(5, 1, 2) == [1, 3]
    assert sorted(sum((parallel.shard_penalties(20, r, 8) for r in range(8)), [])) == list(range(20))
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npy"); b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)
    w = make_workload("tiny", seed=3, penalization=[0.1, 0.15, 0.2])
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init)
    assert np.array_equal(a, np.stack([r.k_final[0] for r in ref]))


def test_gather_single_rank():
    from scregclust_b200 import parallel
    k = [np.arange(5, dtype=np.int32), np.arange(5, dtype=np.int32) + 1]
    out = parallel.gather_assignments(k, 2, 5, 0, 1)
    assert np.array_equal(out, np.stack(k))


def _worker_dynamic(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from scregclust_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    store = dist.TCPStore("127.0.0.1", port + 17, world, rank == 0, wait_for_workers=True)
    counter = parallel.GridCounter(7, store, key="scrc_next_1")
    mine = {}
    while True:
        i = counter()
        if i is None:
            break
        mine[i] = np.full(5, i + 1, dtype=np.int32)          # stand-in for a fitted assignment vector
    table = parallel.merge_assignments(mine, 7, 5, world)
    np.save(os.path.join(out, f"dyn{rank}.npy"), table)
    np.save(os.path.join(out, f"mine{rank}.npy"), np.array(sorted(mine), dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_dynamic_counter_and_merge_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker_dynamic, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "dyn0.npy"); b = np.load(tmp_path / "dyn1.npy")
    assert np.array_equal(a, b)
    assert np.array_equal(a, np.repeat(np.arange(1, 8, dtype=np.int32)[:, None], 5, axis=1))   # every value fitted once
    m0 = set(np.load(tmp_path / "mine0.npy").tolist()); m1 = set(np.load(tmp_path / "mine1.npy").tolist())
    assert not (m0 & m1) and (m0 | m1) == set(range(7))


def _worker_driver(rank, world, port, out):
    """The multi-rank flow of bench.py on CPU: each rank runs the REAL driver loop (scregclust_fit with a shared
    atomic counter and 1 value in flight), with the device session replaced by the oracle-backed one."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from scregclust_b200 import parallel
    from scregclust_b200.driver import scregclust_fit
    from scregclust_b200.synthetic import make_workload
    from test_driver_host_logic import OracleSession
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    store = dist.TCPStore("127.0.0.1", port + 23, world, rank == 0, wait_for_workers=True)
    w = make_workload("tiny", seed=3, penalization=[0.1, 0.15, 0.2, 0.3])
    counter = parallel.GridCounter(len(w.penalization), store, key="scrc_next_1")
    res = scregclust_fit(penalization=w.penalization, n_modules=w.n_modules, initial_target_modules=w.k_init,
                         session=OracleSession(w), claim=counter, concurrency=1, min_module_size=2)
    mine = {i: r.k_final[0] for i, r in enumerate(res) if r is not None}
    table = parallel.merge_assignments(mine, len(w.penalization), w.z1_target.shape[1], world)
    np.save(os.path.join(out, f"drv{rank}.npy"), table)
    np.save(os.path.join(out, f"drvmine{rank}.npy"), np.array(sorted(mine), dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_driver_with_dynamic_claiming_world2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import scregclust_oracle as orc
    from scregclust_b200.synthetic import make_workload
    port = _free_port()
    mp.spawn(_worker_driver, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "drv0.npy"); b = np.load(tmp_path / "drv1.npy")
    assert np.array_equal(a, b)                                   # every rank ends with the whole table
    m0 = set(np.load(tmp_path / "drvmine0.npy").tolist()); m1 = set(np.load(tmp_path / "drvmine1.npy").tolist())
    assert not (m0 & m1) and (m0 | m1) == set(range(4))           # every value fitted exactly once
    w = make_workload("tiny", seed=3, penalization=[0.1, 0.15, 0.2, 0.3])
    ref = orc.scregclust_fit(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.penalization,
                             w.n_modules, w.k_init, min_module_size=2)
    assert np.array_equal(a, np.stack([r.k_final[0] for r in ref]))   # and the fits do not depend on who ran them
