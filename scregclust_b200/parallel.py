"""Multi-GPU plumbing of the penalty grid.

Penalization values are independent fits (R/scregclust.R:1223,1245), so the grid is sharded over
the ranks of a ``torch.distributed`` job (one process per GPU, NCCL over NVLink; gloo on CPU for
tests) with no data-path collective.  The only exchange is the per-penalty module-assignment
vector (int32 x T) -- and optionally the converged-module summaries -- all-gathered at the end.
"""
from __future__ import annotations

import numpy as np


def shard_penalties(n_penalties: int, rank: int, world: int):
    """Round-robin shard: small penalties (more selected regulators, more work) are interleaved."""
    return list(range(n_penalties))[rank::world]


def slots_per_rank(n_penalties: int, world: int) -> int:
    return max(1, (n_penalties + world - 1) // world)


def gather_assignments(local_k, n_penalties: int, n_targets: int, rank: int, world: int, device="cpu"):
    """All-gather the final module assignment of every penalization value.

    ``local_k``: list of int32 arrays (T,), one per local penalty in shard order.  Returns an
    ``(n_penalties, T)`` int32 array identical on every rank (rows of penalties this job did not
    fit stay at the sentinel -9)."""
    import torch
    import torch.distributed as dist

    slots = slots_per_rank(n_penalties, world)
    mine = torch.full((slots, n_targets), -9, dtype=torch.int32, device=device)
    for i, k in enumerate(local_k):
        mine[i] = torch.as_tensor(np.ascontiguousarray(k, dtype=np.int32)).to(device)
    if world > 1:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
    else:
        parts = [mine]
    out = np.full((n_penalties, n_targets), -9, dtype=np.int32)
    for r in range(world):
        for i, p in enumerate(shard_penalties(n_penalties, r, world)):
            out[p] = parts[r][i].cpu().numpy()
    return out


class GridCounter:
    """Atomic "next penalization value" counter shared by all ranks (dynamic load balancing).

    The cost of a fit is not known in advance (small penalties select more regulators and need
    more cycles), so ranks claim values one at a time, cheapest last (ascending penalty =
    descending expected cost), from a counter kept in the job's TCP store.  Host side only; no
    GPU data moves."""

    def __init__(self, n_penalties: int, store=None, key="scrc_next"):
        self.n = int(n_penalties)
        self.store = store
        self.key = key
        self._local = 0

    def reset(self, key: str):
        self.key = key
        self._local = 0

    def __call__(self):
        if self.store is None:
            i = self._local
            self._local += 1
        else:
            i = int(self.store.add(self.key, 1)) - 1
        return i if i < self.n else None


def merge_assignments(local: dict, n_penalties: int, n_targets: int, world: int, device="cpu"):
    """Every rank contributes the rows (penalization values) it fitted; an all-reduce(MAX) over an
    (L, T) int32 table filled with the sentinel -9 gives every rank the complete table."""
    import torch
    import torch.distributed as dist

    tab = torch.full((n_penalties, n_targets), -9, dtype=torch.int32, device=device)
    for i, k in local.items():
        tab[i] = torch.as_tensor(np.ascontiguousarray(k, dtype=np.int32)).to(device)
    if world > 1:
        dist.all_reduce(tab, op=dist.ReduceOp.MAX)
    return tab.cpu().numpy()
