"""Launcher glue of a multi-GPU job.

The multi-GPU machinery lives in the library: a session created with a communicator shards every
cycle over the ranks (Step-1 problems by (penalization value, module), Step 2 by target-gene slice) and
merges the module summaries and assignment vectors with NCCL collectives (csrc/session.cu,
csrc/comm.cu).  What is left on the Python side is handing NCCL's unique id from rank 0 to the other
ranks of a ``torchrun`` job -- any broadcast the launcher offers will do; here ``torch.distributed``.
"""
from __future__ import annotations


def comm_from_torch(device: int):
    """One :class:`scregclust_b200.driver.Comm` per rank of the initialised ``torch.distributed`` job
    (``None`` when the job has a single rank)."""
    import torch
    import torch.distributed as dist

    from .driver import Comm

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    token = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        token = torch.frombuffer(bytearray(Comm.new_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(token, src=0)
    return Comm(bytes(token.cpu().numpy().tobytes()), world, rank, device)
