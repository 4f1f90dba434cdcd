"""Multi-GPU plumbing of the energy path for hosts that do the reduction themselves with torch.distributed
(api.Engine.energy_distributed / first_order_distributed without an in-library communicator): one process per GPU, the
work items of the tile pass are dealt over the ranks (work stealing inside each GPU), and ONE all-reduce sums the packed
accumulators [E2 (grid multiple), screening counters..., E2 (rest)] -- the replacement of the reference's four MPI
all-reduces per energy (xm_equalize / xm_equalize_scalar, /root/reference/src/xm_module.F90:853-910).  With
Engine.attach_comm the library does all of this itself over NCCL (csrc/vb_nccl.cpp)."""
from __future__ import annotations

from typing import List


def shard(ntiles: int, rank: int, nranks: int) -> range:
    """Work items owned by `rank` when items are dealt round-robin (make_items, csrc/vb_tilelist.cpp)."""
    return range(rank, ntiles, nranks)


def allreduce_sum(t):
    """Sum a tensor over all ranks in place (NCCL for CUDA tensors, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def combine_partials(partials: List[List[float]]) -> List[float]:
    """Reference semantics of xm_equalize_scalar: plain sums of the rank partials."""
    return [sum(col) for col in zip(*partials)]
