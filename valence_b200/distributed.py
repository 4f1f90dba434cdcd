"""Multi-GPU plumbing of the energy path: one process per GPU, the tile list of the fused
ERI + contraction pass is sharded block-cyclically over ranks (tile k -> rank k mod N, work
stealing inside each GPU), and ONE all-reduce sums the packed accumulators
[E2, screening counters...] -- the replacement of the reference's four MPI all-reduces per
energy (xm_equalize / xm_equalize_scalar, /root/reference/src/xm_module.F90:853-910)."""
from __future__ import annotations

from typing import List


def shard(ntiles: int, rank: int, nranks: int) -> range:
    """Tiles owned by `rank` (mirrors TileArgs.tile_first / tile_stride in csrc/vb_tile.cuh)."""
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
