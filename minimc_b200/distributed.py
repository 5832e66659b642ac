"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL (gloo in CPU tests).

Histories of a fixed-source problem are independent (history i is fully determined by seed + i,
FixedSource.cpp:61), so the path shards with NO data-path collective: rank r of P owns the contiguous range
[r*N/P, (r+1)*N/P) and the only exchange is one all-reduce of the integer tallies and counters at the end
(the multi-GPU form of `solver_estimator_set += worker_estimator_set.get()`, FixedSource.cpp:31-33).  The sums
are exact 64-bit integers, so the result is bit-identical for any P.
"""
from __future__ import annotations


def shard(first: int, n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous history range of `rank`: [first + rank*n//P, first + (rank+1)*n//P)."""
    lo = rank * n // world_size
    hi = (rank + 1) * n // world_size
    return first + lo, hi - lo


def allreduce_sum_(*tensors, group=None) -> None:
    """In-place SUM all-reduce of int64 tally / counter tensors (device tensors under NCCL)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
