"""Batch data parallelism for the biGRU path: one process per GPU, replicated parameters, the batch
sharded across ranks, ONE all-reduce of the flat gradient vector per step (NCCL over NVLink /
NVSwitch on the GPU box, gloo in the CPU tests).  The reference has no distributed code at all
(SURVEY.md section 2.1); sequences in a batch are independent (biGRU_model.py:63-138 has no
cross-sample operation), so the only exchange the path needs is the gradient sum."""
from __future__ import annotations

import torch
import torch.distributed as dist


def allreduce_flat_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a contiguous tensor."""
    if not t.is_contiguous():
        raise ValueError("allreduce_flat_ needs a contiguous tensor")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, equal shards (the global batch must divide evenly so that every rank's mean-loss
    gradient carries the same weight)."""
    if n_items % world:
        raise ValueError(f"global batch {n_items} is not divisible by world size {world}")
    per = n_items // world
    return rank * per, (rank + 1) * per


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-side timing reduction used by bench.py (max over ranks)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
