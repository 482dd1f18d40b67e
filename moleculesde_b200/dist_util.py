"""Multi-GPU plumbing of the sampling path (one process per GPU, `torch.distributed`).

Sampling groups are independent (no message crosses a molecule, and the Langevin step size is a
per-group mean), so the path shards by group with NO data-path collective: every rank samples its
own groups.  The only communication is the timing / bookkeeping reduction below.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_groups(num_groups: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) range of sampling groups owned by `rank` (strong scaling)."""
    base, rem = divmod(num_groups, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(values: Sequence[float], device) -> List[float]:
    """Element-wise max of per-rank scalars (device times): the job is as slow as its slowest rank."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def sum_over_ranks(values: Sequence[float], device) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]
