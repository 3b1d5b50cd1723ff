"""Multi-GPU plumbing of the sampling path: trajectories are independent, so ranks
only need (a) a disjoint shard of the trajectory list and (b) a max-over-ranks reduction of the
device time. No data-path collective exists on this path. Backend-agnostic (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `total` trajectories owned by `rank` (sizes differ by <= 1)
    — the same split the reference's --chunk_idx/--n_chunks gives (tps_inference.py:160-161)."""
    if not (0 <= rank < world):
        raise ValueError("bad rank")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_seed(base_seed: int, rank: int) -> int:
    """Per-rank seed for weak-scaling runs (each rank samples its own trajectories)."""
    return base_seed + 1000003 * rank


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over all ranks (timing is reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(local_count: int, device=None) -> List[int]:
    """How many trajectories every rank processed (rank 0 sums them for the whole-job throughput)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(local_count)]
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]
