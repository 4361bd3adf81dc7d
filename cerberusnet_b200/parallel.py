"""Multi-GPU plumbing for the hot path.

The op is per-sample (every output element depends on one batch item only: blockIdx.x = n in the
reference, correlation_cuda_kernel.cu:35), so it shards over image pairs / batch items with no
exchange step: one process per GPU, ``torch.distributed`` only for the barrier and for reducing
timings (bench.py) or, in training, DDP's gradient all-reduce of the surrounding model.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``total`` independent units (image pairs, batch items)
    owned by ``rank``; the first ``total % world`` ranks take one extra unit."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks (never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, elapsed_ms_this_rank: float, device=None) -> float:
    """Whole-job throughput: units processed by all ranks / max-over-ranks time (units per second)."""
    total = sum_over_ranks(units_this_rank, device)
    ms = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (ms * 1e-3)
