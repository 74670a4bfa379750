"""Multi-GPU plumbing of the forward path: scenes are independent, so ranks are replicas; the only
exchange is the reduction of per-rank timings (max over ranks) that bench.py reports."""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as dist


def rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def scene_seed(base: int, rank: int) -> int:
    """Every rank renders / encodes its own scenes (weak scaling): distinct, reproducible seeds."""
    return base + 7919 * rank


def max_over_ranks(values: Sequence[float], device) -> list[float]:
    """Element-wise maximum of `values` over all ranks (no-op without a process group)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_throughput(units_per_rank: float, world: int, max_seconds: float) -> float:
    """Whole-job throughput: units all ranks processed / slowest rank's time."""
    return units_per_rank * world / max_seconds
