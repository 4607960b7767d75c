"""Multi-GPU plumbing: envs are independent, so ranks own contiguous env ranges with NO per-step
collective; the only exchange is one all-reduce (sum) of an 8-element float64 episode-statistics
vector per logging interval (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """GPU g of N owns envs [g*E/N, (g+1)*E/N)."""
    return total_envs * rank // world, total_envs * (rank + 1) // world


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank seed; per-env Philox streams are further keyed by the env id on the device."""
    return seed + rank


def reduce_episode_stats(stats: torch.Tensor) -> Dict[str, float]:
    """stats = [sum return, sum length, episodes, successes, falls, timeouts, physics errors, -]."""
    s = stats.clone()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(s, op=torch.distributed.ReduceOp.SUM)
    v = s.cpu().tolist()
    n = max(v[2], 1.0)
    return dict(episodes=v[2], mean_return=v[0] / n, mean_length=v[1] / n, successes=v[3], falls=v[4],
                timeouts=v[5], physics_errors=v[6])
