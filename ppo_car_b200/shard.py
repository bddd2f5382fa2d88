"""Multi-GPU sharding of the environment axis.

Environments are independent, so GPU g of G owns the contiguous env range
[g*N/G, (g+1)*N/G) with its own copy of the track tables, its own state arrays and its own
slice of the action stream; the step and GAE kernels need NO collective.  Only rollout
statistics (train.py:272-274, avg_reward) cross the shards: one small all-reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first (n_total % world_size) ranks own one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_rollout_stats(reward_sum: torch.Tensor, n_steps: torch.Tensor, episodes: torch.Tensor | None = None):
    """Sum (reward_sum, n_steps[, episodes]) over all ranks; returns python floats.
    avg_reward of train.py:272 = reward_sum / n_steps after the reduction."""
    parts = [reward_sum.reshape(1).double(), n_steps.reshape(1).double()]
    if episodes is not None:
        parts.append(episodes.reshape(1).double())
    buf = torch.cat(parts)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return tuple(float(x) for x in buf.tolist())
