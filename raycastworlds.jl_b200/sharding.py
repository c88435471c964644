"""Multi-GPU host logic: independent env shards, one process per GPU, no collective on the step
path.  torch.distributed is plumbing only: a barrier for timing and an optional
all-reduce of three episode-statistics scalars (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_envs(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of global env ids: returns (env_id_offset, num_envs) of `rank`.
    The first `total_envs % world_size` ranks own one extra env.  Global env ids key the Philox
    streams, so trajectories do not depend on world_size."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    if total_envs < 0:
        raise ValueError("total_envs must be non-negative")
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def reduce_episode_stats(stats, group=None, device=None):
    """Sum (episodes, sum_return, sum_length) over all ranks.  `stats` is the tuple returned by
    BatchedSingleRoom.episode_stats().  Works with any initialised torch.distributed backend."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return stats
    t = torch.tensor([float(stats[0]), float(stats[1]), float(stats[2])], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    ep, sr, sl = t.tolist()
    return int(round(ep)), sr, int(round(sl))


def max_over_ranks(value: float, group=None, device=None) -> float:
    """Max of a per-rank scalar (used for device-timed step durations)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
