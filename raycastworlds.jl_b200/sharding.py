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


class ShardedSingleRoom:
    """One process, several GPUs: a batch cut into contiguous blocks of global env ids, one handle per device, all
    driven from this one host thread through the library's sharded entry points (rcw_create_sharded,
    rcw_step_sharded, rcw_step_random_sharded, rcw_sync_sharded, rcw_reduce_episode_stats; include/rcw_b200.h).
    `shards[k]` is the BatchedSingleRoom of block k (state access, observations, top views per shard).
    Trajectories are bit-identical to a single handle holding the whole batch and to the one-process-per-GPU
    layout: the Philox streams are keyed by global env id."""

    def __init__(self, total_envs: int, devices=None, **kw):
        import ctypes as C

        from . import _capi
        from .single_room import BatchedSingleRoom, _ptr, make_config

        n = len(devices) if devices is not None else 1
        self._lib = _capi.load()
        cfg, dirs, obs_format = make_config(total_envs, **kw)
        self._handles = (C.c_void_p * n)()
        dev = (C.c_int32 * n)(*[int(d) for d in devices]) if devices is not None else None
        _capi.check(self._lib.rcw_create_sharded(C.byref(cfg), _ptr(dirs), dev, n, self._handles))
        self.n_shards, self.total_envs, self.shards, self.offsets = n, int(total_envs), [], []
        for k in range(n):
            off, cnt = C.c_int64(), C.c_int64()
            _capi.check(self._lib.rcw_shard_envs(total_envs, n, k, C.byref(off), C.byref(cnt)))
            c = _capi.RcwConfig.from_buffer_copy(cfg)
            c.num_envs, c.env_id_offset = cnt.value, cfg.env_id_offset + off.value
            c.device = int(devices[k]) if devices is not None else k
            if c.obs_window_envs > cnt.value:
                c.obs_window_envs = 0
            self.shards.append(BatchedSingleRoom._from_handle(self._handles[k], c, obs_format))
            self.offsets.append(off.value)

    def act(self, actions):
        """rcw_step_sharded: `actions` is the host array of the whole batch in global env order."""
        import numpy as np

        from . import _capi

        a = np.ascontiguousarray(actions, np.uint8)
        if a.shape != (self.total_envs,):
            raise ValueError(f"actions must have shape ({self.total_envs},)")
        _capi.check(self._lib.rcw_step_sharded(self._handles, self.n_shards, a.ctypes.data))

    def step_random(self, n_steps: int = 1):
        from . import _capi

        _capi.check(self._lib.rcw_step_random_sharded(self._handles, self.n_shards, int(n_steps)))

    def sync(self):
        from . import _capi

        _capi.check(self._lib.rcw_sync_sharded(self._handles, self.n_shards))

    def episode_stats(self, reset_counters: bool = False):
        import ctypes as C

        from . import _capi

        ep, sr, sl = C.c_int64(), C.c_double(), C.c_int64()
        _capi.check(self._lib.rcw_reduce_episode_stats(self._handles, self.n_shards, C.byref(ep), C.byref(sr),
                                                       C.byref(sl), int(reset_counters)))
        return ep.value, sr.value, sl.value

    def reset(self):
        """reset!(env) on every shard (layouts drawn on the devices, keyed by global env id)."""
        for sh in self.shards:
            sh.reset()

    def copy_obs(self):
        """The shards' observations concatenated in global env order (blocking device -> host copies)."""
        import numpy as np

        return np.concatenate([sh.copy_obs() for sh in self.shards])

    def reward_done(self):
        import numpy as np

        parts = [sh.reward_done() for sh in self.shards]
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def get_state(self):
        """The shards' states concatenated in global env order."""
        import numpy as np

        parts = [s.get_state() for s in self.shards]
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}

    def close(self):
        for s in self.shards:
            s.close()
        self.shards = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
