"""raycastworlds.jl_b200 — B200-native batched engine for the RayCastWorlds.jl SingleRoom hot path.

Host-side mirror of the reference API over the C ABI of lib/librcw_b200.so (include/rcw_b200.h).
Because the directory name contains a dot, import it through the repo-root shim:

    import raycastworlds_jl_b200 as rcw
"""
from . import _capi, sharding  # noqa: F401
from ._capi import InvalidActionError, RcwError  # noqa: F401
from .single_room import (  # noqa: F401
    ACTION_KEYS, ACTION_NAMES, CAMERA_VIEW, NUM_ACTIONS, NUM_VIEWS, TOP_VIEW, AbstractGame, BatchedSingleRoom,
    RLBaseEnv, SingleRoom, act, action_space, get_action_keys, get_action_names, is_terminated, play, reset, reward,
    state, state_space,
)
from .sharding import ShardedSingleRoom, max_over_ranks, reduce_episode_stats, shard_envs  # noqa: F401
