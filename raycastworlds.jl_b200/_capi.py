"""ctypes binding of librcw_b200.so (include/rcw_b200.h).

This is the same binding a Julia host makes with `ccall` (julia/BatchedRayCastWorlds.jl); Python
is used here because the image has no Julia.  The library is the product: if it is missing or
fails to load, importing this module raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# RCW_LIB: development knob to A/B a differently compiled build of the same library
LIB_PATH = os.environ.get("RCW_LIB") or os.path.join(_PKG, "lib", "librcw_b200.so")

RCW_OK, RCW_EINVAL, RCW_EACTION, RCW_ECUDA, RCW_ENOMEM, RCW_ESIZE = 0, -1, -2, -3, -4, -5
RCW_OBS_RGB8, RCW_OBS_XRGB32, RCW_OBS_GRAY8, RCW_OBS_COLUMNS, RCW_OBS_GRAY16F, RCW_OBS_GRAY8_HALF = 0, 1, 2, 3, 4, 5
RCW_DDA_TIE_LE, RCW_DDA_DIST_POST = 1, 2
ABI_VERSION = 4
RCW_MAX_EXTRA_LAYERS = 4
RCW_LAYER_BLOCKING, RCW_LAYER_TERMINAL = 0, 1

# every symbol include/rcw_b200.h declares (tests check the .so exports exactly these)
SYMBOLS = (
    "rcw_version", "rcw_config_init", "rcw_create", "rcw_destroy", "rcw_set_wall_map", "rcw_set_layer", "rcw_set_wall_maps", "rcw_reset",
    "rcw_step", "rcw_step_async", "rcw_wait", "rcw_step_range", "rcw_step_random", "rcw_step_tape", "rcw_render",
    "rcw_render_top_view", "rcw_top_view_device_ptr", "rcw_copy_top_view", "rcw_get_state", "rcw_set_state", "rcw_get_rays",
    "rcw_checkpoint_size", "rcw_save_checkpoint", "rcw_load_checkpoint",
    "rcw_obs_device_ptr", "rcw_obs_layout", "rcw_obs_frames", "rcw_copy_obs_frame", "rcw_copy_obs", "rcw_expand_columns", "rcw_expanded_layout", "rcw_episode_stats", "rcw_launch_count", "rcw_stream",
    "rcw_sync", "rcw_last_error",
    "rcw_shard_envs", "rcw_create_sharded", "rcw_destroy_sharded", "rcw_step_sharded", "rcw_step_random_sharded",
    "rcw_sync_sharded", "rcw_reduce_episode_stats",
)


class RcwConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("num_envs", C.c_int64),
        ("env_id_offset", C.c_int64),
        ("height_tile_map_tu", C.c_int32),
        ("width_tile_map_tu", C.c_int32),
        ("num_directions", C.c_int32),
        ("num_rays", C.c_int32),
        ("height_camera_view_pu", C.c_int32),
        ("player_radius_wu", C.c_float),
        ("position_increment_wu", C.c_float),
        ("semi_field_of_view_wu", C.c_float),
        ("camera_height_tile_wu", C.c_float),
        ("goal_reward", C.c_float),
        ("obs_format", C.c_int32),
        ("auto_reset", C.c_int32),
        ("seed", C.c_uint64),
        ("palette", C.c_uint32 * 6),
        ("dda_flags", C.c_uint32),
        ("obs_window_envs", C.c_int32),
        ("top_view", C.c_int32),
        ("pu_per_tu", C.c_int32),
        ("top_palette", C.c_uint32 * 6),
        ("frame_stack", C.c_int32),
        ("result_ring", C.c_int32),
        ("num_object_layers", C.c_int32),
        ("layer_kind", C.c_int32 * 4),
        ("layer_reward", C.c_float * 4),
        ("layer_palette", (C.c_uint32 * 2) * 4),
        ("layer_top_color", C.c_uint32 * 4),
        ("reserved", C.c_uint32 * 3),
    ]


class RcwError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"librcw_b200 error {code}: {message}")
        self.code = code


class InvalidActionError(RcwError, AssertionError):
    """RCW_EACTION — the reference's `@assert action in Base.OneTo(NUM_ACTIONS)` (single_room.jl:140)."""


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  raycastworlds.jl_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    P = C.POINTER
    sig = {
        "rcw_version": (i32, []),
        "rcw_config_init": (i32, [P(RcwConfig)]),
        "rcw_create": (i32, [P(RcwConfig), vp, P(vp)]),
        "rcw_destroy": (i32, [vp]),
        "rcw_set_wall_map": (i32, [vp, vp]),
        "rcw_set_wall_maps": (i32, [vp, vp]),
        "rcw_set_layer": (i32, [vp, i32, vp]),
        "rcw_reset": (i32, [vp, vp, vp, vp, vp]),
        "rcw_step": (i32, [vp, vp]),
        "rcw_step_async": (i32, [vp, vp, P(i64)]),
        "rcw_wait": (i32, [vp, i64, P(vp), P(vp)]),
        "rcw_step_range": (i32, [vp, vp, i64, i64]),
        "rcw_step_random": (i32, [vp, i32]),
        "rcw_step_tape": (i32, [vp, vp, i32]),
        "rcw_render": (i32, [vp]),
        "rcw_render_top_view": (i32, [vp]),
        "rcw_top_view_device_ptr": (i32, [vp, P(vp), P(C.c_size_t), P(C.c_size_t)]),
        "rcw_copy_top_view": (i32, [vp, i64, i64, vp]),
        "rcw_get_state": (i32, [vp, vp, vp, vp, vp, vp]),
        "rcw_set_state": (i32, [vp, vp, vp, vp, vp, vp]),
        "rcw_checkpoint_size": (i32, [vp, P(C.c_size_t)]),
        "rcw_save_checkpoint": (i32, [vp, vp, C.c_size_t]),
        "rcw_load_checkpoint": (i32, [vp, vp, C.c_size_t]),
        "rcw_get_rays": (i32, [vp, i64, i64, vp, vp, vp, vp]),
        "rcw_obs_device_ptr": (i32, [vp, P(vp), P(C.c_size_t), P(C.c_size_t)]),
        "rcw_obs_layout": (i32, [vp, P(C.c_size_t), P(C.c_size_t), P(C.c_size_t), P(i32)]),
        "rcw_copy_obs": (i32, [vp, i64, i64, vp]),
        "rcw_expand_columns": (i32, [vp, vp, C.c_size_t, i64, i32, vp]),
        "rcw_expanded_layout": (i32, [vp, i32, P(C.c_size_t), P(C.c_size_t), P(C.c_size_t)]),
        "rcw_obs_frames": (i32, [vp, P(i32), P(i32), P(C.c_size_t)]),
        "rcw_copy_obs_frame": (i32, [vp, i64, i64, i32, vp]),
        "rcw_episode_stats": (i32, [vp, P(i64), P(C.c_double), P(i64), i32]),
        "rcw_launch_count": (i32, [vp, P(i64)]),
        "rcw_stream": (i32, [vp, P(vp)]),
        "rcw_sync": (i32, [vp]),
        "rcw_last_error": (C.c_char_p, []),
        "rcw_shard_envs": (i32, [i64, i32, i32, P(i64), P(i64)]),
        "rcw_create_sharded": (i32, [P(RcwConfig), vp, P(i32), i32, P(vp)]),
        "rcw_destroy_sharded": (i32, [P(vp), i32]),
        "rcw_step_sharded": (i32, [P(vp), i32, vp]),
        "rcw_step_random_sharded": (i32, [P(vp), i32, i32]),
        "rcw_sync_sharded": (i32, [P(vp), i32]),
        "rcw_reduce_episode_stats": (i32, [P(vp), i32, P(i64), P(C.c_double), P(i64), i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.rcw_version() != ABI_VERSION:
        raise ImportError(f"librcw_b200 ABI {L.rcw_version()} != binding ABI {ABI_VERSION}")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc == RCW_OK:
        return
    msg = load().rcw_last_error().decode("utf-8", "replace")
    if rc == RCW_EACTION:
        raise InvalidActionError(rc, msg)
    raise RcwError(rc, msg)


def default_config() -> RcwConfig:
    cfg = RcwConfig()
    check(load().rcw_config_init(C.byref(cfg)))
    return cfg
