"""ctypes front-end of the CPU oracle (oracle/librcw_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.  PARITY UNPINNED versus Julia — see
oracle/rcw_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librcw_oracle.so")
# RCW_ORACLE_LIB: load another build of the same source instead (tests/test_sanitizers.py: -fsanitize=address,undefined)
_LIB_OVERRIDE = os.environ.get("RCW_ORACLE_LIB")


class OrcConfig(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("N", C.c_int32), ("R", C.c_int32), ("P", C.c_int32),
        ("radius", C.c_float), ("incr", C.c_float), ("sfov", C.c_float), ("cam_h", C.c_float),
        ("goal_reward", C.c_float),
        ("palette", C.c_uint32 * 6),
        ("tie_le", C.c_int32), ("dist_post", C.c_int32),
        ("pu_per_tu", C.c_int32), ("top_palette", C.c_uint32 * 6),
        ("num_layers", C.c_int32), ("layer_kind", C.c_int32 * 4), ("layer_reward", C.c_float * 4),
        ("layer_palette", C.c_uint32 * 8), ("layer_top_color", C.c_uint32 * 4),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile) if the .so is missing or stale."""
    src = os.path.join(_HERE, "rcw_oracle.c")
    hdr = os.path.join(_HERE, "rcw_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_OVERRIDE or _LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    P = C.POINTER
    sig = {
        "orc_config_default": (None, [P(OrcConfig)]),
        "orc_create": (vp, [P(OrcConfig), vp]),
        "orc_destroy": (None, [vp]),
        "orc_directions": (None, [i32, vp]),
        "orc_set_wall_map": (None, [vp, vp]),
        "orc_set_layer": (i32, [vp, i32, vp]),
        "orc_set_state": (None, [vp, f32, f32, i32, i32, i32, f32, i32]),
        "orc_get_state": (None, [vp, vp, vp, vp, vp, vp]),
        "orc_reset_to": (None, [vp, i32, i32, i32, i32, i32]),
        "orc_is_player_colliding": (i32, [vp, i32, f32, f32]),
        "orc_act": (i32, [vp, i32]),
        "orc_cast_rays": (None, [vp]),
        "orc_update_camera_view": (None, [vp]),
        "orc_step": (i32, [vp, i32]),
        "orc_update_top_view": (None, [vp]),
        "orc_top_view": (vp, [vp]),
        "orc_cast_ray": (None, [vp, f32, f32, f32, f32, vp, vp, vp, vp]),
        "orc_ray_stop": (vp, [vp]),
        "orc_ray_dim": (vp, [vp]),
        "orc_ray_dist": (vp, [vp]),
        "orc_ray_dir": (vp, [vp]),
        "orc_camera_view": (vp, [vp]),
        "orc_wall_heights": (None, [vp, vp]),
        "orc_camera_columns": (None, [vp, vp]),
        "orc_obs_rgb8": (None, [vp, vp]),
        "orc_update_camera_view_bytes": (None, [vp, i32]),
        "orc_frame_bytes": (vp, [vp]),
        "orc_philox4x32_10": (None, [vp, vp, vp]),
        "orc_draw_layout": (None, [vp, u64, u64, C.c_uint32, vp, vp, vp]),
        "orc_draw_action": (i32, [u64, u64, u64]),
        "orc_batch_create": (vp, [P(OrcConfig), vp, i64, i64, u64, i32]),
        "orc_batch_destroy": (None, [vp]),
        "orc_batch_world": (vp, [vp, i64]),
        "orc_batch_reset": (None, [vp]),
        "orc_batch_step": (i32, [vp, vp, i32]),
        "orc_batch_rollout": (None, [vp, i32, i32, i32]),
        "orc_batch_episode_stats": (None, [vp, vp, vp, vp]),
        "orc_batch_get_reward_done": (None, [vp, vp, vp]),
        "orc_batch_step_index": (u64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def default_config(**kw) -> OrcConfig:
    cfg = OrcConfig()
    lib().orc_config_default(C.byref(cfg))
    for k, v in kw.items():
        if k in ("palette", "top_palette", "layer_kind", "layer_reward", "layer_palette", "layer_top_color"):
            for i, c in enumerate(v):
                getattr(cfg, k)[i] = float(c) if k == "layer_reward" else int(c)
        else:
            setattr(cfg, k, v)
    return cfg


def directions(n: int) -> np.ndarray:
    out = np.empty((n, 2), np.float32)
    lib().orc_directions(n, out.ctypes.data)
    return out


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.empty(4, np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def draw_action(seed: int, env_id: int, step: int) -> int:
    return int(lib().orc_draw_action(seed, env_id, step))


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class World:
    """One SingleRoomWorld + its camera view, driven like the mutable Julia struct."""

    def __init__(self, cfg: OrcConfig | None = None, dirs: np.ndarray | None = None, _ptr=None,
                 _owner=None):
        self.L = lib()
        self.cfg = cfg if cfg is not None else default_config()
        self._owner = _owner
        if _ptr is not None:
            self.p = _ptr
        else:
            d = None if dirs is None else np.ascontiguousarray(dirs, np.float32)
            self.p = self.L.orc_create(C.byref(self.cfg), None if d is None else d.ctypes.data)
            if not self.p:
                raise MemoryError("orc_create failed")

    def __del__(self):
        if getattr(self, "_owner", None) is None and getattr(self, "p", None):
            self.L.orc_destroy(self.p)
            self.p = None

    # -- state ------------------------------------------------------------------------
    def set_wall_map(self, wall_hw: np.ndarray):
        """wall_hw: bool [H, W]"""
        a = np.asfortranarray(np.asarray(wall_hw, np.uint8))
        flat = np.ascontiguousarray(a.T.reshape(-1))  # [W][H], i fastest
        self.L.orc_set_wall_map(self.p, flat.ctypes.data)

    def set_layer(self, layer: int, tiles_hw: np.ndarray):
        """tile_map[layer, :, :] = tiles (bool [H, W]); layer 1 = WALL, 3.. = the extra object layers."""
        a = np.asfortranarray(np.asarray(tiles_hw, np.uint8))
        flat = np.ascontiguousarray(a.T.reshape(-1))
        if self.L.orc_set_layer(self.p, int(layer), flat.ctypes.data) != 0:
            raise ValueError(f"no settable object layer {layer} (num_layers = {self.cfg.num_layers})")

    def set_state(self, x, y, au, gi, gj, reward=0.0, done=0):
        self.L.orc_set_state(self.p, np.float32(x), np.float32(y), int(au), int(gi), int(gj),
                             np.float32(reward), int(done))

    def reset_to(self, gi, gj, pi, pj, au):
        self.L.orc_reset_to(self.p, int(gi), int(gj), int(pi), int(pj), int(au))

    def state(self):
        xy = np.empty(2, np.float32)
        au = C.c_int32()
        g = np.empty(2, np.int32)
        r = C.c_float()
        d = C.c_int32()
        self.L.orc_get_state(self.p, xy.ctypes.data, C.addressof(au), g.ctypes.data, C.addressof(r),
                             C.addressof(d))
        return dict(pos=xy, au=au.value, goal=g, reward=r.value, done=bool(d.value))

    # -- the reference's functions ------------------------------------------------------
    def is_player_colliding(self, layer, x, y) -> bool:
        return bool(self.L.orc_is_player_colliding(self.p, layer, np.float32(x), np.float32(y)))

    def act(self, a) -> int:
        return int(self.L.orc_act(self.p, int(a)))

    def cast_rays(self):
        self.L.orc_cast_rays(self.p)

    def update_camera_view(self):
        self.L.orc_update_camera_view(self.p)

    def step(self, a) -> int:
        return int(self.L.orc_step(self.p, int(a)))

    def cast_ray(self, x, y, dx, dy):
        i, j, d = C.c_int32(), C.c_int32(), C.c_int32()
        dist = C.c_float()
        self.L.orc_cast_ray(self.p, np.float32(x), np.float32(y), np.float32(dx), np.float32(dy),
                            C.addressof(i), C.addressof(j), C.addressof(d), C.addressof(dist))
        return i.value, j.value, d.value, np.float32(dist.value)

    # -- outputs (copies) ----------------------------------------------------------------
    @property
    def ray_stop(self):
        return _view(self.L.orc_ray_stop(self.p), (self.cfg.R, 2), np.int32).copy()

    @property
    def ray_dim(self):
        return _view(self.L.orc_ray_dim(self.p), (self.cfg.R,), np.int32).copy()

    @property
    def ray_dist(self):
        return _view(self.L.orc_ray_dist(self.p), (self.cfg.R,), np.float32).copy()

    @property
    def ray_dir(self):
        return _view(self.L.orc_ray_dir(self.p), (self.cfg.R, 2), np.float32).copy()

    @property
    def camera_view(self):
        """uint32 [R columns, P rows] (the Julia Array{UInt32}(P, R), transposed view)."""
        return _view(self.L.orc_camera_view(self.p), (self.cfg.R, self.cfg.P), np.uint32).copy()

    def update_top_view(self):
        self.L.orc_update_top_view(self.p)

    @property
    def top_view(self):
        """uint32 [W*pu columns, H*pu rows] (the Julia Array{UInt32}(H*pu, W*pu), transposed view)."""
        pu = self.cfg.pu_per_tu
        return _view(self.L.orc_top_view(self.p), (self.cfg.W * pu, self.cfg.H * pu), np.uint32).copy()

    def wall_heights(self):
        out = np.empty(self.cfg.R, np.int32)
        self.L.orc_wall_heights(self.p, out.ctypes.data)
        return out

    def camera_columns(self):
        """[R] uint32, pad | palette index << 16 per image column: the camera view before it becomes pixels."""
        out = np.empty(self.cfg.R, np.uint32)
        self.L.orc_camera_columns(self.p, out.ctypes.data)
        return out

    def obs_rgb8(self):
        out = np.empty((self.cfg.R, self.cfg.P, 3), np.uint8)
        self.L.orc_obs_rgb8(self.p, out.ctypes.data)
        return out

    def frame_bytes(self, fmt: str):
        """The camera view rendered directly in the engine's byte formats ("rgb8": [R, P, 3], "gray8": [R, P])."""
        code = {"rgb8": 2, "gray8": 3}[fmt]
        self.L.orc_update_camera_view_bytes(self.p, code)
        shape = (self.cfg.R, self.cfg.P, 3) if fmt == "rgb8" else (self.cfg.R, self.cfg.P)
        return _view(self.L.orc_frame_bytes(self.p), shape, np.uint8).copy()

    def draw_layout(self, seed, env_id, episode):
        g = np.empty(2, np.int32)
        p = np.empty(2, np.int32)
        au = C.c_int32()
        self.L.orc_draw_layout(self.p, seed, env_id, episode, g.ctypes.data, p.ctypes.data,
                               C.addressof(au))
        return g, p, au.value


class Batch:
    """CPU statement of the batched engine (Philox layouts/actions, same-step auto-reset)."""

    def __init__(self, num_envs, cfg: OrcConfig | None = None, dirs=None, env_id_offset=0, seed=0,
                 auto_reset=True):
        self.L = lib()
        self.cfg = cfg if cfg is not None else default_config()
        self.num_envs = int(num_envs)
        d = None if dirs is None else np.ascontiguousarray(dirs, np.float32)
        self.p = self.L.orc_batch_create(C.byref(self.cfg), None if d is None else d.ctypes.data,
                                         self.num_envs, int(env_id_offset), int(seed),
                                         int(bool(auto_reset)))

    def __del__(self):
        if getattr(self, "p", None):
            self.L.orc_batch_destroy(self.p)
            self.p = None

    def world(self, e) -> World:
        return World(self.cfg, _ptr=self.L.orc_batch_world(self.p, int(e)), _owner=self)

    def reset(self):
        self.L.orc_batch_reset(self.p)

    def step(self, actions=None, threads=1) -> int:
        if actions is None:
            return int(self.L.orc_batch_step(self.p, None, threads))
        a = np.ascontiguousarray(actions, np.uint8)
        return int(self.L.orc_batch_step(self.p, a.ctypes.data, threads))

    RENDER = {False: 0, True: 1, None: 0, "none": 0, "xrgb32": 1, "rgb8": 2, "gray8": 3}

    def rollout(self, n_steps, threads=1, render=True):
        """render: False / True (the reference's UInt32 camera view) or "xrgb32" | "rgb8" | "gray8" (bench arm)."""
        self.L.orc_batch_rollout(self.p, int(n_steps), int(threads), self.RENDER[render])

    def reward_done(self):
        r = np.empty(self.num_envs, np.float32)
        d = np.empty(self.num_envs, np.uint8)
        self.L.orc_batch_get_reward_done(self.p, r.ctypes.data, d.ctypes.data)
        return r, d

    def episode_stats(self):
        ep = C.c_int64()
        sr = C.c_double()
        sl = C.c_int64()
        self.L.orc_batch_episode_stats(self.p, C.addressof(ep), C.addressof(sr), C.addressof(sl))
        return ep.value, sr.value, sl.value

    def states(self):
        pos = np.empty((self.num_envs, 2), np.float32)
        au = np.empty(self.num_envs, np.int32)
        goal = np.empty((self.num_envs, 2), np.int32)
        for e in range(self.num_envs):
            s = self.world(e).state()
            pos[e] = s["pos"]
            au[e] = s["au"]
            goal[e] = s["goal"]
        return pos, au, goal

    def obs_rgb8(self):
        out = np.empty((self.num_envs, self.cfg.R, self.cfg.P, 3), np.uint8)
        for e in range(self.num_envs):
            out[e] = self.world(e).obs_rgb8()
        return out

    def obs_columns(self):
        out = np.empty((self.num_envs, self.cfg.R), np.uint32)
        for e in range(self.num_envs):
            out[e] = self.world(e).camera_columns()
        return out

    def obs_gray8(self):
        """BT.601 luma of the reference pixels (the engine's learner-facing format, no reference counterpart)."""
        c = self.obs_u32()
        r, g, b = (c >> 16) & 255, (c >> 8) & 255, c & 255
        return ((77 * r + 150 * g + 29 * b + 128) >> 8).astype(np.uint8)

    def obs_gray8_half(self):
        """The GRAY8 frame under a 2 x 2 box filter, (a + b + c + d + 2) >> 2 (RCW_OBS_GRAY8_HALF): [n, R / 2, P / 2]."""
        g = self.obs_gray8().astype(np.uint32)
        return ((g[:, 0::2, 0::2] + g[:, 0::2, 1::2] + g[:, 1::2, 0::2] + g[:, 1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def obs_gray16f(self):
        """The GRAY8 luma / 255 in IEEE binary16: one binary32 division, one rounding to half (RCW_OBS_GRAY16F)."""
        return (self.obs_gray8().astype(np.float32) / np.float32(255)).astype(np.float16)

    def obs_u32(self):
        out = np.empty((self.num_envs, self.cfg.R, self.cfg.P), np.uint32)
        for e in range(self.num_envs):
            out[e] = self.world(e).camera_view
        return out
