"""Host-side mirror of the reference's game API for the SingleRoom hot path.

The reference's host language is Julia (absent from this image), so the host side above the C ABI
is written in Python and keeps the reference's names and meaning:

    reference (src/RayCastWorlds.jl:5-14, src/single_room.jl)      here
    ---------------------------------------------------------      -----------------------------
    SingleRoom(; kw...)                      :258-324              SingleRoom(**kw)
    RCW.reset!(env)                          :326-331              reset(env) / env.reset()
    RCW.act!(env, action)                    :333-340              act(env, action) / env.act(a)
    RCW.cast_rays!, RCW.update_camera_view!  :195-231, 374-444     env.render()  (both, fused)
    RCW.get_action_names(env)                :486                  get_action_names(env)
    env.world.reward / .done / ...           :21-40                env.world.reward / .done / ...
    env.camera_view                          :300                  env.camera_view  (uint32 [P, R])
    RLBaseEnv(env), RLBase.state/reward/...  rlbase.jl:1-7, :574-584   RLBaseEnv(env), state(...)...
    (new) BatchedSingleRoom(; num_envs, ...)                       BatchedSingleRoom(num_envs, ...)

Everything that computes runs in librcw_b200.so on the GPU; this file only moves pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi

NUM_ACTIONS = 4  # single_room.jl:19
ACTION_NAMES = ("MOVE_FORWARD", "MOVE_BACKWARD", "TURN_LEFT", "TURN_RIGHT")  # single_room.jl:486
ACTION_KEYS = ("W", "S", "A", "D")  # single_room.jl:485 (MFB.KB_KEY_W, KB_KEY_S, KB_KEY_A, KB_KEY_D)
NUM_VIEWS, CAMERA_VIEW, TOP_VIEW = 2, 1, 2  # single_room.jl:237-239


_FORMATS = (("rgb8", _capi.RCW_OBS_RGB8), ("xrgb32", _capi.RCW_OBS_XRGB32), ("gray8", _capi.RCW_OBS_GRAY8),
            ("columns", _capi.RCW_OBS_COLUMNS), ("gray16f", _capi.RCW_OBS_GRAY16F), ("gray8_half", _capi.RCW_OBS_GRAY8_HALF))


class AbstractGame:
    """src/RayCastWorlds.jl:5"""


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def make_config(num_envs: int = 1, *, device: int = 0, height_tile_map_tu: int = 8,
                width_tile_map_tu: int = 16, num_directions: int = 128,
                player_radius_wu: float = 1 / 8, position_increment_wu: float = 1 / 8,
                semi_field_of_view_wu: float = 2 / 3, num_rays: int = 512,
                camera_height_tile_wu: float = 1.0, height_camera_view_pu: int = 256,
                goal_reward: float = 1.0, obs_format: str = "rgb8", auto_reset: bool = True,
                seed: int = 0, env_id_offset: int = 0,
                directions_wu: Optional[np.ndarray] = None, palette: Optional[Sequence[int]] = None,
                dda_tie_le: bool = False, dda_dist_post: bool = False, obs_window_envs: int = 0,
                top_view: bool = False, pu_per_tu: int = 32, top_palette: Optional[Sequence[int]] = None,
                frame_stack: int = 1, result_ring: int = 0, num_object_layers: int = 2,
                layer_kind: Sequence[int] = (), layer_reward: Sequence[float] = (),
                layer_palette: Sequence[Sequence[int]] = (), layer_top_color: Sequence[int] = ()):
    """rcw_config from the keyword arguments of SingleRoom(...) (single_room.jl:258-272) plus the batch fields.
    Returns (cfg, directions or None, obs_format name)."""
    cfg = _capi.default_config()
    cfg.device = int(device)
    cfg.num_envs = int(num_envs)
    cfg.env_id_offset = int(env_id_offset)
    cfg.height_tile_map_tu = int(height_tile_map_tu)
    cfg.width_tile_map_tu = int(width_tile_map_tu)
    cfg.num_directions = int(num_directions)
    cfg.num_rays = int(num_rays)
    cfg.height_camera_view_pu = int(height_camera_view_pu)
    cfg.player_radius_wu = float(np.float32(player_radius_wu))
    cfg.position_increment_wu = float(np.float32(position_increment_wu))
    cfg.semi_field_of_view_wu = float(np.float32(semi_field_of_view_wu))
    cfg.camera_height_tile_wu = float(np.float32(camera_height_tile_wu))
    cfg.goal_reward = float(np.float32(goal_reward))
    fmt = dict(_FORMATS)
    if obs_format not in fmt:
        raise ValueError(f"obs_format must be one of {sorted(fmt)}")
    cfg.obs_format = fmt[obs_format]
    cfg.auto_reset = int(bool(auto_reset))
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    if palette is not None:
        for i, c in enumerate(palette):
            cfg.palette[i] = int(c)
    cfg.dda_flags = (_capi.RCW_DDA_TIE_LE if dda_tie_le else 0) | (
        _capi.RCW_DDA_DIST_POST if dda_dist_post else 0)
    cfg.obs_window_envs = int(obs_window_envs)
    cfg.frame_stack = int(frame_stack)
    cfg.result_ring = int(result_ring)
    cfg.top_view = int(bool(top_view))
    cfg.pu_per_tu = int(pu_per_tu)
    if top_palette is not None:
        for i, c in enumerate(top_palette):
            cfg.top_palette[i] = int(c)
    # object layers beyond WALL and GOAL (NUM_OBJECTS > 2, single_room.jl:16-18): per extra layer its kind ("blocking" /
    # "terminal" or RCW_LAYER_*), reward (terminal layers), camera colours (dim 1, dim 2) and top-view colour
    cfg.num_object_layers = int(num_object_layers)
    kinds = {"blocking": _capi.RCW_LAYER_BLOCKING, "terminal": _capi.RCW_LAYER_TERMINAL}
    for k, v in enumerate(layer_kind):
        cfg.layer_kind[k] = kinds[v] if isinstance(v, str) else int(v)
    for k, v in enumerate(layer_reward):
        cfg.layer_reward[k] = float(np.float32(v))
    for k, pair in enumerate(layer_palette):
        cfg.layer_palette[k][0], cfg.layer_palette[k][1] = int(pair[0]), int(pair[1])
    for k, v in enumerate(layer_top_color):
        cfg.layer_top_color[k] = int(v)
    dirs = None
    if directions_wu is not None:
        dirs = np.ascontiguousarray(directions_wu, np.float32)
        if dirs.shape != (cfg.num_directions, 2):
            raise ValueError("directions_wu must have shape [num_directions, 2]")
    return cfg, dirs, obs_format


class BatchedSingleRoom(AbstractGame):
    """`num_envs` independent SingleRoom games advanced by one kernel launch per step.

    Keyword arguments are those of the reference constructor (single_room.jl:258-272); the
    additional ones are `num_envs`, `device`, `obs_format` ("rgb8" | "xrgb32" | "gray8" | "gray16f" | "columns"), `auto_reset`,
    `seed`, `env_id_offset` (global id of env 0 when a batch is sharded over GPUs),
    `directions_wu` (the host's own [N, 2] float32 direction table), the two switches for the
    unpinned RayCaster.cast_ray decisions (`dda_tie_le`, `dda_dist_post`) and `obs_window_envs`
    (the observation buffer holds only that many env slots, env e in slot e mod K — for batches whose
    observations exceed HBM; see `act_range`), `top_view` (redraw the top view inside every step /
    reset / render like the reference's act!(env), single_room.jl:333-340; off by default for a batch),
    `pu_per_tu` (:269) and `top_palette`, and `frame_stack` (K > 1: the observation buffer keeps the K most
    recent frames of every env in a ring that every step advances; see `obs_frames()`), and `result_ring`
    (D >= 1: `act_async` / `wait` — the step kernel writes rewards and terminations straight into a pinned host
    ring, and with D >= 2 the host can enqueue step k + 1 before it reads the results of step k).
    """

    def __init__(self, num_envs: int = 1, **kw):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        cfg, dirs, obs_format = make_config(num_envs, **kw)
        _capi.check(self._lib.rcw_create(C.byref(cfg), _ptr(dirs), C.byref(self._h)))
        self._finish_init(cfg, obs_format)

    @classmethod
    def _from_handle(cls, handle, cfg, obs_format):
        """Wrap a handle that rcw_create_sharded made (the wrapper owns it from here on)."""
        self = cls.__new__(cls)
        self._lib = _capi.load()
        self._h = C.c_void_p(handle)
        self._finish_init(cfg, obs_format)
        return self

    def _finish_init(self, cfg, obs_format):
        self.cfg = cfg
        self.num_envs = int(cfg.num_envs)
        self.obs_window = int(cfg.obs_window_envs) if 0 < int(cfg.obs_window_envs) < self.num_envs else self.num_envs
        self.frame_stack = max(1, int(cfg.frame_stack))
        self.result_ring = int(cfg.result_ring)
        self._result_views = {}
        self._ticket = C.c_int64()
        self._wait_r, self._wait_d = C.c_void_p(), C.c_void_p()
        self.obs_format = obs_format
        self.bytes_per_pixel = {"rgb8": 3, "xrgb32": 4, "gray8": 1, "columns": 4, "gray16f": 2, "gray8_half": 1}[obs_format]
        # columns x rows of one observation: the camera's, or half of each under the 2 x 2 box filter ("gray8_half")
        half = obs_format == "gray8_half"
        self.obs_cols = int(cfg.num_rays) // 2 if half else int(cfg.num_rays)
        self.obs_rows = int(cfg.height_camera_view_pu) // 2 if half else int(cfg.height_camera_view_pu)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._result_views = {}   # views of the pinned result ring die with the handle
            self._lib.rcw_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the reference's generic functions -----------------------------------------------------
    def reset(self, goal_ij=None, player_ij=None, dir_au=None, mask=None):
        """reset!(env): NULL layouts => drawn on the device."""
        g = None if goal_ij is None else np.ascontiguousarray(goal_ij, np.int32).reshape(self.num_envs, 2)
        p = None if player_ij is None else np.ascontiguousarray(player_ij, np.int32).reshape(self.num_envs, 2)
        d = None if dir_au is None else np.ascontiguousarray(dir_au, np.int32).reshape(self.num_envs)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).reshape(self.num_envs)
        _capi.check(self._lib.rcw_reset(self._h, _ptr(g), _ptr(p), _ptr(d), _ptr(m)))

    def act(self, actions):
        """act!(env, actions): one action in 1..4 per env.  Accepts a host array (numpy / list) or
        a CUDA uint8 tensor (anything with __cuda_array_interface__)."""
        cai = getattr(actions, "__cuda_array_interface__", None)
        if cai is not None:
            if cai["typestr"] not in ("|u1", "<u1") or int(np.prod(cai["shape"])) != self.num_envs:
                raise ValueError("device actions must be uint8 [num_envs]")
            _capi.check(self._lib.rcw_step(self._h, C.c_void_p(cai["data"][0])))
            return
        a = np.asarray(actions)
        if a.shape != (self.num_envs,):
            a = a.reshape(self.num_envs)
        if a.dtype != np.uint8:
            if ((a < 1) | (a > NUM_ACTIONS)).any():
                bad = a[(a < 1) | (a > NUM_ACTIONS)][0]
                raise _capi.InvalidActionError(_capi.RCW_EACTION, f"Invalid action: {bad}")
            a = a.astype(np.uint8)
        a = np.ascontiguousarray(a)
        _capi.check(self._lib.rcw_step(self._h, _ptr(a)))

    def act_async(self, actions) -> int:
        """`act` whose rewards / terminations also land in the pinned result ring (rcw_step_async; needs
        `result_ring >= 1`).  Returns the step's ticket for `wait`.  `actions`: uint8 host array [num_envs]
        (or anything `act` accepts)."""
        cai = getattr(actions, "__cuda_array_interface__", None)
        if cai is not None:
            if cai["typestr"] not in ("|u1", "<u1") or int(np.prod(cai["shape"])) != self.num_envs:
                raise ValueError("device actions must be uint8 [num_envs]")
            ptr = C.c_void_p(cai["data"][0])
        else:
            a = actions if isinstance(actions, np.ndarray) else np.asarray(actions)
            if a.dtype != np.uint8:
                if ((a < 1) | (a > NUM_ACTIONS)).any():
                    bad = a[(a < 1) | (a > NUM_ACTIONS)].reshape(-1)[0]
                    raise _capi.InvalidActionError(_capi.RCW_EACTION, f"Invalid action: {bad}")
                a = a.astype(np.uint8)
            if a.size != self.num_envs:
                raise ValueError("actions must hold one value per env")
            a = np.ascontiguousarray(a)
            ptr = a.ctypes.data
        rc = self._lib.rcw_step_async(self._h, ptr, C.byref(self._ticket))
        if rc:
            _capi.check(rc)
        return self._ticket.value

    def wait(self, ticket: int):
        """Blocks until the step of `ticket` has finished; returns (reward f32 [num_envs], done u8 [num_envs]) as
        read-only views of the pinned result ring, valid until `result_ring` further `act_async` calls."""
        rc = self._lib.rcw_wait(self._h, int(ticket), C.byref(self._wait_r), C.byref(self._wait_d))
        if rc:
            _capi.check(rc)
        key = self._wait_r.value
        views = self._result_views.get(key)
        if views is None:
            n = self.num_envs
            r = np.ctypeslib.as_array(C.cast(self._wait_r, C.POINTER(C.c_float)), shape=(n,))
            d = np.ctypeslib.as_array(C.cast(self._wait_d, C.POINTER(C.c_uint8)), shape=(n,))
            r.flags.writeable = False
            d.flags.writeable = False
            views = self._result_views[key] = (r, d)
        return views

    def act_range(self, actions, env0: int, n: Optional[int] = None):
        """act! for the envs [env0, env0 + n) only (rcw_step_range): `actions` holds n values in 1..4,
        host array or CUDA uint8 tensor.  With an observation window the range's observations are in
        slots (env0 + k) mod obs_window afterwards."""
        cai = getattr(actions, "__cuda_array_interface__", None)
        if cai is not None:
            n = int(np.prod(cai["shape"])) if n is None else int(n)
            if cai["typestr"] not in ("|u1", "<u1") or int(np.prod(cai["shape"])) != n:
                raise ValueError("device actions must be uint8 [n]")
            _capi.check(self._lib.rcw_step_range(self._h, C.c_void_p(cai["data"][0]), int(env0), n))
            return
        a = np.asarray(actions).reshape(-1)
        n = a.shape[0] if n is None else int(n)
        if a.shape[0] != n:
            raise ValueError("actions must hold one value per env of the range")
        if a.dtype != np.uint8:
            if ((a < 1) | (a > NUM_ACTIONS)).any():
                bad = a[(a < 1) | (a > NUM_ACTIONS)][0]
                raise _capi.InvalidActionError(_capi.RCW_EACTION, f"Invalid action: {bad}")
            a = a.astype(np.uint8)
        a = np.ascontiguousarray(a)
        _capi.check(self._lib.rcw_step_range(self._h, _ptr(a), int(env0), n))

    def act_tape(self, actions):
        """rcw_step_tape: len(actions) steps driven by an action tape, uint8 [n_steps, num_envs] (numpy, or a CUDA torch
        tensor).  A multi-step call: consecutive steps may overlap on the device (two half-batches on two streams)."""
        if hasattr(actions, "data_ptr"):
            if tuple(actions.shape[1:]) != (self.num_envs,) or not actions.is_contiguous() or actions.element_size() != 1:
                raise ValueError("a device tape must be a contiguous uint8 [n_steps, num_envs] tensor")
            _capi.check(self._lib.rcw_step_tape(self._h, C.c_void_p(actions.data_ptr()), int(actions.shape[0])))
            return
        a = np.ascontiguousarray(actions, np.uint8)
        if a.ndim != 2 or a.shape[1] != self.num_envs:
            raise ValueError(f"actions must have shape (n_steps, {self.num_envs})")
        _capi.check(self._lib.rcw_step_tape(self._h, _ptr(a), a.shape[0]))

    def step_random(self, n_steps: int = 1):
        _capi.check(self._lib.rcw_step_random(self._h, int(n_steps)))

    def render(self):
        """cast_rays! + update_camera_view! from the current state."""
        _capi.check(self._lib.rcw_render(self._h))

    def get_action_names(self):
        return ACTION_NAMES

    # -- state ---------------------------------------------------------------------------------
    def get_state(self):
        n = self.num_envs
        out = dict(pos=np.empty((n, 2), np.float32), dir_au=np.empty(n, np.int32),
                   goal=np.empty((n, 2), np.int32), reward=np.empty(n, np.float32),
                   done=np.empty(n, np.uint8))
        _capi.check(self._lib.rcw_get_state(self._h, _ptr(out["pos"]), _ptr(out["dir_au"]),
                                            _ptr(out["goal"]), _ptr(out["reward"]), _ptr(out["done"])))
        return out

    def reward_done(self, reward: Optional[np.ndarray] = None, done: Optional[np.ndarray] = None):
        """reward / is_terminated of every env (blocking device->host read)."""
        n = self.num_envs
        r = np.empty(n, np.float32) if reward is None else reward
        d = np.empty(n, np.uint8) if done is None else done
        _capi.check(self._lib.rcw_get_state(self._h, None, None, None, _ptr(r), _ptr(d)))
        return r, d

    def set_state(self, pos=None, dir_au=None, goal=None, reward=None, done=None):
        n = self.num_envs
        p = None if pos is None else np.ascontiguousarray(pos, np.float32).reshape(n, 2)
        a = None if dir_au is None else np.ascontiguousarray(dir_au, np.int32).reshape(n)
        g = None if goal is None else np.ascontiguousarray(goal, np.int32).reshape(n, 2)
        r = None if reward is None else np.ascontiguousarray(reward, np.float32).reshape(n)
        d = None if done is None else np.ascontiguousarray(done, np.uint8).reshape(n)
        _capi.check(self._lib.rcw_set_state(self._h, _ptr(p), _ptr(a), _ptr(g), _ptr(r), _ptr(d)))

    def save_checkpoint(self) -> np.ndarray:
        """Exact snapshot of the dynamic state (rcw_save_checkpoint) as a uint8 array."""
        n = C.c_size_t()
        _capi.check(self._lib.rcw_checkpoint_size(self._h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        _capi.check(self._lib.rcw_save_checkpoint(self._h, _ptr(buf), n.value))
        return buf

    def load_checkpoint(self, buf):
        """Restore a snapshot made by save_checkpoint() on a batch with the same configuration; re-renders."""
        b = np.ascontiguousarray(np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray)) else buf, np.uint8)
        _capi.check(self._lib.rcw_load_checkpoint(self._h, _ptr(b), b.size))

    def set_wall_map(self, wall_hw):
        """wall_hw: bool [H, W] — replaces tile_map[WALL, :, :] for every env of the batch."""
        w = np.asarray(wall_hw).astype(np.uint8)
        if w.shape != (self.cfg.height_tile_map_tu, self.cfg.width_tile_map_tu):
            raise ValueError("wall map must be [height_tile_map_tu, width_tile_map_tu]")
        flat = np.ascontiguousarray(w.T.reshape(-1))  # [W][H], i fastest (Julia column-major)
        _capi.check(self._lib.rcw_set_wall_map(self._h, _ptr(flat)))

    def set_layer(self, layer: int, tiles_hw):
        """tile_map[layer, :, :] = tiles_hw (bool [H, W]) for every env: layer 1 = WALL, 3 .. num_object_layers = the
        extra object layers (single_room.jl:16-18 with NUM_OBJECTS > 2)."""
        t = np.asarray(tiles_hw).astype(np.uint8)
        if t.shape != (self.cfg.height_tile_map_tu, self.cfg.width_tile_map_tu):
            raise ValueError("tiles must be [height_tile_map_tu, width_tile_map_tu]")
        flat = np.ascontiguousarray(t.T.reshape(-1))
        _capi.check(self._lib.rcw_set_layer(self._h, int(layer), _ptr(flat)))

    def set_wall_maps(self, walls_ehw):
        """walls_ehw: bool [num_envs, H, W] — one wall layer per env (tile_map[WALL, :, :] of env e)."""
        w = np.asarray(walls_ehw).astype(np.uint8)
        if w.shape != (self.num_envs, self.cfg.height_tile_map_tu, self.cfg.width_tile_map_tu):
            raise ValueError("wall maps must be [num_envs, height_tile_map_tu, width_tile_map_tu]")
        flat = np.ascontiguousarray(w.transpose(0, 2, 1).reshape(-1))  # [E][W][H], i fastest
        _capi.check(self._lib.rcw_set_wall_maps(self._h, _ptr(flat)))

    def get_rays(self, env0: int = 0, n: Optional[int] = None):
        n = self.num_envs - env0 if n is None else n
        R = self.cfg.num_rays
        out = dict(hit=np.empty((n, R, 2), np.int32), dim=np.empty((n, R), np.int32),
                   dist=np.empty((n, R), np.float32), ray_dir=np.empty((n, R, 2), np.float32))
        _capi.check(self._lib.rcw_get_rays(self._h, env0, n, _ptr(out["hit"]), _ptr(out["dim"]),
                                           _ptr(out["dist"]), _ptr(out["ray_dir"])))
        return out

    # -- observations -----------------------------------------------------------------------------
    @property
    def obs_shape(self):
        """[env slots, width (= num_rays columns), height_camera_view_pu, (3)]; the pixel row is the
        fastest index, as in the reference's Array{UInt32}(P, R).  env slots = num_envs unless an
        observation window was configured."""
        R, P = self.obs_cols, self.obs_rows
        if self.obs_format == "columns":
            return (self.obs_window, R)      # one uint32 word per column: pad | palette index << 16
        return (self.obs_window, R, P, 3) if self.obs_format == "rgb8" else (self.obs_window, R, P)

    def obs_device_ptr(self):
        ptr, total, stride = C.c_void_p(), C.c_size_t(), C.c_size_t()
        _capi.check(self._lib.rcw_obs_device_ptr(self._h, C.byref(ptr), C.byref(total), C.byref(stride)))
        return ptr.value, total.value, stride.value

    def obs_layout(self):
        """(env_stride_bytes, column_stride_bytes, column_bytes, bytes_per_pixel) of the device buffer."""
        es, cs, cbytes, bpp = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int32()
        _capi.check(self._lib.rcw_obs_layout(self._h, C.byref(es), C.byref(cs), C.byref(cbytes), C.byref(bpp)))
        return es.value, cs.value, cbytes.value, bpp.value

    def obs_frames(self):
        """(frame_stack, newest ring position, frame_stride_bytes) — rcw_obs_frames."""
        k, newest, stride = C.c_int32(), C.c_int32(), C.c_size_t()
        _capi.check(self._lib.rcw_obs_frames(self._h, C.byref(k), C.byref(newest), C.byref(stride)))
        return k.value, newest.value, stride.value

    def obs_tensor(self):
        """Zero-copy torch view of the device observation buffer, shape obs_shape (borrowed: valid
        until the next act/reset/render, like the reference's aliased `state`, single_room.jl:576).
        Columns whose byte length is not a multiple of 32 are pitched (rcw_obs_layout): the view is
        then strided; `.contiguous()` or copy_obs() give a dense copy."""
        import torch

        ptr, total, _ = self.obs_device_ptr()
        env_stride, col_stride, _, _ = self.obs_layout()
        R, P = self.obs_cols, self.obs_rows
        holder = _CudaBuffer(ptr, total, self)
        flat = torch.as_tensor(holder, device=torch.device("cuda", self.cfg.device))
        slots = self.obs_window
        if self.obs_format == "columns":
            words = flat.view(torch.int32)
            if self.frame_stack > 1:
                k, _, fs = self.obs_frames()
                return torch.as_strided(words, (slots, k, R), (env_stride // 4, fs // 4, 1))
            return torch.as_strided(words, (slots, R), (env_stride // 4, 1))
        if self.frame_stack > 1:
            # the whole ring: [env, ring position, ...]; obs_frames()[1] is the newest position
            k, _, fs = self.obs_frames()
            if self.obs_format == "rgb8":
                return torch.as_strided(flat, (slots, k, R, P, 3), (env_stride, fs, col_stride, 3, 1))
            if self.obs_format in ("gray8", "gray8_half"):
                return torch.as_strided(flat, (slots, k, R, P), (env_stride, fs, col_stride, 1))
            if self.obs_format == "gray16f":
                return torch.as_strided(flat.view(torch.float16), (slots, k, R, P), (env_stride // 2, fs // 2, col_stride // 2, 1))
            return torch.as_strided(flat.view(torch.int32), (slots, k, R, P), (env_stride // 4, fs // 4, col_stride // 4, 1))
        if self.obs_format == "rgb8":
            return torch.as_strided(flat, (slots, R, P, 3), (env_stride, col_stride, 3, 1))
        if self.obs_format in ("gray8", "gray8_half"):
            return torch.as_strided(flat, (slots, R, P), (env_stride, col_stride, 1))
        if self.obs_format == "gray16f":
            return torch.as_strided(flat.view(torch.float16), (slots, R, P), (env_stride // 2, col_stride // 2, 1))
        words = flat.view(torch.int32)
        return torch.as_strided(words, (slots, R, P), (env_stride // 4, col_stride // 4, 1))

    def obs_tensor_nchw(self):
        """The same buffer as a [N, C, num_rays, height_px] torch view for convolutional learners: the
        image is presented transposed (camera columns along torch's "height"), which is exactly torch's
        channels_last memory format whenever the columns are not pitched — no copy, no permute kernel."""
        if self.obs_format == "columns":
            raise ValueError("the columns format holds no pixels; expand_columns() rasterises it")
        t = self.obs_tensor()
        return t.permute(0, 3, 1, 2) if self.obs_format == "rgb8" else t.unsqueeze(1)

    def expand_columns(self, columns=None, pixel_format: str = "rgb8"):
        """Rasterise camera views kept as column words (obs_format="columns") into pixels on the device
        (rcw_expand_columns).  `columns`: CUDA int32 / uint32 tensor [n, num_rays] (rows may be strided) — e.g. a
        minibatch gathered from a replay buffer; None: this handle's own newest observations.  Returns a torch
        view [n, num_rays, height_px(, 3)] (uint8; int32 for "xrgb32") of a freshly allocated device buffer, on
        the handle's stream."""
        import torch

        if pixel_format not in ("rgb8", "xrgb32", "gray8", "gray16f"):
            raise ValueError("pixel_format must be rgb8, xrgb32, gray8 or gray16f")
        dev = torch.device("cuda", self.cfg.device)
        R, P = self.cfg.num_rays, self.cfg.height_camera_view_pu
        if columns is None:
            if self.obs_format != "columns":
                raise ValueError("this handle's observations are pixels already")
            ptr, _, env_stride = self.obs_device_ptr()
            k, newest, fs = self.obs_frames()
            src, src_stride, n = ptr + newest * fs, env_stride, self.obs_window
        else:
            if columns.dim() != 2 or columns.shape[1] != R or columns.stride(1) != 1 or columns.element_size() != 4:
                raise ValueError("columns must be a 32-bit [n, num_rays] CUDA tensor with contiguous rows")
            src, src_stride, n = columns.data_ptr(), columns.stride(0) * 4, columns.shape[0]
        es, cs, cb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _capi.check(self._lib.rcw_expanded_layout(self._h, dict(_FORMATS)[pixel_format], C.byref(es), C.byref(cs), C.byref(cb)))
        es, cs = es.value, cs.value
        stream = torch.cuda.ExternalStream(self.cuda_stream(), device=dev)
        with torch.cuda.stream(stream):
            buf = torch.empty(n * es, dtype=torch.uint8, device=dev)
        _capi.check(self._lib.rcw_expand_columns(self._h, C.c_void_p(src), src_stride, n, dict(_FORMATS)[pixel_format],
                                                 C.c_void_p(buf.data_ptr())))
        if pixel_format == "rgb8":
            return torch.as_strided(buf, (n, R, P, 3), (es, cs, 3, 1))
        if pixel_format == "gray8":
            return torch.as_strided(buf, (n, R, P), (es, cs, 1))
        if pixel_format == "gray16f":
            return torch.as_strided(buf.view(torch.float16), (n, R, P), (es // 2, cs // 2, 1))
        return torch.as_strided(buf.view(torch.int32), (n, R, P), (es // 4, cs // 4, 1))

    def copy_obs(self, env0: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None, age: int = 0):
        """Blocking device->host copy of the observations of envs [env0, env0+n); with a frame ring, `age`
        selects the frame (0 = newest)."""
        n = min(self.num_envs - env0, self.obs_window) if n is None else n
        R, P = self.obs_cols, self.obs_rows
        shape = (n, R, P, 3) if self.obs_format == "rgb8" else ((n, R) if self.obs_format == "columns" else (n, R, P))
        dtype = np.uint32 if self.obs_format in ("xrgb32", "columns") else (np.float16 if self.obs_format == "gray16f" else np.uint8)
        if out is None:
            out = np.empty(shape, dtype)
        _capi.check(self._lib.rcw_copy_obs_frame(self._h, env0, n, int(age), _ptr(out)))
        return out

    # -- top view (single_room.jl:342-372, 446-483) -----------------------------------------------------
    @property
    def top_view_shape(self):
        """[env slots, width_tu * pu (columns), height_tu * pu (rows)] uint32; row fastest, as in the
        reference's Array{UInt32}(H * pu, W * pu) (:302)."""
        pu = self.cfg.pu_per_tu
        return (self.obs_window, self.cfg.width_tile_map_tu * pu, self.cfg.height_tile_map_tu * pu)

    def render_top_view(self):
        """update_top_view!(env) from the current state of every env."""
        _capi.check(self._lib.rcw_render_top_view(self._h))

    def copy_top_view(self, env0: int = 0, n: Optional[int] = None):
        n = min(self.num_envs - env0, self.obs_window) if n is None else n
        out = np.empty((n,) + self.top_view_shape[1:], np.uint32)
        _capi.check(self._lib.rcw_copy_top_view(self._h, env0, n, _ptr(out)))
        return out

    def top_view_tensor(self):
        """Zero-copy torch view (int32) of the device top views, shape top_view_shape; borrowed like obs_tensor()."""
        import torch

        ptr, total, stride = C.c_void_p(), C.c_size_t(), C.c_size_t()
        _capi.check(self._lib.rcw_top_view_device_ptr(self._h, C.byref(ptr), C.byref(total), C.byref(stride)))
        if not ptr.value:
            raise RuntimeError("no top view has been drawn yet")
        flat = torch.as_tensor(_CudaBuffer(ptr.value, total.value, self), device=torch.device("cuda", self.cfg.device))
        slots, wp, hp = self.top_view_shape
        return torch.as_strided(flat.view(torch.int32), (slots, wp, hp), (stride.value // 4, hp, 1))

    # -- bookkeeping ------------------------------------------------------------------------------
    def episode_stats(self, reset_counters: bool = False):
        ep, sr, sl = C.c_int64(), C.c_double(), C.c_int64()
        _capi.check(self._lib.rcw_episode_stats(self._h, C.byref(ep), C.byref(sr), C.byref(sl),
                                                int(reset_counters)))
        return ep.value, sr.value, sl.value

    def launch_count(self) -> int:
        n = C.c_int64()
        _capi.check(self._lib.rcw_launch_count(self._h, C.byref(n)))
        return n.value

    def cuda_stream(self) -> int:
        s = C.c_void_p()
        _capi.check(self._lib.rcw_stream(self._h, C.byref(s)))
        return s.value or 0

    def sync(self):
        _capi.check(self._lib.rcw_sync(self._h))


class _CudaBuffer:
    """Minimal __cuda_array_interface__ holder for a borrowed device pointer."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }


class _WorldView:
    """Field access in the style of `env.world.<field>` (single_room.jl:21-40) for one env."""

    def __init__(self, batch: BatchedSingleRoom, index: int = 0):
        self._b, self._i = batch, index

    def _s(self):
        return self._b.get_state()

    @property
    def reward(self):
        return np.float32(self._s()["reward"][self._i])

    @property
    def done(self):
        return bool(self._s()["done"][self._i])

    @property
    def goal_reward(self):
        return np.float32(self._b.cfg.goal_reward)

    @property
    def player_position_wu(self):
        return self._s()["pos"][self._i].copy()

    @property
    def player_direction_au(self):
        return int(self._s()["dir_au"][self._i])

    @property
    def goal_position(self):
        g = self._s()["goal"][self._i]
        return int(g[0]), int(g[1])

    @property
    def num_directions(self):
        return self._b.cfg.num_directions

    def _rays(self):
        return self._b.get_rays(self._i, 1)

    @property
    def ray_stop_position_tu(self):
        return self._rays()["hit"][0].T.copy()  # [2, R] like the reference

    @property
    def ray_hit_dimension(self):
        return self._rays()["dim"][0]

    @property
    def ray_distance_wu(self):
        return self._rays()["dist"][0]

    @property
    def ray_directions_wu(self):
        return self._rays()["ray_dir"][0]


class SingleRoom(BatchedSingleRoom):
    """The reference's single game (single_room.jl:241-324): one env, no auto-reset, UInt32 pixels, the
    top view redrawn by every act! / reset! (:329,337)."""

    def __init__(self, **kw):
        kw.setdefault("auto_reset", False)
        kw.setdefault("obs_format", "xrgb32")
        kw.setdefault("top_view", True)
        super().__init__(1, **kw)
        self.world = _WorldView(self, 0)

    def act(self, action):  # act!(env::SingleRoom, action)
        if isinstance(action, (int, np.integer)):
            if not 1 <= int(action) <= NUM_ACTIONS:
                raise _capi.InvalidActionError(_capi.RCW_EACTION, f"Invalid action: {action}")
            action = np.array([action], np.uint8)
        super().act(action)

    @property
    def camera_view(self):
        """uint32 [height_camera_view_pu, num_rays], the reference's camera_view (single_room.jl:300)."""
        if self.obs_format != "xrgb32":
            raise ValueError("camera_view needs obs_format='xrgb32'")
        return self.copy_obs(0, 1)[0].T

    @property
    def top_view(self):
        """uint32 [height_tu * pu, width_tu * pu], the reference's top_view (single_room.jl:302)."""
        if not self.cfg.top_view:
            self.render_top_view()
        return self.copy_top_view(0, 1)[0].T


class RLBaseEnv:
    """rlbase.jl:1-3 — wraps a game; RLBase methods below follow single_room.jl:574-584."""

    def __init__(self, env: BatchedSingleRoom):
        self.env = env

    def __call__(self, action):  # (env::RLBaseEnv)(action) = act!(env.env, action)   :581
        self.env.act(action)


# generic functions, named as in the reference without the `!`
def reset(env):
    (env.env if isinstance(env, RLBaseEnv) else env).reset()


def act(env, action):
    (env.env if isinstance(env, RLBaseEnv) else env).act(action)


def get_action_names(env):
    return (env.env if isinstance(env, RLBaseEnv) else env).get_action_names()


def get_action_keys(env):
    """get_action_keys(env) (single_room.jl:485): the keys play! binds to the four actions."""
    return ACTION_KEYS


def play(game: "SingleRoom", keys, on_frame=None):
    """play!(game) (single_room.jl:488-572) without the MiniFB window: the same key handling, driven by an iterable
    of key names instead of keyboard events.  W / S / A / D act, R resets (and zeroes the step
    count), V switches between camera view and top view (clearing the frame buffer, :533-534), Q closes; any
    other key is reported and ignored (:540).  After every key the current view is copied into the frame buffer
    exactly as copy_image_to_frame_buffer! does — transposed, frame_buffer[j, i] = image[i, j] (utils.jl:64-73) —
    and `on_frame(frame_buffer, info)` is called, info = what the reference prints with @show (:549-551).
    Returns (frame_buffer, info) as they stand when the keys run out or Q is pressed.
    frame_buffer: uint32 [max(view widths), max(view heights)] like the reference's zeros(UInt32, width_image,
    height_image) (:503-506), in Fortran order like the Julia array, so its memory is the row-major width x height
    pixel buffer MiniFB is handed (mfb_update, :559)."""
    if not isinstance(game, SingleRoom):
        raise TypeError("play drives one SingleRoom, like the reference")
    camera, top = game.camera_view, game.top_view
    frame_buffer = np.zeros((max(top.shape[1], camera.shape[1]), max(top.shape[0], camera.shape[0])), np.uint32, order="F")
    current_view, steps_taken = CAMERA_VIEW, 0

    def blit():
        image = game.camera_view if current_view == CAMERA_VIEW else game.top_view
        frame_buffer[:image.shape[1], :image.shape[0]] = image.T      # frame_buffer[j, i] = image[i, j]

    blit()                                                        # :512-517
    info = dict(key=None, steps_taken=0, reward=game.world.reward, done=game.world.done, view=current_view, warning=None)
    for key in keys:
        key, warning = str(key).upper(), None
        if key == "Q":                                            # :525-527
            break
        elif key == "R":                                          # :528-530
            game.reset()
            steps_taken = 0
        elif key == "V":                                          # :531-534
            current_view = current_view % NUM_VIEWS + 1           # mod1(current_view + 1, NUM_VIEWS)
            frame_buffer[:] = 0
        elif key in ACTION_KEYS:                                  # :535-538
            game.act(ACTION_KEYS.index(key) + 1)
            steps_taken += 1
        else:
            warning = f"No keybinding exists for {key}"           # :540
        blit()                                                    # :543-547
        info = dict(key=key, steps_taken=steps_taken, reward=game.world.reward, done=game.world.done,
                    view=current_view, warning=warning)
        if on_frame is not None:
            on_frame(frame_buffer, info)
    return frame_buffer, info


def state(env: RLBaseEnv):
    """RLBase.state(env) = env.env.camera_view (:576).  Batched: the device tensor of all envs."""
    e = env.env
    return e.camera_view if isinstance(e, SingleRoom) else e.obs_tensor()


def state_space(env: RLBaseEnv):
    return None  # :575


def action_space(env: RLBaseEnv):
    return range(1, NUM_ACTIONS + 1)  # Base.OneTo(NUM_ACTIONS)  :580


def reward(env: RLBaseEnv):
    r, _ = env.env.reward_done()
    return np.float32(r[0]) if isinstance(env.env, SingleRoom) else r  # :583


def is_terminated(env: RLBaseEnv):
    _, d = env.env.reward_done()
    return bool(d[0]) if isinstance(env.env, SingleRoom) else d.astype(bool)  # :584
