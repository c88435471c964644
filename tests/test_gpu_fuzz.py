"""Seeded random configurations, CUDA path versus the oracle: geometry, camera, formats, palettes, wall layers,
DDA switches, auto-reset, both step kernels (item kernel / env kernel), the top view.  Every comparison is
bit-exact.  The seeds are fixed, so a failure names a reproducible case."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# RCW_FUZZ_BASE=k shifts every seed by 1000 k: fresh cases for an extended run (the default set stays reproducible)
BASE = 1000 * int(os.environ.get("RCW_FUZZ_BASE", "0"))


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def draw_case(seed):
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(3, 40)), int(rng.integers(3, 70))
    if rng.random() < 0.3:
        H, W = int(rng.integers(3, 9)), int(rng.integers(3, 9))
    N = int(rng.choice([4, 7, 16, 36, 128, 256, 600]))
    R = int(rng.choice([1, 2, 31, 32, 33, 45, 64, 84, 100, 128, 160, 257]))
    P = int(rng.choice([1, 2, 7, 16, 21, 32, 51, 64, 84, 96, 120, 200]))
    fmt = str(rng.choice(["rgb8", "xrgb32", "gray8"]))
    radius = float(np.float32(rng.uniform(0.05, 0.45)))
    incr = float(np.float32(rng.uniform(0.03, 0.6)))
    sfov = float(np.float32(rng.uniform(0.2, 1.5)))
    cam_h = float(np.float32(rng.uniform(0.3, 2.0)))
    pal = None
    if rng.random() < 0.4:
        pal = [int(v) for v in rng.integers(0, 1 << 24, 6)]
    return dict(H=H, W=W, N=N, R=R, P=P, fmt=fmt, radius=radius, incr=incr, sfov=sfov, cam_h=cam_h, pal=pal,
                tie_le=bool(rng.integers(0, 2)), dist_post=bool(rng.integers(0, 2)),
                maps=str(rng.choice(["default", "shared", "per_env"])), open_border=bool(rng.random() < 0.3),
                env_kernel=bool(rng.integers(0, 2)), n=int(rng.integers(1, 41)),
                pu=int(rng.choice([1, 2, 3, 4, 8] if seed < 100000 else [1, 3, 8, 8, 16, 24])),
                seed=seed)


@pytest.mark.parametrize("seed", range(72))
def test_random_configuration_matches_oracle(rcw, oracle, monkeypatch, seed):
    c = draw_case(BASE + 1000 + seed)
    rng = np.random.default_rng(BASE + 5000 + seed)
    monkeypatch.setenv("RCW_ENV_PER_WARP", "1" if c["env_kernel"] else "0")
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    monkeypatch.setenv("RCW_PACKED_ACTIONS", str(seed % 2))      # host actions: kernel parameters / staged copy
    monkeypatch.setenv("RCW_ROOM", str((seed >> 1) & 1))         # border-only maps: counting walk / bit-packed wall layer
    monkeypatch.setenv("RCW_COL_TABLE_KB", "0" if (seed >> 2) & 1 else "64")   # env kernel: painted / ready-made columns
    n, H, W = c["n"], c["H"], c["W"]
    kw = dict(height_tile_map_tu=H, width_tile_map_tu=W, num_directions=c["N"], num_rays=c["R"],
              height_camera_view_pu=c["P"], player_radius_wu=c["radius"], position_increment_wu=c["incr"],
              semi_field_of_view_wu=c["sfov"], camera_height_tile_wu=c["cam_h"], obs_format=c["fmt"],
              dda_tie_le=c["tie_le"], dda_dist_post=c["dist_post"], pu_per_tu=c["pu"])
    okw = dict(H=H, W=W, N=c["N"], R=c["R"], P=c["P"], radius=np.float32(c["radius"]), incr=np.float32(c["incr"]),
               sfov=np.float32(c["sfov"]), cam_h=np.float32(c["cam_h"]), tie_le=int(c["tie_le"]),
               dist_post=int(c["dist_post"]), pu_per_tu=c["pu"])
    if c["pal"] is not None:
        kw["palette"] = c["pal"]
        okw["palette"] = c["pal"]
    cfg = oracle.default_config(**okw)
    want_obs = {"rgb8": "obs_rgb8", "xrgb32": "camera_view", "gray8": None}[c["fmt"]]

    def gray(u32):
        r, g, b = (u32 >> 16) & 255, (u32 >> 8) & 255, u32 & 255
        return ((77 * r + 150 * g + 29 * b + 128) >> 8).astype(np.uint8)

    def obs_of(w):
        return gray(w.camera_view) if want_obs is None else (w.obs_rgb8() if want_obs == "obs_rgb8" else w.camera_view)

    if c["maps"] == "default":
        # Philox layouts, random policy, auto-reset: the batched semantics end to end
        env = rcw.BatchedSingleRoom(n, seed=c["seed"], env_id_offset=77, **kw)
        ref = oracle.Batch(n, cfg=cfg, seed=c["seed"], env_id_offset=77)
        steps = int(rng.integers(1, 200))
        env.step_random(steps)
        ref.rollout(steps)
        worlds = [ref.world(e) for e in range(n)]
        r, d = ref.reward_done()
        st = env.get_state()
        np.testing.assert_array_equal(st["reward"], r)
        np.testing.assert_array_equal(st["done"], d)
        assert env.episode_stats() == ref.episode_stats()
    else:
        # host-supplied wall layers (one for the batch, or one per env), injected states, explicit actions
        walls = np.zeros((n, H, W), bool)
        walls[:, 0, :] = walls[:, -1, :] = walls[:, :, 0] = walls[:, :, -1] = True
        if H > 4 and W > 4:
            walls[:, 2:-2, 2:-2] |= rng.random((n, H - 4, W - 4)) < 0.15
        if c["open_border"]:
            walls[:, 0, 1:-1] &= rng.random((n, W - 2)) < 0.5
            walls[:, 1:-1, -1] &= rng.random((n, H - 2)) < 0.5
        if c["maps"] == "shared":
            walls[:] = walls[0]
        pos = np.stack([rng.uniform(0.05, H - 0.05, n), rng.uniform(0.05, W - 0.05, n)], 1).astype(np.float32)
        au = rng.integers(0, c["N"], n).astype(np.int32)
        goal = np.stack([rng.integers(1, H + 1, n), rng.integers(1, W + 1, n)], 1).astype(np.int32)
        env = rcw.BatchedSingleRoom(n, seed=c["seed"], auto_reset=False, **kw)
        if c["maps"] == "shared":
            env.set_wall_map(walls[0])
        else:
            env.set_wall_maps(walls)
        env.set_state(pos=pos, dir_au=au, goal=goal)
        env.render()
        worlds = []
        for e in range(n):
            w = oracle.World(cfg)
            w.set_wall_map(walls[e])
            w.set_state(pos[e, 0], pos[e, 1], au[e], goal[e, 0], goal[e, 1])
            w.cast_rays()
            w.update_camera_view()
            worlds.append(w)
        for _ in range(int(rng.integers(0, 25))):
            a = rng.integers(1, 5, n).astype(np.uint8)
            env.act(a)
            for e in range(n):
                assert worlds[e].step(int(a[e])) == 0
        st = env.get_state()
        np.testing.assert_array_equal(st["reward"], np.array([w.state()["reward"] for w in worlds], np.float32))
        np.testing.assert_array_equal(st["done"], np.array([w.state()["done"] for w in worlds], np.uint8))

    np.testing.assert_array_equal(bits(st["pos"]), bits(np.stack([w.state()["pos"] for w in worlds])), err_msg=str(c))
    np.testing.assert_array_equal(st["dir_au"], np.array([w.state()["au"] for w in worlds], np.int32))
    np.testing.assert_array_equal(st["goal"], np.stack([w.state()["goal"] for w in worlds]))
    obs = env.copy_obs()
    rays = env.get_rays()
    for e in range(n):
        np.testing.assert_array_equal(obs[e], obs_of(worlds[e]), err_msg=f"env {e} of {c}")
        np.testing.assert_array_equal(rays["hit"][e], worlds[e].ray_stop, err_msg=f"env {e} of {c}")
        np.testing.assert_array_equal(rays["dim"][e], worlds[e].ray_dim)
        np.testing.assert_array_equal(bits(rays["dist"][e]), bits(worlds[e].ray_dist))
    if H * W * c["pu"] ** 2 > 600_000:       # beyond the top view renderer's shared memory (RCW_ESIZE): not this test
        env.close()
        return
    env.render_top_view()
    top = env.copy_top_view()
    for e in range(n):
        worlds[e].update_top_view()
        np.testing.assert_array_equal(top[e], worlds[e].top_view, err_msg=f"top view, env {e} of {c}")
    env.close()


@pytest.mark.parametrize("room", [0, 1])
@pytest.mark.parametrize("env_kernel", [0, 1])
def test_player_carried_outside_the_map_is_defined(rcw, oracle, monkeypatch, env_kernel, room):
    """An increment larger than a tile carries the player over the border wall (the collision test only looks at
    the candidate position, collision_detection.jl:21-42), and a state injected onto a border tile can step
    outwards.  Outside the map everything counts as wall: rays stop at once, the column is full height.  The
    unchecked closed-border DDA must not be taken from there (found by the seeded fuzz test above)."""
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    monkeypatch.setenv("RCW_ROOM", str(room))
    n = 6
    kw = dict(num_rays=64, height_camera_view_pu=32, position_increment_wu=2.5, player_radius_wu=0.1, pu_per_tu=4)
    env = rcw.BatchedSingleRoom(n, auto_reset=False, **kw)
    cfg = oracle.default_config(R=64, P=32, incr=np.float32(2.5), radius=np.float32(0.1), pu_per_tu=4)
    pos = np.array([[1.5, 1.5], [1.5, 1.5], [6.5, 14.5], [0.5, 7.5], [4.5, 0.2], [7.9, 15.9]], np.float32)
    au = np.array([64, 96, 0, 64, 96, 16], np.int32)          # facing -i, -j, +i, -i, -j, diagonal
    goal = np.array([[4, 8]] * n, np.int32)
    env.set_state(pos=pos, dir_au=au, goal=goal)
    env.render()
    worlds = []
    for e in range(n):
        w = oracle.World(cfg)
        w.set_state(pos[e, 0], pos[e, 1], au[e], 4, 8)
        w.cast_rays()
        w.update_camera_view()
        worlds.append(w)
    for a in (1, 1, 3, 1, 2, 2, 4, 1):
        env.act(np.full(n, a, np.uint8))
        for w in worlds:
            assert w.step(a) == 0
    st = env.get_state()
    got = np.stack([w.state()["pos"] for w in worlds])
    np.testing.assert_array_equal(bits(st["pos"]), bits(got))
    assert (got < 0).any() or (got[:, 0] >= 8).any() or (got[:, 1] >= 16).any(), "somebody should have left the map"
    obs = env.copy_obs()
    rays = env.get_rays()
    env.render_top_view()
    top = env.copy_top_view()
    for e, w in enumerate(worlds):
        np.testing.assert_array_equal(obs[e], w.obs_rgb8())
        np.testing.assert_array_equal(rays["hit"][e], w.ray_stop)
        np.testing.assert_array_equal(bits(rays["dist"][e]), bits(w.ray_dist))
        w.update_top_view()
        np.testing.assert_array_equal(top[e], w.top_view)
    env.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_operation_sequences_with_windows(rcw, oracle, monkeypatch, seed):
    """Random sequences of whole-batch steps, range steps (unaligned, wrapping the observation window), masked
    resets to host layouts, partial set_state, and a checkpoint round trip through a fresh handle; afterwards the
    whole state and every observation slot must equal the oracle's for the env rendered into it last."""
    rng = np.random.default_rng(BASE + 9000 + seed)
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(int(rng.integers(0, 2))))
    monkeypatch.setenv("RCW_ROOM", str(seed & 1))
    monkeypatch.setenv("RCW_COL_TABLE_KB", "0" if seed & 2 else "64")
    monkeypatch.setenv("RCW_TWO_STREAMS_MIN", "8")
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    monkeypatch.setenv("RCW_PACKED_ACTIONS", str(int(rng.integers(0, 2))))
    n = int(rng.integers(2, 60))
    K = int(rng.integers(1, n + 1)) if rng.random() < 0.8 else 0
    R, P = int(rng.choice([20, 64, 84, 100, 160])), int(rng.choice([8, 21, 32, 64]))
    fmt = str(rng.choice(["rgb8", "xrgb32", "gray8"]))
    top = bool(rng.integers(0, 2))
    kw = dict(num_rays=R, height_camera_view_pu=P, obs_format=fmt, auto_reset=False, obs_window_envs=K,
              top_view=top, pu_per_tu=4)
    cfg = oracle.default_config(R=R, P=P, pu_per_tu=4)
    env = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    win = env.obs_window
    # start from host layouts so that the oracle knows the state
    def layouts():
        g = np.stack([rng.integers(2, 8, n), rng.integers(2, 16, n)], 1).astype(np.int32)
        p = np.stack([rng.integers(2, 8, n), rng.integers(2, 16, n)], 1).astype(np.int32)
        p[(p == g).all(1)] = [2, 2]
        g[(p == g).all(1)] = [3, 3]
        return g, p, rng.integers(0, 128, n).astype(np.int32)
    g, p, a = layouts()
    env.reset(g, p, a)
    worlds = []
    for e in range(n):
        w = oracle.World(cfg)
        w.reset_to(g[e, 0], g[e, 1], p[e, 0], p[e, 1], a[e])
        w.cast_rays()
        w.update_camera_view()
        worlds.append(w)
    owner = {}
    def rendered(envs):
        for e in envs:
            owner[e % win] = e
    rendered(range(n))
    for _ in range(int(rng.integers(3, 14))):
        op = rng.choice(["act", "range", "range", "reset", "state", "checkpoint", "tape"])
        if op == "tape":      # a multi-step call (two streams / the step + top-view pipeline when the batch is not windowed)
            T = int(rng.integers(2, 6))
            tape = rng.choice([1, 1, 2, 3, 4], size=(T, n)).astype(np.uint8)
            env.act_tape(tape)
            for t in range(T):
                for e in range(n):
                    assert worlds[e].step(int(tape[t, e])) == 0
            rendered(range(n))
        elif op == "act":
            acts = rng.choice([1, 1, 2, 3, 4], size=n).astype(np.uint8)
            env.act(acts)
            for e in range(n):
                assert worlds[e].step(int(acts[e])) == 0
            rendered(range(n))
        elif op == "range":
            m = int(rng.integers(1, win + 1))
            e0 = int(rng.integers(0, n - m + 1))
            acts = rng.choice([1, 1, 2, 3, 4], size=m).astype(np.uint8)
            env.act_range(acts, e0)
            for k in range(m):
                assert worlds[e0 + k].step(int(acts[k])) == 0
            rendered(range(e0, e0 + m))
        elif op == "reset":
            g, p, a = layouts()
            mask = (rng.random(n) < 0.4).astype(np.uint8)
            env.reset(g, p, a, mask)
            for e in np.nonzero(mask)[0]:
                worlds[e].reset_to(g[e, 0], g[e, 1], p[e, 0], p[e, 1], a[e])
                worlds[e].cast_rays()
                worlds[e].update_camera_view()
            rendered(np.nonzero(mask)[0].tolist())  # a masked reset re-renders the envs it reset
        elif op == "state":
            au = rng.integers(0, 128, n).astype(np.int32)
            env.set_state(dir_au=au)
            env.render()
            for e in range(n):
                s = worlds[e].state()
                worlds[e].set_state(s["pos"][0], s["pos"][1], au[e], s["goal"][0], s["goal"][1], s["reward"], int(s["done"]))
                worlds[e].cast_rays()
                worlds[e].update_camera_view()
            rendered(range(n))
        else:
            blob = env.save_checkpoint()
            env.close()
            env = rcw.BatchedSingleRoom(n, seed=seed + 1, **kw)
            env.load_checkpoint(blob)
            rendered(range(n))
    st = env.get_state()
    np.testing.assert_array_equal(bits(st["pos"]), bits(np.stack([w.state()["pos"] for w in worlds])))
    np.testing.assert_array_equal(st["dir_au"], np.array([w.state()["au"] for w in worlds], np.int32))
    np.testing.assert_array_equal(st["goal"], np.stack([w.state()["goal"] for w in worlds]))
    np.testing.assert_array_equal(st["reward"], np.array([w.state()["reward"] for w in worlds], np.float32))
    np.testing.assert_array_equal(st["done"], np.array([w.state()["done"] for w in worlds], np.uint8))

    def gray(u32):
        r, g_, b = (u32 >> 16) & 255, (u32 >> 8) & 255, u32 & 255
        return ((77 * r + 150 * g_ + 29 * b + 128) >> 8).astype(np.uint8)

    for slot, e in owner.items():
        w = worlds[e]
        want = {"rgb8": w.obs_rgb8, "xrgb32": lambda: w.camera_view, "gray8": lambda: gray(w.camera_view)}[fmt]()
        np.testing.assert_array_equal(env.copy_obs(e, 1)[0], want, err_msg=f"slot {slot} env {e}")
        if top:
            w.update_top_view()
            np.testing.assert_array_equal(env.copy_top_view(e, 1)[0], w.top_view, err_msg=f"top view slot {slot} env {e}")
    env.close()


@pytest.mark.parametrize("env_kernel", [0, 1])
def test_huge_tile_map_needs_opt_in_shared_memory(rcw, oracle, monkeypatch, env_kernel):
    """1200 x 1000 tiles: the bit-packed wall layer is 150 KB, above the 48 KB of shared memory a kernel gets
    without opting in.  Every step kernel variant (item / env kernel, actions from the parameters, the random
    policy, render) and the reset draw must handle it; rays walk hundreds of tiles."""
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n, H, W, seed = 5, 1200, 1000, 2
    kw = dict(height_tile_map_tu=H, width_tile_map_tu=W, num_directions=64, num_rays=70, height_camera_view_pu=24,
              position_increment_wu=0.45)
    env = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=H, W=W, N=64, R=70, P=24, incr=np.float32(0.45)), seed=seed)
    env.step_random(12)
    ref.rollout(12)
    rng = np.random.default_rng(1)
    for _ in range(6):
        a = rng.integers(1, 5, n).astype(np.uint8)
        env.act(a)
        assert ref.step(a) == 0
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
    np.testing.assert_array_equal(st["goal"], goal)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    rays = env.get_rays()
    assert rays["dist"].max() > 100, "some ray should cross a good part of the map"
    for e in range(n):
        np.testing.assert_array_equal(rays["hit"][e], ref.world(e).ray_stop)
        np.testing.assert_array_equal(bits(rays["dist"][e]), bits(ref.world(e).ray_dist))
    env.close()
    with pytest.raises(rcw.RcwError) as ei:                       # 3000 x 3000 tiles: 1.1 MB, does not fit an SM
        rcw.BatchedSingleRoom(1, height_tile_map_tu=3000, width_tile_map_tu=3000)
    assert ei.value.code == rcw._capi.RCW_ESIZE


@pytest.mark.parametrize("seed", range(32))
def test_random_layered_configuration_matches_oracle(rcw, oracle, monkeypatch, seed):
    """Seeded random maps with 1-4 extra object layers (NUM_OBJECTS 3-6; blocking / terminal kinds, rewards, colours
    drawn at random, objects allowed to overlap each other, walls and goals), every observation format, both step
    kernels: injected states driven by explicit actions, or Philox episodes with auto-reset."""
    rng = np.random.default_rng(BASE + 70000 + seed)
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(seed & 1))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    monkeypatch.setenv("RCW_COL_TABLE_KB", "0" if seed & 2 else "64")
    monkeypatch.setenv("RCW_TWO_STREAMS_MIN", "8")
    H, W = int(rng.integers(4, 20)), int(rng.integers(4, 40))
    N = int(rng.choice([8, 32, 128]))
    R = int(rng.choice([2, 32, 34, 64, 84, 130]))
    P = int(rng.choice([2, 16, 32, 50, 84, 96]))
    fmt = str(rng.choice(["rgb8", "xrgb32", "gray8", "gray16f", "gray8_half", "columns"]))
    n_extra = int(rng.integers(1, 5))
    n = int(rng.integers(2, 30))
    kinds = [int(v) for v in rng.integers(0, 2, n_extra)]
    rewards = [float(np.float32(v)) for v in rng.uniform(-2, 2, n_extra)]
    pal2 = [(int(a), int(b)) for a, b in rng.integers(0, 1 << 24, (n_extra, 2))]
    tops = [int(v) for v in rng.integers(0, 1 << 24, n_extra)]
    radius, incr = float(np.float32(rng.uniform(0.05, 0.4))), float(np.float32(rng.uniform(0.05, 0.5)))
    walls = np.zeros((H, W), bool)
    walls[0, :] = walls[-1, :] = walls[:, 0] = walls[:, -1] = True
    walls[1:-1, 1:-1] |= rng.random((H - 2, W - 2)) < 0.08
    if rng.random() < 0.25:                              # an open border: rays may leave the map (painted as wall)
        walls[0, 1:-1] &= rng.random(W - 2) < 0.5
    extra = rng.random((n_extra, H, W)) < 0.08
    kw = dict(height_tile_map_tu=H, width_tile_map_tu=W, num_directions=N, num_rays=R, height_camera_view_pu=P,
              player_radius_wu=radius, position_increment_wu=incr, obs_format=fmt, pu_per_tu=int(rng.choice([1, 3, 4, 8])),
              num_object_layers=2 + n_extra, layer_kind=kinds, layer_reward=rewards, layer_palette=pal2, layer_top_color=tops)
    flat_pal = [c for pair in pal2 for c in pair] + [0] * (8 - 2 * n_extra)
    cfg = oracle.default_config(H=H, W=W, N=N, R=R, P=P, radius=np.float32(radius), incr=np.float32(incr), pu_per_tu=kw["pu_per_tu"],
                                num_layers=2 + n_extra, layer_kind=kinds + [0] * (4 - n_extra),
                                layer_reward=rewards + [0.0] * (4 - n_extra), layer_palette=flat_pal,
                                layer_top_color=tops + [0] * (4 - n_extra))

    def furnish_world(w):
        w.set_layer(1, walls)
        for k in range(n_extra):
            w.set_layer(3 + k, extra[k])

    philox = bool(rng.integers(0, 2)) and (~(walls | extra.any(0)))[1:-1, 1:-1].sum() >= 2
    if philox:
        env = rcw.BatchedSingleRoom(n, seed=seed, env_id_offset=5, **kw)
        env.set_layer(1, walls)
        for k in range(n_extra):
            env.set_layer(3 + k, extra[k])
        env.reset()
        ref = oracle.Batch(n, cfg=cfg, seed=seed, env_id_offset=5)
        for e in range(n):
            furnish_world(ref.world(e))
        ref.reset()
        steps = int(rng.integers(2, 250))
        env.step_random(steps)
        ref.rollout(steps)
        worlds = [ref.world(e) for e in range(n)]
        assert env.episode_stats() == ref.episode_stats()
    else:
        pos = np.stack([rng.uniform(0.05, H - 0.05, n), rng.uniform(0.05, W - 0.05, n)], 1).astype(np.float32)
        au = rng.integers(0, N, n).astype(np.int32)
        goal = np.stack([rng.integers(1, H + 1, n), rng.integers(1, W + 1, n)], 1).astype(np.int32)
        env = rcw.BatchedSingleRoom(n, seed=seed, auto_reset=False, **kw)
        env.set_layer(1, walls)
        for k in range(n_extra):
            env.set_layer(3 + k, extra[k])
        env.set_state(pos=pos, dir_au=au, goal=goal)
        env.render()
        worlds = []
        for e in range(n):
            w = oracle.World(cfg)
            furnish_world(w)
            w.set_state(pos[e, 0], pos[e, 1], au[e], goal[e, 0], goal[e, 1])
            w.cast_rays()
            w.update_camera_view()
            worlds.append(w)
        for _ in range(int(rng.integers(0, 30))):
            a = rng.choice([1, 1, 2, 3, 4], size=n).astype(np.uint8)
            env.act(a)
            for e in range(n):
                assert worlds[e].step(int(a[e])) == 0
    st = env.get_state()
    case = f"seed {seed}: {H}x{W} N={N} R={R} P={P} {fmt} extra={n_extra} kinds={kinds} philox={philox}"
    np.testing.assert_array_equal(bits(st["pos"]), bits(np.stack([w.state()["pos"] for w in worlds])), err_msg=case)
    np.testing.assert_array_equal(st["dir_au"], np.array([w.state()["au"] for w in worlds], np.int32), err_msg=case)
    np.testing.assert_array_equal(st["goal"], np.stack([w.state()["goal"] for w in worlds]), err_msg=case)
    if philox:      # the step's reward / done survive the same-step auto-reset: the batch keeps them, not the re-drawn world
        want_r, want_d = ref.reward_done()
    else:
        want_r = np.array([w.state()["reward"] for w in worlds], np.float32)
        want_d = np.array([w.state()["done"] for w in worlds], np.uint8)
    np.testing.assert_array_equal(st["reward"], want_r, err_msg=case)
    np.testing.assert_array_equal(st["done"], want_d, err_msg=case)

    def gray(u32):
        r, g, b = (u32 >> 16) & 255, (u32 >> 8) & 255, u32 & 255
        return ((77 * r + 150 * g + 29 * b + 128) >> 8).astype(np.uint32)

    def want(w):
        if fmt == "rgb8":
            return w.obs_rgb8()
        if fmt == "xrgb32":
            return w.camera_view
        if fmt == "columns":
            return w.camera_columns()
        g = gray(w.camera_view)
        if fmt == "gray8":
            return g.astype(np.uint8)
        if fmt == "gray16f":
            return (g.astype(np.float32) / np.float32(255)).astype(np.float16)
        return ((g[0::2, 0::2] + g[0::2, 1::2] + g[1::2, 0::2] + g[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    obs = env.copy_obs()
    rays = env.get_rays()
    for e in range(n):
        np.testing.assert_array_equal(obs[e], want(worlds[e]), err_msg=f"env {e}, {case}")
        np.testing.assert_array_equal(rays["hit"][e], worlds[e].ray_stop, err_msg=f"env {e}, {case}")
        np.testing.assert_array_equal(rays["dim"][e], worlds[e].ray_dim, err_msg=case)
        np.testing.assert_array_equal(bits(rays["dist"][e]), bits(worlds[e].ray_dist), err_msg=case)
    env.render_top_view()
    top = env.copy_top_view()
    for e in range(n):
        worlds[e].update_top_view()
        np.testing.assert_array_equal(top[e], worlds[e].top_view, err_msg=f"top view, env {e}, {case}")
    env.close()
