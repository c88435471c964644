"""GPU parity of the two front-end shortcuts of round 2, each against the oracle AND against the general path:

* RoomMap kernels — when the wall layer is exactly the border of the map (the only map a reference SingleRoom ever
  has, single_room.jl:57-60) the DDA counts the steps left to the border instead of probing a bit-packed layer in
  shared memory (dda_walk_room), and act! / reset! test the border in registers.  RCW_ROOM=0 keeps the general
  kernels; both must give identical rays (tile, dimension, distance bit for bit), states and observations.
* ready-made columns — env_kernel copies small columns out of a table of every possible column instead of painting
  them (RCW_COL_TABLE_KB=0 keeps the painter); identical observation bytes in every format.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def check_against_oracle(env, ref, fmt, n):
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
    np.testing.assert_array_equal(st["dir_au"], au)
    np.testing.assert_array_equal(st["goal"], goal)
    r, d = ref.reward_done()
    np.testing.assert_array_equal(st["reward"], r)
    np.testing.assert_array_equal(st["done"], d)
    want = {"rgb8": ref.obs_rgb8, "xrgb32": ref.obs_u32, "gray8": ref.obs_gray8}[fmt]()
    np.testing.assert_array_equal(env.copy_obs(), want)
    rays = env.get_rays()
    for e in range(n):
        w = ref.world(e)
        np.testing.assert_array_equal(rays["hit"][e], w.ray_stop)
        np.testing.assert_array_equal(rays["dim"][e], w.ray_dim)
        np.testing.assert_array_equal(bits(rays["dist"][e]), bits(w.ray_dist))
    assert env.episode_stats() == ref.episode_stats()


GEOMETRIES = [(84, 84, "gray8"), (84, 84, "rgb8"), (20, 33, "rgb8"), (45, 51, "xrgb32"), (130, 37, "gray8"),
              (64, 64, "xrgb32"), (96, 30, "gray8"), (33, 1, "gray8")]


@pytest.mark.parametrize("room,table_kb", [(0, 0), (1, 0), (0, 64), (1, 64)])
@pytest.mark.parametrize("R,P,fmt", GEOMETRIES)
def test_env_kernel_variants_match_oracle(rcw, oracle, monkeypatch, room, table_kb, R, P, fmt):
    monkeypatch.setenv("RCW_ENV_PER_WARP", "1")
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    monkeypatch.setenv("RCW_ROOM", str(room))
    monkeypatch.setenv("RCW_COL_TABLE_KB", str(table_kb))
    n, seed, steps = 41, 7 * R + P, 150
    kw = dict(height_tile_map_tu=6, width_tile_map_tu=9, num_directions=64)
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format=fmt, num_rays=R, height_camera_view_pu=P, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=6, W=9, N=64, R=R, P=P), seed=seed)
    check_against_oracle(env, ref, fmt, n)
    env.step_random(steps)
    ref.rollout(steps)
    check_against_oracle(env, ref, fmt, n)
    rng = np.random.default_rng(seed)
    for _ in range(12):
        a = rng.choice([1, 1, 1, 2, 3, 4], size=n).astype(np.uint8)
        env.act(a)
        assert ref.step(a) == 0
    check_against_oracle(env, ref, fmt, n)
    assert ref.episode_stats()[0] > 0, "the small map should have finished some episodes (auto-reset path)"
    env.close()


@pytest.mark.parametrize("room", [0, 1])
@pytest.mark.parametrize("flags", [dict(), dict(dda_tie_le=True), dict(dda_dist_post=True), dict(dda_tie_le=True, dda_dist_post=True)])
def test_item_kernel_room_and_bits_walks_agree(rcw, oracle, monkeypatch, room, flags):
    """Default camera (item kernel), all four settings of the unpinned DDA decisions D1 / D2.  (RCW_ROOM=2 forces the
    RoomMap kernel where the handle would keep the bit-packed one because the stores bound the step.)"""
    monkeypatch.setenv("RCW_ROOM", str(2 * room))
    n, seed, steps = 9, 77, 60
    env = rcw.BatchedSingleRoom(n, seed=seed, **flags)
    cfg = oracle.default_config(tie_le=flags.get("dda_tie_le", False), dist_post=flags.get("dda_dist_post", False))
    ref = oracle.Batch(n, cfg=cfg, seed=seed)
    env.step_random(steps)
    ref.rollout(steps)
    check_against_oracle(env, ref, "rgb8", n)
    env.close()


def test_axis_aligned_and_corner_rays(rcw, oracle, monkeypatch):
    """Tile-centre and tile-corner starts with axis-aligned / diagonal directions: ties between the two side distances,
    zero direction components (delta = Inf), both walks, both tie rules."""
    results = {}
    for room in (0, 1):
        for tie in (False, True):
            monkeypatch.setenv("RCW_ROOM", str(room))
            n = 8
            env = rcw.BatchedSingleRoom(n, num_directions=8, num_rays=33, height_camera_view_pu=40, dda_tie_le=tie,
                                        height_tile_map_tu=7, width_tile_map_tu=7, auto_reset=False)
            pos = np.array([[3.5, 3.5]] * 4 + [[3.0, 3.0], [2.0, 4.0], [3.5, 3.0], [1.0, 1.0]], np.float32)
            au = np.array([0, 1, 2, 3, 1, 5, 7, 1], np.int32)
            goal = np.array([[5, 5], [2, 2], [4, 6], [6, 2], [5, 5], [4, 2], [2, 6], [3, 3]], np.int32)
            env.set_state(pos=pos, dir_au=au, goal=goal)
            env.render()
            rays = env.get_rays()
            obs = env.copy_obs()
            cfg = oracle.default_config(H=7, W=7, N=8, R=33, P=40, tie_le=tie)
            for e in range(n):
                w = oracle.World(cfg)
                w.set_state(pos[e, 0], pos[e, 1], au[e], goal[e, 0], goal[e, 1])
                w.cast_rays()
                w.update_camera_view()
                np.testing.assert_array_equal(rays["hit"][e], w.ray_stop, err_msg=f"room={room} tie={tie} env {e}")
                np.testing.assert_array_equal(rays["dim"][e], w.ray_dim)
                np.testing.assert_array_equal(bits(rays["dist"][e]), bits(w.ray_dist))
                np.testing.assert_array_equal(obs[e], w.obs_rgb8())
            results[(room, tie)] = (rays["hit"].copy(), rays["dim"].copy(), bits(rays["dist"]).copy(), obs.copy())
            env.close()
    for tie in (False, True):
        for a, b in zip(results[(0, tie)], results[(1, tie)]):
            np.testing.assert_array_equal(a, b)


def test_interior_wall_turns_the_room_kernels_off(rcw, oracle):
    """rcw_set_wall_map with an interior wall must fall back to the bit-packed layer (and back again for a plain room)."""
    n, H, W = 5, 8, 16
    env = rcw.BatchedSingleRoom(n, seed=3, num_rays=96, height_camera_view_pu=64)
    walls = np.zeros((H, W), bool)
    walls[0, :] = walls[-1, :] = walls[:, 0] = walls[:, -1] = True
    pillar = walls.copy()
    pillar[3, 5] = pillar[4, 9] = True
    for wall_map in (pillar, walls):
        env.set_wall_map(wall_map)
        env.reset()
        cfg = oracle.default_config(R=96, P=64)
        st = env.get_state()
        env.step_random(40)
        st2 = env.get_state()
        rays = env.get_rays()
        for e in range(n):
            w = oracle.World(cfg)
            w.set_wall_map(wall_map)
            w.set_state(st2["pos"][e, 0], st2["pos"][e, 1], st2["dir_au"][e], st2["goal"][e, 0], st2["goal"][e, 1])
            w.cast_rays()
            np.testing.assert_array_equal(rays["hit"][e], w.ray_stop)
            np.testing.assert_array_equal(bits(rays["dist"][e]), bits(w.ray_dist))
    env.close()
