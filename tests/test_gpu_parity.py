"""GPU parity tests: the CUDA path, called through the C ABI, versus the CPU oracle and the golden
fixtures on the same inputs.  Bars (BASELINE.json north_star): hit tiles, hit sides, positions,
directions, rewards, terminations bit-exact; ray distances within 1e-5 relative (we assert
bit-equality, which is stronger); wall-column heights within 1 px (we assert identical images).

Neither the oracle nor the fixtures are outputs of the Julia reference: parity is unpinned.
"""
import zlib

import numpy as np
import pytest

from conftest import GOLDEN_CONFIGS

pytestmark = pytest.mark.gpu

RCW_KW = {
    "A": dict(),
    "B": dict(height_tile_map_tu=64, width_tile_map_tu=64, num_directions=256, num_rays=128,
              height_camera_view_pu=96, pu_per_tu=4),
    "C": dict(height_tile_map_tu=5, width_tile_map_tu=7, num_directions=36, num_rays=45,
              height_camera_view_pu=51, player_radius_wu=np.float32(0.2),
              position_increment_wu=np.float32(0.3), semi_field_of_view_wu=np.float32(0.5),
              camera_height_tile_wu=np.float32(0.8), pu_per_tu=7),
    "D": dict(dda_tie_le=True, dda_dist_post=True),
    "T": dict(height_tile_map_tu=7, width_tile_map_tu=7, num_directions=8, num_rays=33, height_camera_view_pu=40, pu_per_tu=4),
    "U": dict(height_tile_map_tu=7, width_tile_map_tu=7, num_directions=8, num_rays=33, height_camera_view_pu=40, pu_per_tu=4,
              dda_tie_le=True),
}


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def rgb8_of(img_u32):
    return np.stack([(img_u32 >> 16) & 255, (img_u32 >> 8) & 255, img_u32 & 255], -1).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# golden fixtures
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["A", "B", "C", "D", "T", "U"])
@pytest.mark.parametrize("fmt", ["rgb8", "xrgb32"])
def test_cast_and_render_match_golden(rcw, oracle, golden, case, fmt):
    states, au, goal = golden[f"{case}_states"], golden[f"{case}_au"], golden[f"{case}_goal"]
    n = len(states)
    env = rcw.BatchedSingleRoom(n, obs_format=fmt, auto_reset=False, **RCW_KW[case])
    env.set_state(pos=states, dir_au=au, goal=goal)
    env.render()
    rays = env.get_rays()
    np.testing.assert_array_equal(rays["hit"], golden[f"{case}_hit"])
    np.testing.assert_array_equal(rays["dim"], golden[f"{case}_dim"])
    np.testing.assert_array_equal(bits(rays["dist"]), bits(golden[f"{case}_dist"]))
    np.testing.assert_array_equal(bits(rays["ray_dir"]), bits(golden[f"{case}_ray_dir"]))
    obs = env.copy_obs()
    # the fixture stores a CRC of every reference-format image and a few full images; the oracle
    # (already pinned to the same fixture on CPU) supplies the full image for every state
    w = oracle.World(oracle.default_config(**GOLDEN_CONFIGS[case]))
    for k in range(n):
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_camera_view()
        img = w.camera_view
        assert zlib.crc32(img.tobytes()) == int(golden[f"{case}_crc"][k])
        if fmt == "xrgb32":
            np.testing.assert_array_equal(obs[k], img)
            assert zlib.crc32(np.ascontiguousarray(obs[k]).tobytes()) == int(golden[f"{case}_crc"][k])
        else:
            np.testing.assert_array_equal(obs[k], rgb8_of(img))
    env.close()


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_act_trajectories_match_golden(rcw, golden, case):
    init, actions = golden[f"{case}_act_init"], golden[f"{case}_act_actions"]
    n, T = actions.shape
    env = rcw.BatchedSingleRoom(n, auto_reset=False, **RCW_KW[case])
    env.reset(goal_ij=init[:, 0:2], player_ij=init[:, 2:4], dir_au=init[:, 4])
    st = env.get_state()
    assert not st["done"].any() and not st["reward"].any()
    for t in range(T):
        env.act(actions[:, t])
        st = env.get_state()
        np.testing.assert_array_equal(bits(st["pos"]), bits(golden[f"{case}_act_pos"][:, t]))
        np.testing.assert_array_equal(st["dir_au"], golden[f"{case}_act_au"][:, t])
        np.testing.assert_array_equal(st["reward"], golden[f"{case}_act_reward"][:, t])
        np.testing.assert_array_equal(st["done"], golden[f"{case}_act_done"][:, t])
    env.close()


# ------------------------------------------------------------------------------------------------
# oracle, same seeded inputs
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("fmt", ["rgb8", "xrgb32"])
def test_random_rollout_matches_oracle(rcw, oracle, fmt):
    """Philox layouts + random policy + same-step auto-reset, 96 envs x 400 steps, compared every
    50 steps: state, reward/done, rays and the full observation."""
    n, seed, offset = 96, 0x5EED, 1000
    env = rcw.BatchedSingleRoom(n, seed=seed, env_id_offset=offset, obs_format=fmt)
    ref = oracle.Batch(n, seed=seed, env_id_offset=offset)
    for chunk in range(8):
        env.step_random(50)
        ref.rollout(50, threads=4)
        st = env.get_state()
        pos, au, goal = ref.states()
        np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
        np.testing.assert_array_equal(st["dir_au"], au)
        np.testing.assert_array_equal(st["goal"], goal)
        r, d = ref.reward_done()
        np.testing.assert_array_equal(st["reward"], r)
        np.testing.assert_array_equal(st["done"], d)
        obs = env.copy_obs()
        np.testing.assert_array_equal(obs, ref.obs_rgb8() if fmt == "rgb8" else ref.obs_u32())
    rays = env.get_rays()
    for e in range(n):
        w = ref.world(e)
        np.testing.assert_array_equal(rays["hit"][e], w.ray_stop)
        np.testing.assert_array_equal(rays["dim"][e], w.ray_dim)
        np.testing.assert_array_equal(bits(rays["dist"][e]), bits(w.ray_dist))
    assert env.episode_stats() == ref.episode_stats()
    env.close()


def test_goal_seeking_episodes_match_oracle(rcw, oracle):
    """Forward-heavy host actions so that many episodes terminate and auto-reset inside the run."""
    n, seed, T = 64, 7, 600
    env = rcw.BatchedSingleRoom(n, seed=seed)
    ref = oracle.Batch(n, seed=seed)
    rng = np.random.default_rng(3)
    total_done = 0
    for t in range(T):
        a = rng.choice(np.array([1, 1, 1, 1, 3, 4], np.uint8), size=n)
        env.act(a)
        assert ref.step(a) == 0
        if t % 40 == 39 or t == T - 1:
            st = env.get_state()
            pos, au, goal = ref.states()
            r, d = ref.reward_done()
            np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
            np.testing.assert_array_equal(st["dir_au"], au)
            np.testing.assert_array_equal(st["goal"], goal)
            np.testing.assert_array_equal(st["reward"], r)
            np.testing.assert_array_equal(st["done"], d)
            total_done += int(d.sum())
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    stats = env.episode_stats()
    assert stats == ref.episode_stats()
    assert stats[0] > 0, "the run should contain finished episodes"
    assert stats[1] == stats[0] * 1.0, "return at termination == goal_reward (test/runtests.jl:33)"
    env.close()


def test_large_map_config5_matches_oracle(rcw, oracle):
    """64x64 tiles, 256 directions (BASELINE config 5 geometry), default camera."""
    n, seed = 24, 99
    kw = dict(height_tile_map_tu=64, width_tile_map_tu=64, num_directions=256)
    env = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=64, W=64, N=256), seed=seed)
    env.step_random(120)
    ref.rollout(120, threads=4)
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
    np.testing.assert_array_equal(st["dir_au"], au)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    env.close()


def test_reference_test_strategy(rcw):
    """test/runtests.jl:15-44 on the mirrored API: after reset reward == 0 and not terminated; the
    return at termination is goal_reward."""
    env = rcw.RLBaseEnv(rcw.SingleRoom(seed=5))
    rng = np.random.default_rng(0)
    for _ in range(3):
        rcw.reset(env)
        assert rcw.reward(env) == 0
        assert rcw.is_terminated(env) is False
        total = 0.0
        for _ in range(1500):
            s = rcw.state(env)
            assert s.shape == (256, 512) and s.dtype == np.uint32
            env(int(rng.choice(list(rcw.action_space(env)))))
            total += float(rcw.reward(env))
            if rcw.is_terminated(env):
                assert total == float(env.env.world.goal_reward)
                break
    assert rcw.get_action_names(env) == ("MOVE_FORWARD", "MOVE_BACKWARD", "TURN_LEFT", "TURN_RIGHT")
    env.env.close()


# ------------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------------

def test_invalid_actions(rcw):
    import torch

    env = rcw.BatchedSingleRoom(8, seed=1)
    before = env.get_state()
    with pytest.raises(AssertionError):
        env.act(np.array([1, 2, 3, 4, 5, 1, 1, 1], np.uint8))
    with pytest.raises(rcw.InvalidActionError):
        env.act(np.array([0, 2, 3, 4, 1, 1, 1, 1], np.uint8))
    after = env.get_state()
    for k in before:
        np.testing.assert_array_equal(before[k], after[k])  # nothing was enqueued
    # device-side array: the offending env is skipped, the rest step, the error surfaces on sync
    a = torch.tensor([3, 3, 9, 3, 3, 3, 3, 3], dtype=torch.uint8, device="cuda")
    env.act(a)
    with pytest.raises(rcw.InvalidActionError):
        env.sync()
    st = env.get_state()
    expect = (before["dir_au"] + 1) % 128
    expect[2] = before["dir_au"][2]
    np.testing.assert_array_equal(st["dir_au"], expect)
    env.close()


def test_wall_edge_f6_is_defined(rcw, oracle):
    """walking axis-aligned into the bottom / right wall makes the reference index tile
    H+1 / W+1; here tiles outside the map are empty, the wall tile itself blocks the move."""
    env = rcw.BatchedSingleRoom(2, auto_reset=False)
    # env 0 faces +x (au 0) next to the bottom wall, env 1 faces +y (au 32) next to the right wall
    env.reset(goal_ij=[[2, 2], [2, 2]], player_ij=[[7, 8], [4, 15]], dir_au=[0, 32])
    ref = [oracle.World(), oracle.World()]
    ref[0].reset_to(2, 2, 7, 8, 0)
    ref[1].reset_to(2, 2, 4, 15, 32)
    for _ in range(8):
        env.act(np.array([1, 1], np.uint8))
        for w in ref:
            w.act(1)
    st = env.get_state()
    for e, w in enumerate(ref):
        np.testing.assert_array_equal(bits(st["pos"][e]), bits(w.state()["pos"]))
    assert st["pos"][0, 0] < 7.0 - 0.125 + 1e-6 and st["pos"][1, 1] < 15.0 - 0.125 + 1e-6
    assert not st["done"].any()
    env.close()


def test_masked_reset_and_host_layouts(rcw):
    env = rcw.BatchedSingleRoom(6, seed=11, auto_reset=False)
    before = env.get_state()
    mask = np.array([1, 0, 1, 0, 0, 1], np.uint8)
    env.reset(mask=mask)
    after = env.get_state()
    for e in range(6):
        same = np.array_equal(before["pos"][e], after["pos"][e]) and before["dir_au"][e] == after["dir_au"][e] \
            and np.array_equal(before["goal"][e], after["goal"][e])
        if not mask[e]:
            assert same
    # a fresh Philox episode for the masked envs: at least one of them changed
    assert any(not (np.array_equal(before["pos"][e], after["pos"][e])
                    and before["dir_au"][e] == after["dir_au"][e]
                    and np.array_equal(before["goal"][e], after["goal"][e])) for e in (0, 2, 5))
    with pytest.raises(rcw.RcwError):
        env.reset(goal_ij=np.full((6, 2), 99), player_ij=np.full((6, 2), 2), dir_au=np.zeros(6))
    env.close()


def test_custom_wall_map_matches_oracle(rcw, oracle):
    """A host-supplied wall layer (pillars inside the room)."""
    H, W = 12, 20
    wall = np.zeros((H, W), bool)
    wall[0, :] = wall[-1, :] = wall[:, 0] = wall[:, -1] = True
    wall[3:9:2, 4:16:3] = True
    n, seed = 16, 21
    env = rcw.BatchedSingleRoom(n, seed=seed, height_tile_map_tu=H, width_tile_map_tu=W, num_rays=256,
                                height_camera_view_pu=128)
    env.set_wall_map(wall)
    env.reset()
    cfg = oracle.default_config(H=H, W=W, R=256, P=128)
    ref = oracle.Batch(n, cfg=cfg, seed=seed)
    for e in range(n):
        ref.world(e).set_wall_map(wall)
    ref.reset()  # episode 2 on both sides
    env.step_random(200)
    ref.rollout(200, threads=4)
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
    np.testing.assert_array_equal(st["goal"], goal)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    env.close()


@pytest.mark.parametrize("R", [32, 96, 512])
def test_per_env_wall_maps_match_oracle(rcw, oracle, R):
    """Every env has its own wall layer (random pillars), staged per env by TMA; R = 32 puts eight
    envs in one CTA, R = 512 spreads one env over two CTAs."""
    H, W, n, seed = 16, 24, 40, 5
    rng = np.random.default_rng(R)
    walls = np.zeros((n, H, W), bool)
    walls[:, 0, :] = walls[:, -1, :] = walls[:, :, 0] = walls[:, :, -1] = True
    walls[:, 2:-2, 2:-2] |= rng.random((n, H - 4, W - 4)) < 0.12
    env = rcw.BatchedSingleRoom(n, seed=seed, height_tile_map_tu=H, width_tile_map_tu=W, num_rays=R,
                                height_camera_view_pu=64)
    env.set_wall_maps(walls)
    env.reset()
    ref = oracle.Batch(n, cfg=oracle.default_config(H=H, W=W, R=R, P=64), seed=seed)
    for e in range(n):
        ref.world(e).set_wall_map(walls[e])
    ref.reset()
    for _ in range(3):
        env.step_random(100)
        ref.rollout(100, threads=4)
        st = env.get_state()
        pos, au, goal = ref.states()
        np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
        np.testing.assert_array_equal(st["dir_au"], au)
        np.testing.assert_array_equal(st["goal"], goal)
        np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    rays = env.get_rays()
    for e in range(n):
        np.testing.assert_array_equal(rays["hit"][e], ref.world(e).ray_stop)
    # back to one shared layer
    env.set_wall_map(walls[0])
    env.reset()
    for e in range(n):
        ref.world(e).set_wall_map(walls[0])
    ref.reset()
    env.step_random(50)
    ref.rollout(50, threads=4)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    env.close()


@pytest.mark.parametrize("R,P", [(1, 1), (2, 3), (31, 7), (33, 16), (64, 84), (100, 17)])
@pytest.mark.parametrize("fmt", ["rgb8", "xrgb32"])
def test_ragged_observation_sizes(rcw, oracle, R, P, fmt):
    """num_rays not a multiple of 32, column sizes not a multiple of 16 bytes, tiny images."""
    n, seed = 5, 17
    env = rcw.BatchedSingleRoom(n, seed=seed, num_rays=R, height_camera_view_pu=P, obs_format=fmt)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=R, P=P), seed=seed)
    env.step_random(30)
    ref.rollout(30)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8() if fmt == "rgb8" else ref.obs_u32())
    env.close()


@pytest.mark.parametrize("R,P", [(64, 64), (40, 50), (96, 256)])
@pytest.mark.parametrize("fmt", ["rgb8", "xrgb32"])
def test_custom_palette_takes_the_phase_rotated_path(rcw, oracle, R, P, fmt):
    """A palette whose colours have three different bytes: single-colour runs are no longer one byte
    repeated, so every sector goes through the funnel-shift path (mirror-pair and pitched renderers)."""
    pal = [0x00102030, 0x00A0B0C0, 0x00112233, 0x00445566, 0x00778899, 0x00AABBCC]
    n, seed = 6, 31
    env = rcw.BatchedSingleRoom(n, seed=seed, num_rays=R, height_camera_view_pu=P, obs_format=fmt, palette=pal)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=R, P=P, palette=pal), seed=seed)
    env.step_random(40)
    ref.rollout(40)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8() if fmt == "rgb8" else ref.obs_u32())
    env.close()


@pytest.mark.parametrize("R,P", [(512, 256), (84, 84), (33, 50)])
def test_gray8_format_matches_oracle(rcw, oracle, R, P):
    """Learner-facing one-byte format: luma of the reference pixel."""
    n, seed = 12, 77
    env = rcw.BatchedSingleRoom(n, seed=seed, num_rays=R, height_camera_view_pu=P, obs_format="gray8")
    ref = oracle.Batch(n, cfg=oracle.default_config(R=R, P=P), seed=seed)
    env.step_random(60)
    ref.rollout(60, threads=4)
    obs = env.copy_obs()
    assert obs.shape == (n, R, P) and obs.dtype == np.uint8
    np.testing.assert_array_equal(obs, ref.obs_gray8())
    assert set(np.unique(obs).tolist()) <= {255, 64, 128, 192, 39, 58}
    env.close()


def test_device_tensor_view_matches_host_copy(rcw):
    """obs_tensor() is a zero-copy (possibly pitched) view of the same bytes copy_obs() returns."""
    import torch

    for kw in (dict(), dict(num_rays=45, height_camera_view_pu=51), dict(num_rays=64, height_camera_view_pu=84)):
        for fmt in ("rgb8", "xrgb32", "gray8"):
            env = rcw.BatchedSingleRoom(5, seed=2, obs_format=fmt, **kw)
            env.step_random(7)
            env.sync()
            t = env.obs_tensor()
            assert tuple(t.shape) == env.obs_shape
            host = env.copy_obs()
            dev = t.contiguous().cpu().numpy()
            if fmt == "xrgb32":
                dev = dev.view(np.uint32)
            np.testing.assert_array_equal(dev, host)
            es, cs, cbytes, bpp = env.obs_layout()
            assert cs % 32 == 0 and es % 128 == 0 and cs >= cbytes and bpp == env.bytes_per_pixel
            env.close()


def test_looping_grid_matches_oracle(rcw, oracle, monkeypatch):
    """Small items run on a capped grid whose CTAs loop over several rounds of eight items (shared
    act! pose double-buffered across rounds).  Force one CTA per SM so that 700 envs x 3 groups need
    two rounds, with env boundaries inside CTAs and a ragged last group."""
    monkeypatch.setenv("RCW_CTAS_PER_SM", "1")
    n, seed, R, P = 700, 13, 80, 40
    env = rcw.BatchedSingleRoom(n, seed=seed, num_rays=R, height_camera_view_pu=P)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=R, P=P), seed=seed)
    for _ in range(2):
        env.step_random(60)
        ref.rollout(60, threads=4)
        st = env.get_state()
        pos, au, goal = ref.states()
        np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
        np.testing.assert_array_equal(st["dir_au"], au)
        np.testing.assert_array_equal(st["goal"], goal)
        np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    assert env.episode_stats() == ref.episode_stats()
    env.close()


def test_direction_table_fallbacks_match_oracle(rcw, oracle):
    """The direction table normally sits in one of 8 constant-memory slots of 512 entries; more
    handles than slots, or more directions than a slot holds, read it from global memory instead."""
    seed = 41
    handles = [rcw.BatchedSingleRoom(3, seed=seed + k, num_rays=64, height_camera_view_pu=32) for k in range(11)]
    for k, env in enumerate(handles):
        env.step_random(25)
        ref = oracle.Batch(3, cfg=oracle.default_config(R=64, P=32), seed=seed + k)
        ref.rollout(25)
        np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    for env in handles:
        env.close()
    env = rcw.BatchedSingleRoom(4, seed=seed, num_directions=600, num_rays=64, height_camera_view_pu=32)
    ref = oracle.Batch(4, cfg=oracle.default_config(N=600, R=64, P=32), seed=seed)
    env.step_random(80)
    ref.rollout(80)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    np.testing.assert_array_equal(env.get_state()["dir_au"], ref.states()[1])
    env.close()


def test_wide_map_and_large_env_ids_match_oracle(rcw, oracle):
    """A 40 x 100 tile map (four 32-bit words per bit-packed row) and global env ids above 2^32
    (the high counter word of the Philox streams)."""
    n, seed, off = 10, 3, (1 << 33) + 5
    kw = dict(height_tile_map_tu=40, width_tile_map_tu=100, num_rays=96, height_camera_view_pu=64)
    env = rcw.BatchedSingleRoom(n, seed=seed, env_id_offset=off, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=40, W=100, R=96, P=64), seed=seed, env_id_offset=off)
    env.step_random(150)
    ref.rollout(150, threads=4)
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
    np.testing.assert_array_equal(st["goal"], goal)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    env.close()


def test_two_host_threads_drive_two_handles(rcw, oracle):
    """Threading contract of the ABI: distinct handles may be driven concurrently from different host
    threads (each has its own stream); results are those of the sequential oracle."""
    import threading

    results, errors = {}, []

    def worker(k):
        try:
            env = rcw.BatchedSingleRoom(48, seed=100 + k, num_rays=128, height_camera_view_pu=64)
            rng = np.random.default_rng(k)
            acts = rng.integers(1, 5, size=(120, 48)).astype(np.uint8)
            for t in range(120):
                env.act(acts[t])
                if t % 7 == 0:
                    env.reward_done()
            results[k] = (acts, env.copy_obs(), env.get_state(), env.episode_stats())
            env.close()
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(2):
        acts, obs, st, stats = results[k]
        ref = oracle.Batch(48, cfg=oracle.default_config(R=128, P=64), seed=100 + k)
        for t in range(120):
            assert ref.step(acts[t]) == 0
        np.testing.assert_array_equal(obs, ref.obs_rgb8())
        np.testing.assert_array_equal(bits(st["pos"]), bits(ref.states()[0]))
        assert stats == ref.episode_stats()


def test_learner_view_is_channels_last_and_feeds_a_conv(rcw):
    """obs_tensor_nchw(): zero-copy NCHW view in channels_last format, consumable by torch convs."""
    import torch

    env = rcw.BatchedSingleRoom(6, seed=8)
    env.step_random(5)
    env.sync()
    x = env.obs_tensor_nchw()
    assert tuple(x.shape) == (6, 3, 512, 256)
    assert x.is_contiguous(memory_format=torch.channels_last)
    assert x.data_ptr() == env.obs_device_ptr()[0]
    host = env.copy_obs()                                    # [N, R, P, 3]
    np.testing.assert_array_equal(x.cpu().numpy(), host.transpose(0, 3, 1, 2))
    conv = torch.nn.Conv2d(3, 4, 3, padding=1).cuda().to(memory_format=torch.channels_last)
    y = conv(x.float() / 255)
    assert tuple(y.shape) == (6, 4, 512, 256) and torch.isfinite(y).all()
    g = rcw.BatchedSingleRoom(2, seed=8, obs_format="gray8")
    assert tuple(g.obs_tensor_nchw().shape) == (2, 1, 512, 256) and g.obs_tensor_nchw().is_contiguous()
    g.close()
    env.close()


def test_range_errors(rcw):
    env = rcw.BatchedSingleRoom(4, seed=1)
    with pytest.raises(rcw.RcwError):
        env.get_rays(3, 2)
    with pytest.raises(rcw.RcwError):
        env.copy_obs(-1, 1)
    with pytest.raises(rcw.RcwError):
        env.set_state(pos=np.full((4, 2), 99.0, np.float32))
    with pytest.raises(rcw.RcwError):
        env.set_state(dir_au=np.full(4, 128, np.int32))
    with pytest.raises(rcw.RcwError):
        env.reset(goal_ij=np.full((4, 2), 3))          # the three layout arrays go together
    env.close()


def test_create_errors(rcw):
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(0)
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(1, height_tile_map_tu=2)
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(1, player_radius_wu=0.75)
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(1, device=99)


# ------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: 4096 envs, default camera)
# ------------------------------------------------------------------------------------------------

def test_full_size_properties(rcw, oracle):
    import torch

    n, seed = 4096, 0x5EED
    env = rcw.BatchedSingleRoom(n, seed=seed)
    env.step_random(20)
    env.sync()
    obs = env.obs_tensor()
    assert tuple(obs.shape) == (n, 512, 256, 3) and obs.dtype == torch.uint8
    # (1) only palette colours
    px = obs.view(-1, 3).to(torch.int32)
    packed = (px[:, 0] << 16) | (px[:, 1] << 8) | px[:, 2]
    present = set(torch.unique(packed).tolist())
    assert present <= {0xFFFFFF, 0x404040, 0x808080, 0xC0C0C0, 0x800000, 0xC00000}
    assert {0xFFFFFF, 0x404040} < present
    # (2) every column is ceiling^pad, colour^(P-2pad), floor^pad: the picture is its own mirror
    #     about the horizontal centre line once ceiling and floor are identified
    grey = packed.view(n, 512, 256)
    top, bottom = grey[:, :, :128], torch.flip(grey[:, :, 128:], dims=[2])
    is_ceiling, is_floor = top == 0xFFFFFF, bottom == 0x404040
    assert torch.equal(is_ceiling, is_floor)
    assert torch.equal(top[~is_ceiling], bottom[~is_floor])
    # ceiling run is a prefix of the column
    run = is_ceiling.to(torch.int8)
    assert bool((run[:, :, 1:] <= run[:, :, :-1]).all())
    # (3) determinism and shard independence: envs [1024, 1088) stepped as their own shard
    sub = rcw.BatchedSingleRoom(64, seed=seed, env_id_offset=1024)
    sub.step_random(20)
    np.testing.assert_array_equal(sub.copy_obs(), env.copy_obs(1024, 64))
    a, b = sub.get_state(), env.get_state()
    np.testing.assert_array_equal(a["pos"], b["pos"][1024:1088])
    # (4) a sample of envs against the oracle at full batch size
    ref = oracle.Batch(64, seed=seed, env_id_offset=1024)
    ref.rollout(20, threads=4)
    np.testing.assert_array_equal(sub.copy_obs(), ref.obs_rgb8())
    sub.close()
    env.close()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs at their full sizes
# ------------------------------------------------------------------------------------------------

def test_config1_single_env_10k_random_steps(rcw, oracle):
    """configs[0]: one default SingleRoom, reset + 10,000 uniformly random act! steps with the camera
    view rendered every step.  Reward and termination are compared at every step, the UInt32 camera
    view (the reference's own pixel format) every 250 steps; no auto-reset, like the reference."""
    env = rcw.SingleRoom(seed=4)
    ref = oracle.Batch(1, seed=4, auto_reset=False)
    rng = np.random.default_rng(10)
    actions = rng.integers(1, 5, 10_000)
    n_done = 0
    for t, a in enumerate(actions):
        env.act(int(a))
        assert ref.step(np.array([a], np.uint8)) == 0
        r, d = env.reward_done()
        rr, rd = ref.reward_done()
        assert r[0] == rr[0] and d[0] == rd[0]
        n_done += int(d[0])
        if t % 250 == 249:
            np.testing.assert_array_equal(env.camera_view, ref.world(0).camera_view.T)
            st = env.get_state()
            assert bits(st["pos"]).tolist() == bits(ref.states()[0]).tolist()
    env.close()


def _sampled_shards_match_oracle(rcw, oracle, env, kw_oracle, n_total, seed, steps, starts, width=8):
    for s0 in starts:
        ref = oracle.Batch(width, cfg=oracle.default_config(**kw_oracle), seed=seed, env_id_offset=s0)
        ref.rollout(steps, threads=4)
        np.testing.assert_array_equal(env.copy_obs(s0, width), ref.obs_rgb8())
        pos, au, goal = ref.states()
        st = env.get_state()
        np.testing.assert_array_equal(bits(st["pos"][s0:s0 + width]), bits(pos))
        np.testing.assert_array_equal(st["goal"][s0:s0 + width], goal)


def test_long_horizon_20k_steps_sampled_against_oracle(rcw, oracle):
    """configs[1] for 20,000 steps (about one auto-reset per env on average under the random policy): a
    window of eight envs is replayed by the oracle and compared bit for bit; the episode counters obey
    return == episodes (goal_reward 1)."""
    n, seed, steps, s0 = 4096, 0x5EED, 20_000, 2040
    env = rcw.BatchedSingleRoom(n, seed=seed)
    env.step_random(steps)
    ref = oracle.Batch(8, seed=seed, env_id_offset=s0)
    ref.rollout(steps, threads=4, render=False)
    for e in range(8):
        ref.world(e).update_camera_view()
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"][s0:s0 + 8]), bits(pos))
    np.testing.assert_array_equal(st["dir_au"][s0:s0 + 8], au)
    np.testing.assert_array_equal(st["goal"][s0:s0 + 8], goal)
    np.testing.assert_array_equal(env.copy_obs(s0, 8), ref.obs_rgb8())
    assert ref.episode_stats()[0] >= 4, "the window should have gone through several episodes"
    ep, sr, sl = env.episode_stats()
    assert ep > n // 2 and sr == float(ep) and ep <= sl <= n * steps
    env.close()


def test_config4_65536_envs_sampled_against_oracle(rcw, oracle):
    """configs[3]: 65,536 envs at 512 rays x 256 px (25.8 GB of observations per step).  Windows of
    eight envs spread over the batch, including both ends, are replayed by the oracle."""
    n, seed, steps = 65536, 0x5EED, 12
    env = rcw.BatchedSingleRoom(n, seed=seed)
    env.step_random(steps)
    _sampled_shards_match_oracle(rcw, oracle, env, {}, n, seed, steps, [0, 4093, 32768, 50001, n - 8])
    ep, sr, sl = env.episode_stats()
    assert sr == float(ep) and sl >= ep          # every finished episode returned goal_reward = 1
    env.close()


def test_config5_262144_envs_large_map_sampled_against_oracle(rcw, oracle):
    """configs[4]: 262,144 envs on 64x64 tile maps with 256 directions (103 GB of observations)."""
    n, seed, steps = 262144, 5, 6
    kw = dict(height_tile_map_tu=64, width_tile_map_tu=64, num_directions=256)
    env = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    env.step_random(steps)
    _sampled_shards_match_oracle(rcw, oracle, env, dict(H=64, W=64, N=256), n, seed, steps,
                                 [0, 131071, n - 8])
    env.close()
