"""Object layers beyond WALL and GOAL (NUM_OBJECTS > 2; single_room.jl:16-18,148-163,209,355-360,417-429; SURVEY.md
8(f) N2): the CUDA path against the oracle and against the golden case L (made by the independent Python
restatement).  Every object stops rays, a column takes the colours of the first object on the hit tile, the top view
shows findfirst over the layers, terminal layers end the episode with their own reward, blocking layers refuse the
move, and resets never place the player on an object."""
import zlib

import numpy as np
import pytest

from conftest import LAYERED_CONFIG, layered_oracle_world

pytestmark = pytest.mark.gpu

LAYERED_KW = dict(height_tile_map_tu=9, width_tile_map_tu=12, num_directions=32, num_rays=64, height_camera_view_pu=48,
                  player_radius_wu=np.float32(0.15), position_increment_wu=np.float32(0.2), pu_per_tu=4,
                  num_object_layers=5, layer_kind=["blocking", "terminal", "terminal"], layer_reward=[0.0, -1.0, 0.5],
                  layer_palette=[(0x00205080, 0x003070A0), (0x00A04000, 0x00C06000), (0x0000A040, 0x0000C060)],
                  layer_top_color=[0x000000FF, 0x00FF8000, 0x0000FF00])


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def furnish(env, golden):
    env.set_layer(1, golden["L_wall"])
    for k in range(3):
        env.set_layer(3 + k, golden["L_extra"][k])


def rgb8_of(img_u32):
    return np.stack([(img_u32 >> 16) & 255, (img_u32 >> 8) & 255, img_u32 & 255], -1).astype(np.uint8)


def gray8_of(c):
    r, g, b = (c >> 16) & 255, (c >> 8) & 255, c & 255
    return ((77 * r + 150 * g + 29 * b + 128) >> 8).astype(np.uint8)


@pytest.mark.parametrize("env_kernel", [0, 1])
@pytest.mark.parametrize("fmt", ["rgb8", "xrgb32", "gray8", "columns"])
def test_layered_cast_render_and_top_view_match_golden(rcw, oracle, golden, monkeypatch, fmt, env_kernel):
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    states, au, goal = golden["L_states"], golden["L_au"], golden["L_goal"]
    n = len(states)
    env = rcw.BatchedSingleRoom(n, obs_format=fmt, auto_reset=False, **LAYERED_KW)
    furnish(env, golden)
    env.set_state(pos=states, dir_au=au, goal=goal)
    env.render()
    rays = env.get_rays()
    np.testing.assert_array_equal(rays["hit"], golden["L_hit"])
    np.testing.assert_array_equal(rays["dim"], golden["L_dim"])
    np.testing.assert_array_equal(bits(rays["dist"]), bits(golden["L_dist"]))
    obs = env.copy_obs()
    img = golden["L_image"]
    if fmt == "xrgb32":
        np.testing.assert_array_equal(obs, img)
    elif fmt == "rgb8":
        np.testing.assert_array_equal(obs, rgb8_of(img))
    elif fmt == "gray8":
        np.testing.assert_array_equal(obs, gray8_of(img))
    else:
        w = layered_oracle_world(oracle, golden)
        for k in range(n):
            w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
            w.cast_rays()
            np.testing.assert_array_equal(obs[k], w.camera_columns())
        assert (obs >> 16).max() >= 6, "some column should show an extra object (colour ids 6..)"
    env.render_top_view()
    np.testing.assert_array_equal(env.copy_top_view(), golden["L_top_image"])
    env.close()


@pytest.mark.parametrize("env_kernel", [0, 1])
def test_layered_act_trajectories_match_golden(rcw, golden, monkeypatch, env_kernel):
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    init, actions = golden["L_act_init"], golden["L_act_actions"]
    n, T = actions.shape
    env = rcw.BatchedSingleRoom(n, auto_reset=False, **LAYERED_KW)
    furnish(env, golden)
    env.reset(goal_ij=init[:, 0:2], player_ij=init[:, 2:4], dir_au=init[:, 4])
    for t in range(T):
        env.act(actions[:, t])
        st = env.get_state()
        np.testing.assert_array_equal(bits(st["pos"]), bits(golden["L_act_pos"][:, t]))
        np.testing.assert_array_equal(st["dir_au"], golden["L_act_au"][:, t])
        np.testing.assert_array_equal(st["reward"], golden["L_act_reward"][:, t])
        np.testing.assert_array_equal(st["done"], golden["L_act_done"][:, t])
    env.close()


@pytest.mark.parametrize("env_kernel,fmt", [(0, "rgb8"), (1, "gray8"), (1, "rgb8"), (0, "xrgb32")])
def test_layered_random_rollout_with_auto_reset_matches_oracle(rcw, oracle, golden, monkeypatch, env_kernel, fmt):
    """Philox resets (never on an object), random policy, terminal layers with rewards -1 / 0.5, episode statistics."""
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n, seed, steps = 53, 31, 700
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format=fmt, **LAYERED_KW)
    furnish(env, golden)
    env.reset()
    ref = oracle.Batch(n, cfg=oracle.default_config(**LAYERED_CONFIG), seed=seed)
    for e in range(n):
        w = ref.world(e)
        w.set_layer(1, golden["L_wall"])
        for k in range(3):
            w.set_layer(3 + k, golden["L_extra"][k])
    ref.reset()
    occupied = golden["L_wall"] | golden["L_extra"].any(axis=0)
    st = env.get_state()
    tiles = np.floor(st["pos"]).astype(int)
    assert not occupied[tiles[:, 0], tiles[:, 1]].any(), "a reset placed a player on an object"
    env.step_random(steps)
    ref.rollout(steps, threads=4)
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
    ep, ret, length = env.episode_stats()
    assert (ep, ret, length) == ref.episode_stats()
    assert ep > 20 and ret != ep, "episodes should have ended on the goal AND on the terminal layers (returns -1 / 0.5 / 1)"
    env.render_top_view()
    top = env.copy_top_view()
    for e in (0, n // 2, n - 1):
        w = ref.world(e)
        w.update_top_view()
        np.testing.assert_array_equal(top[e], w.top_view)
    env.close()


def test_expand_columns_paints_extra_object_colours(rcw, oracle, golden):
    states, au, goal = golden["L_states"], golden["L_au"], golden["L_goal"]
    n = len(states)
    words_env = rcw.BatchedSingleRoom(n, obs_format="columns", auto_reset=False, **LAYERED_KW)
    furnish(words_env, golden)
    words_env.set_state(pos=states, dir_au=au, goal=goal)
    words_env.render()
    got = words_env.expand_columns(pixel_format="xrgb32")
    words_env.sync()
    np.testing.assert_array_equal(got.cpu().numpy().view(np.uint32), golden["L_image"])
    words_env.close()


def test_layer_errors(rcw, golden):
    env = rcw.BatchedSingleRoom(3, **LAYERED_KW)
    with pytest.raises(rcw.RcwError):
        env.set_layer(2, golden["L_wall"])            # GOAL is a position per env, not a map
    with pytest.raises(rcw.RcwError):
        env.set_layer(6, golden["L_wall"])            # num_object_layers = 5
    with pytest.raises(rcw.RcwError):
        env.set_wall_maps(np.stack([golden["L_wall"]] * 3))   # per-env wall layers + extra objects
    env.close()
    plain = rcw.BatchedSingleRoom(2, height_tile_map_tu=9, width_tile_map_tu=12)
    with pytest.raises(rcw.RcwError):
        plain.set_layer(3, golden["L_wall"])          # NUM_OBJECTS = 2: there is no layer 3
    plain.set_layer(1, golden["L_wall"])              # layer 1 is rcw_set_wall_map
    plain.close()
