"""RCW_OBS_GRAY8_HALF: the GRAY8 frame under a 2 x 2 box filter, composed by env_kernel from the two
rays' column decisions without writing the full-resolution frame.  Bit-exact against the oracle's definition
((a + b + c + d + 2) >> 2 over the full GRAY8 frame) for aligned and ragged geometries, custom palettes, object
layers, both map views, the frame ring, observation windows and multi-step calls on two streams."""
import numpy as np
import pytest

from conftest import LAYERED_CONFIG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def box(g):
    g = g.astype(np.uint32)
    return ((g[:, 0::2, 0::2] + g[:, 0::2, 1::2] + g[:, 1::2, 0::2] + g[:, 1::2, 1::2] + 2) >> 2).astype(np.uint8)


@pytest.mark.parametrize("room", [0, 1])
@pytest.mark.parametrize("R,P", [(512, 256), (84, 84), (64, 64), (34, 50), (2, 2), (130, 6), (96, 200)])
def test_half_rollout_matches_oracle(rcw, oracle, monkeypatch, room, R, P):
    monkeypatch.setenv("RCW_ROOM", str(room))
    monkeypatch.setenv("RCW_ENV_PER_WARP", "0")           # the format must take env_kernel whatever the switches say
    monkeypatch.setenv("RCW_TWO_STREAMS_MIN", "8")
    n, seed = 13, 5 * R + P
    kw = dict(height_tile_map_tu=6, width_tile_map_tu=9, num_directions=64)
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format="gray8_half", num_rays=R, height_camera_view_pu=P, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=6, W=9, N=64, R=R, P=P), seed=seed)
    for steps in (0, 1, 60):
        env.step_random(steps)
        ref.rollout(steps)
        obs = env.copy_obs()
        assert obs.dtype == np.uint8 and obs.shape == (n, R // 2, P // 2)
        np.testing.assert_array_equal(obs, ref.obs_gray8_half())
        np.testing.assert_array_equal(obs, box(ref.obs_gray8()))
    rng = np.random.default_rng(seed)
    for _ in range(8):
        a = rng.choice([1, 1, 2, 3, 4], size=n).astype(np.uint8)
        env.act(a)
        assert ref.step(a) == 0
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_gray8_half())
    t = env.obs_tensor()
    env.sync()
    np.testing.assert_array_equal(t.cpu().numpy(), ref.obs_gray8_half())
    st = env.get_state()
    np.testing.assert_array_equal(st["dir_au"], ref.states()[1])
    rays = env.get_rays()                                   # the ray dump is independent of the observation format
    np.testing.assert_array_equal(rays["hit"][0], ref.world(0).ray_stop)
    env.close()


def test_half_needs_even_sizes(rcw):
    for R, P in ((33, 50), (64, 51)):
        with pytest.raises(rcw.RcwError):
            rcw.BatchedSingleRoom(2, obs_format="gray8_half", num_rays=R, height_camera_view_pu=P)


def test_half_with_palette_layers_ring_and_window(rcw, oracle, golden):
    from test_gpu_layers import LAYERED_KW, furnish

    states, au, goal = golden["L_states"], golden["L_au"], golden["L_goal"]
    n = len(states)
    pal = [0x00F0E0D0, 0x00102030, 0x00806040, 0x00A08060, 0x00C02010, 0x00E04020]
    env = rcw.BatchedSingleRoom(n, obs_format="gray8_half", auto_reset=False, palette=pal, frame_stack=2, **LAYERED_KW)
    furnish(env, golden)
    env.set_state(pos=states, dir_au=au, goal=goal)
    env.render()
    w = oracle.World(oracle.default_config(**dict(LAYERED_CONFIG, palette=pal)))
    w.set_layer(1, golden["L_wall"])
    for k in range(3):
        w.set_layer(3 + k, golden["L_extra"][k])
    want = []
    for k in range(n):
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_camera_view()
        c = w.camera_view
        want.append((77 * ((c >> 16) & 255) + 150 * ((c >> 8) & 255) + 29 * (c & 255) + 128) >> 8)
    want = box(np.stack(want))
    np.testing.assert_array_equal(env.copy_obs(), want)
    env.act(np.full(n, 3, np.uint8))                        # turn: the previous frame moves to age 1
    np.testing.assert_array_equal(env.copy_obs(age=1), want)
    env.close()
    # observation window: 10 envs through 4 slots
    win = rcw.BatchedSingleRoom(10, seed=2, obs_format="gray8_half", num_rays=64, height_camera_view_pu=32, obs_window_envs=4)
    ref = oracle.Batch(10, cfg=oracle.default_config(R=64, P=32), seed=2)
    win.step_random(5)
    ref.rollout(5)
    np.testing.assert_array_equal(win.copy_obs(8, 2), ref.obs_gray8_half()[8:10])   # the last window holds envs 8, 9 in slots 0, 1
    win.close()
