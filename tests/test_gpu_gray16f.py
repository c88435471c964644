"""RCW_OBS_GRAY16F: normalised float16 frames — the GRAY8 luma / 255 rounded to IEEE binary16 —
written by the renderer itself, in both step kernels, the table renderer, pitched geometries, with extra object
layers, through the frame ring and through rcw_expand_columns.  Bit-exact against the oracle's definition."""
import numpy as np
import pytest

from conftest import LAYERED_CONFIG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def h16(a):
    return np.ascontiguousarray(a, np.float16).view(np.uint16)


@pytest.mark.parametrize("env_kernel,table_kb", [(0, 64), (1, 64), (1, 0)])
@pytest.mark.parametrize("R,P", [(512, 256), (84, 84), (33, 50), (128, 128), (20, 7)])
def test_gray16f_rollout_matches_oracle(rcw, oracle, monkeypatch, env_kernel, table_kb, R, P):
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    monkeypatch.setenv("RCW_COL_TABLE_KB", str(table_kb))
    n, seed = 11, 3 * R + P
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format="gray16f", num_rays=R, height_camera_view_pu=P)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=R, P=P), seed=seed)
    for steps in (0, 1, 40):
        env.step_random(steps)
        ref.rollout(steps)
        obs = env.copy_obs()
        assert obs.dtype == np.float16 and obs.shape == (n, R, P)
        np.testing.assert_array_equal(h16(obs), h16(ref.obs_gray16f()))
    lumas = [(77 * (c >> 16) + 150 * ((c >> 8) & 255) + 29 * (c & 255) + 128) >> 8
             for c in (0xFFFFFF, 0x404040, 0x808080, 0xC0C0C0, 0x800000, 0xC00000)]        # the palette, single_room.jl:291-296
    assert set(np.unique(obs).tolist()) <= {float(np.float16(np.float32(v) / np.float32(255))) for v in lumas}
    t = env.obs_tensor()
    assert str(t.dtype) == "torch.float16" and tuple(t.shape) == (n, R, P)
    env.sync()
    np.testing.assert_array_equal(h16(t.cpu().numpy()), h16(obs))
    env.close()


def test_gray16f_custom_palette_layers_and_expansion(rcw, oracle, golden):
    from test_gpu_layers import LAYERED_KW, furnish

    states, au, goal = golden["L_states"], golden["L_au"], golden["L_goal"]
    n = len(states)
    pal = [0x00F0E0D0, 0x00102030, 0x00806040, 0x00A08060, 0x00C02010, 0x00E04020]
    env = rcw.BatchedSingleRoom(n, obs_format="gray16f", auto_reset=False, palette=pal, **LAYERED_KW)
    furnish(env, golden)
    env.set_state(pos=states, dir_au=au, goal=goal)
    env.render()
    from conftest import layered_oracle_world

    cfg = dict(LAYERED_CONFIG, palette=pal)
    w = oracle.World(oracle.default_config(**cfg))
    w.set_layer(1, golden["L_wall"])
    for k in range(3):
        w.set_layer(3 + k, golden["L_extra"][k])
    want = []
    for k in range(n):
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_camera_view()
        c = w.camera_view
        luma = (77 * ((c >> 16) & 255) + 150 * ((c >> 8) & 255) + 29 * (c & 255) + 128) >> 8
        want.append((luma.astype(np.float32) / np.float32(255)).astype(np.float16))
    want = np.stack(want)
    np.testing.assert_array_equal(h16(env.copy_obs()), h16(want))
    env.close()
    words = rcw.BatchedSingleRoom(n, obs_format="columns", auto_reset=False, palette=pal, **LAYERED_KW)
    furnish(words, golden)
    words.set_state(pos=states, dir_au=au, goal=goal)
    words.render()
    got = words.expand_columns(pixel_format="gray16f")
    words.sync()
    np.testing.assert_array_equal(h16(got.cpu().numpy()), h16(want))
    words.close()


def test_gray16f_frame_ring(rcw, oracle):
    n, K, seed = 6, 3, 12
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format="gray16f", num_rays=64, height_camera_view_pu=48, frame_stack=K)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=64, P=48), seed=seed)
    frames = []
    for _ in range(K):
        env.step_random(1)
        ref.rollout(1)
        frames.append(ref.obs_gray16f())
    for age in range(K):
        np.testing.assert_array_equal(h16(env.copy_obs(age=age)), h16(frames[K - 1 - age]))
    env.close()
