"""GPU tests of the column-word observation format (RCW_OBS_COLUMNS) and of rcw_expand_columns:
the words must equal the oracle's per-column decisions of update_camera_view! (single_room.jl:404-439), and their
expansion must reproduce the oracle's pixel image bit for bit in every pixel format."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def _oracle_pixels(ref, fmt):
    return {"rgb8": ref.obs_rgb8, "xrgb32": ref.obs_u32, "gray8": ref.obs_gray8}[fmt]()


def _as_numpy(t, fmt):
    a = t.contiguous().cpu().numpy()
    return a.view(np.uint32) if fmt == "xrgb32" else a


GEOMETRIES = [
    dict(),                                                                     # the default camera, 8x16 room
    dict(H=5, W=7, N=36, R=45, P=51, radius=np.float32(0.2), incr=np.float32(0.3), sfov=np.float32(0.5),
         cam_h=np.float32(0.8)),                                                # ragged: 45 rays, 51 rows (pitched columns)
    dict(H=64, W=64, N=256, R=128, P=96),                                       # config 5 geometry, reduced camera
    dict(R=84, P=84),
]


def _engine_kwargs(g):
    names = dict(H="height_tile_map_tu", W="width_tile_map_tu", N="num_directions", R="num_rays",
                 P="height_camera_view_pu", radius="player_radius_wu", incr="position_increment_wu",
                 sfov="semi_field_of_view_wu", cam_h="camera_height_tile_wu")
    return {names[k]: (float(v) if isinstance(v, np.floating) else v) for k, v in g.items()}


@pytest.mark.parametrize("env_kernel", [0, 1])
@pytest.mark.parametrize("geo", GEOMETRIES)
def test_columns_match_oracle_and_expand_to_the_pixel_image(rcw, oracle, monkeypatch, geo, env_kernel):
    import torch

    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))      # both step kernels: one warp per env / per 32 rays
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n, seed = 37, 21
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format="columns", **_engine_kwargs(geo))
    ref = oracle.Batch(n, cfg=oracle.default_config(**geo), seed=seed)
    assert env.obs_shape == (n, ref.cfg.R)
    assert env.obs_layout()[1:] == (4, 4, 4)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_columns())          # after the reset of rcw_create
    rng = np.random.default_rng(2)
    for t in range(25):
        if t % 4 == 3:
            env.step_random(1)
            ref.rollout(1)
        else:
            a = rng.integers(1, 5, n).astype(np.uint8)
            (env.act if t % 2 else lambda x: env.act(torch.from_numpy(x).cuda()))(a)
            assert ref.step(a) == 0
        np.testing.assert_array_equal(env.copy_obs(), ref.obs_columns(), err_msg=f"step {t}")
    words = ref.obs_columns()
    assert ((words >> 16) >= 2).all() and ((words >> 16) <= 5).all()
    np.testing.assert_array_equal(env.obs_tensor().cpu().numpy().view(np.uint32), words)
    for fmt in ("rgb8", "xrgb32", "gray8"):
        got = _as_numpy(env.expand_columns(pixel_format=fmt), fmt)
        np.testing.assert_array_equal(got, _oracle_pixels(ref, fmt), err_msg=fmt)
    # the state is the oracle's too (act! is the same code path as for pixel formats)
    s = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(s["pos"], pos)
    np.testing.assert_array_equal(s["dir_au"], au)
    r, d = ref.reward_done()
    np.testing.assert_array_equal(s["reward"], r)
    np.testing.assert_array_equal(s["done"], d)
    env.close()


def test_expand_records_from_a_replay_buffer_on_a_pixel_handle(rcw, oracle):
    """Column words kept elsewhere (here: the oracle's, uploaded with a row stride) are expanded by any handle of
    the same geometry — the handle only supplies num_rays, height_px and the palette."""
    import torch

    n, seed = 19, 4
    palette = [0x112233, 0x445566, 0x778899, 0xAABBCC, 0xDD1122, 0x3344EE]     # non-grey: the rotating RGB path
    ref = oracle.Batch(n, cfg=oracle.default_config(R=96, P=40, palette=palette), seed=seed)
    ref.rollout(7)
    pixels = rcw.BatchedSingleRoom(2, num_rays=96, height_camera_view_pu=40, palette=palette)   # an unrelated batch
    store = torch.zeros((n, 96 + 32), dtype=torch.int32, device="cuda")
    store[:, :96] = torch.from_numpy(ref.obs_columns().view(np.int32)).cuda()
    for fmt in ("rgb8", "xrgb32", "gray8"):
        got = _as_numpy(pixels.expand_columns(store[:, :96], pixel_format=fmt), fmt)
        np.testing.assert_array_equal(got, _oracle_pixels(ref, fmt), err_msg=fmt)
    picks = torch.tensor([3, 3, 18, 0], device="cuda")
    got = _as_numpy(pixels.expand_columns(store[picks][:, :96].contiguous(), pixel_format="rgb8"), "rgb8")
    np.testing.assert_array_equal(got, ref.obs_rgb8()[[3, 3, 18, 0]])
    pixels.close()


@pytest.mark.parametrize("env_kernel", [0, 1])
def test_columns_with_frame_ring_window_masked_reset_and_checkpoint(rcw, oracle, monkeypatch, env_kernel):
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n, seed, K = 24, 8, 3
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format="columns", num_rays=64, height_camera_view_pu=32, frame_stack=K)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=64, P=32), seed=seed)
    history = [ref.obs_columns()]
    for t in range(5):
        env.step_random(1)
        ref.rollout(1)
        history.append(ref.obs_columns())
        for age in range(min(K, len(history))):
            np.testing.assert_array_equal(env.copy_obs(age=age), history[-1 - age])
    blob = env.save_checkpoint()
    env.close()
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format="columns", num_rays=64, height_camera_view_pu=32)
    env.load_checkpoint(blob)
    np.testing.assert_array_equal(env.copy_obs(), history[-1])
    env.step_random(2)
    ref.rollout(2)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_columns())
    # a masked reset redraws only the chosen envs
    before = env.copy_obs()
    mask = np.zeros(n, np.uint8)
    mask[[1, 20]] = 1
    env.reset(mask=mask)
    after = env.copy_obs()
    np.testing.assert_array_equal(after[mask == 0], before[mask == 0])
    env.close()
    # observation window: env e lives in slot e mod 8
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format="columns", num_rays=64, height_camera_view_pu=32,
                                obs_window_envs=8)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=64, P=32), seed=seed)
    a = np.random.default_rng(0).integers(1, 5, n).astype(np.uint8)
    env.act_range(a[8:16], 8)
    assert ref.step(a) == 0                                  # the oracle steps everybody; the range is compared
    np.testing.assert_array_equal(env.copy_obs(8, 8), ref.obs_columns()[8:16])
    env.close()


def test_columns_errors(rcw):
    import torch

    env = rcw.BatchedSingleRoom(4, num_rays=32, height_camera_view_pu=16)
    with pytest.raises(ValueError):
        env.expand_columns()                                   # pixels already
    with pytest.raises(ValueError):
        env.expand_columns(torch.zeros((2, 31), dtype=torch.int32, device="cuda"))
    lib, h = env._lib, env._h
    buf = torch.zeros(4096, dtype=torch.int32, device="cuda")
    host = np.zeros(64, np.uint32)
    assert lib.rcw_expand_columns(h, host.ctypes.data, 0, 1, 0, buf.data_ptr()) == rcw._capi.RCW_EINVAL   # host pointer
    assert lib.rcw_expand_columns(h, buf.data_ptr(), 0, 1, 3, buf.data_ptr() + 2048) == rcw._capi.RCW_EINVAL  # not a pixel format
    assert lib.rcw_expand_columns(h, buf.data_ptr(), 64, 1, 0, buf.data_ptr() + 2048) == rcw._capi.RCW_EINVAL  # stride < num_rays * 4
    assert lib.rcw_expand_columns(h, buf.data_ptr(), 0, 0, 0, buf.data_ptr() + 2048) == rcw._capi.RCW_ESIZE
    env.close()
    col = rcw.BatchedSingleRoom(4, num_rays=32, height_camera_view_pu=16, obs_format="columns")
    with pytest.raises(ValueError):
        col.obs_tensor_nchw()
    col.close()
