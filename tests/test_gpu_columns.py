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


@pytest.mark.parametrize("fmt", ["rgb8", "xrgb32", "gray8"])
@pytest.mark.parametrize("shape", [(64, 32), (45, 51), (512, 256)])
def test_expand_clamps_garbage_words_and_reports_them(rcw, fmt, shape):
    """Words from a caller's replay buffer may be stale or uninitialised: a palette index outside 2..5 or a pad
    above height_px / 2 is clamped (nothing is written outside the word's own column: guard bytes around the
    destination stay intact) and the next blocking call reports RCW_EINVAL once."""
    import ctypes as C

    import torch

    R, P = shape
    n = 5
    env = rcw.BatchedSingleRoom(2, num_rays=R, height_camera_view_pu=P)
    rng = np.random.default_rng(R * 1000 + P)
    words = rng.integers(0, 2 ** 32, size=(n, R), dtype=np.uint64).astype(np.uint32)
    words[0, 0] = 0xFFFFFFFF
    words[1, R - 1] = (200 << 16) | 0xFFFF
    words[2, 1] = (3 << 16) | (P // 2 + 1)                      # valid colour, pad one row too many
    es, cs, cb = C.c_size_t(), C.c_size_t(), C.c_size_t()
    code = {"rgb8": 0, "xrgb32": 1, "gray8": 2}[fmt]
    assert env._lib.rcw_expanded_layout(env._h, code, C.byref(es), C.byref(cs), C.byref(cb)) == 0
    guard = 1 << 20
    buf = torch.full((guard + n * es.value + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    src = torch.from_numpy(words.view(np.int32)).cuda()
    dst = buf.data_ptr() + guard
    assert dst % 32 == 0
    assert env._lib.rcw_expand_columns(env._h, src.data_ptr(), 0, n, code, dst) == 0
    rc = env._lib.rcw_sync(env._h)
    assert rc == rcw._capi.RCW_EINVAL, env._lib.rcw_last_error()
    assert env._lib.rcw_sync(env._h) == 0                      # reported once
    host = buf.cpu().numpy()
    assert (host[:guard] == 0xA5).all() and (host[guard + n * es.value:] == 0xA5).all()
    # the picture is that of the clamped words
    cid = np.clip(words >> 16, 2, 5)
    pad = np.minimum(words & 0xFFFF, P // 2)
    clamped = (pad | (cid << 16)).astype(np.uint32)
    src2 = torch.from_numpy(clamped.view(np.int32)).cuda()
    want = env.expand_columns(src2, pixel_format=fmt)
    assert env._lib.rcw_sync(env._h) == 0                      # valid words: no report
    bpp = {"rgb8": 3, "xrgb32": 4, "gray8": 1}[fmt]
    got = torch.as_strided(buf[guard:guard + n * es.value], (n, R, P * bpp), (es.value, cs.value, 1))
    want_b = want.contiguous().view(torch.uint8).reshape(n, R, P * bpp)
    assert torch.equal(got, want_b)
    env.close()
