"""GPU parity tests of the top view (update_top_view!, single_room.jl:342-372, 446-483):
the CUDA path through the C ABI versus the golden fixtures and the CPU oracle, pixel for pixel (UInt32 images).
The shapes (line, circle) are SimpleDraw.jl's, which is not vendored: unpinned versus Julia."""
import zlib

import numpy as np
import pytest

from conftest import GOLDEN_CONFIGS
from test_gpu_parity import RCW_KW

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def oracle_top_views(ref, n):
    out = []
    for e in range(n):
        w = ref.world(e)
        w.update_top_view()
        out.append(w.top_view)
    return np.stack(out)


@pytest.mark.parametrize("case", ["A", "B", "C", "D", "T", "U"])
def test_top_view_matches_golden_and_oracle(rcw, oracle, golden, case):
    states, au, goal = golden[f"{case}_states"], golden[f"{case}_au"], golden[f"{case}_goal"]
    n = len(states)
    env = rcw.BatchedSingleRoom(n, auto_reset=False, **RCW_KW[case])
    env.set_state(pos=states, dir_au=au, goal=goal)
    env.render_top_view()
    top = env.copy_top_view()
    w = oracle.World(oracle.default_config(**GOLDEN_CONFIGS[case]))
    assert top.shape[1:] == (w.cfg.W * w.cfg.pu_per_tu, w.cfg.H * w.cfg.pu_per_tu)
    for k in range(n):
        assert zlib.crc32(np.ascontiguousarray(top[k]).tobytes()) == int(golden[f"{case}_top_crc"][k]), k
        w.set_state(states[k, 0], states[k, 1], au[k], goal[k, 0], goal[k, 1])
        w.cast_rays()
        w.update_top_view()
        np.testing.assert_array_equal(top[k], w.top_view)
    if f"{case}_top_image" in golden:
        np.testing.assert_array_equal(top[0], golden[f"{case}_top_image"][0])
    env.close()


def test_single_room_redraws_the_top_view_every_act(rcw, oracle):
    """The reference's act!(env) sequence (single_room.jl:333-340): act, cast, top view, camera view."""
    env = rcw.SingleRoom(seed=11)
    ref = oracle.Batch(1, seed=11, auto_reset=False)
    np.testing.assert_array_equal(env.top_view, oracle_top_views(ref, 1)[0].T)   # after the constructor's reset
    rng = np.random.default_rng(2)
    for t in range(300):
        a = int(rng.integers(1, 5))
        env.act(a)
        ref.step(np.array([a], np.uint8))
        if t % 25 == 24:
            tv = env.top_view
            assert tv.shape == (8 * 32, 16 * 32) and tv.dtype == np.uint32          # Array{UInt32}(H*pu, W*pu), :302
            np.testing.assert_array_equal(tv, oracle_top_views(ref, 1)[0].T)
            np.testing.assert_array_equal(env.camera_view, ref.world(0).camera_view.T)
    launches = env.launch_count()
    env.act(1)
    assert env.launch_count() - launches == 2                                        # frame + top view
    env.close()


@pytest.mark.parametrize("kw,okw", [
    (dict(), dict()),
    (dict(height_tile_map_tu=5, width_tile_map_tu=7, num_directions=36, num_rays=45, height_camera_view_pu=51, pu_per_tu=5),
     dict(H=5, W=7, N=36, R=45, P=51, pu_per_tu=5)),       # 25 x 35 pixels: not a multiple of 8, ragged last sector
    (dict(height_tile_map_tu=12, width_tile_map_tu=9, num_rays=200, height_camera_view_pu=64, pu_per_tu=3),
     dict(H=12, W=9, R=200, P=64, pu_per_tu=3)),
    (dict(height_tile_map_tu=6, width_tile_map_tu=5, num_rays=100, height_camera_view_pu=48, pu_per_tu=24),
     dict(H=6, W=5, R=100, P=48, pu_per_tu=24)),            # one-tile sectors with a tile height that is not a power of two
    (dict(height_tile_map_tu=7, width_tile_map_tu=10, num_rays=128, height_camera_view_pu=48, pu_per_tu=16),
     dict(H=7, W=10, R=128, P=48, pu_per_tu=16)),
])
def test_batched_rollout_with_top_view(rcw, oracle, kw, okw):
    n, seed, steps = 24, 21, 60
    env = rcw.BatchedSingleRoom(n, seed=seed, top_view=True, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(**okw), seed=seed)
    np.testing.assert_array_equal(env.copy_top_view(), oracle_top_views(ref, n))
    env.step_random(steps)
    ref.rollout(steps)
    np.testing.assert_array_equal(env.copy_top_view(), oracle_top_views(ref, n))
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    got = env.top_view_tensor().cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(got, oracle_top_views(ref, n))
    env.close()


def test_top_view_of_custom_and_open_maps(rcw, oracle):
    """Host-supplied wall layers, one per env; an open border lets rays leave the map, so their segments
    end outside the image and are clipped.  A custom palette is honoured."""
    n, H, W, seed = 6, 9, 11, 4
    rng = np.random.default_rng(8)
    walls = np.zeros((n, H, W), bool)
    walls[:, 0, :] = walls[:, -1, :] = walls[:, :, 0] = walls[:, :, -1] = True
    walls[:, 2:-2, 2:-2] |= rng.random((n, H - 4, W - 4)) < 0.25
    walls[1, 0, 3:8] = False                                   # open stretches of the border
    walls[2, 3:6, -1] = False
    pal = [0x112233, 0x445566, 0x778899, 0xAABBCC, 0xDDEEFF, 0x010203]
    kw = dict(height_tile_map_tu=H, width_tile_map_tu=W, num_rays=160, height_camera_view_pu=64, pu_per_tu=16)
    env = rcw.BatchedSingleRoom(n, seed=seed, auto_reset=False, top_palette=pal, **kw)
    cfg = oracle.default_config(H=H, W=W, R=160, P=64, pu_per_tu=16, top_palette=pal)
    pos = np.array([[1.5, 4.5], [1.3, 5.2], [4.4, 9.6], [7.5, 9.5], [2.5, 1.5], [6.2, 3.3]], np.float32)
    au = np.array([0, 64, 32, 100, 17, 90], np.int32)
    goal = np.array([[8, 2]] * n, np.int32)
    for e in range(n):
        walls[e, int(pos[e, 0]), int(pos[e, 1])] = False       # the player stands on a free tile
    env.set_wall_maps(walls)
    env.set_state(pos=pos, dir_au=au, goal=goal)
    env.render_top_view()
    top = env.copy_top_view()
    for e in range(n):
        w = oracle.World(cfg)
        w.set_wall_map(walls[e])
        w.set_state(pos[e, 0], pos[e, 1], au[e], goal[e, 0], goal[e, 1])
        w.cast_rays()
        w.update_top_view()
        np.testing.assert_array_equal(top[e], w.top_view, err_msg=f"env {e}")
    assert set(np.unique(top)) <= set(pal)
    env.close()


def test_top_view_window_and_errors(rcw, oracle):
    n, k, seed = 20, 8, 6
    kw = dict(num_rays=64, height_camera_view_pu=32, pu_per_tu=8)
    env = rcw.BatchedSingleRoom(n, seed=seed, top_view=True, obs_window_envs=k, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=64, P=32, pu_per_tu=8), seed=seed)
    env.step_random(10)
    ref.rollout(10)
    want = oracle_top_views(ref, n)
    np.testing.assert_array_equal(env.copy_top_view(16, 4), want[16:20])     # slots 0..3: the ragged last window
    np.testing.assert_array_equal(env.copy_top_view(12, 4), want[12:16])     # slots 4..7: the window before it
    a = np.full(k, 1, np.uint8)
    env.act_range(a, 8)
    full = np.full(n, 3, np.uint8)
    full[8:16] = a
    # the oracle steps everybody; only the range is compared
    ref.step(full)
    np.testing.assert_array_equal(env.copy_top_view(8, 8), oracle_top_views(ref, n)[8:16])
    env.close()
    plain = rcw.BatchedSingleRoom(2, **kw)
    with pytest.raises(rcw.RcwError):
        plain.copy_top_view()                                                 # nothing drawn yet
    plain.close()
    with pytest.raises(rcw.RcwError) as ei:                                   # 2048 x 2048 pixels: does not fit an SM
        rcw.BatchedSingleRoom(1, height_tile_map_tu=64, width_tile_map_tu=64, top_view=True)
    assert ei.value.code == rcw._capi.RCW_ESIZE
    big = rcw.BatchedSingleRoom(1, height_tile_map_tu=64, width_tile_map_tu=64)
    with pytest.raises(rcw.RcwError):
        big.render_top_view()
    big.close()


@pytest.mark.parametrize("H,W,pu,radius,rays", [
    (4, 6, 32, 0.125, 96),      # 128 rows = 16 sectors per column: a thread keeps its rows (256 % 16 == 0)
    (8, 5, 8, 0.3, 64),         # 64 rows = 8 sectors per column
    (5, 7, 40, 0.45, 130),      # 200 rows = 25 sectors per column (256 % 25 != 0); circle radius 19 px: two bitmap words per column
    (3, 9, 24, 0.4, 40),        # 72 rows = 9 sectors per column, tiles three sectors high
    (9, 4, 64, 0.49, 33),       # 576 rows; circle radius 32 px
])
def test_one_tile_sector_sweeps_and_circle_rewrite(rcw, oracle, H, W, pu, radius, rays):
    """The image sizes whose sectors lie inside one tile (pu and H * pu multiples of 8) take the fast sweeps and
    write the circle afterwards (the sectors under its bounding box a second time).  Borderless maps let the
    player stand at the image's edges: the circle is clipped there."""
    kw = dict(height_tile_map_tu=H, width_tile_map_tu=W, num_rays=rays, height_camera_view_pu=32, pu_per_tu=pu,
              player_radius_wu=radius, num_directions=24)
    cfg = oracle.default_config(H=H, W=W, R=rays, P=32, pu_per_tu=pu, radius=np.float32(radius), N=24)
    pos = np.array([[H / 2, W / 2], [0.05, 0.07], [H - 0.02, W - 0.3], [0.3, W - 0.01], [H - 0.4, 0.2],
                    [H / 2 + 0.01, 0.5], [0.5, W / 2], [0.0, 0.0], [H - 0.001, W - 0.001]], np.float32)
    n = len(pos)
    au = (np.arange(n, dtype=np.int32) * 5) % 24
    goal = np.tile(np.array([[2, 2]], np.int32), (n, 1))
    walls = np.zeros((n, H, W), bool)                           # no walls at all: every ray leaves the map
    walls[0, 0, :] = walls[0, -1, :] = walls[0, :, 0] = walls[0, :, -1] = True   # env 0: the usual closed room
    env = rcw.BatchedSingleRoom(n, seed=3, auto_reset=False, **kw)
    env.set_wall_maps(walls)
    env.set_state(pos=pos, dir_au=au, goal=goal)
    env.render_top_view()
    top = env.copy_top_view()
    for e in range(n):
        w = oracle.World(cfg)
        w.set_wall_map(walls[e])
        w.set_state(pos[e, 0], pos[e, 1], au[e], goal[e, 0], goal[e, 1])
        w.cast_rays()
        w.update_top_view()
        np.testing.assert_array_equal(top[e], w.top_view, err_msg=f"env {e}")
    env.close()


def test_play_drives_the_key_bindings_without_a_window(rcw, oracle):
    """play!(game) (single_room.jl:488-572) headless: W/S/A/D act, R resets, V toggles the view and clears the frame
    buffer, unknown keys are reported, Q stops; the frame buffer holds the current view like
    copy_image_to_frame_buffer! leaves it (utils.jl:64-73)."""
    assert rcw.get_action_keys(None) == ("W", "S", "A", "D")                          # :485
    game = rcw.SingleRoom(seed=5, num_rays=96, height_camera_view_pu=40, pu_per_tu=8)   # camera 40 x 96, top view 64 x 128
    ref = oracle.Batch(1, cfg=oracle.default_config(R=96, P=40, pu_per_tu=8), seed=5, auto_reset=False)
    frames = []
    fb, info = rcw.play(game, "wwdxsaVwwa", on_frame=lambda f, i: frames.append((f.copy(), i)))
    assert fb.shape == (128, 64) and fb.dtype == np.uint32                            # (width, height): max of the two views, :503-506
    assert fb.flags.f_contiguous                                                      # the Julia array's memory: row-major pixels
    assert [i["key"] for _, i in frames] == list("WWDXSAVWWA")
    assert frames[3][1]["warning"] == "No keybinding exists for X" and frames[3][1]["steps_taken"] == 3
    assert info["steps_taken"] == 8 and info["view"] == rcw.TOP_VIEW
    for a in [1, 1, 4, 2, 3]:                                                        # W W D (X) S A
        ref.step(np.array([a], np.uint8))
    w = ref.world(0)
    np.testing.assert_array_equal(frames[5][0][:96, :40], w.camera_view)             # fb[j, i] = image[i, j] (utils.jl:69)
    assert not frames[5][0][96:].any() and not frames[5][0][:, 40:].any()            # the rest of the buffer stays zero
    for a in [1, 1, 3]:
        ref.step(np.array([a], np.uint8))
    w = ref.world(0)
    w.update_top_view()
    np.testing.assert_array_equal(fb, w.top_view)                                     # V: top view fills the buffer
    assert info["reward"] == w.state()["reward"] and bool(info["done"]) == bool(w.state()["done"])
    fb2, info2 = rcw.play(game, ["r", "q", "w"])                                      # R resets the count, Q stops before W
    assert info2["key"] == "R" and info2["steps_taken"] == 0 and info2["view"] == rcw.CAMERA_VIEW
    np.testing.assert_array_equal(fb2[:96, :40], game.camera_view.T)
    with pytest.raises(TypeError):
        rcw.play(rcw.BatchedSingleRoom, "w")
    game.close()
