"""Known answers derived by hand from the reference source, checked against the CPU oracle.
(The reference's own tests pin only `reward == 0 / not done after reset` and `return at
termination == goal_reward`, test/runtests.jl:22-23,33 — both are here too.)"""
import math

import numpy as np
import pytest


def test_direction_table(oracle):
    d = oracle.directions(128)
    assert d[0, 0] == 1.0 and d[0, 1] == 0.0                       # cos 0, sin 0
    assert d[32, 1] == 1.0 and d[32, 0] == np.float32(math.cos(math.pi / 2))   # 6.1e-17, not 0
    assert d[64, 0] == -1.0
    np.testing.assert_allclose(np.hypot(d[:, 0], d[:, 1]), 1.0, atol=1e-6)


def test_philox_known_answer_vectors(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]).tolist() == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_axis_aligned_rays_from_a_tile_centre(oracle):
    """(2.5, 2.5) facing +x in the default 8x16 room: the central rays run along x and stop on the
    bottom wall row i = 8, crossing dimension 1, after 4.5 / ray_x world units."""
    w = oracle.World()
    w.set_state(2.5, 2.5, 0, 7, 15)
    w.cast_rays()
    hit, dim, dist, rd = w.ray_stop, w.ray_dim, w.ray_dist, w.ray_dir
    for i in (255, 256):
        assert hit[i].tolist() == [8, 3] and dim[i] == 1
        assert dist[i] == pytest.approx(4.5 / rd[i, 0], rel=1e-6)
    # first ray = dir + s * (d2, -d1) = (1, -2/3) normalised: heads to smaller j, hits the left wall j = 1
    assert rd[0, 0] == pytest.approx(1 / math.hypot(1, 2 / 3), rel=1e-6)
    assert rd[0, 1] == pytest.approx(-(2 / 3) / math.hypot(1, 2 / 3), rel=1e-6)
    assert hit[0].tolist()[1] == 1 and dim[0] == 2
    assert hit[511].tolist()[0] == 8                       # last ray: long way to the right, ends on i = 8
    # every ray ends on an obstacle tile and the distance is at least the distance to its box
    for i in range(512):
        ih, jh = hit[i]
        assert ih in (1, 8) or jh in (1, 16) or (ih, jh) == (7, 15)


def test_column_rasteriser_rules(oracle):
    """height_line = cam_h * R / (2 s proj); full column when h >= P - 1, else pad = (P - h) // 2,
    painted wall height P - 2 pad in {h, h + 1} (single_room.jl:406-439)."""
    w = oracle.World()
    w.set_state(2.5, 2.5, 0, 7, 15)
    w.cast_rays()
    w.update_camera_view()
    img, h = w.camera_view, w.wall_heights()
    # centre ray: proj = 4.5 -> 512 / (4/3 * 4.5) = 85.33 -> h = 85, pad = 85
    assert h[255] == 85 and h[256] == 85
    col = img[512 - 256]                                   # ray index 256 (1-based) paints column 512-256+1
    assert (col[:85] == 0xFFFFFF).all() and (col[85:171] == 0x808080).all() and (col[171:] == 0x404040).all()
    for i in range(512):
        c = img[511 - i]
        if h[i] >= 255:
            assert len(set(c.tolist())) == 1
        else:
            pad = (256 - h[i]) // 2
            assert (c[:pad] == 0xFFFFFF).all() and (c[256 - pad:] == 0x404040).all()
            assert 256 - 2 * pad in (h[i], h[i] + 1)
            assert len(set(c[pad:256 - pad].tolist())) == 1


def test_goal_collision_gives_reward_and_no_move(oracle):
    """single_room.jl:162-168: goal checked first; reward = goal_reward, done, position unchanged."""
    w = oracle.World()
    w.reset_to(4, 8, 4, 7, 32)              # goal tile (4, 8); player at the centre of (4, 7) facing +y
    s0 = w.state()
    assert s0["reward"] == 0 and not s0["done"]                      # runtests.jl:22-23
    total, steps = 0.0, 0
    while True:
        w.act(1)
        steps += 1
        s = w.state()
        total += s["reward"]
        if s["done"]:
            break
        assert steps < 10
    assert total == 1.0                                               # runtests.jl:33
    # circle radius 1/8, tile edge at y = 7.0: 6.5 -> 6.625 -> 6.75 -> 6.875 moves, the 4th is blocked
    assert steps == 4 and s["pos"].tolist() == [3.5, 6.875]
    w.act(3)                                                          # done is overwritten by the next act
    assert not w.state()["done"] and w.state()["reward"] == 0


def test_wall_collision_is_strict_and_blocks(oracle):
    w = oracle.World()
    w.reset_to(7, 15, 2, 2, 64)             # facing -x towards the top wall row i = 1 (x in [0, 1))
    for _ in range(10):
        w.act(1)
    s = w.state()
    # |x - 1.0| < 1/8 collides (strict <): x = 1.125 is allowed, 1.0 is not
    assert s["pos"].tolist() == [1.125, 1.5] and not s["done"]
    assert not w.is_player_colliding(1, 1.125, 1.5) and w.is_player_colliding(1, 1.124, 1.5)


def test_turning_wraps(oracle):
    w = oracle.World()
    w.reset_to(2, 2, 4, 4, 0)
    w.act(4)
    assert w.state()["au"] == 127
    w.act(3)
    w.act(3)
    assert w.state()["au"] == 1
    assert w.act(0) == -2 and w.act(5) == -2                          # @assert action in 1:4


def test_layout_draws_are_valid_and_deterministic(oracle):
    w = oracle.World()
    seen = set()
    for env in range(200):
        g, p, au = w.draw_layout(7, env, 1)
        assert 2 <= g[0] <= 7 and 2 <= g[1] <= 15 and 0 <= au < 128
        assert 2 <= p[0] <= 7 and 2 <= p[1] <= 15 and tuple(p) != tuple(g)   # empty tile: no wall, no goal
        seen.add((tuple(g), tuple(p), au))
        g2, p2, au2 = w.draw_layout(7, env, 1)
        assert (tuple(g2), tuple(p2), au2) == (tuple(g), tuple(p), au)
    assert len(seen) > 150
    acts = [oracle.draw_action(7, 0, s) for s in range(4000)]
    assert set(acts) == {1, 2, 3, 4} and abs(acts.count(1) - 1000) < 150


def test_batch_is_independent_of_threads_and_shards(oracle):
    a = oracle.Batch(12, seed=3)
    b = oracle.Batch(12, seed=3)
    a.rollout(150, threads=1)
    for _ in range(150):
        b.step(None, threads=3)
    pa, aa, ga = a.states()
    pb, ab, gb = b.states()
    assert np.array_equal(pa, pb) and np.array_equal(aa, ab) and np.array_equal(ga, gb)
    c = oracle.Batch(6, seed=3, env_id_offset=6)
    c.rollout(150)
    pc, ac, gc = c.states()
    assert np.array_equal(pa[6:], pc) and np.array_equal(aa[6:], ac)
    assert np.array_equal(a.obs_rgb8()[6:], c.obs_rgb8())


def test_top_view_hand_derived(oracle):
    """update_top_view! (single_room.jl:342-372, 446-483) from (2.5, 2.5) facing +i (au 0) on the default map:
    player pixel = floor(2.5 * 32) + 1 = 81 (utils.jl:6), radius floor(0.125 * 32) + 1 = 5 pixels."""
    w = oracle.World()
    w.set_state(2.5, 2.5, 0, 6, 14)
    w.cast_rays()
    w.update_top_view()
    top = w.top_view.T                       # [i, j], 0-based here
    WALL, GOAL, EMPTY, BORDER, RAY, PLAYER = 0xFFFFFF, 0xFF0000, 0x000000, 0xCCCCCC, 0x808080, 0xC0C0C0
    assert top.shape == (256, 512)
    assert top[0, 0] == BORDER and top[31, 31] == BORDER and top[1, 1] == WALL          # wall tile (1, 1)
    assert top[5 * 32 + 1, 13 * 32 + 1] == GOAL and top[5 * 32, 13 * 32 + 5] == BORDER  # goal tile (6, 14)
    assert top[3 * 32 + 1, 7 * 32 + 1] == EMPTY
    # the centre ray runs along +i from the player pixel (81, 81) to the wall at x = 7: pixel floor(7 * 32) + 1 = 225
    for i in range(81 + 6, 225 + 1):
        assert top[i - 1, 81 - 1] == RAY, i
    assert top[226 - 1, 81 - 1] in (BORDER, WALL) and top[226 - 1, 81 - 1] != RAY
    # the circle of radius 5 about (81, 81): its four axis points, and the centre is not part of the outline
    for (i, j) in ((86, 81), (76, 81), (81, 86), (81, 76)):
        assert top[i - 1, j - 1] == PLAYER
    assert top[81 - 1, 81 - 1] == RAY          # every ray starts at the player pixel
    # nothing is drawn behind the player: the field of view is +-atan(2/3) about +i
    assert top[60 - 1, 81 - 1] == EMPTY
    assert int((top == PLAYER).sum()) in range(20, 41)
