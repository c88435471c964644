"""Property tests of the CPU oracle (hypothesis): invariants any correct grid DDA / column
rasteriser must satisfy, independent of the unpinned RayCaster.jl details."""
import math

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

F = np.float32


def make_world(oracle, H, W, R, P, N):
    return oracle.World(oracle.default_config(H=H, W=W, R=R, P=P, N=N))


states = st.tuples(st.integers(2, 7), st.integers(2, 15), st.floats(0.13, 0.87), st.floats(0.13, 0.87),
                   st.integers(0, 127), st.integers(2, 7), st.integers(2, 15))


@settings(max_examples=60, deadline=None)
@given(states)
def test_every_ray_ends_on_the_first_obstacle_along_it(oracle, s):
    pi, pj, fx, fy, au, gi, gj = s
    if (pi, pj) == (gi, gj):
        return
    w = make_world(oracle, 8, 16, 64, 32, 128)
    x, y = F(pi - 1 + fx), F(pj - 1 + fy)
    w.set_state(x, y, au, gi, gj)
    w.cast_rays()
    hit, dim, dist, rd = w.ray_stop, w.ray_dim, w.ray_dist, w.ray_dir
    for k in range(64):
        i, j = int(hit[k, 0]), int(hit[k, 1])
        # the hit tile is an obstacle: border wall or the goal
        assert i in (1, 8) or j in (1, 16) or (i, j) == (gi, gj)
        assert dim[k] in (1, 2) and dist[k] > 0
        # the crossing point pos + dist * ray lies on the face of the hit tile that the ray enters
        px, py = float(x) + float(dist[k]) * float(rd[k, 0]), float(y) + float(dist[k]) * float(rd[k, 1])
        if dim[k] == 1:
            face = i - 1 if rd[k, 0] > 0 else i
            assert abs(px - face) < 1e-4 and j - 1 - 1e-4 <= py <= j + 1e-4
        else:
            face = j - 1 if rd[k, 1] > 0 else j
            assert abs(py - face) < 1e-4 and i - 1 - 1e-4 <= px <= i + 1e-4
        # no obstacle strictly before the crossing: sample the segment
        for t in np.linspace(0.0, float(dist[k]) * 0.999, 25):
            qi = math.floor(float(x) + t * float(rd[k, 0])) + 1
            qj = math.floor(float(y) + t * float(rd[k, 1])) + 1
            assert not (qi in (1, 8) or qj in (1, 16) or (qi, qj) == (gi, gj))
        assert abs(math.hypot(float(rd[k, 0]), float(rd[k, 1])) - 1.0) < 1e-6


@settings(max_examples=40, deadline=None)
@given(states, st.sampled_from([(32, 24), (45, 51), (64, 84)]))
def test_columns_are_ceiling_wall_floor_and_symmetric(oracle, s, geom):
    pi, pj, fx, fy, au, gi, gj = s
    if (pi, pj) == (gi, gj):
        return
    R, P = geom
    w = make_world(oracle, 8, 16, R, P, 128)
    w.set_state(F(pi - 1 + fx), F(pj - 1 + fy), au, gi, gj)
    w.cast_rays()
    w.update_camera_view()
    img, h = w.camera_view, w.wall_heights()
    for i in range(R):
        col = img[R - 1 - i]
        if h[i] >= P - 1:
            assert len(set(col.tolist())) == 1
            continue
        pad = (P - h[i]) // 2
        assert pad >= 1
        assert (col[:pad] == 0xFFFFFF).all() and (col[P - pad:] == 0x404040).all()
        body = set(col[pad:P - pad].tolist())
        assert len(body) == 1 and body <= {0x808080, 0xC0C0C0, 0x800000, 0xC00000}
        assert P - 2 * pad in (h[i], h[i] + 1)
    # nearer walls are taller: height is non-increasing in the perpendicular distance
    d = w.ray_dist * (w.ray_dir @ np.array(oracle.directions(128)[au], np.float32))
    order = np.argsort(d, kind="stable")
    assert (np.diff(h[order]) <= 0).all()


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 7), st.integers(2, 15), st.integers(0, 127), st.lists(st.integers(1, 4), min_size=1, max_size=80))
def test_player_never_enters_an_obstacle_and_reward_only_at_the_goal(oracle, pi, pj, au, actions):
    w = make_world(oracle, 8, 16, 32, 16, 128)
    gi, gj = (4, 8) if (pi, pj) != (4, 8) else (5, 9)
    w.reset_to(gi, gj, pi, pj, au)
    for a in actions:
        before = w.state()
        assert w.act(a) == 0
        s = w.state()
        x, y = float(s["pos"][0]), float(s["pos"][1])
        # the circle of radius 1/8 stays clear of the border walls and of the goal tile
        assert 1.0 + 0.125 <= x <= 7.0 - 0.125 and 1.0 + 0.125 <= y <= 15.0 - 0.125
        assert not w.is_player_colliding(1, x, y) and not w.is_player_colliding(2, x, y)
        assert s["reward"] in (0.0, 1.0) and (s["reward"] == 1.0) == s["done"]
        if s["done"]:
            assert np.array_equal(s["pos"], before["pos"])      # goal hit: no move (single_room.jl:166-168)
        if a >= 3:
            assert np.array_equal(s["pos"], before["pos"]) and (s["au"] - before["au"]) % 128 in (1, 127)
