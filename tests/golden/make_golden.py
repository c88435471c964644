#!/usr/bin/env python
"""Generate tests/golden/singleroom_golden.npz.

The reference (/root/reference, pure Julia) cannot run in this image and its DDA lives in the
un-vendored RayCaster.jl 0.1, so these fixtures are NOT outputs of the reference: PARITY IS
UNPINNED versus Julia.  They are produced by a second, independent restatement of the reference
source written in Python with numpy.float32 scalars (one IEEE rounding per operation), so that
the C oracle (oracle/rcw_oracle.c) and the CUDA path are checked against something that was
not derived from either.  Each function cites the reference lines it follows.

Run:  python tests/golden/make_golden.py        (writes the .npz next to this file)
"""
from __future__ import annotations

import math
import os
import zlib

import numpy as np

F = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))

PALETTE = dict(ceiling=0x00FFFFFF, floor=0x00404040, wall1=0x00808080, wall2=0x00C0C0C0,
               goal1=0x00800000, goal2=0x00C00000)  # single_room.jl:291-296


class PyWorld:
    """SingleRoomWorld + camera view, src/single_room.jl:21-108,258-324."""

    def __init__(self, H=8, W=16, N=128, R=512, P=256, radius=1 / 8, incr=1 / 8, sfov=2 / 3, cam_h=1.0,
                 goal_reward=1.0, tie_le=False, dist_post=False, extras=()):
        self.H, self.W, self.N, self.R, self.P = H, W, N, R, P
        self.radius, self.incr, self.sfov, self.cam_h = F(radius), F(incr), F(sfov), F(cam_h)
        self.goal_reward = F(goal_reward)
        self.tie_le, self.dist_post = tie_le, dist_post
        # object layers 3.. (NUM_OBJECTS > 2): dicts {tiles: bool [H+1, W+1] 1-based, terminal: bool,
        # reward, colors: (dim 1, dim 2), top: top-view colour}
        self.extras = list(extras)
        self.wall = np.zeros((H + 1, W + 1), bool)  # 1-based
        self.wall[1:, 1] = True   # :57
        self.wall[1:, W] = True   # :58
        self.wall[1, 1:] = True   # :59
        self.wall[H, 1:] = True   # :60
        # :65-69
        self.dirs = []
        for i in range(1, N + 1):
            theta = (i - 1) * 2 * math.pi / N
            self.dirs.append((F(math.cos(theta)), F(math.sin(theta))))
        self.goal = (2, 2)
        self.pos = (F(1.5), F(1.5))
        self.au = 0
        self.reward = F(0)
        self.done = False

    # ---- utils.jl:5
    @staticmethod
    def wu_to_tu(x):
        return int(math.floor(float(x))) + 1

    def layer(self, which, i, j):
        if i < 1 or i > self.H or j < 1 or j > self.W:
            return False  # new-engine rule for F6 (the reference throws)
        if which == "wall":
            return bool(self.wall[i, j])
        if which == "goal":
            return (i, j) == self.goal
        return bool(self.extras[which]["tiles"][i, j])      # which = index of an extra object layer

    def first_object(self, i, j):
        """findfirst(tile_map[:, i, j]) (:355): 1 wall, 2 goal, 3 + k extra object k, 0 none; outside the map = wall."""
        if i < 1 or i > self.H or j < 1 or j > self.W:
            return 1
        if self.wall[i, j]:
            return 1
        if (i, j) == self.goal:
            return 2
        for k, ex in enumerate(self.extras):
            if ex["tiles"][i, j]:
                return 3 + k
        return 0

    # ---- collision_detection.jl:9-42
    def is_player_colliding(self, which, x, y):
        half = F(0.5)
        ip, jp = self.wu_to_tu(x), self.wu_to_tu(y)
        for j in range(jp - 1, jp + 2):
            for i in range(ip - 1, ip + 2):
                if not self.layer(which, i, j):
                    continue
                px = F(x - F(F(i) - half))
                py = F(y - F(F(j) - half))
                qx = min(max(px, -half), half)
                qy = min(max(py, -half), half)
                vx, vy = F(px - qx), F(py - qy)
                if F(F(vx * vx) + F(vy * vy)) < F(self.radius * self.radius):
                    return True
        return False

    # ---- single_room.jl:139-191
    def act(self, a):
        assert a in (1, 2, 3, 4)
        if a <= 2:
            dx, dy = self.dirs[self.au]
            sx, sy = F(self.incr * dx), F(self.incr * dy)
            if a == 1:
                nx, ny = F(self.pos[0] + sx), F(self.pos[1] + sy)
            else:
                nx, ny = F(self.pos[0] - sx), F(self.pos[1] - sy)
            hit_goal = self.is_player_colliding("goal", nx, ny)
            hit_wall = self.is_player_colliding("wall", nx, ny)
            goal_reward = self.goal_reward
            for k, ex in enumerate(self.extras):       # terminal layers act like GOAL (in object order), the others like WALL
                if self.is_player_colliding(k, nx, ny):
                    if ex["terminal"]:
                        if not hit_goal:
                            hit_goal, goal_reward = True, F(ex["reward"])
                    else:
                        hit_wall = True
            if hit_goal:
                self.reward, self.done = goal_reward, True
            elif hit_wall:
                self.reward, self.done = F(0), False
            else:
                self.pos = (nx, ny)
                self.reward, self.done = F(0), False
        else:
            self.au = (self.au + 1) % self.N if a == 3 else (self.au - 1) % self.N
            self.reward, self.done = F(0), False

    # ---- [EXT] RayCaster.cast_ray, SURVEY.md §8(a) a10
    def cast_ray(self, x, y, dx, dy):
        one = F(1)
        i, j = self.wu_to_tu(x), self.wu_to_tu(y)
        with np.errstate(divide="ignore"):
            ddx, ddy = F(abs(F(one / dx))), F(abs(F(one / dy)))
        if dx < 0:
            sx, tx = -1, F(F(x - F(i - 1)) * ddx)
        else:
            sx, tx = 1, F(F(F(i) - x) * ddx)
        if dy < 0:
            sy, ty = -1, F(F(y - F(j - 1)) * ddy)
        else:
            sy, ty = 1, F(F(F(j) - y) * ddy)
        dim, d = 0, F(0)

        def obstacle(i, j):
            if i < 1 or i > self.H or j < 1 or j > self.W:
                return True
            return self.first_object(i, j) != 0  # any(tile_map, dims=1) :209

        while not obstacle(i, j):
            take_x = (tx <= ty) if self.tie_le else (tx < ty)
            if take_x:
                d, tx, i, dim = tx, F(tx + ddx), i + sx, 1
            else:
                d, ty, j, dim = ty, F(ty + ddy), j + sy, 2
        if self.dist_post and dim:
            d = F(tx - ddx) if dim == 1 else F(ty - ddy)
        return i, j, dim, d

    # ---- single_room.jl:193-231
    def cast_rays(self):
        R, s = self.R, self.sfov
        d0, d1 = self.dirs[self.au]
        c0, c1 = d1, F(-d0)
        f0, f1 = F(d0 + F(s * c0)), F(d1 + F(s * c1))
        l0, l1 = F(d0 - F(s * c0)), F(d1 - F(s * c1))
        lendiv = max(R - 1, 1)
        self.ray_dir, self.hit, self.dim, self.dist = [], [], [], []
        for i in range(1, R + 1):
            t = (i - 1) / lendiv
            u0 = F((1 - t) * float(f0) + t * float(l0))
            u1 = F((1 - t) * float(f1) + t * float(l1))
            n = F(np.sqrt(F(F(u0 * u0) + F(u1 * u1))))
            q = F(F(1) / n)
            r0, r1 = F(q * u0), F(q * u1)
            ih, jh, dim, dist = self.cast_ray(self.pos[0], self.pos[1], r0, r1)
            self.ray_dir.append((r0, r1))
            self.hit.append((ih, jh))
            self.dim.append(dim)
            self.dist.append(dist)

    # ---- single_room.jl:374-444
    def height_line(self, i0):
        d0, d1 = self.dirs[self.au]
        r0, r1 = self.ray_dir[i0]
        dot = F(F(d0 * r0) + F(d1 * r1))
        proj = F(self.dist[i0] * dot)
        num = F(self.cam_h * F(self.R))
        den = F(F(F(2) * self.sfov) * proj)
        with np.errstate(divide="ignore", invalid="ignore"):
            hl = F(num / den)
        if not np.isfinite(hl):
            return self.P
        if hl >= self.P:
            return self.P
        return max(int(math.floor(float(hl))), 0)

    def camera_view(self):
        """uint32 [R columns][P rows]"""
        R, P = self.R, self.P
        img = np.zeros((R, P), np.uint32)
        for i in range(1, R + 1):
            h = self.height_line(i - 1)
            ih, jh = self.hit[i - 1]
            obj = self.first_object(ih, jh)
            if obj <= 1:
                color = PALETTE["wall1"] if self.dim[i - 1] == 1 else PALETTE["wall2"]
            elif obj == 2:
                color = PALETTE["goal1"] if self.dim[i - 1] == 1 else PALETTE["goal2"]
            else:
                color = self.extras[obj - 3]["colors"][0 if self.dim[i - 1] == 1 else 1]
            k = R - i + 1
            col = img[k - 1]
            if h >= P - 1:
                col[:] = color
            else:
                pad = (P - h) // 2
                col[:pad] = PALETTE["ceiling"]
                col[pad:P - pad] = color
                col[P - pad:] = PALETTE["floor"]
        return img


    # ---- single_room.jl:342-372, 446-483 (shapes: [EXT] SimpleDraw.jl 0.3 — Bresenham line, midpoint circle, clipped)
    def top_view(self, pu):
        """uint32 [W*pu columns][H*pu rows] after cast_rays()."""
        H, W = self.H, self.W
        img = np.zeros((H * pu + 2, W * pu + 2), np.uint32)   # 1-based [i, j], one spare row / column

        def put(i, j, c):
            if 1 <= i <= H * pu and 1 <= j <= W * pu:
                img[i, j] = c

        for j in range(1, W + 1):
            for i in range(1, H + 1):
                it, jt = (i - 1) * pu + 1, (j - 1) * pu + 1                      # :350-351
                obj = self.first_object(i, j)                                    # findfirst, :355-360
                if obj == 1:
                    color = 0x00FFFFFF
                elif obj == 2:
                    color = 0x00FF0000
                elif obj >= 3:
                    color = self.extras[obj - 3]["top"]
                else:
                    color = 0x00000000
                img[it:it + pu, jt:jt + pu] = color                              # :353,362
                img[it, jt:jt + pu] = 0x00CCCCCC                                 # :364
                img[it + pu - 1, jt:jt + pu] = 0x00CCCCCC                        # :365
                img[it:it + pu, jt] = 0x00CCCCCC                                 # :366
                img[it:it + pu, jt + pu - 1] = 0x00CCCCCC                        # :367

        def wu_to_pu(x):                                                         # utils.jl:6
            return int(math.floor(float(F(F(x) * F(pu))))) + 1

        ip, jp = wu_to_pu(self.pos[0]), wu_to_pu(self.pos[1])                    # :469
        rp = wu_to_pu(self.radius)                                               # :470
        for k in range(self.R):                                                  # :474-478
            r0, r1 = self.ray_dir[k]
            i2 = wu_to_pu(F(self.pos[0] + F(self.dist[k] * r0)))
            j2 = wu_to_pu(F(self.pos[1] + F(self.dist[k] * r1)))
            i, j = ip, jp
            di, dj = abs(i2 - i), -abs(j2 - j)
            si, sj = (1 if i < i2 else -1), (1 if j < j2 else -1)
            err = di + dj
            while True:
                put(i, j, 0x00808080)
                if (i, j) == (i2, j2):
                    break
                e2 = 2 * err
                if e2 >= dj:
                    err += dj
                    i += si
                if e2 <= di:
                    err += di
                    j += sj
        a, b, d = 0, rp, 1 - rp                                                  # :480
        while a <= b:
            for (u, v) in ((a, b), (b, a)):
                for su in (1, -1):
                    for sv in (1, -1):
                        put(ip + su * u, jp + sv * v, 0x00C0C0C0)
            if d < 0:
                d += 2 * a + 3
            else:
                d += 2 * (a - b) + 5
                b -= 1
            a += 1
        return np.ascontiguousarray(img[1:H * pu + 1, 1:W * pu + 1].T)


def random_walk_states(w: PyWorld, rng, n_states, steps_between):
    """Reachable states: start at a tile centre, take random actions (never entering the goal)."""
    states = []
    while len(states) < n_states:
        gi, gj = int(rng.integers(2, w.H)), int(rng.integers(2, w.W))
        while True:
            pi, pj = int(rng.integers(2, w.H)), int(rng.integers(2, w.W))
            if (pi, pj) != (gi, gj):
                break
        w.goal = (gi, gj)
        w.pos = (F(pi - 0.5), F(pj - 0.5))
        w.au = int(rng.integers(0, w.N))
        for _ in range(int(rng.integers(0, steps_between))):
            w.act(int(rng.integers(1, 5)))
        states.append((w.pos[0], w.pos[1], w.au, gi, gj))
    return states


def cast_case(w: PyWorld, states, full_images=0, pu=32):
    out = dict(states=np.array([[s[0], s[1]] for s in states], np.float32),
               au=np.array([s[2] for s in states], np.int32),
               goal=np.array([[s[3], s[4]] for s in states], np.int32))
    hits, dims, dists, rdirs, heights, crcs, imgs, top_crcs, tops = [], [], [], [], [], [], [], [], []
    for k, s in enumerate(states):
        w.pos, w.au, w.goal = (F(s[0]), F(s[1])), int(s[2]), (int(s[3]), int(s[4]))
        w.cast_rays()
        img = w.camera_view()
        hits.append(w.hit)
        dims.append(w.dim)
        dists.append(w.dist)
        rdirs.append(w.ray_dir)
        heights.append([w.height_line(i) for i in range(w.R)])
        crcs.append(zlib.crc32(img.tobytes()))
        top = w.top_view(pu)
        top_crcs.append(zlib.crc32(top.tobytes()))
        if k < full_images:
            imgs.append(img)
            if top.size <= 1 << 16:
                tops.append(top)
    out.update(hit=np.array(hits, np.int32), dim=np.array(dims, np.int32),
               dist=np.array(dists, np.float32), ray_dir=np.array(rdirs, np.float32),
               height=np.array(heights, np.int32), crc=np.array(crcs, np.uint32))
    out["top_crc"] = np.array(top_crcs, np.uint32)
    if imgs:
        out["image"] = np.array(imgs, np.uint32)
    if tops:
        out["top_image"] = np.array(tops, np.uint32)
    return out


def act_case(w: PyWorld, rng, n_episodes, n_steps):
    """Trajectories: (initial layout, action list) -> per-step pos/au/reward/done."""
    init, actions, pos, au, rew, done = [], [], [], [], [], []
    for _ in range(n_episodes):
        gi, gj = int(rng.integers(2, w.H)), int(rng.integers(2, w.W))
        while True:
            pi, pj = int(rng.integers(2, w.H)), int(rng.integers(2, w.W))
            if (pi, pj) != (gi, gj):
                break
        a0 = int(rng.integers(0, w.N))
        w.goal, w.pos, w.au = (gi, gj), (F(pi - 0.5), F(pj - 0.5)), a0
        init.append((gi, gj, pi, pj, a0))
        # biased towards moving so walls and goals get hit
        acts = rng.choice([1, 1, 1, 2, 3, 4], size=n_steps)
        tp, ta, tr, td = [], [], [], []
        for a in acts:
            w.act(int(a))
            tp.append((w.pos[0], w.pos[1]))
            ta.append(w.au)
            tr.append(w.reward)
            td.append(w.done)
        actions.append(acts)
        pos.append(tp)
        au.append(ta)
        rew.append(tr)
        done.append(td)
    return dict(init=np.array(init, np.int32), actions=np.array(actions, np.uint8),
                pos=np.array(pos, np.float32), au=np.array(au, np.int32),
                reward=np.array(rew, np.float32), done=np.array(done, np.uint8))


def main():
    rng = np.random.default_rng(20261018)
    out = {}

    # A: reference default geometry (8x16, N=128, R=512, P=256)
    w = PyWorld()
    st = random_walk_states(w, rng, 24, 400)
    # hand-picked: tile centre facing +x (au 0), facing the goal, hugging a corner
    st += [(F(2.5), F(2.5), 0, 6, 14), (F(3.5), F(7.5), 32, 3, 9), (F(1.125), F(1.125), 80, 7, 15),
           (F(6.875), F(14.875), 16, 2, 2), (F(4.5), F(8.5), 64, 4, 10), (F(4.5), F(8.5), 96, 4, 7)]
    for k, v in cast_case(w, st, full_images=2).items():
        out["A_" + k] = v
    for k, v in act_case(w, rng, 12, 600).items():
        out["A_act_" + k] = v

    # B: config 5 geometry (64x64, N=256), fewer rays to keep the file small
    w = PyWorld(H=64, W=64, N=256, R=128, P=96)
    st = random_walk_states(w, rng, 12, 300)
    for k, v in cast_case(w, st, pu=4).items():
        out["B_" + k] = v
    for k, v in act_case(w, rng, 4, 400).items():
        out["B_act_" + k] = v

    # C: odd sizes (R not a multiple of 32, P odd, other radius / increment / fov)
    w = PyWorld(H=5, W=7, N=36, R=45, P=51, radius=0.2, incr=0.3, sfov=0.5, cam_h=0.8)
    st = random_walk_states(w, rng, 10, 100)
    for k, v in cast_case(w, st, full_images=1, pu=7).items():
        out["C_" + k] = v
    for k, v in act_case(w, rng, 4, 300).items():
        out["C_act_" + k] = v

    # D: the other DDA contract (D1 tie <=, D2 post-loop distance) on the default geometry
    w = PyWorld(tie_le=True, dist_post=True)
    st = random_walk_states(w, rng, 6, 400)
    for k, v in cast_case(w, st).items():
        out["D_" + k] = v

    # T / U: exact ties between the two side distances — tile centres and tile corners, 8 directions (so the diagonal
    # directions have two equal components) and an odd ray count (so the central ray IS the player's direction):
    # the states that separate decision D1 (T: advance along dimension 1 only when side_x < side_y; U: when <=).
    # No random numbers are drawn here, so cases A-D above stay byte-identical.
    st = [(F(3.5), F(3.5), 0, 5, 5), (F(3.5), F(3.5), 1, 2, 2), (F(3.5), F(3.5), 2, 4, 6), (F(3.5), F(3.5), 3, 6, 2),
          (F(3.0), F(3.0), 1, 5, 5), (F(2.0), F(4.0), 5, 4, 2), (F(3.5), F(3.0), 7, 2, 6), (F(1.0), F(1.0), 1, 3, 3),
          (F(2.5), F(4.5), 5, 6, 6), (F(4.5), F(2.5), 3, 2, 5), (F(3.5), F(3.5), 1, 5, 5), (F(3.5), F(3.5), 5, 2, 2)]
    for name, tie in (("T", False), ("U", True)):
        w = PyWorld(H=7, W=7, N=8, R=33, P=40, tie_le=tie)
        for k, v in cast_case(w, st, full_images=len(st), pu=4).items():
            out[name + "_" + k] = v

    # L: NUM_OBJECTS = 5: an interior wall, blocking pillars (object 3), a terminal layer with a
    # negative reward (object 4) and a terminal layer with reward 0.5 (object 5) on a 9 x 12 map.  A fresh generator,
    # so the cases above stay byte-identical.
    rng_l = np.random.default_rng(20261019)
    H, W = 9, 12
    tiles = [np.zeros((H + 1, W + 1), bool) for _ in range(3)]
    for (i, j) in [(3, 4), (6, 8), (4, 9), (7, 3)]:
        tiles[0][i, j] = True
    for (i, j) in [(5, 5), (2, 10), (8, 6)]:
        tiles[1][i, j] = True
    for (i, j) in [(6, 2), (3, 7), (5, 5)]:          # (5, 5) carries objects 4 and 5: findfirst shows 4
        tiles[2][i, j] = True
    w = PyWorld(H=H, W=W, N=32, R=64, P=48, radius=0.15, incr=0.2,
                extras=[dict(tiles=tiles[0], terminal=False, reward=0.0, colors=(0x00205080, 0x003070A0), top=0x000000FF),
                        dict(tiles=tiles[1], terminal=True, reward=-1.0, colors=(0x00A04000, 0x00C06000), top=0x00FF8000),
                        dict(tiles=tiles[2], terminal=True, reward=0.5, colors=(0x0000A040, 0x0000C060), top=0x0000FF00)])
    w.wall[4, 5:8] = True                             # an interior wall segment
    out["L_wall"] = w.wall[1:, 1:].copy()
    out["L_extra"] = np.stack([t[1:, 1:] for t in tiles])
    free = [(i, j) for i in range(2, H) for j in range(2, W) if w.first_object(i, j) == 0 or (i, j) == w.goal]
    st = []
    while len(st) < 16:
        gi, gj = int(rng_l.integers(2, H)), int(rng_l.integers(2, W))
        w.goal = (gi, gj)
        pi, pj = free[int(rng_l.integers(0, len(free)))]
        if w.first_object(pi, pj) != 0:
            continue
        w.pos, w.au = (F(pi - 0.5), F(pj - 0.5)), int(rng_l.integers(0, w.N))
        for _ in range(int(rng_l.integers(0, 60))):
            w.act(int(rng_l.choice([1, 1, 2, 3, 4])))
            if w.done:
                break
        st.append((w.pos[0], w.pos[1], w.au, gi, gj))
    for k, v in cast_case(w, st, full_images=len(st), pu=4).items():
        out["L_" + k] = v
    # act! trajectories that run into every kind of object
    init, actions, pos, au, rew, done = [], [], [], [], [], []
    while len(init) < 24:
        gi, gj = int(rng_l.integers(2, H)), int(rng_l.integers(2, W))
        pi, pj = free[int(rng_l.integers(0, len(free)))]
        w.goal = (gi, gj)
        if w.first_object(pi, pj) != 0:
            continue
        a0 = int(rng_l.integers(0, w.N))
        w.pos, w.au = (F(pi - 0.5), F(pj - 0.5)), a0
        acts = rng_l.choice([1, 1, 1, 1, 2, 3, 4], size=150)
        tp, ta, tr, td = [], [], [], []
        for a in acts:
            w.act(int(a))
            tp.append((w.pos[0], w.pos[1])), ta.append(w.au), tr.append(w.reward), td.append(w.done)
        init.append((gi, gj, pi, pj, a0)), actions.append(acts), pos.append(tp), au.append(ta), rew.append(tr), done.append(td)
    out["L_act_init"], out["L_act_actions"] = np.array(init, np.int32), np.array(actions, np.uint8)
    out["L_act_pos"], out["L_act_au"] = np.array(pos, np.float32), np.array(au, np.int32)
    out["L_act_reward"], out["L_act_done"] = np.array(rew, np.float32), np.array(done, np.uint8)

    path = os.path.join(HERE, "singleroom_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
