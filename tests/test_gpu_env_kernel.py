"""GPU parity tests of env_kernel (narrow cameras, one warp per env): the same comparisons against the oracle as
the item kernel, with the kernel forced on small batches (RCW_ENV_PER_WARP_MIN=1), plus a batch large enough to
take it by default, plus a check that both kernels produce identical bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


@pytest.fixture()
def forced(monkeypatch):
    monkeypatch.setenv("RCW_ENV_PER_WARP", "1")
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_equals_oracle(env, ref, fmt="rgb8"):
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
    assert env.episode_stats() == ref.episode_stats()


@pytest.mark.parametrize("R,P,fmt", [(84, 84, "rgb8"), (84, 84, "gray8"), (64, 64, "rgb8"), (128, 128, "xrgb32"),
                                     (20, 33, "rgb8"), (45, 51, "xrgb32"), (100, 70, "gray8"), (97, 64, "rgb8")])
def test_random_rollout_matches_oracle(rcw, oracle, forced, R, P, fmt):
    n, seed, steps = 37, 100 + R, 120                       # 37 envs: the last CTA is ragged
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format=fmt, num_rays=R, height_camera_view_pu=P)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=R, P=P), seed=seed)
    assert_equals_oracle(env, ref, fmt)                     # the constructor's reset + render
    env.step_random(steps)
    ref.rollout(steps)
    assert_equals_oracle(env, ref, fmt)
    rng = np.random.default_rng(R)
    for _ in range(10):                                     # explicit actions, biased towards moving forward
        a = rng.choice([1, 1, 1, 2, 3, 4], size=n).astype(np.uint8)
        env.act(a)
        assert ref.step(a) == 0
    assert_equals_oracle(env, ref, fmt)
    rays = env.get_rays()                                   # the dump path stays on the item kernel
    for e in (0, n - 1):
        np.testing.assert_array_equal(rays["hit"][e], ref.world(e).ray_stop)
    env.close()


def test_goal_seeking_episodes_with_auto_reset(rcw, oracle, forced):
    """Long rollout with many auto-resets on a small map (goals are hit often)."""
    n, seed, steps = 64, 5, 1500
    kw = dict(height_tile_map_tu=5, width_tile_map_tu=6, num_directions=32, num_rays=96, height_camera_view_pu=40)
    env = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=5, W=6, N=32, R=96, P=40), seed=seed)
    env.step_random(steps)
    ref.rollout(steps, threads=4)
    assert ref.episode_stats()[0] > n
    assert_equals_oracle(env, ref)
    env.close()


def test_per_env_maps_custom_palette_and_window(rcw, oracle, forced):
    n, H, W, R, P, seed = 24, 9, 11, 90, 60, 13
    rng = np.random.default_rng(4)
    walls = np.zeros((n, H, W), bool)
    walls[:, 0, :] = walls[:, -1, :] = walls[:, :, 0] = walls[:, :, -1] = True
    walls[:, 2:-2, 2:-2] |= rng.random((n, H - 4, W - 4)) < 0.2
    walls[3, 0, 2:6] = False                                  # an open border: the bounds-checked DDA variant
    pal = [0x102030, 0x405060, 0x708090, 0xA0B0C0, 0xD0E0F0, 0x112233]   # no byte-replicated colour: phase-rotated path
    kw = dict(height_tile_map_tu=H, width_tile_map_tu=W, num_rays=R, height_camera_view_pu=P)
    env = rcw.BatchedSingleRoom(n, seed=seed, auto_reset=False, palette=pal, obs_window_envs=16, **kw)
    env.set_wall_maps(walls)
    cfg = oracle.default_config(H=H, W=W, R=R, P=P, palette=pal)
    pos = np.stack([rng.uniform(1.2, H - 1.2, n), rng.uniform(1.2, W - 1.2, n)], 1).astype(np.float32)
    au = rng.integers(0, 128, n).astype(np.int32)
    goal = np.array([[H - 1, W - 1]] * n, np.int32)
    for e in range(n):
        walls[e, int(pos[e, 0]), int(pos[e, 1])] = False
        walls[e, H - 2, W - 2] = False
    env.set_wall_maps(walls)
    env.set_state(pos=pos, dir_au=au, goal=goal)
    worlds = []
    for e in range(n):
        w = oracle.World(cfg)
        w.set_wall_map(walls[e])
        w.set_state(pos[e, 0], pos[e, 1], au[e], goal[e, 0], goal[e, 1])
        worlds.append(w)
    for t in range(6):
        a = rng.integers(1, 5, n).astype(np.uint8)
        for w0 in (0, 16):                                    # window by window, like a learner
            k = min(16, n - w0)
            env.act_range(a[w0:w0 + k], w0)
            got = env.copy_obs(w0, k)
            for e in range(w0, w0 + k):
                assert worlds[e].step(int(a[e])) == 0
                np.testing.assert_array_equal(got[e - w0], worlds[e].obs_rgb8(), err_msg=f"step {t} env {e}")
    st = env.get_state()
    for e in range(n):
        assert bits(st["pos"][e]).tolist() == bits(worlds[e].state()["pos"]).tolist()
    env.close()


def test_both_kernels_write_identical_bytes_and_default_selection(rcw, oracle, monkeypatch):
    """4096 envs at 84 x 84 take env_kernel by default; RCW_ENV_PER_WARP=0 forces the item kernel."""
    n, seed, steps = 4096, 3, 40
    kw = dict(num_rays=84, height_camera_view_pu=84, obs_format="gray8")
    a = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    monkeypatch.setenv("RCW_ENV_PER_WARP", "0")
    b = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    a.step_random(steps)
    b.step_random(steps)
    np.testing.assert_array_equal(a.copy_obs(), b.copy_obs())
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k])
    ref = oracle.Batch(8, cfg=oracle.default_config(R=84, P=84), seed=seed, env_id_offset=2000)
    ref.rollout(steps)
    np.testing.assert_array_equal(a.copy_obs(2000, 8), ref.obs_gray8())
    a.close()
    b.close()


@pytest.mark.parametrize("n", [32768, 32769, 40000])
def test_host_actions_around_the_parameter_capacity(rcw, oracle, n):
    """Host action arrays ride in the kernel parameters up to 32,768 envs per launch (2 bits each); larger
    launches stage them through a device copy.  Both sides of the limit against the oracle, sampled."""
    seed, steps = 17, 6
    kw = dict(num_rays=32, height_camera_view_pu=16, obs_format="gray8")
    env = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    rng = np.random.default_rng(n)
    actions = rng.choice([1, 1, 2, 3, 4], size=(steps, n)).astype(np.uint8)
    for t in range(steps):
        env.act(actions[t])
    st = env.get_state()
    for s0 in (0, 16380, 32760, n - 8):
        ref = oracle.Batch(8, cfg=oracle.default_config(R=32, P=16), seed=seed, env_id_offset=s0)
        for t in range(steps):
            assert ref.step(actions[t, s0:s0 + 8]) == 0
        pos, au, goal = ref.states()
        np.testing.assert_array_equal(bits(st["pos"][s0:s0 + 8]), bits(pos))
        np.testing.assert_array_equal(st["dir_au"][s0:s0 + 8], au)
        np.testing.assert_array_equal(env.copy_obs(s0, 8), ref.obs_gray8())
    bad = actions[0].copy()
    bad[n - 1] = 5
    with pytest.raises(AssertionError):
        env.act(bad)
    env.close()


@pytest.mark.parametrize("env_kernel", [0, 1])
def test_masked_reset_redraws_only_the_reset_envs(rcw, oracle, monkeypatch, env_kernel):
    """rcw_reset with a mask redraws the envs it reset and leaves the other observations (and top views)
    byte for byte as they were; it does not block and reuses its device scratch from call to call."""
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n = 50
    kw = dict(num_rays=96, height_camera_view_pu=40, top_view=True, pu_per_tu=4, auto_reset=False)
    env = rcw.BatchedSingleRoom(n, seed=8, **kw)
    cfg = oracle.default_config(R=96, P=40, pu_per_tu=4)
    env.step_random(30)
    rng = np.random.default_rng(0)
    for _ in range(4):
        before_obs, before_top, before = env.copy_obs(), env.copy_top_view(), env.get_state()
        mask = (rng.random(n) < 0.3).astype(np.uint8)
        g = np.stack([rng.integers(2, 8, n), rng.integers(2, 16, n)], 1).astype(np.int32)
        p = np.stack([rng.integers(2, 8, n), rng.integers(2, 16, n)], 1).astype(np.int32)
        p[(p == g).all(1)] = [2, 2]
        g[(p == g).all(1)] = [3, 3]
        a = rng.integers(0, 128, n).astype(np.int32)
        env.reset(g, p, a, mask)
        obs, top, st = env.copy_obs(), env.copy_top_view(), env.get_state()
        keep = mask == 0
        np.testing.assert_array_equal(obs[keep], before_obs[keep])
        np.testing.assert_array_equal(top[keep], before_top[keep])
        np.testing.assert_array_equal(bits(st["pos"][keep]), bits(before["pos"][keep]))
        for e in np.nonzero(mask)[0]:
            w = oracle.World(cfg)
            w.reset_to(g[e, 0], g[e, 1], p[e, 0], p[e, 1], a[e])
            w.cast_rays()
            w.update_camera_view()
            w.update_top_view()
            np.testing.assert_array_equal(obs[e], w.obs_rgb8())
            np.testing.assert_array_equal(top[e], w.top_view)
            assert st["reward"][e] == 0 and st["done"][e] == 0
        env.step_random(5)
    # a mask of zeros redraws nothing; a Philox reset under a mask draws new layouts only there
    before_obs = env.copy_obs()
    env.reset(mask=np.zeros(n, np.uint8))
    np.testing.assert_array_equal(env.copy_obs(), before_obs)
    mask = np.zeros(n, np.uint8)
    mask[7] = mask[49] = 1
    env.reset(mask=mask)
    after = env.copy_obs()
    np.testing.assert_array_equal(after[mask == 0], before_obs[mask == 0])
    env.close()
