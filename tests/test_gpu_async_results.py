"""GPU tests of the result ring (rcw_config.result_ring, rcw_step_async / rcw_wait): the step kernel writes every
env's reward and done straight into pinned host memory; each ticket's slot must hold exactly what the oracle
computed for that step, also when later steps have already been enqueued."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def _goal_heavy_batch(rcw, oracle, n, seed, **kw):
    """A 5x5 room (3x3 interior): episodes end every few steps, so rewards / terminations are not all zero."""
    geo = dict(height_tile_map_tu=5, width_tile_map_tu=5, num_directions=8, num_rays=64, height_camera_view_pu=32,
               position_increment_wu=0.3)
    env = rcw.BatchedSingleRoom(n, seed=seed, **geo, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=5, W=5, N=8, R=64, P=32, incr=np.float32(0.3)), seed=seed)
    return env, ref


@pytest.mark.parametrize("depth,lag", [(1, 0), (2, 1), (4, 3), (4, 0)])
@pytest.mark.parametrize("env_kernel", [0, 1])
def test_ring_slots_hold_the_oracles_results(rcw, oracle, monkeypatch, depth, lag, env_kernel):
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n, steps = 300, 60
    env, ref = _goal_heavy_batch(rcw, oracle, n, seed=5, result_ring=depth)
    rng = np.random.default_rng(1)
    expected = {}
    waited = 0
    total_done = 0
    for t in range(steps):
        a = rng.integers(1, 3, n).astype(np.uint8) if t % 2 else rng.integers(1, 5, n).astype(np.uint8)
        ticket = env.act_async(a)
        assert ticket == t
        assert ref.step(a) == 0
        r, d = ref.reward_done()
        expected[t] = (r.copy(), d.copy())
        # read `lag` steps behind the newest ticket: the later steps are already enqueued (or running)
        while waited <= t - lag:
            r_got, d_got = env.wait(waited)
            np.testing.assert_array_equal(r_got, expected[waited][0], err_msg=f"reward of step {waited}")
            np.testing.assert_array_equal(d_got, expected[waited][1], err_msg=f"done of step {waited}")
            total_done += int(d_got.sum())
            assert not r_got.flags.writeable
            waited += 1
    while waited < steps:
        r_got, d_got = env.wait(waited)
        np.testing.assert_array_equal(r_got, expected[waited][0])
        np.testing.assert_array_equal(d_got, expected[waited][1])
        waited += 1
    assert total_done > 20, "the scenario must actually finish episodes"
    # the device copies agree with the newest slot, and the state with the oracle
    s = env.get_state()
    np.testing.assert_array_equal(s["reward"], expected[steps - 1][0])
    np.testing.assert_array_equal(s["done"], expected[steps - 1][1])
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(s["pos"], pos)
    np.testing.assert_array_equal(s["dir_au"], au)
    np.testing.assert_array_equal(s["goal"], goal)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    env.close()


def test_async_and_plain_steps_interleave(rcw, oracle):
    """rcw_step / rcw_step_random between async steps do not consume tickets or touch the ring."""
    n = 64
    env, ref = _goal_heavy_batch(rcw, oracle, n, seed=9, result_ring=2)
    rng = np.random.default_rng(3)
    a0 = rng.integers(1, 5, n).astype(np.uint8)
    t0 = env.act_async(a0)
    ref.step(a0)
    r0 = tuple(x.copy() for x in ref.reward_done())
    for _ in range(5):
        a = rng.integers(1, 3, n).astype(np.uint8)
        env.act(a)
        ref.step(a)
    env.step_random(3)
    ref.rollout(3)
    a1 = rng.integers(1, 3, n).astype(np.uint8)
    t1 = env.act_async(a1)
    ref.step(a1)
    assert (t0, t1) == (0, 1)
    np.testing.assert_array_equal(env.wait(t0)[0], r0[0])
    np.testing.assert_array_equal(env.wait(t0)[1], r0[1])
    r1, d1 = ref.reward_done()
    np.testing.assert_array_equal(env.wait(t1)[0], r1)
    np.testing.assert_array_equal(env.wait(t1)[1], d1)
    env.close()


def test_async_with_observation_window_and_large_batch(rcw, oracle):
    """Windowed launches (several kernels per step) all write the same slot; a batch above the 32,768-env
    parameter capacity takes the staged-action path."""
    n = 1000
    env, ref = _goal_heavy_batch(rcw, oracle, n, seed=2, result_ring=3, obs_window_envs=256)
    rng = np.random.default_rng(8)
    for t in range(12):
        a = rng.integers(1, 3, n).astype(np.uint8)
        tk = env.act_async(a)
        ref.step(a)
        r, d = ref.reward_done()
        r_got, d_got = env.wait(tk)
        np.testing.assert_array_equal(r_got, r)
        np.testing.assert_array_equal(d_got, d)
    env.close()
    n = 40000
    env = rcw.BatchedSingleRoom(n, seed=4, height_tile_map_tu=5, width_tile_map_tu=5, num_directions=8, num_rays=32,
                                height_camera_view_pu=16, position_increment_wu=0.3, result_ring=2)
    ref = oracle.Batch(n, cfg=oracle.default_config(H=5, W=5, N=8, R=32, P=16, incr=np.float32(0.3)), seed=4)
    for t in range(6):
        a = rng.integers(1, 3, n).astype(np.uint8)
        tk = env.act_async(a)
        ref.step(a, threads=8)
        r, d = ref.reward_done()
        r_got, d_got = env.wait(tk)
        np.testing.assert_array_equal(r_got, r)
        np.testing.assert_array_equal(d_got, d)
    assert d.sum() + r.sum() >= 0
    env.close()


def test_device_actions_and_invalid_device_action(rcw, oracle):
    import torch

    n = 128
    env, ref = _goal_heavy_batch(rcw, oracle, n, seed=6, result_ring=2)
    rng = np.random.default_rng(4)
    prev = None
    for t in range(8):
        a = rng.integers(1, 3, n).astype(np.uint8)
        tk = env.act_async(torch.from_numpy(a).cuda())
        ref.step(a)
        prev = tuple(x.copy() for x in ref.reward_done())
        np.testing.assert_array_equal(env.wait(tk)[0], prev[0])
        np.testing.assert_array_equal(env.wait(tk)[1], prev[1])
    # env 7 gets an invalid device-side action: it is not stepped and keeps its previous reward / done
    a = rng.integers(1, 3, n).astype(np.uint8)
    bad = a.copy()
    bad[7] = 9
    tk = env.act_async(torch.from_numpy(bad).cuda())
    r_got, d_got = env.wait(tk)
    assert r_got[7] == prev[0][7] and d_got[7] == prev[1][7]
    with pytest.raises(rcw.InvalidActionError):
        env.sync()
    env.close()


def test_result_ring_errors(rcw):
    geo = dict(num_rays=32, height_camera_view_pu=16)
    env = rcw.BatchedSingleRoom(8, **geo)
    with pytest.raises(rcw.RcwError) as ei:
        env.act_async(np.ones(8, np.uint8))               # no ring configured
    assert ei.value.code == rcw._capi.RCW_EINVAL
    with pytest.raises(rcw.RcwError):
        env.wait(0)
    env.close()
    env = rcw.BatchedSingleRoom(8, result_ring=2, **geo)
    with pytest.raises(rcw.RcwError):
        env.wait(0)                                        # not issued yet
    with pytest.raises(rcw.InvalidActionError):
        env.act_async(np.array([1, 2, 3, 4, 5, 1, 1, 1], np.uint8))
    assert env.act_async(np.ones(8, np.uint8)) == 0        # the failed call consumed no ticket
    assert env.act_async(np.ones(8, np.uint8)) == 1
    env.wait(0)
    assert env.act_async(np.ones(8, np.uint8)) == 2        # reuses the slot of ticket 0
    with pytest.raises(rcw.RcwError) as ei:
        env.wait(0)
    assert "too old" in str(ei.value)
    env.wait(1)
    env.wait(2)
    with pytest.raises(rcw.RcwError):
        env.wait(3)
    env.close()
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(8, result_ring=65, **geo)
