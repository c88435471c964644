"""GPU tests of the observation window (rcw_config.obs_window_envs) and of rcw_step_range: batches whose
observations exceed HBM (BASELINE.json configs[2]: 2^20 default-camera envs = 412 GB) are rendered window
by window into K env slots, env e in slot e mod K.  State, reward and termination must not depend on the
window; the observations a window holds must equal the oracle's for the envs rendered into it last.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SMALL = dict(num_rays=96, height_camera_view_pu=64)
SMALL_ORC = dict(R=96, P=64)


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _assert_state_equals_oracle(env, ref, lo=0, hi=None):
    hi = env.num_envs if hi is None else hi
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"][lo:hi]), bits(pos))
    np.testing.assert_array_equal(st["dir_au"][lo:hi], au)
    np.testing.assert_array_equal(st["goal"][lo:hi], goal)
    r, d = ref.reward_done()
    np.testing.assert_array_equal(st["reward"][lo:hi], r)
    np.testing.assert_array_equal(st["done"][lo:hi], d)


@pytest.mark.parametrize("fmt", ["rgb8", "xrgb32"])
def test_windowed_random_rollout_matches_oracle(rcw, oracle, fmt):
    """100 envs in a window of 32 slots (a ragged last window of 4): 40 random-policy steps with
    auto-reset.  After a step, slots 0..3 hold envs 96..99 and slots 4..31 envs 68..95."""
    n, k, seed, steps = 100, 32, 77, 40
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_format=fmt, obs_window_envs=k, **SMALL)
    assert env.obs_shape[0] == k and env.obs_device_ptr()[1] == k * env.obs_layout()[0]
    ref = oracle.Batch(n, cfg=oracle.default_config(**SMALL_ORC), seed=seed)
    launches0 = env.launch_count()
    env.step_random(steps)
    assert env.launch_count() - launches0 == steps * 4          # one launch per window
    ref.rollout(steps)
    _assert_state_equals_oracle(env, ref)
    want = ref.obs_rgb8() if fmt == "rgb8" else ref.obs_u32()
    np.testing.assert_array_equal(env.copy_obs(96, 4), want[96:100])
    np.testing.assert_array_equal(env.copy_obs(68, 28), want[68:96])
    # the zero-copy device view shows the slots themselves
    import torch
    t = env.obs_tensor()
    assert tuple(t.shape)[0] == k
    got = t.cpu().numpy()
    if fmt == "xrgb32":
        got = got.view(np.uint32)
    np.testing.assert_array_equal(got[:4], want[96:100])
    np.testing.assert_array_equal(got[4:], want[68:96])
    assert env.episode_stats() == ref.episode_stats()
    env.close()
    del torch


def test_step_range_walks_a_batch_window_by_window(rcw, oracle):
    """A learner's loop over a batch that does not fit: step one window with rcw_step_range, read its
    observations, step the next.  Every window's observations, and the state of the whole batch after
    the sweep, equal the oracle stepped with the same actions."""
    n, k, seed, steps = 96, 32, 5, 25
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_window_envs=k, **SMALL)
    ref = oracle.Batch(n, cfg=oracle.default_config(**SMALL_ORC), seed=seed)
    rng = np.random.default_rng(3)
    for _ in range(steps):
        actions = rng.integers(1, 5, n).astype(np.uint8)
        got = []
        for w0 in range(0, n, k):
            env.act_range(actions[w0:w0 + k], w0)
            got.append(env.copy_obs(w0, k))
        assert ref.step(actions) == 0
        np.testing.assert_array_equal(np.concatenate(got), ref.obs_rgb8())
        _assert_state_equals_oracle(env, ref)
    assert env.episode_stats() == ref.episode_stats()
    env.close()


def test_step_range_leaves_other_envs_alone_and_wraps_the_window(rcw, oracle):
    """A range that is not aligned to the window wraps around it (slots 20..31, 0..19); envs outside the
    range keep state, reward and done.  Device-side action arrays work as for rcw_step."""
    import torch
    n, k, seed = 64, 32, 9
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_window_envs=k, auto_reset=False, **SMALL)
    ref = oracle.Batch(n, cfg=oracle.default_config(**SMALL_ORC), seed=seed, auto_reset=False)
    before = env.get_state()
    rng = np.random.default_rng(1)
    for _ in range(12):
        a = rng.integers(1, 5, k).astype(np.uint8)
        env.act_range(torch.from_numpy(a).cuda(), 20)
        full = np.full(n, 3, np.uint8)                     # the oracle steps everybody ...
        full[20:52] = a
        ref.step(full)
    env.sync()
    after = env.get_state()
    for key in ("pos", "dir_au", "goal", "reward", "done"):   # ... so only the range is compared with it
        np.testing.assert_array_equal(after[key][:20], before[key][:20])
        np.testing.assert_array_equal(after[key][52:], before[key][52:])
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(after["pos"][20:52]), bits(pos[20:52]))
    np.testing.assert_array_equal(after["dir_au"][20:52], au[20:52])
    np.testing.assert_array_equal(env.copy_obs(20, 32), ref.obs_rgb8()[20:52])
    env.close()


def test_step_range_without_a_window_and_errors(rcw, oracle):
    n, seed = 40, 2
    env = rcw.BatchedSingleRoom(n, seed=seed, **SMALL)             # full observation buffer
    ref = oracle.Batch(n, cfg=oracle.default_config(**SMALL_ORC), seed=seed)
    actions = np.random.default_rng(0).integers(1, 5, n).astype(np.uint8)
    env.act_range(actions[:15], 0)
    env.act_range(actions[15:], 15)
    ref.step(actions)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    _assert_state_equals_oracle(env, ref)
    with pytest.raises(rcw.RcwError) as ei:
        env.act_range(actions[:10], 35)                            # range leaves the batch
    assert ei.value.code == rcw._capi.RCW_ESIZE
    with pytest.raises(AssertionError):
        env.act_range(np.array([1, 5, 2], np.uint8), 0)            # the reference's @assert
    _assert_state_equals_oracle(env, ref)                          # nothing was enqueued
    env.close()
    win = rcw.BatchedSingleRoom(n, seed=seed, obs_window_envs=8, **SMALL)
    with pytest.raises(rcw.RcwError) as ei:
        win.act_range(actions[:9], 0)                              # more envs than slots
    assert ei.value.code == rcw._capi.RCW_ESIZE
    with pytest.raises(rcw.RcwError):
        win.copy_obs(0, 9)
    win.close()
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(4, obs_window_envs=-1)


def test_config3_2pow20_envs_on_one_gpu_through_a_window(rcw, oracle):
    """configs[2]: 2^20 default-camera envs (412 GB of observations per step) — more than one B200
    holds, and more than each of 2 GPUs holds when the batch is sharded over them.  With a window of
    16,384 slots (6.4 GB) one step is 64 launches and still writes every frame to HBM.  Sampled windows
    of eight envs are replayed by the oracle: state everywhere, observations in the last window."""
    n, k, seed, steps = 1 << 20, 1 << 14, 0x5EED, 3
    env = rcw.BatchedSingleRoom(n, seed=seed, obs_window_envs=k)
    env.step_random(steps)
    st = env.get_state()
    for s0 in (0, 16380, 524287, 777777, n - k, n - 8):
        ref = oracle.Batch(8, seed=seed, env_id_offset=s0)
        ref.rollout(steps, threads=4)
        pos, au, goal = ref.states()
        np.testing.assert_array_equal(bits(st["pos"][s0:s0 + 8]), bits(pos))
        np.testing.assert_array_equal(st["dir_au"][s0:s0 + 8], au)
        np.testing.assert_array_equal(st["goal"][s0:s0 + 8], goal)
        if s0 >= n - k:
            np.testing.assert_array_equal(env.copy_obs(s0, 8), ref.obs_rgb8())
    env.close()
