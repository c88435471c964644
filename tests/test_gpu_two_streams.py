"""rcw_step_random(n >= 2) runs its steps as two half-batches on two streams (VERDICT r01 #4a): the halves overlap each
other's launch ramp and tail.  Nothing about the results may change: states, observations (every ring position),
top views and episode totals equal the oracle's, whatever the split, and equal the single-stream run's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


CASES = [
    dict(n=9, kw=dict(), okw=dict()),                                                              # default camera, odd batch
    dict(n=64, kw=dict(num_rays=84, height_camera_view_pu=84, obs_format="gray8"), okw=dict(R=84, P=84)),   # env kernel + table
    dict(n=37, kw=dict(num_rays=96, height_camera_view_pu=40, top_view=True, pu_per_tu=4), okw=dict(R=96, P=40, pu_per_tu=4)),
    dict(n=21, kw=dict(num_rays=128, height_camera_view_pu=64, obs_format="xrgb32", height_tile_map_tu=5, width_tile_map_tu=6,
                       num_directions=16), okw=dict(R=128, P=64, H=5, W=6, N=16)),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("two", [0, 1, 2])
def test_multi_step_calls_match_oracle(rcw, oracle, monkeypatch, case, two):
    c = CASES[case]
    if two == 2:                                         # top views as two half-batches like the steps, not as a pipeline
        if not c["kw"].get("top_view"):
            pytest.skip("only differs with top views")
        monkeypatch.setenv("RCW_TOP_PIPELINE", "0")
        two = 1
    pipelined = two == 1 and bool(c["kw"].get("top_view")) and __import__("os").environ.get("RCW_TOP_PIPELINE") != "0"
    monkeypatch.setenv("RCW_TWO_STREAMS", str(two))
    monkeypatch.setenv("RCW_TWO_STREAMS_MIN", "8")
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n, seed = c["n"], 40 + case
    env = rcw.BatchedSingleRoom(n, seed=seed, **c["kw"])
    ref = oracle.Batch(n, cfg=oracle.default_config(**c["okw"]), seed=seed)
    launches0 = env.launch_count()
    total = 0
    for steps in (2, 1, 7, 150, 3):                      # multi-step calls of several lengths, a single step in between
        env.step_random(steps)
        ref.rollout(steps)
        total += steps
        st = env.get_state()
        pos, au, goal = ref.states()
        np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
        np.testing.assert_array_equal(st["dir_au"], au)
        np.testing.assert_array_equal(st["goal"], goal)
        r, d = ref.reward_done()
        np.testing.assert_array_equal(st["reward"], r)
        np.testing.assert_array_equal(st["done"], d)
        fmt = c["kw"].get("obs_format", "rgb8")
        want = {"rgb8": ref.obs_rgb8, "xrgb32": ref.obs_u32, "gray8": ref.obs_gray8}[fmt]()
        np.testing.assert_array_equal(env.copy_obs(), want)
        if c["kw"].get("top_view"):
            top = env.copy_top_view()
            for e in (0, n - 1):
                w = ref.world(e)
                w.update_top_view()
                np.testing.assert_array_equal(top[e], w.top_view)
        assert env.episode_stats() == ref.episode_stats()
    per_step = 2 if c["kw"].get("top_view") else 1
    launched = env.launch_count() - launches0
    if two and not pipelined:
        assert launched > total * per_step, "the multi-step calls should have been split into two launches per step"
    else:       # one stream — or, with top views, the two-stage pipeline: step kernels on one stream, top view kernels on the other
        assert launched == total * per_step
    env.close()


def test_frame_ring_with_two_streams(rcw, oracle, monkeypatch):
    monkeypatch.setenv("RCW_TWO_STREAMS_MIN", "8")
    n, K, seed = 24, 4, 3
    env = rcw.BatchedSingleRoom(n, seed=seed, num_rays=64, height_camera_view_pu=32, frame_stack=K)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=64, P=32), seed=seed)
    frames = []
    for _ in range(3):
        ref.rollout(1)
        frames.append(ref.obs_rgb8())
    env.step_random(3)
    for age in range(3):
        np.testing.assert_array_equal(env.copy_obs(age=age), frames[2 - age])
    env.close()


def test_default_split_at_bench_size_sampled_against_oracle(rcw, oracle, monkeypatch):
    """The bench's own shape — 4096 default-camera envs stepped by one rcw_step_random(K) call, which the handle runs as
    two half-batches on two streams without any switch set — against oracle batches of 48 envs placed at the start, around
    the split and at the end of the batch (the Philox streams are keyed by global env id, so a sub-range can be replayed
    on its own)."""
    for var in ("RCW_TWO_STREAMS", "RCW_TWO_STREAMS_MIN"):        # the handle's own defaults, whatever the caller exported
        monkeypatch.delenv(var, raising=False)
    n, steps, seed = 4096, 1500, 24301
    env = rcw.BatchedSingleRoom(n, seed=seed)
    launches0 = env.launch_count()
    env.step_random(steps)
    assert env.launch_count() - launches0 == 2 * steps
    st = env.get_state()
    stats = env.episode_stats()
    assert stats[0] > 0
    for off in (0, 2048 - 24, n - 48):
        ref = oracle.Batch(48, seed=seed, env_id_offset=off)
        ref.rollout(steps, threads=4, render=False)
        pos, au, goal = ref.states()
        np.testing.assert_array_equal(bits(st["pos"][off:off + 48]), bits(pos))
        np.testing.assert_array_equal(st["dir_au"][off:off + 48], au)
        np.testing.assert_array_equal(st["goal"][off:off + 48], goal)
        r, d = ref.reward_done()
        np.testing.assert_array_equal(st["reward"][off:off + 48], r)
        np.testing.assert_array_equal(st["done"][off:off + 48], d)
        for e in (0, 23, 24, 47):
            w = ref.world(e)
            w.cast_rays()
            w.update_camera_view()
            np.testing.assert_array_equal(env.copy_obs(off + e, 1)[0], w.obs_rgb8())
    env.close()


@pytest.mark.parametrize("two,device_tape", [(0, False), (1, False), (1, True)])
def test_action_tape_matches_oracle(rcw, oracle, monkeypatch, two, device_tape):
    """rcw_step_tape: the steps of a host or device action tape, run as two half-batches on two streams (or one launch
    per step with RCW_TWO_STREAMS=0), equal the oracle stepping the same rows one by one."""
    monkeypatch.setenv("RCW_TWO_STREAMS", str(two))
    monkeypatch.setenv("RCW_TWO_STREAMS_MIN", "8")
    n, T, seed = 45, 120, 19
    kw = dict(num_rays=96, height_camera_view_pu=40, top_view=True, pu_per_tu=4)
    env = rcw.BatchedSingleRoom(n, seed=seed, **kw)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=96, P=40, pu_per_tu=4), seed=seed)
    rng = np.random.default_rng(3)
    tape = rng.choice([1, 1, 1, 2, 3, 4], size=(T, n)).astype(np.uint8)
    if device_tape:
        import torch

        env.act_tape(torch.from_numpy(tape).cuda())
    else:
        env.act_tape(tape[:1])                     # a one-step tape, then the rest
        env.act_tape(tape[1:])
    for t in range(T):
        assert ref.step(tape[t]) == 0
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
    np.testing.assert_array_equal(st["dir_au"], au)
    np.testing.assert_array_equal(st["goal"], goal)
    r, d = ref.reward_done()
    np.testing.assert_array_equal(st["reward"], r)
    np.testing.assert_array_equal(st["done"], d)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    top = env.copy_top_view()
    for e in (0, n - 1):
        w = ref.world(e)
        w.update_top_view()
        np.testing.assert_array_equal(top[e], w.top_view)
    assert env.episode_stats() == ref.episode_stats()
    # an invalid action anywhere on a host tape: nothing is enqueued (the reference's @assert)
    before = env.launch_count()
    bad = tape[:3].copy()
    bad[2, n - 1] = 0
    with pytest.raises(rcw.InvalidActionError):
        env.act_tape(bad)
    assert env.launch_count() == before
    env.close()
