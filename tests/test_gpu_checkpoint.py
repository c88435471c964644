"""GPU tests of checkpoint / resume (rcw_save_checkpoint / rcw_load_checkpoint): a batch
restored from a snapshot continues bit-identically to the uninterrupted run and to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KW = dict(num_rays=96, height_camera_view_pu=64)


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def snapshot(env):
    st = env.get_state()
    return (bits(st["pos"]).copy(), st["dir_au"].copy(), st["goal"].copy(), st["reward"].copy(), st["done"].copy(),
            env.copy_obs(), env.episode_stats())


def assert_same(a, b):
    for x, y in zip(a[:-1], b[:-1]):
        np.testing.assert_array_equal(x, y)
    assert a[-1] == b[-1]


def test_resume_continues_bit_identically(rcw, oracle):
    n, seed = 64, 31
    a = rcw.BatchedSingleRoom(n, seed=seed, env_id_offset=1000, **KW)
    a.step_random(400)                                   # long enough for a number of auto-resets
    ckpt = a.save_checkpoint()
    at_save = snapshot(a)
    a.step_random(300)
    final = snapshot(a)
    assert final[-1][0] > at_save[-1][0] > 0, "episodes should finish before and after the snapshot"

    b = rcw.BatchedSingleRoom(n, seed=999, env_id_offset=5, **KW)   # key and offset come back with the snapshot
    b.load_checkpoint(ckpt.tobytes())
    assert_same(snapshot(b), at_save)                    # state, counters and the re-rendered observations
    b.step_random(300)
    assert_same(snapshot(b), final)

    ref = oracle.Batch(n, cfg=oracle.default_config(R=96, P=64), seed=seed, env_id_offset=1000)
    ref.rollout(700)
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(final[0], bits(pos))
    np.testing.assert_array_equal(final[1], au)
    np.testing.assert_array_equal(final[5], ref.obs_rgb8())
    assert final[-1] == ref.episode_stats()

    # explicit actions after a resume, and a second save of the resumed run equals a save of the original
    acts = np.random.default_rng(0).integers(1, 5, n).astype(np.uint8)
    a.act(acts)
    b.act(acts)
    assert_same(snapshot(a), snapshot(b))
    np.testing.assert_array_equal(a.save_checkpoint(), b.save_checkpoint())
    a.close()
    b.close()


def test_checkpoint_errors(rcw):
    env = rcw.BatchedSingleRoom(8, seed=1, **KW)
    ckpt = env.save_checkpoint()
    with pytest.raises(rcw.RcwError) as ei:
        env.load_checkpoint(ckpt[:-1])                   # truncated
    assert ei.value.code == rcw._capi.RCW_ESIZE
    bad = ckpt.copy()
    bad[0] ^= 0xFF                                       # magic
    with pytest.raises(rcw.RcwError) as ei:
        env.load_checkpoint(bad)
    assert ei.value.code == rcw._capi.RCW_EINVAL
    other = rcw.BatchedSingleRoom(9, seed=1, **KW)       # another batch size
    with pytest.raises(rcw.RcwError) as ei:
        other.load_checkpoint(ckpt)
    assert ei.value.code == rcw._capi.RCW_ESIZE
    other.close()
    bad = ckpt.copy()
    hdr = bad.size - 8 * (8 * 4 + 1)
    bad[hdr + 2 * 8 * 4: hdr + 2 * 8 * 4 + 4] = np.frombuffer(np.int32(4000).tobytes(), np.uint8)   # dir_au[0] = 4000
    with pytest.raises(rcw.RcwError) as ei:
        env.load_checkpoint(bad)
    assert ei.value.code == rcw._capi.RCW_EINVAL
    before = snapshot(env)
    env.load_checkpoint(ckpt)                            # the failed loads left the batch untouched
    assert_same(snapshot(env), before)
    env.close()
