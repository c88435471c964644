"""GPU tests of the frame ring (rcw_config.frame_stack): the K most recent frames of every
env stay in the observation buffer; each of them must equal the oracle's frame of that step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


@pytest.mark.parametrize("R,P,fmt,env_kernel", [(64, 32, "rgb8", 1), (45, 21, "rgb8", 0), (96, 40, "gray8", 1),
                                                (160, 64, "xrgb32", 0)])
def test_ring_holds_the_last_k_frames(rcw, oracle, monkeypatch, R, P, fmt, env_kernel):
    monkeypatch.setenv("RCW_ENV_PER_WARP", str(env_kernel))
    monkeypatch.setenv("RCW_ENV_PER_WARP_MIN", "1")
    n, K, seed = 20, 4, 12
    env = rcw.BatchedSingleRoom(n, seed=seed, num_rays=R, height_camera_view_pu=P, obs_format=fmt, frame_stack=K)
    ref = oracle.Batch(n, cfg=oracle.default_config(R=R, P=P), seed=seed)
    frame = {"rgb8": ref.obs_rgb8, "xrgb32": ref.obs_u32, "gray8": ref.obs_gray8}[fmt]
    history = [frame()]
    assert env.obs_frames()[:2] == (K, 0)
    np.testing.assert_array_equal(env.copy_obs(), history[-1])
    assert not env.copy_obs(age=1).any(), "older ring positions start black"
    rng = np.random.default_rng(0)
    for t in range(11):
        if t % 3 == 0:
            env.step_random(1)
            ref.rollout(1)
        else:
            a = rng.integers(1, 5, n).astype(np.uint8)
            env.act(a)
            assert ref.step(a) == 0
        history.append(frame())
        k, newest, stride = env.obs_frames()
        assert (k, newest) == (K, (t + 1) % K)
        for age in range(min(K, len(history))):
            np.testing.assert_array_equal(env.copy_obs(age=age), history[-1 - age], err_msg=f"step {t} age {age}")
    # the zero-copy view shows the ring itself: [env, ring position, columns, rows(, 3)]
    ring = env.obs_tensor().cpu().numpy()
    if fmt == "xrgb32":
        ring = ring.view(np.uint32)
    _, newest, _ = env.obs_frames()
    for age in range(K):
        np.testing.assert_array_equal(ring[:, (newest - age) % K], history[-1 - age])
    # render / masked reset overwrite the newest frame only
    env.render()
    np.testing.assert_array_equal(env.copy_obs(age=0), history[-1])
    mask = np.zeros(n, np.uint8)
    mask[[3, 17]] = 1
    env.reset(mask=mask)
    assert env.obs_frames()[1] == newest
    got = env.copy_obs()
    np.testing.assert_array_equal(got[mask == 0], history[-1][mask == 0])
    assert not np.array_equal(got[3], history[-1][3]) or not np.array_equal(got[17], history[-1][17])
    np.testing.assert_array_equal(env.copy_obs(age=1), history[-2])
    env.close()


def test_frame_ring_errors(rcw):
    env = rcw.BatchedSingleRoom(8, num_rays=32, height_camera_view_pu=16, frame_stack=3)
    with pytest.raises(rcw.RcwError) as ei:
        env.act_range(np.ones(4, np.uint8), 0)             # the ring position is shared by the batch
    assert ei.value.code == rcw._capi.RCW_EINVAL
    with pytest.raises(rcw.RcwError) as ei:
        env.copy_obs(age=3)
    assert ei.value.code == rcw._capi.RCW_ESIZE
    env.close()
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(8, frame_stack=2, obs_window_envs=4)
    with pytest.raises(rcw.RcwError):
        rcw.BatchedSingleRoom(8, frame_stack=65)
    one = rcw.BatchedSingleRoom(2, num_rays=32, height_camera_view_pu=16)
    assert one.obs_frames()[:2] == (1, 0)
    one.step_random(3)
    assert one.obs_frames()[:2] == (1, 0)
    one.close()
