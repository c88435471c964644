"""One process, several GPUs through the C ABI (include/rcw_b200.h "one process, several GPUs").

rcw_create_sharded / rcw_step_sharded / rcw_step_random_sharded / rcw_sync_sharded / rcw_reduce_episode_stats cut a
batch into contiguous blocks of global env ids, one handle per device.  Every trajectory, observation and episode total
must be bit-identical to the oracle's single batch — and therefore to a single handle and to the one-process-per-GPU
layout bench.py uses.  With one visible device the shards share it (the logic is the same); with several (gpurun
--gpus N) every shard gets its own device, driven from one host thread and from one host thread per handle."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rcw():
    import raycastworlds_jl_b200 as m
    return m


def device_count():
    import torch
    return torch.cuda.device_count()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def check_against_oracle(env, ref, total):
    st = env.get_state()
    pos, au, goal = ref.states()
    np.testing.assert_array_equal(bits(st["pos"]), bits(pos))
    np.testing.assert_array_equal(st["dir_au"], au)
    np.testing.assert_array_equal(st["goal"], goal)
    r, d = ref.reward_done()
    np.testing.assert_array_equal(st["reward"], r)
    np.testing.assert_array_equal(st["done"], d)
    want = ref.obs_rgb8()
    for shard, off in zip(env.shards, env.offsets):
        np.testing.assert_array_equal(shard.copy_obs(), want[off:off + shard.num_envs])
    assert env.episode_stats() == ref.episode_stats()


@pytest.mark.parametrize("n_shards,total", [(1, 37), (2, 37), (3, 64), (5, 23)])
def test_sharded_batch_matches_oracle(rcw, oracle, n_shards, total):
    nd = device_count()
    devices = [k % nd for k in range(n_shards)]
    kw = dict(height_tile_map_tu=6, width_tile_map_tu=9, num_directions=64, num_rays=96, height_camera_view_pu=64)
    env = rcw.ShardedSingleRoom(total, devices=devices, seed=21, env_id_offset=1000, **kw)
    assert [s.num_envs for s in env.shards] == [rcw.shard_envs(total, n_shards, k)[1] for k in range(n_shards)]
    assert [s.cfg.device for s in env.shards] == devices
    ref = oracle.Batch(total, cfg=oracle.default_config(H=6, W=9, N=64, R=96, P=64), seed=21, env_id_offset=1000)
    check_against_oracle(env, ref, total)
    env.step_random(150)
    ref.rollout(150)
    check_against_oracle(env, ref, total)
    rng = np.random.default_rng(total)
    for _ in range(20):
        a = rng.choice([1, 1, 1, 2, 3, 4], size=total).astype(np.uint8)
        env.act(a)
        assert ref.step(a) == 0
    env.sync()
    check_against_oracle(env, ref, total)
    np.testing.assert_array_equal(env.copy_obs(), ref.obs_rgb8())
    r, d = env.reward_done()
    np.testing.assert_array_equal(r, ref.reward_done()[0])
    np.testing.assert_array_equal(d, ref.reward_done()[1])
    assert ref.episode_stats()[0] > 0
    # an invalid action anywhere: nothing is enqueued on any shard (the reference's @assert, single_room.jl:140)
    before = [s.launch_count() for s in env.shards]
    a = np.ones(total, np.uint8)
    a[total - 1] = 5
    with pytest.raises(rcw.InvalidActionError):
        env.act(a)
    assert [s.launch_count() for s in env.shards] == before
    ep = env.episode_stats(reset_counters=True)
    assert ep == ref.episode_stats() and env.episode_stats()[0] == 0
    env.close()


def test_sharded_create_errors(rcw):
    import ctypes as C

    from raycastworlds_jl_b200 import _capi

    lib = _capi.load()
    cfg = _capi.default_config()
    cfg.num_envs = 2
    handles = (C.c_void_p * 4)()
    assert lib.rcw_create_sharded(C.byref(cfg), None, None, 4, handles) == _capi.RCW_EINVAL      # fewer envs than shards
    assert not any(h for h in handles)
    cfg.num_envs = 8
    dev = (C.c_int32 * 2)(0, 63)                                                                 # the second device does not exist
    assert lib.rcw_create_sharded(C.byref(cfg), None, dev, 2, handles) != _capi.RCW_OK
    assert not any(h for h in handles) and b"shard 1" in lib.rcw_last_error()
    assert lib.rcw_step_sharded(handles, 2, None) == _capi.RCW_EINVAL
    assert lib.rcw_reduce_episode_stats(None, 0, None, None, None, 0) == _capi.RCW_EINVAL


def test_one_host_thread_per_device(rcw, oracle):
    """N handles on N devices (all visible ones; one device twice if there is only one), each driven from its own host
    thread in one process: the result is the oracle's whole batch, i.e. what N processes produce (bench.py's layout)."""
    nd = device_count()
    n_shards = max(nd, 2)
    total, steps, seed = 50 * n_shards + 3, 200, 77
    kw = dict(num_rays=128, height_camera_view_pu=64)
    env = rcw.ShardedSingleRoom(total, devices=[k % nd for k in range(n_shards)], seed=seed, **kw)
    rng = np.random.default_rng(5)
    acts = rng.integers(1, 5, size=(steps, total)).astype(np.uint8)
    errors = []

    def worker(k):
        try:
            shard, off = env.shards[k], env.offsets[k]
            for t in range(steps):
                shard.act(acts[t, off:off + shard.num_envs])
                if t % 9 == 0:
                    shard.reward_done()
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(n_shards)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    ref = oracle.Batch(total, cfg=oracle.default_config(R=128, P=64), seed=seed)
    for t in range(steps):
        assert ref.step(acts[t]) == 0
    check_against_oracle(env, ref, total)
    env.close()
