"""Host logic of the multi-GPU path on CPU: env sharding and the episode-statistics reduction
over a 2-rank gloo group (the step path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import raycastworlds_jl_b200 as rcw


def test_shard_envs_partitions_exactly():
    for total in (0, 1, 7, 4096, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            cover, prev_end = 0, 0
            for r in range(world):
                off, n = rcw.shard_envs(total, world, r)
                assert off == prev_end and n >= 0
                prev_end = off + n
                cover += n
            assert cover == total
            sizes = [rcw.shard_envs(total, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rcw.shard_envs(10, 2, 2)


def test_reduce_without_process_group_is_identity():
    assert rcw.reduce_episode_stats((3, 3.0, 90)) == (3, 3.0, 90)
    assert rcw.max_over_ranks(1.5) == 1.5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, n = rcw.shard_envs(1001, world, rank)
        # every rank reports the statistics of its own shard; the reduction sums them
        stats = (n, float(off), 10 * n)
        total = rcw.reduce_episode_stats(stats)
        slowest = rcw.max_over_ranks(1.0 + rank)
        if rank == 0:
            torch.save({"total": total, "slowest": slowest}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_reduction(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    offs = [rcw.shard_envs(1001, 2, k) for k in range(2)]
    assert r["total"] == (1001, float(sum(o for o, _ in offs)), 10010)
    assert r["slowest"] == 2.0


def test_library_shard_blocks_equal_the_python_helper():
    """rcw_shard_envs (the C ABI's block partition) == sharding.shard_envs, no GPU needed."""
    import ctypes as C

    from raycastworlds_jl_b200 import _capi

    lib = _capi.load()
    for total in (0, 1, 7, 64, 1000003, 1 << 20):
        for n in (1, 2, 3, 8, 64):
            covered = 0
            for k in range(n):
                off, cnt = C.c_int64(), C.c_int64()
                assert lib.rcw_shard_envs(total, n, k, C.byref(off), C.byref(cnt)) == _capi.RCW_OK
                assert (off.value, cnt.value) == rcw.shard_envs(total, n, k)
                assert off.value == covered
                covered += cnt.value
            assert covered == total
    assert lib.rcw_shard_envs(10, 0, 0, None, None) == _capi.RCW_EINVAL
    assert lib.rcw_shard_envs(10, 2, 2, None, None) == _capi.RCW_EINVAL
