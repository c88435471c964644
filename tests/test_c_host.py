"""The C ABI driven from plain C (tests/c/abi_host.c), as a foreign-language binding would drive it:
compiled with gcc against include/rcw_b200.h, linked with the library only.  The CPU part checks that
the program builds and that, without a GPU, it fails loudly; the GPU part compares what the C host
computed with the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "raycastworlds.jl_b200", "lib")
SRC = os.path.join(ROOT, "tests", "c", "abi_host.c")


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("c_host") / "abi_host")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    SRC, "-o", out, "-L", LIBDIR, "-lrcw_b200", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return out


def actions_of(n, s):
    e = np.arange(n)
    return (1 + np.where((e + s) % 7 == 0, 2, 0) + np.where((e * 31 + s) % 11 == 0, 1, 0)).astype(np.uint8)


def fnv1a(buf):
    h = 2166136261
    for v in buf.tobytes():
        h = ((h ^ v) * 16777619) & 0xFFFFFFFF
    return h


def test_c_host_builds_and_fails_loudly_without_a_gpu(host_binary):
    import torch

    r = subprocess.run([host_binary, "4", "3", "1"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr and "rcw_create" in r.stderr


@pytest.mark.gpu
def test_c_host_matches_oracle(host_binary, oracle):
    n, steps, seed = 24, 60, 9
    r = subprocess.run([host_binary, str(n), str(steps), str(seed)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = [ln.split() for ln in r.stdout.strip().splitlines()]
    sharded = lines.pop()
    ring = lines.pop()
    assert [int(ln[0]) for ln in lines] == [10, 20, 30, 40, 50, 60]
    ref = oracle.Batch(n, cfg=oracle.default_config(R=96, P=64), seed=seed)
    k = 0
    sum_reward, sum_done = 0.0, 0
    for s in range(1, steps + 1):
        assert ref.step(actions_of(n, s)) == 0
        rew, don = ref.reward_done()
        sum_reward += float(rew.sum())
        sum_done += int(don.sum())
        if s % 10 == 0:
            pos, au, _ = ref.states()
            step, sx, sy, sd, episodes, crc = lines[k]
            assert float(sx) == pytest.approx(float(pos[:, 0].astype(np.float64).sum()), abs=1e-4)
            assert float(sy) == pytest.approx(float(pos[:, 1].astype(np.float64).sum()), abs=1e-4)
            assert int(sd) == int(au.sum())
            assert int(episodes) == ref.episode_stats()[0]
            assert int(crc) == fnv1a(ref.world(0).obs_rgb8())
            k += 1
    # every step's rewards / terminations as read from the pinned result ring (rcw_step_async / rcw_wait)
    assert ring[0] == "ring" and float(ring[1]) == sum_reward and int(ring[2]) == sum_done
    # the same batch as three shards driven through rcw_*_sharded from C: the program itself compares state, pixels and
    # episode totals with the single handle (exit codes 8-11); the totals it prints are the oracle's
    ep, ret, length = ref.episode_stats()
    assert sharded[0] == "sharded" and (int(sharded[1]), float(sharded[2]), int(sharded[3])) == (ep, ret, length)
