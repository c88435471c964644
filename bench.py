#!/usr/bin/env python
"""bench.py — rendered env-steps/s of the batched SingleRoom hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 5 --warmup 1      # CPU arm (oracle port, all cores)

A step = one random-policy env-step of every env of the batch: act! (+ same-step auto-reset), the
512-ray DDA and the full 256x512 RGB8 observation written to HBM (BASELINE.json configs[1]:
4096 envs per GPU, default camera).  One step is exactly one kernel launch.  Weak scaling: every
rank owns `--envs-per-gpu` envs, global env ids key the RNG, no collective on the step path.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rendered env-steps/sec"
UNIT = "env-steps/s"
SEED = 0x5EED


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--obs-format", choices=["rgb8", "xrgb32", "gray8", "columns"], default="rgb8")
    ap.add_argument("--map", choices=["default", "large"], default="default",
                    help="default: 8x16 tiles / 128 directions; large: 64x64 / 256 (BASELINE config 5)")
    ap.add_argument("--rays", type=int, default=512, help="num_rays = observation width (default 512)")
    ap.add_argument("--height", type=int, default=256, help="height_camera_view_pu (default 256)")
    ap.add_argument("--obs-window-envs", type=int, default=0,
                    help="observation buffer of this many env slots (0 = every env); for batches whose "
                         "observations exceed HBM, e.g. BASELINE config 3 at 2 GPUs: 524288 envs per GPU")
    ap.add_argument("--top-view", action="store_true",
                    help="also redraw the top view inside every step (the reference's act!(env), single_room.jl:337); "
                         "not part of the north-star path, a second kernel launch and 512 KB more per env-step")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(args):
    kw = dict(height_tile_map_tu=8, width_tile_map_tu=16, num_directions=128)
    if args.map == "large":
        kw = dict(height_tile_map_tu=64, width_tile_map_tu=64, num_directions=256)
    kw.update(num_rays=args.rays, height_camera_view_pu=args.height)
    return kw


def mem_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, measured)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic(n_envs, kw, fmt):
    """DRAM bytes (read + write) of one launch of the step kernel, from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    key = f"{n_envs}x{kw['num_rays']}x{kw['height_camera_view_pu']}x{fmt}x{kw['height_tile_map_tu']}x{kw['width_tile_map_tu']}"
    try:
        with open(path) as f:
            return json.load(f).get(key)
    except OSError:
        return None


def cpu_port_rate(n_envs, kw, seconds, threads):
    """Env-steps/s of the CPU oracle (C restatement of the reference, pthreads over envs) on a
    bounded sample: the same batch, as many whole steps as fit in about `seconds`."""
    from oracle import oracle as orc

    cfg = orc.default_config(H=kw["height_tile_map_tu"], W=kw["width_tile_map_tu"],
                             N=kw["num_directions"], R=kw["num_rays"], P=kw["height_camera_view_pu"])
    b = orc.Batch(n_envs, cfg=cfg, seed=SEED)
    b.rollout(2, threads=threads)                      # cold: thread start-up, first touch of the images
    steps, chunk = 0, 8
    t0 = time.perf_counter()
    while True:
        b.rollout(chunk, threads=threads)
        steps += chunk
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
        chunk = int(max(8, min(4096, 0.25 * seconds * steps / dt)))   # about four more chunks
    return n_envs * steps / dt, steps, dt


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia is absent from this image and the DDA
    lives in un-vendored RayCaster.jl, so this times the oracle port (C restatement, pthreads over
    envs = the analogue of Threads.@threads over envs) with every host core, on the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    kw = workload(args)
    n = args.envs_per_gpu
    threads = os.cpu_count() or 1
    cfg = orc.default_config(H=kw["height_tile_map_tu"], W=kw["width_tile_map_tu"],
                             N=kw["num_directions"], R=kw["num_rays"], P=kw["height_camera_view_pu"])
    # bounded sample: a step is one random-policy step of `m` of the batch's n envs, m chosen so that
    # warmup + steps stay within about two minutes on this host
    probe = orc.Batch(min(n, 4 * threads), cfg=cfg, seed=SEED)
    t0 = time.perf_counter()
    probe.rollout(2, threads=threads)
    per_env_step = (time.perf_counter() - t0) / (2 * probe.num_envs)
    del probe
    budget = 120.0
    m = int(min(n, max(threads, budget / (per_env_step * max(1, args.steps + args.warmup)))))
    b = orc.Batch(m, cfg=cfg, seed=SEED)
    for _ in range(args.warmup):
        b.rollout(1, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.rollout(1, threads=threads)
    dt = time.perf_counter() - t0
    value = m * args.steps / dt
    sample = (f"{m} of the {n} envs per step x {args.steps} steps in {dt:.1f} s, UInt32 camera view, "
              f"{threads} pthreads over envs")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"BatchedSingleRoom {n} envs, {kw['height_tile_map_tu']}x{kw['width_tile_map_tu']} tiles, "
                               f"{kw['num_rays']} rays x {kw['height_camera_view_pu']} px, random policy + auto-reset",
                   "note": "CPU oracle port of the reference (C, -O2, pthreads over envs); Julia is not in the image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import raycastworlds_jl_b200 as rcw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    kw = workload(args)
    n = args.envs_per_gpu
    offset = rank * n
    env = rcw.BatchedSingleRoom(n, device=local, seed=SEED, env_id_offset=offset,
                                obs_format=args.obs_format, obs_window_envs=args.obs_window_envs,
                                top_view=args.top_view, result_ring=2, **kw)
    windowed = env.obs_window < n
    stream = torch.cuda.ExternalStream(env.cuda_stream(), device=dev)
    K, W = args.steps, args.warmup
    bytes_per_step_env = kw["num_rays"] * kw["height_camera_view_pu"] * env.bytes_per_pixel
    if args.obs_format == "columns":   # 4 bytes per column: not rendered pixels, the step is bound by act! + DDA
        bytes_per_step_env = kw["num_rays"] * 4
    if args.top_view:
        bytes_per_step_env += 4 * int(np.prod(env.top_view_shape[1:]))

    # ---- device-resident throughput: K launches back to back, CUDA events on the handle's stream
    env.step_random(W)
    env.sync()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    launches0 = env.launch_count()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    events[0].record(stream)
    for k in range(K):
        env.step_random(1)
        events[k + 1].record(stream)
    env.sync()
    barrier()
    launches = env.launch_count() - launches0
    clocks = None
    if rank == 0:
        # nvidia-smi needs ~100 ms to deliver its first sample: when the timed region was shorter, keep
        # the same load running (untimed) until a few samples exist, and say so
        extra_t0 = time.perf_counter()
        while len(sampler.lines) < 3 and time.perf_counter() - extra_t0 < 2.0:
            env.step_random(20)
            env.sync()
        extra_ms = 1e3 * (time.perf_counter() - extra_t0)
        clocks = sampler.stop()
        clocks["sampled"] = ("during the timed region" if extra_ms < 1.0 else
                             f"timed region + {extra_ms:.0f} ms of the same launches right after it")
    total_ms = events[0].elapsed_time(events[K])
    per_launch_ms = [events[k].elapsed_time(events[k + 1]) for k in range(K)]
    total_ms = rcw.max_over_ranks(total_ms, device=dev if world > 1 else None)
    value = world * n * K / (total_ms * 1e-3)

    # ---- end to end through the public API with HOST buffers, every step: pinned host actions in (they ride in
    #      the kernel parameters: H2D inside the launch), the step, reward + done of every env back in host memory
    #      (the kernel writes them through to the pinned result ring: D2H inside the launch) and read by the host.
    #      Observations stay in HBM for the learner (rcw_obs_device_ptr, the reference's aliased `state`).
    #      `e2e`: the host enqueues step k + 1 before it reads the results of step k (result_ring = 2), which a
    #      rollout loop may do because its actions come from the observations, not from the rewards;
    #      `e2e_lockstep`: the host reads the results of step k before it enqueues step k + 1.
    #      A third figure also pulls the whole observation to the host each step (PCIe-bound).
    e2e = None
    e2e_lockstep = None
    e2e_obs = None
    if not args.no_e2e:
        rng = np.random.default_rng(SEED + rank)
        pinned = torch.empty((K, n), dtype=torch.uint8).pin_memory()
        actions = pinned.numpy()
        actions[:] = rng.integers(1, 5, size=(K, n), dtype=np.uint8)
        r_host = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
        d_host = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()

        def run_e2e(lag):
            ret, fin, pending = 0.0, 0, []
            for k in range(K):
                pending.append(env.act_async(actions[k]))
                if len(pending) > lag:
                    r, d = env.wait(pending.pop(0))
                    ret += float(r.sum())
                    fin += int(np.count_nonzero(d))
            for t in pending:
                r, d = env.wait(t)
                ret += float(r.sum())
                fin += int(np.count_nonzero(d))
            return ret, fin

        results = {}
        for name, lag in (("e2e", 1), ("e2e_lockstep", 0)):
            for k in range(min(W, K)):
                env.wait(env.act_async(actions[k]))
            env.sync()
            barrier()
            t0 = time.perf_counter()
            ret, fin = run_e2e(lag)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            dt = rcw.max_over_ranks(dt, device=dev if world > 1 else None)
            results[name] = {"value": world * n * K / dt, "unit": UNIT, "h2d_bytes_per_step": n,
                             "d2h_bytes_per_step": n * 5, "ms_per_step": 1e3 * dt / K,
                             "episodes_finished_rank0": fin, "sum_reward_rank0": ret}
        e2e, e2e_lockstep = results["e2e"], results["e2e_lockstep"]
        e2e["note"] = ("host actions in, reward + done of every env out to host memory and summed by the host, every "
                       "step; step k + 1 is enqueued before the results of step k are read (rcw_step_async / rcw_wait, "
                       "result_ring = 2); observations stay in HBM")
        e2e_lockstep["note"] = "as e2e, but the results of step k are read before step k + 1 is enqueued"
    if not args.no_e2e and not windowed:
        Ko = max(1, min(K, 5))
        obs_host = torch.empty(env.obs_shape, dtype=torch.int32 if args.obs_format in ("xrgb32", "columns") else torch.uint8)
        obs_host = obs_host.pin_memory().numpy()
        if args.obs_format in ("xrgb32", "columns"):
            obs_host = obs_host.view(np.uint32)
        barrier()
        t0 = time.perf_counter()
        for k in range(Ko):
            env.act(actions[k])
            env.reward_done(r_host, d_host)
            env.copy_obs(out=obs_host)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        dt = rcw.max_over_ranks(dt, device=dev if world > 1 else None)
        e2e_obs = {"value": world * n * Ko / dt, "unit": UNIT, "steps": Ko, "h2d_bytes_per_step": n,
                   "d2h_bytes_per_step": n * 5 + n * bytes_per_step_env,
                   "note": "as e2e plus the full observation copied to pinned host memory every step"}

    env_window = env.obs_window
    stats = env.episode_stats()
    stats = rcw.reduce_episode_stats(stats, device=dev if world > 1 else None)
    env.close()

    if rank == 0:
        peak, peak_src = mem_peaks()
        launch_ms = sum(per_launch_ms) / len(per_launch_ms)
        achieved = n * bytes_per_step_env / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"BatchedSingleRoom {n} envs per GPU, {kw['height_tile_map_tu']}x{kw['width_tile_map_tu']} tiles, "
                            f"{kw['num_directions']} directions, {kw['num_rays']} rays x {kw['height_camera_view_pu']} px "
                            f"{args.obs_format}, random policy + auto-reset" + (" (BASELINE.json configs[1])" if (n, args.map, args.rays, args.height) == (4096, "default", 512, 256) else ""),
                "envs_per_gpu": n, "obs_bytes_per_env_step": bytes_per_step_env,
                "obs_window_envs": env_window, "launches_per_step": -(-n // env_window) * (2 if args.top_view else 1),
                "top_view": bool(args.top_view),
                "l2": (f"each step writes {n * bytes_per_step_env / 1e9:.2f} GB of observations, larger than the 126 MB L2; no flush needed"
                       if n * bytes_per_step_env > 126e6 else
                       f"each step writes {n * bytes_per_step_env / 1e6:.1f} MB of observations: L2-resident, the step is bound by act! + DDA (issue), not by HBM"),
                "seed": SEED,
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if args.top_view else measured_traffic(n, kw, args.obs_format),
                "kernel": "rcw::frame_kernel<kModeStep, fused>" + (" + rcw::top_view_kernel" if args.top_view else ""),
                "algorithmic_bytes_per_launch": min(n, env_window) * bytes_per_step_env,
                "launch_ms": launch_ms / (-(-n // env_window)), "peak_source": peak_src,
            },
            "e2e": e2e, "e2e_lockstep": e2e_lockstep, "e2e_obs_to_host": e2e_obs,
            "gpu_launches": launches,
            "clocks": clocks,
            "episodes": {"finished": stats[0], "sum_return": stats[1], "sum_length": stats[2]},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, steps, dt = cpu_port_rate(min(n, 1024), kw, args.cpu_baseline_seconds, threads)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{min(n, 1024)} envs x {steps} steps of the same workload in {dt:.1f} s "
                          "(C restatement of the reference, pthreads over envs, UInt32 camera view)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
