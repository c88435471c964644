#!/usr/bin/env python
"""bench.py — rendered env-steps/s of the batched SingleRoom hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 20 --warmup 5     # CPU arm (oracle port, all cores)

A step = one random-policy env-step of every env of the batch: act! (+ same-step auto-reset), the
512-ray DDA and the full 256x512 RGB8 observation written to HBM (BASELINE.json configs[1]:
4096 envs per GPU, default camera).  One step is exactly one kernel launch.  Weak scaling: every
rank owns `--envs-per-gpu` envs, global env ids key the RNG, no collective on the step path.
Rank 0 prints ONE JSON line.

The timed region is exactly K steps between two CUDA events on the handle's stream (barrier +
synchronize on both sides, max over ranks); it is repeated until at least `--min-seconds` of timed
launches have run and the MEDIAN region is reported (`timing` tells how many and their spread), so that
the clock samples come from timed launches even at the driver's --steps 20 (4.6 ms of work).

The other BASELINE configs ride in the same line under `configs` (short device-timed runs, each with
its own roofline fraction and clocks): at N = 1 config 4 (65,536 envs), config 5 (64x64 map, 256
directions, 262,144 envs), the reference's own UInt32 pixel format, two reduced resolutions and
step + top view; at N > 1 config 3 (2^20 envs sharded over the N GPUs, through an observation window
where one GPU's share does not fit in HBM) and the 128x128 reduced-resolution figure.

`--impl reference` times the CPU port of the reference (oracle/, C, pthreads over envs — Julia is
not in the image) on the SAME config, writing the SAME frame format, under both loop schedules
(step-major: every env once per step; env-major: each thread runs its envs through all K steps, the
analogue of Threads.@threads over envs) and reports the FASTER one as value / e2e.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
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
BPP = {"rgb8": 3, "xrgb32": 4, "gray8": 1, "columns": 4, "gray16f": 2, "gray8_half": 1}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--obs-format", choices=["rgb8", "xrgb32", "gray8", "columns", "gray16f", "gray8_half"], default="rgb8")
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
    ap.add_argument("--min-seconds", type=float, default=0.5,
                    help="repeat the K-step timed region until this much timed work has run; the median is reported")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE configs)")
    ap.add_argument("--reference-seconds", type=float, default=8.0,
                    help="--impl reference: timed work per loop schedule")
    return ap.parse_args()


def geometry(map_name="default", rays=512, height=256):
    kw = dict(height_tile_map_tu=8, width_tile_map_tu=16, num_directions=128)
    if map_name == "large":
        kw = dict(height_tile_map_tu=64, width_tile_map_tu=64, num_directions=256)
    kw.update(num_rays=rays, height_camera_view_pu=height)
    return kw


def frame_bytes(kw, fmt):
    if fmt == "columns":   # 4 bytes per column: not rendered pixels, the step is bound by act! + DDA
        return kw["num_rays"] * 4
    if fmt == "gray8_half":   # the 2 x 2 box-filtered frame: a quarter of the GRAY8 bytes
        return (kw["num_rays"] // 2) * (kw["height_camera_view_pu"] // 2)
    return kw["num_rays"] * kw["height_camera_view_pu"] * BPP[fmt]


def config_block(n_per_gpu, n_gpus, kw, fmt, map_name, top_view=False):
    """The `config` object — built the same way by BOTH arms so that the driver can compare them."""
    baseline = (n_per_gpu, map_name, kw["num_rays"], kw["height_camera_view_pu"], fmt, top_view) == \
               (4096, "default", 512, 256, "rgb8", False)
    return {
        "workload": (f"BatchedSingleRoom {n_per_gpu} envs per GPU x {n_gpus} GPU(s), "
                     f"{kw['height_tile_map_tu']}x{kw['width_tile_map_tu']} tiles, {kw['num_directions']} directions, "
                     f"{kw['num_rays']} rays x {kw['height_camera_view_pu']} px {fmt}, random policy + auto-reset"
                     + (" + top view" if top_view else "")
                     + (" (BASELINE.json configs[1] per GPU)" if baseline else "")),
        "envs_per_gpu": n_per_gpu, "n_gpus": n_gpus, "obs_format": fmt,
        "obs_bytes_per_env_step": frame_bytes(kw, fmt), "seed": SEED,
    }


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
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Number of samples so far (to split one sampler's stream into windows)."""
        return len(self.lines)

    def summary(self, first=0, last=None):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        sm, smax, power, reasons = [], [], [], set()
        for _, ln in self.lines[first:last]:
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
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except subprocess.TimeoutExpired:
            self.proc.kill()


def measured_traffic(n_envs, kw, fmt):
    """DRAM bytes (read + write) of ONE launch of the step kernel at this exact workload, taken from a committed
    `ncu --set full` capture (profiles/traffic.json names the capture), or None when this workload has none."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    key = f"{n_envs}x{kw['num_rays']}x{kw['height_camera_view_pu']}x{fmt}x{kw['height_tile_map_tu']}x{kw['width_tile_map_tu']}"
    try:
        with open(path) as f:
            d = json.load(f)
    except OSError:
        return None, None
    v = d.get(key)
    if isinstance(v, dict):
        return v.get("bytes"), v.get("source")
    return v, ("profiles/r01_step_kernel_final.md" if v is not None else None)


# --------------------------------------------------------------------------------------------
# CPU arm
# --------------------------------------------------------------------------------------------

ORACLE_RENDER = {"rgb8": "rgb8", "xrgb32": "xrgb32", "gray8": "gray8", "columns": "none", "gray16f": "gray8", "gray8_half": "gray8"}


def oracle_cfg(orc, kw):
    return orc.default_config(H=kw["height_tile_map_tu"], W=kw["width_tile_map_tu"], N=kw["num_directions"],
                              R=kw["num_rays"], P=kw["height_camera_view_pu"])


def cpu_schedules(orc, m, kw, fmt, K, W, seconds, threads):
    """Env-steps/s of the CPU port on `m` envs of the workload under both loop schedules.  A timed region is K
    steps of the m envs; regions repeat until `seconds` of timed work per schedule; the median region counts."""
    render = ORACLE_RENDER[fmt]
    out = {}
    for name in ("step_major", "env_major"):
        b = orc.Batch(m, cfg=oracle_cfg(orc, kw), seed=SEED)

        def region(k):
            if name == "step_major":
                for _ in range(k):
                    b.rollout(1, threads=threads, render=render)
            else:
                b.rollout(k, threads=threads, render=render)

        region(max(1, W))
        times, t_all = [], time.perf_counter()
        while True:
            t0 = time.perf_counter()
            region(K)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all >= seconds and len(times) >= 3 or len(times) >= 2000:
                break
        med = statistics.median(times)
        out[name] = {"value": m * K / med, "region_s_median": med, "region_s_min": min(times),
                     "region_s_max": max(times), "regions": len(times)}
        del b
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia is absent from this image and the DDA lives in
    un-vendored RayCaster.jl, so this times the oracle port (C restatement, pthreads over envs = the analogue of
    Threads.@threads over envs) with every host core, on the same config and frame format as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    kw = geometry(args.map, args.rays, args.height)
    n, G = args.envs_per_gpu, args.gpus
    fmt = args.obs_format
    threads = os.cpu_count() or 1
    # bounded sample of the N-GPU workload: m of its n * G envs (the port keeps a UInt32 and a byte image per env
    # in host memory; 8192 envs = 7 GB), every step of the timed region steps all m of them
    m = int(min(n * G, 8192))
    sched = cpu_schedules(orc, m, kw, fmt, args.steps, args.warmup, args.reference_seconds, threads)
    best = max(sched, key=lambda k: sched[k]["value"])
    value = sched[best]["value"]
    sample = (f"{m} of the workload's {n * G} envs, regions of {args.steps} steps repeated for "
              f"{args.reference_seconds:.0f} s per schedule, {fmt} frames written directly, {threads} pthreads over "
              f"envs; value = the faster schedule ({best})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sched[best]["region_s_median"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(n, G, kw, fmt, args.map, args.top_view),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "schedules": sched,
        "note": ("CPU oracle port of the reference (C, -O2, pthreads over envs); Julia is not in the image.  "
                 "step_major = every env once per step (frames stream through the host caches), env_major = each "
                 "thread runs its envs through all steps of a region (cache-resident frames)"),
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------

class Timer:
    """K-step regions on one handle, CUDA events on the handle's stream, max over ranks, median of regions."""

    def __init__(self, torch, dist, rcw, dev, world):
        self.torch, self.dist, self.rcw, self.dev, self.world = torch, dist, rcw, dev, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def device_regions(self, env, K, min_seconds, max_regions=400):
        """Median (over regions, each the max over ranks) milliseconds of K back-to-back step launches."""
        torch = self.torch
        stream = torch.cuda.ExternalStream(env.cuda_stream(), device=self.dev)

        def one():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            e0.record(stream)
            env.step_random(K)
            e1.record(stream)
            env.sync()
            self.barrier()
            return e0.elapsed_time(e1)

        first = self.max_over_ranks([one()])[0]
        n_more = int(min(max_regions, max(2, -(-min_seconds * 1e3 // max(first, 1e-3)))))   # the same on every rank
        times = self.max_over_ranks([first] + [one() for _ in range(n_more)])
        return times


def summarize(times_ms):
    return {"regions": len(times_ms), "region_ms_median": statistics.median(times_ms), "region_ms_min": min(times_ms),
            "region_ms_max": max(times_ms), "timed_total_s": sum(times_ms) * 1e-3}


def run_side_config(T, rcw, sampler, label, n, kw, fmt, K, W, min_seconds, peak, local, offset, window=0,
                    top_view=False, note=None):
    """One entry of the `configs` block: its own handle, W warm-up steps, K-step regions, median."""
    env = rcw.BatchedSingleRoom(n, device=local, seed=SEED, env_id_offset=offset, obs_format=fmt,
                                obs_window_envs=window, top_view=top_view, **kw)
    try:
        env.step_random(W)
        env.sync()
        m0 = sampler.mark() if sampler else 0
        l0 = env.launch_count()
        times = T.device_regions(env, K, min_seconds)
        launches = env.launch_count() - l0
        med = statistics.median(times)
        bytes_env = frame_bytes(kw, fmt)
        if top_view:
            bytes_env += 4 * kw["height_tile_map_tu"] * 32 * kw["width_tile_map_tu"] * 32
        achieved = n * bytes_env / (med / K * 1e-3) / 1e9
        out = {"workload": label, "envs_per_gpu": n, "obs_format": fmt, "obs_bytes_per_env_step": bytes_env,
               "obs_window_envs": env.obs_window, "steps": K, "warmup": W,
               "value": T.world * n * K / (med * 1e-3), "unit": UNIT, "ms_per_step": med / K,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
               "gpu_launches_per_region": launches // len(times), "timing": summarize(times)}
        if sampler:
            out["clocks"] = sampler.summary(m0, sampler.mark())
        if note:
            out["note"] = note
        return out
    finally:
        env.close()


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
    T = Timer(torch, dist, rcw, dev, world)

    kw = geometry(args.map, args.rays, args.height)
    fmt = args.obs_format
    n = args.envs_per_gpu
    offset = rank * n
    K, W = args.steps, args.warmup
    peak, peak_src = mem_peaks()
    env = rcw.BatchedSingleRoom(n, device=local, seed=SEED, env_id_offset=offset, obs_format=fmt,
                                obs_window_envs=args.obs_window_envs, top_view=args.top_view, result_ring=2, **kw)
    windowed = env.obs_window < n
    bytes_per_step_env = frame_bytes(kw, fmt)
    if args.top_view:
        bytes_per_step_env += 4 * int(np.prod(env.top_view_shape[1:]))

    # ---- device-resident throughput: regions of K launches back to back, CUDA events on the handle's stream
    env.step_random(W)
    env.sync()
    sampler = ClockSampler(local).start() if rank == 0 else None
    launches0 = env.launch_count()
    m0 = sampler.mark() if sampler else 0
    times = T.device_regions(env, K, args.min_seconds)
    clocks = None
    if sampler:
        clocks = sampler.summary(m0, sampler.mark())
        clocks["sampled"] = f"during the {len(times)} timed regions ({sum(times) * 1e-3:.2f} s of timed launches)"
    launches_per_region = (env.launch_count() - launches0) // len(times)
    region_ms = statistics.median(times)
    value = world * n * K / (region_ms * 1e-3)

    # ---- end to end through the public API with HOST buffers, every step: pinned host actions in (they ride in
    #      the kernel parameters: H2D inside the launch), the step, reward + done of every env back in host memory
    #      (the kernel writes them through to the pinned result ring: D2H inside the launch) and read by the host.
    #      Observations stay in HBM for the learner (rcw_obs_device_ptr, the reference's aliased `state`).
    #      `e2e`: the host enqueues step k + 1 before it reads the results of step k (result_ring = 2), which a
    #      rollout loop may do because its actions come from the observations, not from the rewards;
    #      `e2e_lockstep`: the host reads the results of step k before it enqueues step k + 1.
    #      A third figure also pulls the whole observation to the host each step (PCIe-bound).
    e2e = e2e_lockstep = e2e_obs = None
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

            def one():
                T.barrier()
                t0 = time.perf_counter()
                ret, fin = run_e2e(lag)
                torch.cuda.synchronize(dev)
                return time.perf_counter() - t0, ret, fin

            first, ret, fin = one()
            first = T.max_over_ranks([first])[0]
            n_more = int(min(400, max(2, -(-args.min_seconds // max(first, 1e-6)))))
            dts = T.max_over_ranks([first] + [one()[0] for _ in range(n_more)])
            dt = statistics.median(dts)
            results[name] = {"value": world * n * K / dt, "unit": UNIT, "h2d_bytes_per_step": n,
                             "d2h_bytes_per_step": n * 5, "ms_per_step": 1e3 * dt / K, "regions": len(dts),
                             "region_ms_min": 1e3 * min(dts), "region_ms_max": 1e3 * max(dts),
                             "episodes_finished_rank0_first_region": fin, "sum_reward_rank0_first_region": ret}
        e2e, e2e_lockstep = results["e2e"], results["e2e_lockstep"]
        e2e["note"] = ("host actions in, reward + done of every env out to host memory and summed by the host, every "
                       "step; step k + 1 is enqueued before the results of step k are read (rcw_step_async / rcw_wait, "
                       "result_ring = 2); observations stay in HBM; median of wall-clock regions of K steps")
        e2e_lockstep["note"] = "as e2e, but the results of step k are read before step k + 1 is enqueued"
    if not args.no_e2e and not windowed:
        Ko = max(1, min(K, 5))
        obs_host = torch.empty(env.obs_shape, dtype=torch.int32 if fmt in ("xrgb32", "columns") else torch.uint8)
        obs_host = obs_host.pin_memory().numpy()
        if fmt in ("xrgb32", "columns"):
            obs_host = obs_host.view(np.uint32)
        T.barrier()
        t0 = time.perf_counter()
        for k in range(Ko):
            env.act(actions[k])
            env.reward_done(r_host, d_host)
            env.copy_obs(out=obs_host)
        torch.cuda.synchronize(dev)
        dt = T.max_over_ranks([time.perf_counter() - t0])[0]
        e2e_obs = {"value": world * n * Ko / dt, "unit": UNIT, "steps": Ko, "h2d_bytes_per_step": n,
                   "d2h_bytes_per_step": n * 5 + n * bytes_per_step_env,
                   "note": "as e2e plus the full observation copied to pinned host memory every step"}

    env_window = env.obs_window
    stats = env.episode_stats()
    stats = rcw.reduce_episode_stats(stats, device=dev if world > 1 else None)
    env.close()

    # ---- the other BASELINE configs, short device-timed runs (every rank takes part at N > 1)
    side = {}
    if not args.no_configs and (args.map, args.rays, args.height, fmt, args.top_view) == ("default", 512, 256, "rgb8", False):
        Ks, Ws, secs = max(5, min(K, 20)), max(3, min(W, 5)), 0.3
        g_def, g_large = geometry(), geometry("large")
        runs = []
        if world == 1:
            runs = [
                ("config4_65536_envs", "BASELINE.json configs[3]: 512 rays x 256 px, 65,536 envs, rgb8 (25.8 GB of frames per step)",
                 65536, g_def, "rgb8", 0, False),
                ("config5_large_map_262144_envs", "BASELINE.json configs[4]: 64x64 tiles, 256 directions, 262,144 envs, rgb8 "
                 "(103 GB of frames per step, all resident)", 262144, g_large, "rgb8", 0, False),
                ("xrgb32_4096_envs", "configs[1] in the reference's own pixel format (UInt32, single_room.jl:300)",
                 4096, g_def, "xrgb32", 0, False),
                ("step_plus_top_view_4096_envs", "configs[1] + update_top_view! in every step: the reference's full "
                 "act!(env) sequence (single_room.jl:333-340); 917,504 B per env-step", 4096, g_def, "rgb8", 0, True),
                ("reduced_128x128_rgb8_65536_envs", "REDUCED RESOLUTION (not the default camera): 128 rays x 128 px rgb8",
                 65536, geometry(rays=128, height=128), "rgb8", 0, False),
                ("reduced_84x84_gray8_65536_envs", "REDUCED RESOLUTION (not the default camera): 84 rays x 84 px gray8",
                 65536, geometry(rays=84, height=84), "gray8", 0, False),
            ]
        else:
            share = (1 << 20) // world
            fits = share * frame_bytes(g_def, "rgb8") < 150e9
            runs = [
                ("config3_1M_envs_sharded", f"BASELINE.json configs[2]: 2^20 envs over {world} GPUs = {share} envs per GPU, "
                 "default camera rgb8" + ("" if fits else "; one GPU's frames (206 GB) exceed HBM: rendered through a "
                                          "131,072-slot observation window, every frame still written to HBM"),
                 share, g_def, "rgb8", 0 if fits else 131072, False),
                ("reduced_128x128_rgb8_65536_envs_per_gpu", "REDUCED RESOLUTION (not the default camera): 128 rays x 128 px "
                 "rgb8, 65,536 envs per GPU", 65536, geometry(rays=128, height=128), "rgb8", 0, False),
            ]
        for key, label, n_s, g, f, win, tv in runs:
            try:
                side[key] = run_side_config(T, rcw, sampler, label, n_s, g, f, Ks, Ws, secs, peak, local,
                                            rank * n_s, window=win, top_view=tv)
            except Exception as exc:   # a side run must never take the headline down with it
                side[key] = {"workload": label, "error": f"{type(exc).__name__}: {exc}"}
    if sampler:
        sampler.stop()

    if rank == 0:
        # launches of the step kernel per step, counted by the library: the observation windows of a step, times two when
        # the multi-step call runs the batch as two half-batches on two streams (their ramps and tails overlap)
        step_launches = max(1, launches_per_region // K // (2 if args.top_view else 1))
        launch_ms = region_ms / K / step_launches
        achieved = n * bytes_per_step_env / step_launches / (launch_ms * 1e-3) / 1e9
        traffic, traffic_src = (None, None) if args.top_view else measured_traffic(n // step_launches, kw, fmt)   # per launch, like achieved
        cfg = config_block(n, world, kw, fmt, args.map, args.top_view)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": region_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "timing": dict(summarize(times), obs_window_envs=env_window,
                           launches_per_step=launches_per_region // K,
                           l2=(f"each step writes {n * bytes_per_step_env / 1e9:.2f} GB of observations, larger than the 126 MB L2; no flush needed"
                               if n * bytes_per_step_env > 126e6 else
                               f"each step writes {n * bytes_per_step_env / 1e6:.1f} MB of observations: L2-resident, the step is bound by act! + DDA (issue), not by HBM"),
                           rule="value = median over regions of (K steps / region time); region time = max over ranks"),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "rcw::frame_kernel<kModeStep, fused>" + (" + rcw::top_view_kernel" if args.top_view else ""),
                "algorithmic_bytes_per_launch": n * bytes_per_step_env // step_launches,
                "launch_ms": launch_ms, "peak_source": peak_src,
                "launch_note": (f"a step is {step_launches} launches of the step kernel"
                                + (" — two half-batches on two streams inside rcw_step_random(K), running concurrently so that one's "
                                   "launch ramp / tail is filled by the other; launch_ms = step time / launches is therefore the "
                                   "EFFECTIVE time per launch (a half-batch launch timed alone, as ncu serialises it, lasts longer: "
                                   "profiles/README.md)" if step_launches == 2 * (-(-n // env_window)) else "")),
            },
            "e2e": e2e, "e2e_lockstep": e2e_lockstep, "e2e_obs_to_host": e2e_obs,
            "gpu_launches": launches_per_region,
            "clocks": clocks,
            "episodes": {"finished": stats[0], "sum_return": stats[1], "sum_length": stats[2]},
            "configs": side,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc

            threads = os.cpu_count() or 1
            m = min(n, 1024)
            sched = cpu_schedules(orc, m, kw, fmt, K, 2, args.cpu_baseline_seconds / 2, threads)
            best = max(sched, key=lambda k: sched[k]["value"])
            line["cpu_baseline"] = {
                "value": sched[best]["value"], "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{m} envs of the same workload, regions of {K} steps for {args.cpu_baseline_seconds / 2:.0f} s per "
                          f"schedule (C restatement of the reference, pthreads over envs, {fmt} frames written "
                          f"directly); value = the faster schedule ({best})",
                "schedules": {k: v["value"] for k, v in sched.items()}}
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
