#!/usr/bin/env python
"""Where the end-to-end step time goes (host actions in, reward/done out): per-call wall clock of act() and
reward_done() and the device time of the kernel.  usage: python tools/e2e_breakdown.py [--rays R --height P --fmt F]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raycastworlds_jl_b200 as rcw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--rays", type=int, default=512)
ap.add_argument("--height", type=int, default=256)
ap.add_argument("--fmt", default="rgb8")
ap.add_argument("--steps", type=int, default=2000)
a = ap.parse_args()
import torch

env = rcw.BatchedSingleRoom(a.envs, seed=1, num_rays=a.rays, height_camera_view_pu=a.height, obs_format=a.fmt)
n = a.envs
acts = torch.randint(1, 5, (a.steps, n), dtype=torch.uint8).pin_memory().numpy()
r = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
d = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
for k in range(50):
    env.act(acts[k]); env.reward_done(r, d)
t_act = t_rd = 0.0
t0 = time.perf_counter()
for k in range(a.steps):
    t1 = time.perf_counter()
    env.act(acts[k])
    t2 = time.perf_counter()
    env.reward_done(r, d)
    t3 = time.perf_counter()
    t_act += t2 - t1
    t_rd += t3 - t2
total = time.perf_counter() - t0
env.sync()
stream = torch.cuda.ExternalStream(env.cuda_stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
env.step_random(200)
e1.record(stream)
env.sync()
kern = e0.elapsed_time(e1) / 200 * 1e3
print(f"{a.envs} envs {a.rays}x{a.height} {a.fmt}: e2e {1e6 * total / a.steps:.1f} us/step "
      f"(act() {1e6 * t_act / a.steps:.1f} us, reward_done() {1e6 * t_rd / a.steps:.1f} us), kernel {kern:.1f} us, "
      f"overhead {1e6 * total / a.steps - kern:.1f} us")
