#!/usr/bin/env python
"""Times rcw_render_top_view (update_top_view! for every env) with CUDA events on the handle's stream.
    python tools/top_view_bench.py --envs 4096 --iters 200
Prints one JSON line: ms per launch, image bytes written in GB/s, fraction of the measured copy peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import raycastworlds_jl_b200 as rcw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--rays", type=int, default=512)
ap.add_argument("--pu", type=int, default=32)
ap.add_argument("--map", choices=["default", "large"], default="default")
ap.add_argument("--warm-steps", type=int, default=200, help="random-policy steps before timing (spreads the poses)")
args = ap.parse_args()

kw = dict(num_rays=args.rays, pu_per_tu=args.pu)
if args.map == "large":
    kw.update(height_tile_map_tu=64, width_tile_map_tu=64, num_directions=256)
env = rcw.BatchedSingleRoom(args.envs, seed=11, **kw)
env.step_random(args.warm_steps)
stream = torch.cuda.ExternalStream(env.cuda_stream())
for _ in range(5):
    env.render_top_view()
env.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(args.iters):
    env.render_top_view()
e1.record(stream)
env.sync()
ms = e0.elapsed_time(e1) / args.iters
nbytes = args.envs * 4 * int(np.prod(env.top_view_shape[1:]))
peak = 6536.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except OSError:
    pass
print(json.dumps({"kernel": "top_view_kernel", "envs": args.envs, "image": list(env.top_view_shape[1:]), "rays": args.rays,
                  "ms_per_launch": ms, "bytes_per_launch": nbytes, "GB/s": nbytes / (ms * 1e-3) / 1e9,
                  "frac_of_copy_peak": nbytes / (ms * 1e-3) / 1e9 / peak}))
