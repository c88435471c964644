#!/usr/bin/env python
"""Times rcw_expand_columns (column words -> pixels, a pure store stream) with CUDA events on the handle's stream.
    python tools/expand_bench.py --envs 4096 --fmt rgb8 --iters 200
Prints one JSON line: ms per launch, pixels written in GB/s, fraction of the measured copy peak."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import raycastworlds_jl_b200 as rcw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--fmt", default="rgb8")
ap.add_argument("--rays", type=int, default=512)
ap.add_argument("--height", type=int, default=256)
ap.add_argument("--iters", type=int, default=200)
args = ap.parse_args()

env = rcw.BatchedSingleRoom(args.envs, seed=3, obs_format="columns", num_rays=args.rays, height_camera_view_pu=args.height)
env.step_random(30)
fmt = dict(rcw.single_room._FORMATS)[args.fmt]
es = C.c_size_t()
rcw._capi.check(env._lib.rcw_expanded_layout(env._h, fmt, C.byref(es), None, None))
dst = torch.empty(args.envs * es.value, dtype=torch.uint8, device="cuda")
src, _, stride = env.obs_device_ptr()
stream = torch.cuda.ExternalStream(env.cuda_stream())


def launch():
    rcw._capi.check(env._lib.rcw_expand_columns(env._h, C.c_void_p(src), stride, args.envs, fmt, C.c_void_p(dst.data_ptr())))


for _ in range(10):
    launch()
env.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(args.iters):
    launch()
e1.record(stream)
env.sync()
ms = e0.elapsed_time(e1) / args.iters
bpp = {"rgb8": 3, "xrgb32": 4, "gray8": 1}[args.fmt]
nbytes = args.envs * args.rays * args.height * bpp
peak = 6536.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except OSError:
    pass
print(json.dumps({"kernel": "rcw_expand_columns", "envs": args.envs, "fmt": args.fmt, "geom": f"{args.rays}x{args.height}",
                  "ms_per_launch": ms, "frames_per_s": args.envs / (ms * 1e-3), "GB/s": nbytes / (ms * 1e-3) / 1e9,
                  "frac_of_copy_peak": nbytes / (ms * 1e-3) / 1e9 / peak}))
