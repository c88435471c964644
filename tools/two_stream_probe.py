#!/usr/bin/env python
"""Probe for VERDICT r01 #4(a): does splitting a batch over two streams hide the launch ramp / tail?
Two handles of E/2 envs each (every handle has its own stream) stepped alternately, against one handle of E envs.
Timed with host wall clock around K steps + sync (K large), device otherwise idle."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raycastworlds_jl_b200 as rcw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--parts", type=int, default=2)
ap.add_argument("--op", choices=["step", "top"], default="step", help="what to time: env steps or rcw_render_top_view")
args = ap.parse_args()


def run(handles, steps):
    for h in handles:
        h.step_random(200 if args.op == "top" else 20)
        if args.op == "top":
            h.render_top_view()
    for h in handles:
        h.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        for h in handles:
            if args.op == "top":
                h.render_top_view()
            else:
                h.step_random(1)
    for h in handles:
        h.sync()
    return (time.perf_counter() - t0) / steps * 1e3


one = [rcw.BatchedSingleRoom(args.envs, seed=1)]
ms1 = run(one, args.steps)
one[0].close()
parts = [rcw.BatchedSingleRoom(args.envs // args.parts, seed=1, env_id_offset=k * (args.envs // args.parts)) for k in range(args.parts)]
ms2 = run(parts, args.steps)
print(f"envs={args.envs}: one stream {ms1:.4f} ms/step, {args.parts} streams {ms2:.4f} ms/step ({(ms1 / ms2 - 1) * 100:+.1f} %)")
