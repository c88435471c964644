#!/usr/bin/env python
"""Development sweep: time the step kernel for library variants (RCW_LIB) and grid shapes
(RCW_CTAS_PER_SM).  Prints one line per configuration.  Usage (on a GPU box):
    python tools/sweep.py --libs default,st1 --ctas 0,4,8 --envs 4096 --steps 100
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ap = argparse.ArgumentParser()
ap.add_argument("--libs", default="default")
ap.add_argument("--ctas", default="auto", help="RCW_CTAS_PER_SM values; auto = library heuristic")
ap.add_argument("--envs", default="4096")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--fmt", default="rgb8")
ap.add_argument("--map", default="default")
ap.add_argument("--geom", default="512x256", help="comma list of RAYSxHEIGHT")
args = ap.parse_args()

for lib in args.libs.split(","):
    for ctas in args.ctas.split(","):
        for envs, geom in [(e, g) for e in args.envs.split(",") for g in args.geom.split(",")]:
            rays, height = geom.split("x")
            env = dict(os.environ)
            if lib != "default":
                env["RCW_LIB"] = os.path.join(ROOT, "raycastworlds.jl_b200", "lib", "variants", f"librcw_b200_{lib}.so")
            if ctas != "auto":
                env["RCW_CTAS_PER_SM"] = ctas
            else:
                env.pop("RCW_CTAS_PER_SM", None)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(args.steps),
                                "--warmup", "10", "--no-e2e", "--no-cpu-baseline", "--envs-per-gpu", envs,
                                "--obs-format", args.fmt, "--map", args.map, "--rays", rays, "--height", height],
                               env=env, capture_output=True, text=True)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                print(f"lib={lib:8s} ctas/sm={ctas:4s} envs={envs:7s} geom={geom:9s} fmt={args.fmt} map={args.map} ms/step={j['ms_per_step']:.4f} "
                      f"steps/s={j['value']:.4g} GB/s={j['roofline']['achieved']:.0f} frac={j['roofline']['frac']:.3f} "
                      f"sm_mhz={j['clocks']['sm_mhz']}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"lib={lib} ctas={ctas} envs={envs} FAILED: {ex}\n{r.stdout[-500:]}\n{r.stderr[-1500:]}", flush=True)
