#!/usr/bin/env python
"""Per-source-line warp-instruction counts of one kernel of an ncu report (needs -lineinfo + --import-source on).

usage: python tools/ncu_lines.py report.ncu-rep [units] [top]
Prints the lines of the .cu / .cuh sources that executed the most warp instructions, as instructions per unit
(units = envs or items of the launch) — the breakdown the front-bound work is steered by."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, total, lines = None, None, 0, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or not r[0].isdigit() or r[2] != "-":     # source rows have "-" in the Address column
        continue
    n = int(r[ie] or 0)
    if n:
        key = (fname, int(r[0]))
        cur = lines.get(key, [0, 0, r[1]])
        cur[0] += n
        cur[1] += int(r[isamp] or 0)
        lines[key] = cur
        total += n
print(f"total warp instructions {total}  ({total / units:.1f} per unit)")
for (f, ln), (n, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / units:9.1f} {s:6d}  {f}:{ln:<5d} {src.strip()[:110]}")
