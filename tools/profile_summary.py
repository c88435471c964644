#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep) into the small text summary committed under profiles/.

usage: python tools/profile_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.md "title" [items]
Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source sass`
(no GPU needed)."""
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
items = int(sys.argv[4]) if len(sys.argv) > 4 else 65536

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "dram__bytes_write.sum", "dram__bytes_read.sum", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
]
lines = [f"# {title}", "", f"source report: `{rep}` (ncu --set full --clock-control none --import-source on)", ""]
for r in rows[2:]:
    lines.append(f"## {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    for k in KEYS:
        if k in hdr:
            lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
    lines.append("")

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sass[sass.index('"Kernel Name"'):])))
secs = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
shdr = srows[1]
ie, isamp = shdr.index("Instructions Executed"), shdr.index("# Samples")
body = srows[2:(secs[1] if len(secs) > 1 else len(srows))]
tot = sum(int(r[ie]) for r in body)
stall_cols = [i for i, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
acc = {}
for r in body:
    for i in stall_cols:
        acc[shdr[i]] = acc.get(shdr[i], 0) + int(r[i] or 0)
s = sum(acc.values()) or 1
lines.append(f"warp instructions: {tot} total, {tot / items:.1f} per (env, 32-ray group) item ({items} items)")
lines.append("")
lines.append("warp stall samples: " + ", ".join(f"{k[6:]} {v / s:.1%}" for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:8]))
lines.append("")
lines.append("most sampled SASS instructions:")
lines.append("")
lines.append("```")
for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:10]:
    st = {shdr[i][6:]: r[i] for i in stall_cols if int(r[i] or 0) > 100}
    lines.append(f"{r[1].strip()[:60]:60s} exec={r[ie]:>9s} samples={r[isamp]:>6s} {st}")
lines.append("```")
lines.append("")
lines.append("SASS store / TMA mnemonics in this kernel: " + ", ".join(sorted({tok for r in body for tok in r[1].replace(',', ' ').split()
                                                                          if tok.startswith(("STG", "UBLKCP", "SYNCS", "LDS", "LDG"))})))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
