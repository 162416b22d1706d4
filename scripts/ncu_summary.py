#!/usr/bin/env python
"""Summarise gpurun_out/launches_<tag>.csv (ncu launch list) and gpurun_out/prof_<tag>.ncu-rep (ncu --set full)
into profiles/<tag>_ncu_summary.txt.   usage: python scripts/ncu_summary.py <tag>"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
p = os.path.join(root, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(p):
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        key = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append(f"# ncu launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:80s} n={v[0]:5d} us={v[1]:12.1f} share={v[1] / tot * 100:5.1f}%")
    out.append(f"total us {tot:.1f}\n")
import glob
for rep in [os.path.join(root, "gpurun_out", f"prof_{tag}.ncu-rep")] + sorted(glob.glob(os.path.join(root, "gpurun_out", f"prof_{tag}_*.ncu-rep"))):
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    pat = re.compile(r"Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write)\.sum$|dram__cycles_active|gpu__dram_throughput|"
                     r"sm__pipe_tensor|sm__warps_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
                     r"sm__throughput.avg.pct|sm__pipe_fma_cycles_active.avg.pct|lts__t_bytes.sum$|lts__throughput.avg.pct|"
                     r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__inst_executed.avg.per_cycle_active|"
                     r"launch__shared_mem_per_block_dynamic|sm__inst_executed_pipe_tensor|smsp__cycles_active.avg$")
    idx = [i for i, h in enumerate(hdr) if pat.search(h)]
    out.append(f"# ncu --set full ({os.path.basename(rep)}), per launch")
    for row in rows[2:]:
        out.append("---")
        for i in idx:
            out.append(f"{hdr[i]:75s} {row[i]:>16s} {units[i]}")
os.makedirs(os.path.join(root, "profiles"), exist_ok=True)
dst = os.path.join(root, "profiles", f"{tag}_ncu_summary.txt")
open(dst, "w").write("\n".join(out) + "\n")
print(dst)
