#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a short text report. Usage: tools/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name"), file=out)
    for k in keys:
        if k in d:
            print(f"  {k:80s} {d[k]:>16s} {units[hdr.index(k)]}", file=out)
    st = [(h, d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    st.sort(key=lambda x: -float(x[1] or 0))
    print("  stalls per issue:", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={float(v):.2f}" for h, v in st[:8]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ia = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ia["# Samples"]].isdigit()]   # several kernels: headers repeat
tot = sum(int(r[ia["# Samples"]]) for r in data); texec = sum(int(r[ia["Instructions Executed"]]) for r in data)
print(f"source page: {tot} samples, {texec} warp instructions", file=out)
c = Counter()
for r in data:
    t = r[ia["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += int(r[ia["Instructions Executed"]])
print("  executed mix:", ", ".join(f"{k}={v/texec:.3f}" for k, v in c.most_common(14)), file=out)
print("  top stall sites:", file=out)
for r in sorted(data, key=lambda r: -int(r[ia["# Samples"]]))[:12]:
    print(f"    {int(r[ia['# Samples']]):7d} samples {int(r[ia['Instructions Executed']]):11d} exec  {r[ia['Source']].strip()[:70]}", file=out)
