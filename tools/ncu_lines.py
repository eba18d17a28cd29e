#!/usr/bin/env python
"""Per-source-line and per-opcode shares of one ncu report (source page).  usage: ncu_lines.py rep [min_pct]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
def I(x):
    try: return int(x)
    except Exception: return 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
for a, b in zip(h, v):
    if a in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "launch__registers_per_thread",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed"):
        print(f"{a} = {b}")
    if a.startswith("smsp__average_warps_issue_stalled") and a.endswith("_per_issue_active.ratio") or a.startswith("smsp__average_warp_latency_issue_stalled"):
        try:
            if float(b) > 0.3: print(f"{a} = {b}")
        except Exception: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(r for r in rows if "# Samples" in r)
isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
lines = [r for r in rows if len(r) > iex and r[0].strip().isdigit()]
sass = [r for r in rows if len(r) > iex and r[0] == "" and r[2].startswith("0x")]
te = sum(I(r[iex]) for r in lines) or 1; ts = sum(I(r[isamp]) for r in lines) or 1
print("total inst", te, "samples", ts)
for r in sorted(lines, key=lambda r: int(r[0])):
    e, s = I(r[iex]) / te * 100, I(r[isamp]) / ts * 100
    if e > thr or s > thr: print(f"{e:5.1f}% ex {s:5.1f}% smp  L{r[0]}: {r[1].strip()[:100]}")
ce, cs = Counter(), Counter()
for r in sass:
    t = r[3].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ce[op] += I(r[iex]); cs[op] += I(r[isamp])
print("opcodes:", ", ".join(f"{op} {n/te*100:.1f}%/{cs[op]/ts*100:.1f}%" for op, n in ce.most_common(16)))
