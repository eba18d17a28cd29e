#!/usr/bin/env python
"""Turns the ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01

  profiles/<round>_launches.csv        raw launch list (gpu__time_duration per launch)
  profiles/<round>_launch_shares.txt   per-kernel totals and shares of the profiled command
  profiles/<round>_<capture>.txt       key metrics of each `ncu --set full` capture
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(rnd):
    src = os.path.join(OUT, f"launches_{rnd}.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, f"{rnd}_launches.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for row in r:
        name = row[ki].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(row[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{rnd}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 2 --warmup 3\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n# (cutlass3x...tf32gemm and at::...normal_ are bench.py's own dense-TF32 peak measurement with torch.matmul, outside the timed region)\n")
        f.write(f"{'total ms':>12s} {'launches':>9s} {'share':>7s}  kernel\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{t / 1e6:12.3f} {c:9d} {100 * t / tot:6.1f}%  {n}\n")


def capture(rnd, name):
    rep = os.path.join(OUT, f"{name}_{rnd}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, f"{rnd}_{name}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source report gpurun_out/{name}_{rnd}.ncu-rep (not tracked)\n")
        for r in rows[2:]:
            kn = r[hdr.index("Kernel Name")]
            vals = [(m, r[hdr.index(m)], units[hdr.index(m)]) for m in METRICS if m in hdr]
            if all(v in ("nan", "-nan", "") for _, v, _ in vals[:2]):
                continue
            f.write(f"\n== {kn}\n")
            for m, v, u in vals:
                if v not in ("nan", "-nan", ""):
                    f.write(f"{m:80s} {v:>20s} {u}\n")


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(rnd)
    for cap in ("prof_fbf_tc", "prof_fbf_simt", "prof_traj_dtw", "prof_traj", "prof_dtw", "prof_dtwbarrier", "prof_argmax", "prof_group", "prof_recheck", "prof_pair", "prof_dtwpipe", "prof_trajtiled"):
        capture(rnd, cap)
    print(sorted(os.listdir(PROF)))
