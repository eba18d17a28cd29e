#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libvcb200.so (cuobjdump -sass): the instructions that prove the
Blackwell paths -- UTCHMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UBLKCP (cp.async.bulk / TMA),
DMMA (FP64 tensor MMA), LDGSTS (cp.async), SYNCS (mbarrier).  usage: sass_summary.py [lib] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "voiceconversion.jl_b200", "libvcb200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP", "UBLKCP.MULTICAST", "SYNCS", "DMMA", "LDGSTS", "FFMA2", "DFMA",
        "DADD", "DMUL", "HMMA", "SHFL", "BAR", "MUFU"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.defaultdict(collections.Counter)
total = collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        base = op.split(".")[0]
        counts[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
        if op.startswith("UBLKCP") and "MULTICAST" in op:
            counts[cur]["UBLKCP.MULTICAST"] += 1
dem = subprocess.run(["cu++filt"] + list(total), capture_output=True, text=True).stdout.splitlines()
pretty = dict(zip(total, dem)) if len(dem) == len(total) else {k: k for k in total}
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  ({len(total)} kernels, sm_100a)")
print("# columns: instructions | " + " ".join(KEYS))
grand = collections.Counter()
for k in sorted(total, key=lambda k: -total[k]):
    row = [counts[k].get(key, 0) for key in KEYS]
    for key, v in zip(KEYS, row):
        grand[key] += v
    if any(row[:9]) or total[k] > 2000:
        name = pretty[k].replace("(anonymous namespace)::", "").replace("vcb::", "").replace("void ", "")
        name = re.sub(r"\((?:const |unsigned |TrajParams|GroupParams|double|int|long|float|void)[^()]*\)$", "", name)
        print(f"{total[k]:7d} | " + " ".join(f"{v:5d}" for v in row) + f" | {name[:110]}")
print("  TOTAL | " + " ".join(f"{grand[key]:5d}" for key in KEYS))
