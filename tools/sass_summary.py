#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 (packed
fp32) -- and HMMA, which would betray a legacy mma.sync path. Runs on the build box (cuobjdump, no GPU):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "weathermodel_b200", "libwm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "MUFU.EX2", "HMMA"]
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    kernels[cur]["_total"] += 1
    for k in KEYS:
        if op.startswith(k) and not (k == "HMMA" and op.startswith("UTCHMMA")):
            kernels[cur][k] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a); static instruction counts per kernel")
print(f"# {'kernel':70s} " + " ".join(f"{k:>8s}" for k in KEYS) + "    total")
for (name, c), dm in zip(kernels.items(), demangled):
    short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("wm::", "")
    print(f"  {short[:70]:70s} " + " ".join(f"{c[k]:8d}" for k in KEYS) + f" {c['_total']:8d}")
tot = Counter()
for c in kernels.values():
    tot.update(c)
print(f"  {'ALL KERNELS':70s} " + " ".join(f"{tot[k]:8d}" for k in KEYS) + f" {tot['_total']:8d}")
