#!/usr/bin/env python
"""Per-kernel-instance stall analysis of an `ncu --page source --csv` export: groups SASS instructions by execution count
(= code regions run by the same set of warps) and prints stall reasons / hottest instructions of the biggest regions.
    ncu -i X.ncu-rep --page source --csv > X_src.csv ; python tools/ncu_bucket.py X_src.csv [instance] [top]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1][:100] if i > 0 and len(rows[i - 1]) > 1 else "?" for i in hdrs]
print("instances:", len(hdrs))
i0 = hdrs[inst]
hdr = rows[i0]
body = rows[i0 + 1:(hdrs[inst + 1] - 1 if inst + 1 < len(hdrs) else None)]
print("kernel:", names[inst])
col = {h: i for i, h in enumerate(hdr)}
seen, recs = set(), []
for r in body:
    if len(r) < len(hdr) or r[0] in seen:
        continue
    seen.add(r[0])
    try:
        ns, ni = int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0)
    except ValueError:
        continue
    recs.append((r[0], r[col["Source"]].strip(), ns, ni,
                 {h[6:]: int(r[col[h]] or 0) for h in hdr if h.startswith("stall_") and "Not" not in h and (r[col[h]] or "0") != "0"}))
tot = sum(x[2] for x in recs)
print("static instrs", len(recs), "samples", tot, "warp-instr", sum(x[3] for x in recs))
by, byi = Counter(), Counter()
for a, s, ns, ni, st in recs:
    by[ni] += ns
    byi[ni] += 1
for ni, ns in by.most_common(4):
    hot = [x for x in recs if x[3] == ni]
    st = Counter()
    for x in hot:
        for k, v in x[4].items():
            st[k] += v
    n = max(1, sum(st.values()))
    print(f"== region executed {ni} times: {byi[ni]} instrs, {ns} samples ({100 * ns / tot:.1f} %)")
    print("   stalls:", ", ".join(f"{k} {100 * v / n:.1f}%" for k, v in st.most_common(8)))
    for x in sorted(hot, key=lambda x: -x[2])[:top]:
        print(f"   {x[2]:5d} {x[1][:84]:84s}", sorted(x[4].items(), key=lambda kv: -kv[1])[:2])
