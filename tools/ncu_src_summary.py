#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export (SASS view): stall reasons in total and the hottest instructions.
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > K_src.csv ; python tools/ncu_src_summary.py K_src.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, body = r, rows[i + 1:]
        break
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter()
ops = Counter()
opinst = Counter()
recs = []
for r in body:
    if len(r) < len(hdr):
        continue
    try:
        ns = int(r[col["# Samples"]] or 0)
        ni = int(r[col["Instructions Executed"]] or 0)
    except ValueError:
        continue
    src = r[col["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.split(".")[0]
    ops[op] += ns
    opinst[op] += ni
    for s in stalls:
        try:
            tot[s] += int(r[col[s]] or 0)
        except ValueError:
            pass
    recs.append((ns, ni, r[col["Address"]], src, {s: int(r[col[s]] or 0) for s in stalls if (r[col[s]] or "0") != "0"}))
n = sum(tot.values())
print("total samples", n, " total warp instructions", sum(opinst.values()))
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / n:.1f}%" for k, v in tot.most_common(10)))
print("samples by opcode:", ", ".join(f"{k} {100 * v / n:.1f}%" for k, v in ops.most_common(14)))
ti = sum(opinst.values())
print("instructions by opcode:", ", ".join(f"{k} {100 * v / ti:.1f}%" for k, v in opinst.most_common(18)))
print(f"top {top} instructions by samples:")
for ns, ni, addr, src, st in sorted(recs, reverse=True)[:top]:
    main = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"  {ns:6d} smp {ni:9d} inst  {addr[-5:]}  {src[:70]:70s} {main}")
