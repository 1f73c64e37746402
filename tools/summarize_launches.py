"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches_summary.txt
Times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        val = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3 if unit in ("ms", "msecond") else val
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("wm::", "")
        rows.append((name, us))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches (first {skip} skipped), total {total / 1e3:.3f} ms of kernel time under ncu")
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:8d} {us / 1e3:10.3f} {us / n:10.1f} {100 * us / total:6.1f}%")


if __name__ == "__main__":
    main()
