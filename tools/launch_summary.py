#!/usr/bin/env python
"""Condense an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
usage: python tools/launch_summary.py gpurun_out/launches.csv "command line that was profiled" > profiles/xxx_launches.txt"""
import csv
import sys
from collections import OrderedDict


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list of `{cmd}` ({sum(a[0] for a in agg.values())} launches)")
    print("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES")
    print(f"{'kernel':76s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:76]:76s} {n:5d} {t:10.1f} {t / n:9.2f} {t / total:7.3f}")


if __name__ == "__main__":
    main()
