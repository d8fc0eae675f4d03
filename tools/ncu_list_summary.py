#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch list: per-kernel-name totals and the launch sequence.
Usage: python tools/ncu_list_summary.py gpurun_out/launches.csv [--seq]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = None
    data = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        key = (int(d["ID"]), d["Kernel Name"][:52])
        data.setdefault(key, {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    agg = collections.OrderedDict()
    for (i, k), m in data.items():
        t = m.get("gpu__time_duration.sum", 0.0) / 1e3
        mb = (m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) / 1e6
        iss = m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
        if "--seq" in sys.argv:
            print(f"{i:4d} {k:54s} {t:8.1f} us {mb:9.1f} MB  {mb / max(t, 1e-9) * 1e-3:5.2f} TB/s  issue {iss:5.1f} %")
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += mb
        a[3] += iss * t
    tot = sum(a[1] for a in agg.values())
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:54s} n={a[0]:3d} {a[1]:8.1f} us {100 * a[1] / tot:5.1f} % {a[2]:9.1f} MB {a[2] / max(a[1], 1e-9) * 1e-3:5.2f} TB/s  issue {a[3] / max(a[1], 1e-9):5.1f} %")


if __name__ == "__main__":
    main()
