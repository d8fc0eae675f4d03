#!/usr/bin/env python
"""profiles/<tag>_traffic.json from ncu launch lists (duration + DRAM bytes per launch), read by bench.py's roofline.traffic.
Usage: python tools/make_traffic_json.py <tag> <reverse_step.csv> <first launch id> <train_step.csv> <first launch id>"""
import collections
import csv
import json
import re
import sys


def load(path, first, last=None):
    rows = list(csv.reader(open(path)))
    hdr, data = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        i = int(d["ID"])
        if i < first or (last is not None and i >= last):
            continue
        name = re.sub(r"^void ", "", d["Kernel Name"]).split("(")[0].split("<")[0]
        data.setdefault(i, {"name": name})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    agg = collections.OrderedDict()
    for i, m in data.items():
        a = agg.setdefault(m["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a[2] += m.get("gpu__time_duration.sum", 0.0) / 1e3
    tot = sum(a[2] for a in agg.values())
    return {k: {"launches": a[0], "bytes_per_launch": a[1] / a[0], "ncu_us_per_launch": a[2] / a[0], "share_of_step": a[2] / tot}
            for k, a in agg.items()}, sum(a[1] for a in agg.values()), tot


def main():
    tag, rs, rs0, ts, ts0 = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
    rev, rev_bytes, rev_us = load(rs, rs0, int(sys.argv[6]) if len(sys.argv) > 6 else None)
    trn, trn_bytes, trn_us = load(ts, ts0)
    out = {"note": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, averaged over the launches of each kernel "
                   "in ONE step at B=256, L=4096, bf16; ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                   "dram__bytes_write.sum --clock-control none (cold caches, serialised launches: compare shares, not absolutes)",
           "sources": {"reverse_step": rs, "train_step": ts},
           "reverse_step_total": {"bytes": rev_bytes, "ncu_us": rev_us}, "train_step_total": {"bytes": trn_bytes, "ncu_us": trn_us},
           "reverse_step": rev, "train_step": trn}
    path = f"profiles/{tag}_traffic.json"
    json.dump(out, open(path, "w"), indent=1)
    print(path, "reverse step", f"{rev_bytes / 1e9:.3f} GB {rev_us:.1f} us;", "train step", f"{trn_bytes / 1e9:.3f} GB {trn_us:.1f} us")


if __name__ == "__main__":
    main()
