#!/usr/bin/env python
"""Key numbers of an .ncu-rep (`ncu --set full`): raw-page metrics, stall reasons, instruction mix and the hottest SASS lines.
Usage: python tools/ncu_rep_summary.py gpurun_out/x.ncu-rep [launch index] [top N]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    d = dict(zip(hdr, rows[2 + idx]))
    print(d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"  {k:72s} {d[k]}")
    st = [(k, float(v)) for k, v in d.items() if "average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio")
          and v not in ("", "n/a")]
    if not st:
        st = [(k, float(v)) for k, v in d.items() if "warp_issue_stalled" in k and "pct" in k and v not in ("", "n/a")]
    print("  stall reasons (warps per issue):")
    for k, v in sorted(st, key=lambda x: -x[1])[:8]:
        print(f"    {k:90s} {v:8.2f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot, tots, ops, lines = 0, 0, collections.Counter(), []
    for r in rows[2:]:
        if len(r) <= ia or not r[ia].isdigit():
            continue
        n, sm = int(r[ia]), int(r[isamp])
        tot += n
        tots += sm
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += n
        lines.append((sm, n, r[0], r[isrc].strip()))
    print(f"  warp instructions executed: {tot}, samples {tots}")
    print("  mix: " + ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(14)))
    print(f"  hottest SASS by stall samples (top {top}):")
    for sm, n, addr, txt in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f"    {100 * sm / max(tots, 1):5.1f}%  exec {n:9d}  {addr[-5:]}  {txt[:90]}")


if __name__ == "__main__":
    main()
