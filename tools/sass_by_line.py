#!/usr/bin/env python
"""Attribute the executed-instruction counts of an ncu source page (SASS) to CUDA source lines through `nvdisasm -g`.
Usage: python tools/sass_by_line.py <ncu-rep> <object.o> <mangled kernel name> [launch index]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kern = sys.argv[1:4]
    idx = sys.argv[4] if len(sys.argv) > 4 else "0"
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    line_of, cur, on = {}, None, False
    for ln in dis:
        if ln.startswith(kern + ":"):
            on = True
            continue
        if on and ln.startswith("\t.section"):
            break
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", idx, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    agg = collections.Counter()
    samp = collections.Counter()
    tot = tots = 0
    for r in rows[2:]:
        if len(r) <= ia or not r[ia].isdigit():
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        key = line_of.get(addr - base)
        agg[key] += int(r[ia])
        samp[key] += int(r[isamp])
        tot += int(r[ia])
        tots += int(r[isamp])
    srcs = {}
    print(f"{tot} warp instructions, {tots} samples")
    for key, n in agg.most_common(40):
        text = ""
        if key:
            f = srcs.get(key[0])
            if f is None:
                for root, _, files in os.walk(os.path.dirname(os.path.abspath(obj)) + "/.."):
                    if key[0] in files:
                        f = open(os.path.join(root, key[0])).read().splitlines()
                        break
                srcs[key[0]] = f or []
            f = srcs[key[0]]
            text = f[key[1] - 1].strip()[:90] if f and key[1] <= len(f) else ""
        print(f"{100 * n / tot:5.1f}% instr {100 * samp[key] / max(tots, 1):5.1f}% stall  {key}  {text}")


if __name__ == "__main__":
    main()
