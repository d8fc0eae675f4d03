#!/usr/bin/env python
"""CUDA-event timing of every kernel launch in one training step (eager launches, GPU box), plus the graph-replayed
step time.  Usage: python tools/train_profile.py --B 256 --L 4096 --cin 7 --dtype bf16"""
import argparse
import collections
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

from weights import make_state_dict, synthetic_chirps  # noqa: E402
from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D  # noqa: E402
from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep  # noqa: E402
from step_profile import TimedLib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--cin", type=int, default=7)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--selfcond", type=int, default=0)
    ap.add_argument("--json", default=None)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--wgrad-variant", type=int, default=-1)
    ap.add_argument("--fuse-gn-bwd", type=int, default=-1, help="1/0: one-pass GroupNorm backward on/off (default: library default)")
    ap.add_argument("--opt", action="append", default=[], help="name=value for gw_set_option (repeatable)")
    ap.add_argument("--overlap-prep", type=int, default=-1, help="1/0: dgrad weight preparation on a forked stream")
    ap.add_argument("--only", default=None, help="print only launches whose name contains this")
    a = ap.parse_args()
    cc = 1 if a.cin == 3 else 5
    model = UNet1D(in_ch=a.cin, cond_in_ch=cc, use_selfcond=True, compute_dtype=a.dtype)
    model.load_state_dict(make_state_dict(a.cin, cc, seed=0))
    model = model.cuda()
    diff = CustomDiffusion(T=1000, device="cuda")
    st = FusedTrainStep(model, diff, a.B, a.L, compute_dtype=a.dtype, seed=3)
    d = synthetic_chirps(a.B, a.L, seed=1)
    cond = d["y_norm"]
    if cc == 5:
        cond = torch.cat([cond, torch.zeros(a.B, 4, a.L)], 1)
    if a.wgrad_variant >= 0:
        st.bwd.wgrad_variant = a.wgrad_variant
    if a.fuse_gn_bwd >= 0:
        st.bwd.fuse_gn_bwd = bool(a.fuse_gn_bwd)
    if a.overlap_prep >= 0:
        st.overlap_prep = bool(a.overlap_prep)
    for o in a.opt:
        k, v = o.split("=")
        assert st.lib.gw_set_option(k.encode(), int(v)) == 0, o
    st.load_batch(d["clean_norm"].cuda(), cond.cuda(), None)
    for _ in range(2):
        st.step(selfcond=bool(a.selfcond), use_graph=False)
    torch.cuda.synchronize()
    tl = TimedLib(st.lib)
    st.lib = st.eng.lib = st.bwd.lib = tl
    tl.on = True
    for _ in range(a.steps):
        st.step(selfcond=bool(a.selfcond), use_graph=False)
    torch.cuda.synchronize()
    tl.on = False
    per_step = len(tl.records) // a.steps
    agg = collections.OrderedDict()
    for i, (name, s, e) in enumerate(tl.records):
        agg.setdefault(f"{i % per_step:02d} {name}", []).append(s.elapsed_time(e) * 1e3)
    tot, out = 0.0, []
    by_name = collections.OrderedDict()
    for k, v in agg.items():
        m = statistics.median(v)
        tot += m
        out.append({"launch": k, "us": m})
        by_name[k.split(" ", 1)[1]] = by_name.get(k.split(" ", 1)[1], 0.0) + m
        if a.only is None or a.only in k:
            print(f"{k:28s} {m:9.1f} us")
    print("--- by entry point")
    for k, v in by_name.items():
        print(f"{k:28s} {v:9.1f} us  {100 * v / tot:5.1f} %")
    print(f"sum of C-ABI calls: {tot:.1f} us per training step (B={a.B}, L={a.L}, cin={a.cin}, dtype={a.dtype}, selfcond={a.selfcond})")
    if a.no_graph:
        return
    # graph-replayed step
    st.lib = st.eng.lib = st.bwd.lib = tl._lib
    for _ in range(3):
        st.step(selfcond=bool(a.selfcond), use_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        st.step(selfcond=bool(a.selfcond), use_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"graph-replayed step: {ms * 1e3:.1f} us -> {a.B / ms * 1e3:.0f} samples/s; loss {float(st.loss):.5f} grad-norm {float(st.info[0]):.4f}")
    if a.json:
        with open(a.json, "w") as fh:
            json.dump({"B": a.B, "L": a.L, "dtype": a.dtype, "sum_us": tot, "graph_step_us": ms * 1e3, "launches": out}, fh, indent=1)


if __name__ == "__main__":
    main()
