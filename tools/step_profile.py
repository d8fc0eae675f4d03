#!/usr/bin/env python
"""CUDA-event timing of every kernel launch in one reverse-diffusion step (eager launches, GPU box).

Wraps the ctypes entry points so each C-ABI call is bracketed by events on the current stream.
"""
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
from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf  # noqa: E402


class TimedLib:
    def __init__(self, lib):
        self._lib = lib
        self.records = []
        self.on = False

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("gw_") or name in ("gw_last_error", "gw_conv_tc_packed_elems", "gw_conv_tc_n_part", "gw_conv_gn_group", "gw_conv_in_gn_group", "gw_conv_gn_sync_bytes", "gw_conv_in_direct_ws_floats", "gw_version"):
            return fn

        def wrapped(*a):
            if not self.on:
                return fn(*a)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*a)
            e.record()
            self.records.append((name, s, e))
            return rc
        return wrapped


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--cin", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    cc = 1 if a.cin == 3 else 5
    model = UNet1D(in_ch=a.cin, cond_in_ch=cc, use_selfcond=True, compute_dtype=a.dtype)
    model.load_state_dict(make_state_dict(a.cin, cc, seed=0))
    model = model.cuda().eval()
    diff = CustomDiffusion(T=1000, device="cuda")
    eng = model.engine(a.dtype)
    if a.variant >= 0:
        eng.tc_variant = a.variant
    plan = inf.make_sampler_plan(model, diff, a.B, a.L, T=1000, steps=50, eta=1.0, compute_dtype=a.dtype, cache=False)
    y = synthetic_chirps(a.B, a.L, seed=1)["y_norm"].cuda()
    if cc == 5:
        y = torch.cat([y, torch.zeros(a.B, 4, a.L, device="cuda")], 1)
    plan.load_inputs(torch.randn(a.B, 1, a.L, device="cuda"), y, torch.zeros_like(y), None)
    for _ in range(3):
        plan.enqueue_step()
    torch.cuda.synchronize()
    tl = TimedLib(eng.lib)
    eng.lib = tl
    tl.on = True
    for _ in range(a.steps):
        plan.enqueue_step()
    torch.cuda.synchronize()
    per_step = len(tl.records) // a.steps
    agg = collections.OrderedDict()
    for i, (name, s, e) in enumerate(tl.records):
        key = f"{i % per_step:02d} {name}"
        agg.setdefault(key, []).append(s.elapsed_time(e) * 1e3)
    tot = 0.0
    out = []
    for k, v in agg.items():
        m = statistics.median(v)
        tot += m
        out.append({"launch": k, "us": m})
        print(f"{k:28s} {m:9.1f} us")
    print(f"sum of launches: {tot:.1f} us per reverse step (B={a.B}, L={a.L}, dtype={a.dtype}, variant={eng.tc_variant})")
    if a.json:
        with open(a.json, "w") as fh:
            json.dump({"B": a.B, "L": a.L, "dtype": a.dtype, "variant": eng.tc_variant, "sum_us": tot, "launches": out}, fh, indent=1)


if __name__ == "__main__":
    main()
