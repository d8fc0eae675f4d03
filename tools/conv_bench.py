#!/usr/bin/env python
"""Time every tcgen05 conv layer under the kernel variants and check they agree (GPU box).

variant bits (csrc/conv_tc.cu): 0 = v1; 1 = v1 sized for 2 CTAs/SM; 2 = v2 persistent super-tile; 6 = v2 + halo-shared A;
14 = v2 + halo + descriptor base_offset.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

from weights import make_state_dict, gaussian  # noqa: E402
from diffusion_models_for_gravitational_waveform_reconstruction_b200.engine import ModelSpec, UNetEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--variants", default="0,1,2,6,14")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    B, L = a.B, a.L
    sd = make_state_dict(3, 1, seed=0)
    spec = ModelSpec(in_ch=3, cond_in_ch=1, use_selfcond=True)
    eng = UNetEngine({k: v.cuda() for k, v in sd.items()}, spec, dtype="bf16", conv_impl="tc")
    x = gaussian((B, 3, L), seed=1).cuda()
    t = torch.full((B,), 500, device="cuda")
    eng.forward(x, t)                       # fills every activation buffer with realistic data
    torch.cuda.synchronize()
    ws = eng.workspace(B, L)
    d = spec.depth
    lc = spec.layer_channels
    res = []
    for li in range(1, 2 * d + 1):
        Lout = ws.lay_len[li]
        if li <= d:
            src0, src1, cin = ws.pooled[li - 1], None, lc[li - 1]
        else:
            i = li - d - 1
            src0, src1, cin = ws.out[li - 1], ws.out[d - 1 - i], lc[li - 1] + spec.chs[d - 1 - i]
        flops = 2.0 * cin * lc[li] * 3 * Lout * B
        ref_raw = ref_part = None
        for v in [int(s) for s in a.variants.split(",")]:
            eng.tc_variant = v
            raw = torch.zeros_like(ws.raw[li])
            part = torch.zeros_like(ws.part)
            try:
                n_part = eng._conv(li, src0, src1, raw, part)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"{spec.layer_names()[li]:12s} variant {v:2d} FAILED {e}")
                continue
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
            for s_, e_ in ev:
                s_.record()
                eng._conv(li, src0, src1, raw, part)
                e_.record()
            torch.cuda.synchronize()
            ms = statistics.median(s_.elapsed_time(e_) for s_, e_ in ev)
            if ref_raw is None:
                ref_raw, ref_part = raw.clone(), part[:, :n_part].clone()
                dr = dp = 0.0
            else:
                dr = float((raw.float() - ref_raw.float()).abs().max())
                dp = float((part[:, :n_part] - ref_part).abs().max() / (ref_part.abs().max() + 1e-30))
            row = {"layer": spec.layer_names()[li], "variant": v, "ms": ms, "tflops": flops / ms / 1e9, "max_diff_raw": dr,
                   "rel_diff_part": dp}
            res.append(row)
            print(f"{row['layer']:12s} variant {v:2d}  {ms*1e3:8.1f} us  {row['tflops']:7.1f} TFLOP/s  diff_raw={dr:.3e} diff_part={dp:.3e}")
    if a.json:
        with open(a.json, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
