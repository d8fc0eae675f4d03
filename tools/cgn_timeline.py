#!/usr/bin/env python
"""Phase timeline of the fused conv+GroupNorm kernel (clock64 stamps written by CTA 0..G-1 for their first 16 samples).

stamps: 0 loop top, 1 accumulator ready, 2 pass 1 done, 3 exchange done, 4 coefficients done, 5 pass 2 done (epilogue);
        6 MMA start, 7 MMA issued (MMA warp).
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

from weights import make_state_dict  # noqa: E402
from diffusion_models_for_gravitational_waveform_reconstruction_b200 import UNet1D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--cin", type=int, default=3)
    ap.add_argument("--layer", type=int, default=1)
    ap.add_argument("--cta", type=int, default=0)
    ap.add_argument("--mode", type=int, default=0, help="ablation bits (results become wrong; timing only)")
    ap.add_argument("--brief", action="store_true")
    a = ap.parse_args()
    cc = 1 if a.cin == 3 else 5
    model = UNet1D(in_ch=a.cin, cond_in_ch=cc, use_selfcond=True, compute_dtype="bf16")
    model.load_state_dict(make_state_dict(a.cin, cc, seed=0))
    model = model.cuda().eval()
    eng = model.engine("bf16")
    x = torch.randn(a.B, a.cin, a.L, device="cuda")
    t = torch.randint(0, 1000, (a.B,), device="cuda")
    eng.forward(x, t)
    for k in list(eng._fuse_ok):
        if k[0] != a.layer:
            eng._fuse_ok[k] = False
    eng.forward(x, t)
    dbg = torch.zeros(148 * 16 * 8, dtype=torch.int64, device="cuda")
    eng.lib.gw_conv_gn_debug.argtypes = [C.c_void_p]
    eng.lib.gw_conv_gn_debug.restype = None
    eng.lib.gw_conv_gn_debug(C.c_void_p(dbg.data_ptr()))
    eng.lib.gw_conv_gn_debug_mode.argtypes = [C.c_int]
    eng.lib.gw_conv_gn_debug_mode.restype = None
    eng.lib.gw_conv_gn_debug_mode(a.mode)
    eng.forward(x, t)
    torch.cuda.synchronize()
    eng.lib.gw_conv_gn_debug(C.c_void_p(0))
    d = dbg.view(148, 16, 8).cpu()
    if a.brief:
        r = d[a.cta]
        n = sum(1 for it in range(16) if int(r[it, 0]))
        med = lambda v: sorted(v)[len(v) // 2]
        ph = [med([int(r[it, k + 1]) - int(r[it, k]) for it in range(1, n)]) for k in range(5)]
        per = med([int(r[it + 1, 0]) - int(r[it, 0]) for it in range(1, n - 1)])
        mma = med([int(r[it, 7]) - int(r[it, 6]) for it in range(1, n)])
        print(f"layer {a.layer} mode {a.mode}: wait {ph[0]} pass1 {ph[1]} xchg {ph[2]} coef {ph[3]} pass2 {ph[4]} | per sample {per} | mma issue {mma}")
        return
    for cta in (a.cta, a.cta + 1):
        r = d[cta]
        t0 = int(r[0, 0])
        print(f"CTA {cta}: cycles relative to the first loop top; columns: top accrdy pass1 xchg coef pass2 | mma_start mma_issued")
        for it in range(16):
            if int(r[it, 0]) == 0:
                break
            print(f"  sample {it:2d}: " + " ".join(f"{int(r[it, k]) - t0:8d}" for k in range(6)) + "  | " +
                  " ".join(f"{int(r[it, k]) - t0:8d}" for k in (6, 7)))
        dur = [int(r[it + 1, 0]) - int(r[it, 0]) for it in range(14) if int(r[it + 1, 0])]
        if dur:
            print(f"  cycles per sample (median): {sorted(dur)[len(dur) // 2]}")


if __name__ == "__main__":
    main()
