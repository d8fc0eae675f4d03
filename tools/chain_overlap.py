#!/usr/bin/env python
"""Do chained layer kernels (gw_conv_gn3, programmatic stream serialization) really overlap?  Runs eager reverse steps with
%globaltimer stamps of every CTA's start / end and prints, per launch, first start, last start, first end, last end (us
relative to the first kernel of the step)."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

from weights import make_state_dict, synthetic_chirps  # noqa: E402
from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D  # noqa: E402
from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--L", type=int, default=4096)
    a = ap.parse_args()
    model = UNet1D(in_ch=3, cond_in_ch=1, use_selfcond=True, compute_dtype="bf16")
    model.load_state_dict(make_state_dict(3, 1, seed=0))
    model = model.cuda().eval()
    diff = CustomDiffusion(T=1000, device="cuda")
    plan = inf.make_sampler_plan(model, diff, a.B, a.L, T=1000, steps=1000, eta=1.0, seed=1, compute_dtype="bf16")
    y = synthetic_chirps(a.B, a.L, snr=10.0, seed=3)["y_norm"].cuda()
    plan.load_inputs(torch.randn(a.B, 1, a.L, device="cuda"), y, torch.zeros_like(y), None)
    for _ in range(3):
        plan.enqueue_step()
    torch.cuda.synchronize()
    lib = plan.eng.lib
    n_steps = 2
    buf = torch.zeros(n_steps * 6 * 160 * 2, dtype=torch.int64, device="cuda")
    lib.gw_conv_gn_tdebug.argtypes = [C.c_void_p]
    lib.gw_conv_gn_tdebug.restype = None
    lib.gw_conv_gn_tdebug(C.c_void_p(buf.data_ptr()))
    lib.gw_conv_gn_debug_mode.argtypes = [C.c_int]
    lib.gw_conv_gn_debug_mode.restype = None
    lib.gw_conv_gn_debug_mode(int(os.environ.get("GWB200_DBG", "0")))
    for _ in range(n_steps):
        plan.enqueue_step()
    torch.cuda.synchronize()
    lib.gw_conv_gn_tdebug(C.c_void_p(0))
    d = buf.view(n_steps * 6, 160, 2).cpu()
    names = ["enc1", "enc2", "mid", "dec0", "dec1", "dec2"]
    for s in range(n_steps):
        t0 = None
        for k in range(6):
            r = d[s * 6 + k]
            m = r[:, 0] > 0
            st, en = r[m, 0], r[m, 1]
            if t0 is None:
                t0 = int(st.min())
            print(f"step {s} {names[k]}: CTAs {int(m.sum()):3d} start {(int(st.min()) - t0) / 1e3:8.1f} .. {(int(st.max()) - t0) / 1e3:8.1f} us"
                  f"   end {(int(en.min()) - t0) / 1e3:8.1f} .. {(int(en.max()) - t0) / 1e3:8.1f} us")


if __name__ == "__main__":
    main()
