#!/usr/bin/env python
"""Per-layer parity diagnostics on the GPU box: CUDA engine vs the CPU oracle (test infrastructure).

Usage: python tools/gpu_diag.py [--mode fp32|bf16_simt|bf16_tc|all] [--L 1024] [--B 2] [--cin 3|7]
Prints one line per tensor with rel-L2 and max-abs errors; exit code 0 even on mismatch (it is a report).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

import oracle  # noqa: E402
from weights import make_state_dict, gaussian  # noqa: E402
from diffusion_models_for_gravitational_waveform_reconstruction_b200.engine import ModelSpec, UNetEngine  # noqa: E402


def rel(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    l2 = float((a - b).norm() / (b.norm() + 1e-30))
    mx = float((a - b).abs().max() / (b.abs().max() + 1e-30))
    return l2, mx


def run(mode, L, B, cin, out):
    cc = 1 if cin == 3 else 5
    sd = make_state_dict(in_ch=cin, cond_in_ch=cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=cin, cond_in_ch=cc, use_selfcond=True)
    spec = ModelSpec(in_ch=cin, cond_in_ch=cc, use_selfcond=True)
    x = gaussian((B, cin, L), seed=100 + L + cin)
    t = torch.tensor(([24, 999, 500, 0] * B)[:B])
    with torch.no_grad():
        taps = oracle.unet_forward_taps(sd, cfg, x, t)
        films = oracle.film_vectors(sd, cfg, t)
    dtype, impl = {"fp32": ("fp32", "simt"), "bf16_simt": ("bf16", "simt"), "bf16_tc": ("bf16", "tc")}[mode]
    params = {k: v.cuda() for k, v in sd.items()}
    eng = UNetEngine(params, spec, dtype=dtype, conv_impl=impl)
    t0 = time.time()
    eps = eng.forward(x.cuda(), t.cuda(), keep_raw=True)
    torch.cuda.synchronize()
    res = {"mode": mode, "L": L, "B": B, "cin": cin, "sec": time.time() - t0, "layers": {}}
    film = eng.film_vectors(t.cuda())
    l2, mx = rel(film, torch.cat(films, dim=1))
    res["layers"]["film"] = (l2, mx)
    ws = eng.workspace(B, L, True)
    names = ["enc0", "enc1", "enc2", "mid", "dec0", "dec1", "dec2"]
    for li, n in enumerate(names):
        res["layers"][n + ".raw"] = rel(ws.raw[li].float().transpose(1, 2), taps[n + ".raw"])
        res["layers"][n + ".out"] = rel(ws.out[li].float().transpose(1, 2), taps[n + ".out"])
    res["layers"]["eps"] = rel(eps, taps["eps"])
    for k, (a, b) in res["layers"].items():
        print(f"[{mode} L={L} B={B} cin={cin}] {k:10s} rel_l2={a:.3e} max_rel={b:.3e}")
    out.append(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="all")
    ap.add_argument("--L", type=int, default=1024)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--cin", type=int, default=3)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    modes = ["fp32", "bf16_simt", "bf16_tc"] if a.mode == "all" else [a.mode]
    out = []
    for m in modes:
        try:
            run(m, a.L, a.B, a.cin, out)
        except Exception as e:  # report and continue
            print(f"[{m}] FAILED: {type(e).__name__}: {e}")
            out.append({"mode": m, "error": str(e)})
    if a.json:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        with open(a.json, "a") as fh:
            for r in out:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
