#!/usr/bin/env python
"""CPU half of BASELINE config 5's parity check: the first K injections of an SNR sweep (saved by
`tools/snr_sweep.py --save-first file`, on any number of GPUs) are re-done with the CPU oracle -- the reference restated --
from the same x_T, and compared: overlap <a,b>/(|a||b|), the reference's Pearson `_corr` (inference.py:15-18) on the tail
window, rel-L2.  Runs anywhere (no GPU): the injections are regenerated from their seeds."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

import oracle  # noqa: E402
from weights import make_state_dict  # noqa: E402
from snr_sweep import corr, injections, overlap  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file")
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=32)
    a = ap.parse_args()
    d = torch.load(a.file)
    k = min(a.k, d["recon"].shape[0])
    sd = make_state_dict(3, 1, seed=0)
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    torch.set_num_threads(os.cpu_count() or 1)
    inj = injections(0, k, d["length"], d["seed"])
    t0 = time.time()
    refs = []
    for c0 in range(0, k, a.chunk):
        sl = slice(c0, min(k, c0 + a.chunk))
        refs.append(oracle.ddim_sample(sd, cfg, ab, inj["y_norm"][sl], T=1000, steps=d["steps"], eta=d["eta"], start_t=d["start_t"],
                                       noise=[d["xT"][sl]]))
    ref = torch.cat(refs)
    rec = d["recon"][:k]
    tail = slice(int(0.6 * d["length"]), d["length"])
    ov = overlap(rec, ref)
    res = {"file": os.path.basename(a.file), "k": k, "sweep_n": d["n"], "sweep_world": d["world"], "steps": d["steps"],
           "dtype": d["dtype"], "cpu_seconds": time.time() - t0, "cpu_cores": os.cpu_count(),
           "overlap_vs_oracle_min": float(ov.min()), "overlap_vs_oracle_mean": float(ov.mean()),
           "tail_corr_vs_oracle_min": float(corr(rec[:, :, tail], ref[:, :, tail]).min()),
           "rel_l2_max": float(((rec - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max()),
           "snr_range": [float(inj["snr"].min()), float(inj["snr"].max())]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
