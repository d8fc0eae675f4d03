"""Batched evaluation harness: the hyper-parameter sweep of `sweep_infer.py` and the per-index evaluation of `grid_infer.py`
(SURVEY.md section 8f.3), with ONE batched reverse chain per hyper-parameter combination instead of the reference's nested
batch-1 loops (sweep_infer.py:203-259, 305-322; grid_infer.py:372-448).

  eval_combo(...)    <- sweep_infer.main.eval_combo (:203-244): J = mean over the samples of r_strain + 0.5 r_white - 0.1 NMAE_sigma
  grid_search(...)   <- the --grid branch (:247-289): grid_results.json, best_cmd.txt
  random_sweep(...)  <- stage A / stage B (:292-354): coarse_top.json, final_results.json, best_cmd.txt
  eval_indices(...)  <- grid_infer.main.eval_index (:372-432): one row of metrics per sample, per_index_metrics.csv

Whitening, sigma, the chain, de-whitening and the scores all run on the device (pipeline.reconstruct_batch, scoring.py); the
host only draws the combinations -- with the same RNG calls, in the same order, as the reference, so a seed reproduces its
combinations -- and writes the JSON / CSV files with the reference's keys.  HDF5 reading is `dataloader.open_h5`'s business:
the functions here take the raw strain arrays of the chosen indices.
"""
from __future__ import annotations

import csv
import json
import math
import os
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import pipeline, scoring
from .inference import t_for_target_snr

__all__ = ["Batch", "eval_combo", "sample_combo", "grid_search", "random_sweep", "eval_indices", "best_command"]


class Batch:
    """The samples a sweep is evaluated on: raw strain y [B, L] (+ clean, metadata, model PSD) of equal length, on the device."""

    def __init__(self, y_raw, clean_raw=None, fs: float = 4096.0, meta=None, P_model=None, device="cuda"):
        dev = torch.device(device)
        self.y_raw = torch.as_tensor(np.asarray(y_raw), dtype=torch.float32).to(dev) if not torch.is_tensor(y_raw) else y_raw.to(dev).float()
        self.clean_raw = None if clean_raw is None else (
            torch.as_tensor(np.asarray(clean_raw), dtype=torch.float32).to(dev) if not torch.is_tensor(clean_raw) else clean_raw.to(dev).float())
        self.fs, self.meta, self.P_model = float(fs), meta, P_model
        self.B, self.L = self.y_raw.shape[0], self.y_raw.shape[-1]


def _reconstruct(model, diffusion, batch: Batch, combo: Dict, steps: int, *, whiten_mode: Optional[str], sigma_mode: str,
                 sigma_fixed: float, amp: bool, seed: Optional[int], start_t: Optional[int] = None, noise=None):
    return pipeline.reconstruct_batch(
        model, diffusion, batch.y_raw, fs=batch.fs, clean_raw=batch.clean_raw, meta=batch.meta, whiten=bool(whiten_mode),
        whiten_mode=whiten_mode or "train", P_model=batch.P_model, sigma_mode=sigma_mode, sigma_fixed=sigma_fixed,
        start_snr=None if start_t is not None else combo.get("start_snr"), start_t=start_t, steps=int(steps),
        eta=float(combo.get("eta", 0.0)), init_mode=str(combo.get("init_mode", "scaled-noise")), x0_std_est=0.14,
        dc_weight=float(combo.get("dc_weight", 0.0)), cfg_scale=float(combo.get("cfg_scale", 1.0)),
        cfg_mode=str(combo.get("cfg_mode", "const")), cfg_center=float(combo.get("cfg_center", 0.7)),
        cfg_width=float(combo.get("cfg_width", 0.12)), cfg_u_only_thresh=0.05, score_secs=0.8,
        seed=0 if seed is None else int(seed), noise=noise, compute_dtype="bf16" if amp else None)


@torch.no_grad()
def eval_combo(model, diffusion, batch: Batch, combo: Dict, steps: int, seed: Optional[int] = None, *, whiten_mode: Optional[str] = "train",
               sigma_mode: str = "std", sigma_fixed: float = 1.0, amp: bool = False, noise=None) -> Tuple[float, List[Tuple]]:
    """sweep_infer.eval_combo: (J averaged over the samples, [(J_i, m_strain_i, m_white_i)]) for one hyper-parameter combination."""
    if seed is not None:
        torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    r = _reconstruct(model, diffusion, batch, combo, steps, whiten_mode=whiten_mode, sigma_mode=sigma_mode, sigma_fixed=sigma_fixed,
                     amp=amp, seed=seed, noise=noise)
    if "strain" not in r:
        return -1e9, []
    ms, mw = r["strain"], r.get("white")
    J = scoring.objective(ms, mw)
    cols_s = {k: ms[k].cpu().tolist() for k in ("corr_last", "mae_last", "nmae_sigma")}
    cols_w = {k: mw[k].cpu().tolist() for k in ("corr_last", "mae_last")} if mw is not None else None
    Jl = J.cpu().tolist()
    scores = [(float(Jl[i]), {k: float(v[i]) for k, v in cols_s.items()},
               ({k: float(v[i]) for k, v in cols_w.items()} if cols_w else None)) for i in range(batch.B)]
    return (float(np.mean(Jl)) if scores else -1e9), scores


def sample_combo(a) -> Dict:
    """One random combination: the RNG calls of sweep_infer.py:293-305 in the same order (global `random` / `np.random`)."""
    cfg_mode = a.cfg_mode
    if cfg_mode == "auto":
        cfg_mode = "gauss" if (random.random() < 0.7) else "const"
    return dict(
        start_snr=10 ** np.random.uniform(math.log10(a.start_snr_min), math.log10(a.start_snr_max)),
        cfg_scale=np.random.uniform(a.cfg_min, a.cfg_max),
        cfg_mode=cfg_mode,
        cfg_center=np.random.uniform(a.cfg_center_min, a.cfg_center_max),
        cfg_width=np.random.uniform(a.cfg_width_min, a.cfg_width_max),
        dc_weight=float(random.choice(a.dc_choices)),
        init_mode=random.choice(a.init_choices),
        eta=float(random.choice(a.eta_choices)),
    )


def best_command(best: Dict, steps: int, *, input_h5: str, index: int, model_path: str, outdir: str, sigma_mode: str,
                 whiten: bool, whiten_mode: str, amp: bool) -> List[str]:
    """The `inference.py` command line the reference writes to best_cmd.txt (sweep_infer.py:267-284, 331-348)."""
    cmd = ["python", "inference.py", "--input-h5", input_h5, "--index", str(index), "--model", model_path, "--outdir",
           os.path.join(outdir, "best"), "--steps", str(steps), "--eta", f"{best['eta']:.2f}", "--start-snr", f"{best['start_snr']:.3f}",
           "--init-mode", best["init_mode"], "--cfg-scale", f"{best['cfg_scale']:.2f}", "--cfg-mode", best["cfg_mode"],
           "--cfg-center", f"{best['cfg_center']:.2f}", "--cfg-width", f"{best['cfg_width']:.2f}", "--dc-weight",
           f"{best['dc_weight']:.2f}", "--sigma-mode", sigma_mode]
    if whiten:
        cmd += ["--whiten", "--whiten-mode", whiten_mode]
    if amp:
        cmd += ["--amp"]
    return cmd


def _jsonable(d: Dict) -> Dict:
    return {k: (float(v) if isinstance(v, (np.floating, np.integer)) else v) for k, v in d.items()}


def grid_search(model, diffusion, batch: Batch, a, outdir: str, **kw) -> List[Dict]:
    """The --grid branch (sweep_infer.py:247-289).  `a` carries grid_snr / grid_cfg / grid_init / grid_dc / grid_eta / grid_steps."""
    os.makedirs(outdir, exist_ok=True)
    grid = []
    for snr in a.grid_snr:
        for cfg in a.grid_cfg:
            for init in a.grid_init:
                for dc in a.grid_dc:
                    for et in a.grid_eta:
                        combo = dict(start_snr=float(snr), cfg_scale=float(cfg), cfg_mode=("gauss" if init == "y-blend" else "const"),
                                     cfg_center=0.70, cfg_width=0.12, dc_weight=float(dc), init_mode=init, eta=float(et))
                        J, _ = eval_combo(model, diffusion, batch, combo, steps=a.grid_steps, **kw)
                        grid.append({**combo, "J": J})
    grid = sorted(grid, key=lambda z: z["J"], reverse=True)
    with open(os.path.join(outdir, "grid_results.json"), "w") as fh:
        json.dump([_jsonable(g) for g in grid], fh, indent=2)
    return grid


def random_sweep(model, diffusion, batch: Batch, a, outdir: str, **kw) -> List[Dict]:
    """Stage A (n_coarse random combinations at steps_coarse) and stage B (top-k refined at steps_refine over seeds_refine
    seeds); sweep_infer.py:292-329."""
    os.makedirs(outdir, exist_ok=True)
    random.seed(a.seed); np.random.seed(a.seed); torch.manual_seed(a.seed)
    coarse = []
    for _ in range(a.n_coarse):
        c = sample_combo(a)
        J, _s = eval_combo(model, diffusion, batch, c, steps=a.steps_coarse, **kw)
        coarse.append({**c, "J_coarse": J})
    coarse = sorted(coarse, key=lambda z: z["J_coarse"], reverse=True)
    top = coarse[: a.topk]
    with open(os.path.join(outdir, "coarse_top.json"), "w") as fh:
        json.dump([_jsonable(t) for t in top], fh, indent=2)
    finals = []
    for c in top:
        JJ = [eval_combo(model, diffusion, batch, {k: v for k, v in c.items() if k != "J_coarse"}, steps=a.steps_refine,
                         seed=a.seed + s, **kw)[0] for s in range(a.seeds_refine)]
        finals.append({**c, "J_refine_mean": float(np.mean(JJ)), "J_refine_std": float(np.std(JJ))})
    finals = sorted(finals, key=lambda z: z["J_refine_mean"], reverse=True)
    with open(os.path.join(outdir, "final_results.json"), "w") as fh:
        json.dump([_jsonable(f) for f in finals], fh, indent=2)
    return finals


@torch.no_grad()
def eval_indices(model, diffusion, batch: Batch, knobs: Dict, *, indices: Optional[Sequence[int]] = None, labels: Optional[Dict] = None,
                 win: str = "tail", tail_secs: float = 0.8, left: float = 0.08, right: float = 0.04, align: str = "none",
                 align_max_shift_s: float = 0.02, whiten_mode: Optional[str] = "train", sigma_mode: str = "std",
                 sigma_fixed: float = 1.0, amp: bool = False, seed: Optional[int] = None, noise=None,
                 csv_path: Optional[str] = None) -> List[Dict]:
    """grid_infer.eval_index for a whole batch: one reverse chain, then per sample the window (full | tail | merger), the
    alignment (none | peak | xcorr), MAE / NMAE_sigma / NMAE_clean and J (grid_infer.py:372-432).  `labels`: arrays m1, m2, q,
    chirp_mass for the CSV columns."""
    r = _reconstruct(model, diffusion, batch, knobs, int(knobs.get("steps", 300)), whiten_mode=whiten_mode, sigma_mode=sigma_mode,
                     sigma_fixed=sigma_fixed, amp=amp, seed=seed, start_t=knobs.get("start_t"), noise=noise)
    xs = r["x0_hat_strain"].double()
    sig = r["sigma"].double().cpu().numpy()
    fs = batch.fs
    rows = []
    for i in range(batch.B):
        idx = int(indices[i]) if indices is not None else i
        lab = {k: (float(labels[k][i]) if labels and k in labels and labels[k] is not None else float("nan"))
               for k in ("m1", "m2", "q", "chirp_mass")}
        if batch.clean_raw is None:
            rows.append(dict(idx=idx, **lab, corr_last=float("nan"), mae_last=float("nan"), nmae_sigma=float("nan"), J=float("nan")))
            continue
        a_, b_ = xs[i], batch.clean_raw[i].double()
        L = a_.numel()
        if win == "full":
            s, e = 0, L
        elif win == "tail":
            W = int(max(1, tail_secs * fs)); s, e = max(0, L - W), L
        else:
            pk = int(torch.argmax(b_.abs()))
            s, e = int(max(0, pk - left * fs)), int(min(L, pk + right * fs))
        aw, bw = a_[s:e], b_[s:e]
        if align != "none":
            if align == "peak":
                k = int(torch.argmax(bw.abs())) - int(torch.argmax(aw.abs()))
                if k > 0:
                    aw, bw = aw[: aw.numel() - k], bw[k:]
                elif k < 0:
                    aw, bw = aw[-k:], bw[: bw.numel() + k]
            else:
                k = int(scoring.best_lag_by_xcorr(aw[None].float(), bw[None].float(), max_shift=int(max(1, align_max_shift_s * fs)))[0])
                if k > 0:
                    aw, bw = aw[k:], bw[: bw.numel() - k]
                elif k < 0:
                    aw, bw = aw[: aw.numel() + k], bw[-k:]
            n = min(aw.numel(), bw.numel())
            aw, bw = aw[:n], bw[:n]
        mae = float((aw - bw).abs().mean())
        nmae_sigma = mae / (float(sig[i]) + 1e-12)
        nmae_clean = mae / (float(bw.abs().mean()) + 1e-12)
        rows.append(dict(idx=idx, **lab, corr_last=float("nan"), mae_last=mae, nmae_sigma=nmae_sigma, nmae_clean=nmae_clean,
                         J=float(0.0 - 0.1 * nmae_sigma)))
    if csv_path:
        os.makedirs(os.path.dirname(csv_path) or ".", exist_ok=True)
        cols = ["idx", "m1", "m2", "q", "chirp_mass", "corr_last", "mae_last", "nmae_sigma", "nmae_clean", "J"]
        with open(csv_path, "w", newline="") as fh:
            w = csv.DictWriter(fh, fieldnames=cols, extrasaction="ignore")
            w.writeheader()
            for row in rows:
                w.writerow(row)
    return rows
