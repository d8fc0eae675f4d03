"""The batched sweep / grid harness (sweep_infer.py:203-354, grid_infer.py:372-432) against the same quantities
computed the reference's way: CPU oracle chain with identical injected noise, then the reference's numpy scoring formulas
(inference.py:11-27, sweep_infer.py:8-13, 225-241) restated in the test."""
import argparse
import json
import os

import numpy as np
import pytest
import torch

import oracle
from weights import gaussian, make_state_dict, synthetic_chirps

pytestmark = pytest.mark.gpu


def _corr(a, b):
    a = a - a.mean(); b = b - b.mean()
    return float(np.dot(a, b) / (np.sqrt((a * a).sum() * (b * b).sum()) + 1e-30))


def _score_last_window(x, c, fs, secs=0.8):
    L = min(len(x), len(c))
    x, c = np.asarray(x[:L], dtype=np.float64), np.asarray(c[:L], dtype=np.float64)
    t = np.arange(L) / fs
    m = t >= (t.max() - secs)
    return {"corr_last": _corr(x[m], c[m]), "mae_last": float(np.mean(np.abs(x[m] - c[m])))}


def _setup(B=3, L=1024):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import sweep as S
    d = synthetic_chirps(B, L, snr=10.0, seed=17)
    y_raw = (d["y_norm"][:, 0] * 3e-3 + 1e-3)
    clean_raw = (d["clean_norm"][:, 0] * 3e-3)
    sd = make_state_dict(3, 1, seed=0)
    m = UNet1D(in_ch=3, cond_in_ch=1, use_selfcond=True)
    m.load_state_dict(sd)
    return S, m.cuda().eval(), CustomDiffusion(T=1000, device="cuda"), S.Batch(y_raw, clean_raw, fs=4096.0), sd, y_raw, clean_raw


def test_eval_combo_matches_reference_style_evaluation():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening as W
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.inference import t_for_target_snr
    S, model, diff, batch, sd, y_raw, clean_raw = _setup()
    B, L, fs = batch.B, batch.L, batch.fs
    combo = dict(start_snr=2.0, cfg_scale=1.5, cfg_mode="gauss", cfg_center=0.7, cfg_width=0.12, dc_weight=0.05, init_mode="y-blend", eta=0.0)
    noise = torch.stack([gaussian((B, 1, L), seed=40 + k) for k in range(3)], 0)
    J, scores = S.eval_combo(model, diff, batch, combo, steps=6, noise=noise.cuda(), whiten_mode="train", sigma_mode="std")
    assert len(scores) == B
    # the reference's way, sample by sample
    y_w, c_w, P = W.whiten_train_like(batch.y_raw, batch.clean_raw)
    sig = W.sigma(y_w, "std").cpu().numpy()
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    start_t = t_for_target_snr(diff, 2.0)
    y_norm = (y_w.cpu() / torch.from_numpy(sig).float()[:, None])[:, None, :]
    x0n = oracle.ddim_sample(sd, cfg, ab, y_norm, T=1000, steps=6, eta=0.0, start_t=start_t, init_mode="y-blend", dc_weight=0.05,
                             cfg_scale=1.5, cfg_mode="gauss", cfg_center=0.7, cfg_width=0.12, cfg_u_only_thresh=0.05, noise=list(noise))
    Js = []
    for i in range(B):
        x0_white = (x0n[i, 0].double().numpy() * float(np.float32(sig[i]))).astype(np.float32)
        x0_strain = W._dewhiten_train_like(x0_white, P[i].cpu().numpy())
        ms = _score_last_window(x0_strain, clean_raw[i].numpy(), fs)
        w = int(fs * 0.8)
        ms["nmae_sigma"] = float(np.mean(np.abs(x0_strain[L - w:] - clean_raw[i].numpy()[L - w:]))) / (sig[i] + 1e-12)
        mw = _score_last_window(x0_white, c_w[i].cpu().numpy(), fs)
        Ji = ms["corr_last"] + 0.5 * mw["corr_last"] - 0.1 * ms["nmae_sigma"]
        Js.append(Ji)
        assert abs(scores[i][1]["corr_last"] - ms["corr_last"]) <= 2e-4 and abs(scores[i][2]["corr_last"] - mw["corr_last"]) <= 2e-4
        assert abs(scores[i][1]["nmae_sigma"] - ms["nmae_sigma"]) <= 1e-3 * abs(ms["nmae_sigma"]) + 1e-12
        assert abs(scores[i][0] - Ji) <= 5e-4 * max(1.0, abs(Ji))
    assert abs(J - float(np.mean(Js))) <= 5e-4 * max(1.0, abs(np.mean(Js)))


def test_grid_and_random_sweep_outputs(tmp_path):
    S, model, diff, batch, *_ = _setup(B=2, L=512)
    a = argparse.Namespace(grid_snr=[0.9, 2.2], grid_cfg=[1.5], grid_init=["y-blend", "scaled-noise"], grid_dc=[0.0], grid_eta=[0.0],
                           grid_steps=4, n_coarse=3, topk=2, steps_coarse=3, steps_refine=4, seeds_refine=2, seed=5,
                           start_snr_min=0.8, start_snr_max=3.0, cfg_min=1.0, cfg_max=3.0, cfg_mode="auto", cfg_center_min=0.55,
                           cfg_center_max=0.80, cfg_width_min=0.08, cfg_width_max=0.18, dc_choices=[0.0, 0.05], init_choices=["y-blend", "scaled-noise"],
                           eta_choices=[0.0])
    grid = S.grid_search(model, diff, batch, a, str(tmp_path / "g"))
    saved = json.load(open(tmp_path / "g" / "grid_results.json"))
    assert len(grid) == 4 and [g["J"] for g in saved] == sorted([g["J"] for g in saved], reverse=True)
    assert set(saved[0]) == {"start_snr", "cfg_scale", "cfg_mode", "cfg_center", "cfg_width", "dc_weight", "init_mode", "eta", "J"}
    assert all(g["cfg_mode"] == ("gauss" if g["init_mode"] == "y-blend" else "const") for g in saved)
    finals = S.random_sweep(model, diff, batch, a, str(tmp_path / "r"))
    top = json.load(open(tmp_path / "r" / "coarse_top.json"))
    fin = json.load(open(tmp_path / "r" / "final_results.json"))
    assert len(top) == 2 and len(fin) == 2 and {"J_coarse", "J_refine_mean", "J_refine_std"} <= set(fin[0])
    assert all(0.8 <= c["start_snr"] <= 3.0 and 1.0 <= c["cfg_scale"] <= 3.0 for c in top)
    again = S.random_sweep(model, diff, batch, a, str(tmp_path / "r2"))                  # a seed reproduces the combinations
    assert [c["start_snr"] for c in again] == [c["start_snr"] for c in finals]
    cmd = S.best_command(finals[0], a.steps_refine, input_h5="d.h5", index=0, model_path="m.pth", outdir="o", sigma_mode="std",
                         whiten=True, whiten_mode="train", amp=False)
    assert cmd[:2] == ["python", "inference.py"] and "--whiten" in cmd and "--start-snr" in cmd


def test_eval_indices_rows_and_csv(tmp_path):
    S, model, diff, batch, sd, y_raw, clean_raw = _setup(B=3, L=1024)
    knobs = dict(steps=5, eta=0.0, start_t=289, init_mode="scaled-noise", cfg_scale=1.0, cfg_mode="const", dc_weight=0.0)
    labels = {"m1": np.array([30.0, 35.0, 40.0]), "m2": np.array([20.0, 25.0, 30.0]), "q": None, "chirp_mass": None}
    rows = S.eval_indices(model, diff, batch, knobs, indices=[7, 8, 9], labels=labels, win="tail", tail_secs=0.1, seed=3,
                          csv_path=str(tmp_path / "per_index_metrics.csv"))
    assert [r["idx"] for r in rows] == [7, 8, 9] and rows[1]["m1"] == 35.0 and np.isnan(rows[0]["q"])
    for r in rows:
        assert np.isfinite(r["mae_last"]) and abs(r["J"] + 0.1 * r["nmae_sigma"]) <= 1e-12
        assert abs(r["nmae_clean"] * 0 + r["nmae_sigma"] * 0) == 0
    merger = S.eval_indices(model, diff, batch, knobs, win="merger", align="xcorr", seed=3)
    assert len(merger) == 3 and all(np.isfinite(r["mae_last"]) for r in merger)
    head = open(tmp_path / "per_index_metrics.csv").readline().strip().split(",")
    assert head == ["idx", "m1", "m2", "q", "chirp_mass", "corr_last", "mae_last", "nmae_sigma", "nmae_clean", "J"]
