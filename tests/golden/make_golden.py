#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Usage (build container only -- /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Recipe: put the reference's flat-import directory on sys.path, stub the
missing `h5py` module, import `models`, `inference`, `train` unchanged.  Noise is injected by
temporarily replacing `torch.randn` / `torch.randn_like` (the reference draws from the global
RNG).  Weights/inputs come from tests/golden/weights.py (numpy PCG64) so the tests
can regenerate them anywhere.  Nothing from the reference is copied into the repo: only its
numerical outputs are stored.
"""
from __future__ import annotations

import argparse
import os
import sys
import types
from copy import deepcopy

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from weights import make_state_dict, synthetic_chirps, gaussian  # noqa: E402

REF = "/root/reference/src/snr_denoising"


def import_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import models as ref_models          # noqa
    import inference as ref_inf          # noqa
    import train as ref_train            # noqa
    return ref_models, ref_inf, ref_train


class NoiseInjector:
    """Replace torch.randn / torch.randn_like by a queue of pre-drawn tensors."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.k = 0

    def _next(self, shape):
        z = self.draws[self.k]
        self.k += 1
        assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        return z.clone()

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *s, **kw: self._next(s[0] if len(s) == 1 and not isinstance(s[0], int) else s)
        torch.randn_like = lambda x, **kw: self._next(x.shape)
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def sub(x: torch.Tensor) -> np.ndarray:
    """Strided subsample that keeps fixtures small: every 4th channel, every 4th position."""
    return x.detach()[:, ::4, ::4].contiguous().numpy()


def build_model(M, in_ch, cond_in_ch, seed=0, base_ch=64):
    model = M.UNet1D(in_ch=in_ch, base_ch=base_ch, cond_in_ch=cond_in_ch, use_selfcond=True)
    sd = make_state_dict(in_ch=in_ch, cond_in_ch=cond_in_ch, base_ch=base_ch, seed=seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, sd


def gen_schedule(M, I, out):
    d = {}
    diff = M.CustomDiffusion(T=1000)
    d["betas"] = diff.betas.numpy()
    d["alpha_bar"] = diff.alpha_bar.numpy()
    d50 = M.CustomDiffusion(T=50)
    d["alpha_bar_T50"] = d50.alpha_bar.numpy()
    for (T, steps, st) in [(1000, 1000, None), (1000, 50, None), (1000, 10, 289), (1000, 7, 529), (1000, 200, 100),
                           (1000, 1, None), (50, 50, None), (1000, 3, 0)]:
        ts = I._build_t_schedule(T, steps, torch.device("cpu"), st)
        d[f"sched_{T}_{steps}_{st}"] = ts.numpy()
    tt = torch.tensor([0, 1, 24, 55, 289, 529, 998, 999])
    d["temb_t"] = tt.numpy()
    d["temb_128"] = M.TimeEmbedding(128, 999.0)(tt).numpy()
    d["temb_33"] = M.TimeEmbedding(33, 49.0)(tt).numpy()
    w = []
    for mode in ["const", "tophat", "gauss"]:
        for i in [0, 3, 5, 9]:
            w.append(I._cfg_weight(i, 10, mode, 1.5, 0.5, 0.3))
    d["cfg_w"] = np.array(w, dtype=np.float64)
    d["t_for_snr"] = np.array([I.t_for_target_snr(diff, s) for s in [0.9, 2.0, 10.0, 20.0]])
    d["snr_tab"] = I.snr_from_alpha_bar(diff.alpha_bar)
    np.savez_compressed(os.path.join(out, "schedule.npz"), **d)


def gen_forward(M, out):
    for tag, in_ch, cc, L, B in [("c3_L256", 3, 1, 256, 2), ("c7_L256", 7, 5, 256, 2), ("c3_L500", 3, 1, 500, 1),
                                 ("c7_L1024", 7, 5, 1024, 1)]:
        model, _ = build_model(M, in_ch, cc, seed=0)
        x = gaussian((B, in_ch, L), seed=100 + L + in_ch)
        if cc == 5:   # metadata channels are tiled scalars (dataloader.py:219-222)
            x[:, 2:6, :] = x[:, 2:6, :1].clone()
        t = torch.tensor([24, 999][:B] if B == 2 else [529])
        rec = {}
        hooks = []

        def mk(name):
            def h(mod, inp, outp):
                rec[name + ".in"] = sub(inp[0])
                rec[name + ".raw"] = sub(outp)
            return h
        for i in range(3):
            hooks.append(model.encoders[i][0].register_forward_hook(mk(f"enc{i}")))
            hooks.append(model.decoders[i][0].register_forward_hook(mk(f"dec{i}")))
        hooks.append(model.mid[0].register_forward_hook(mk("mid")))
        hooks.append(model.final.register_forward_hook(mk("final")))
        with torch.no_grad():
            eps = model(x, t)
        for h in hooks:
            h.remove()
        rec["eps"] = eps.numpy()
        rec["t"] = t.numpy()
        np.savez_compressed(os.path.join(out, f"forward_{tag}.npz"), **rec)


CHAIN_CASES = {
    # tag: kwargs
    "ddim10_s289": dict(steps=10, eta=0.0, start_t=289, init_mode="noise", dc_weight=0.0, cfg_scale=1.0),
    "ddpm12_full": dict(steps=12, eta=1.0, start_t=None, init_mode="noise", dc_weight=0.0, cfg_scale=1.0),
    "ddpm_all_s40": dict(steps=41, eta=1.0, start_t=40, init_mode="scaled-noise", dc_weight=0.0, cfg_scale=1.0),
    "cfg15_dc": dict(steps=10, eta=0.5, start_t=529, init_mode="y-blend", dc_weight=0.05, cfg_scale=1.5),
    "cfg_tophat": dict(steps=8, eta=0.0, start_t=289, init_mode="noise", dc_weight=0.0, cfg_scale=2.0,
                       cfg_mode="tophat", cfg_center=0.5, cfg_width=0.5),
    "uonly": dict(steps=6, eta=1.0, start_t=289, init_mode="noise", dc_weight=0.0, cfg_scale=0.0,
                  cfg_u_only_thresh=0.0),
    "x0pred": dict(steps=6, eta=0.3, start_t=289, init_mode="noise", dc_weight=0.0, cfg_scale=1.0,
                   pred_type="x0", eps_scale=0.9, cond_scale=1.1),
}


def gen_chains(M, I, out):
    L = 256
    for in_ch, cc in [(3, 1), (7, 5)]:
        model, _ = build_model(M, in_ch, cc, seed=1)
        diff = M.CustomDiffusion(T=1000)
        data = synthetic_chirps(2, L, snr=10.0, seed=77)
        y = data["y_norm"]
        if cc == 5:
            meta = gaussian((2, 4, 1), seed=5).expand(2, 4, L).contiguous() * 0.3
            cond = torch.cat([y, meta], dim=1)
        else:
            cond = y
        for tag, kw in CHAIN_CASES.items():
            if cc == 5 and tag not in ("cfg15_dc", "ddim10_s289"):
                continue
            full = dict(T=1000, device=torch.device("cpu"), length=L, debug=False, x0_std_est=0.14,
                        cond_scale=1.0, eps_scale=1.0, pred_type="eps", in_ch=in_ch, cond_in_ch=cc,
                        use_selfcond=True, cfg_mode="const", cfg_center=0.5, cfg_width=0.3,
                        cfg_u_only_thresh=0.0)
            full.update(kw)
            outs, xin, eh = [], [], []
            for b in range(2):   # the reference sampler is batch-1 only
                draws = [gaussian((1, 1, L), seed=9000 + 100 * b + k) for k in range(64)]
                calls_x, calls_o = [], []
                orig_forward = model.forward

                def spy(xx, tt, _f=orig_forward):
                    o = _f(xx, tt)
                    calls_x.append(xx[:, :1].clone())
                    calls_o.append(o.clone())
                    return o
                model.forward = spy
                with NoiseInjector(draws) as inj:
                    xr = I.ddim_sample(model, diff, cond[b:b + 1], **full)
                    n_draws = inj.k
                model.forward = orig_forward
                outs.append(xr)
                xin.append(torch.cat(calls_x, 0))
                eh.append(torch.cat(calls_o, 0))
            np.savez_compressed(os.path.join(out, f"chain_c{in_ch}_{tag}.npz"),
                                x_final=torch.cat(outs, 0).numpy(),
                                fwd_x=torch.stack(xin, 0).numpy(),      # [B, n_forward_calls, 1, L]
                                fwd_out=torch.stack(eh, 0).numpy(),
                                n_draws=np.array(n_draws))


def gen_train(M, TR, out):
    import torch.optim as optim
    L, B = 256, 4
    for in_ch, cc in [(7, 5), (3, 1)]:
        model, sd0 = build_model(M, in_ch, cc, seed=2)
        model.train()
        diff = M.CustomDiffusion(T=1000)
        data = synthetic_chirps(B, L, snr=12.0, seed=31)
        clean, y = data["clean_norm"], data["y_norm"]
        mask = torch.ones(B, 1, L)
        mask[1, :, :37] = 0.0                        # left-padded sample (dataloader.py:248-268)
        if cc == 5:
            meta = gaussian((B, 4, 1), seed=6).expand(B, 4, L).contiguous() * 0.3
            cond = torch.cat([y, meta], dim=1)
        else:
            cond = y
        t = torch.tensor([500, 731, 999, 612])
        eps = gaussian((B, 1, L), seed=41)
        drop = torch.tensor([0.0, 1.0, 0.0, 0.0]).view(B, 1, 1)
        for sc in (False, True):
            m = deepcopy(model)
            ema = deepcopy(m)
            opt = optim.AdamW(m.parameters(), lr=2e-4, weight_decay=1e-4)
            sched = TR.make_warmup_cosine_scheduler(opt, warmup_steps=10, total_steps=100, min_lr_scale=0.1)
            rec = {}
            for step in range(2):
                clean_c = clean.clamp(-10, 10)
                y_c = y.clamp(-10, 10)
                with NoiseInjector([eps]):
                    x_t, e = diff.q_sample(clean_c, t)
                x_t = x_t.clamp(-10, 10)
                if cc == 5:
                    cond_used = torch.cat([y_c * (1.0 - drop), cond[:, 1:]], dim=1)
                else:
                    cond_used = cond * (1.0 - drop)
                if sc:
                    x0_sc = TR._predict_x0_norm(m, diff, x_t, cond_used, t)
                else:
                    x0_sc = torch.zeros_like(x_t)
                eps_hat = m(torch.cat([x_t, cond_used, x0_sc], dim=1), t)
                el = TR._element_loss(eps_hat, e, mask, "huber", 0.5)
                loss = (el.sum(dim=[1, 2]) / mask.sum(dim=[1, 2]).clamp_min(1.0)).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                if step == 0:
                    rec["loss"] = loss.detach().numpy()
                    rec["eps_hat"] = eps_hat.detach().numpy()
                    for k, p in m.named_parameters():
                        g = p.grad.detach()
                        rec["gnorm/" + k] = np.array(float(g.norm()))
                        if g.numel() <= 4096:
                            rec["grad/" + k] = g.numpy().copy()
                        else:
                            rec["grad/" + k] = g.reshape(-1)[::97].numpy().copy()
                gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                rec[f"grad_norm{step}"] = np.array(float(gn))
                rec[f"lr{step}"] = np.array(opt.param_groups[0]["lr"])
                opt.step()
                sched.step()
                TR.update_ema(ema, m, 0.999)
                rec[f"loss{step}"] = loss.detach().numpy()
            for k, p in m.named_parameters():
                if p.numel() <= 4096:
                    rec["p2/" + k] = p.detach().numpy().copy()
                    rec["ema2/" + k] = dict(ema.named_parameters())[k].detach().numpy().copy()
                else:
                    rec["p2/" + k] = p.detach().reshape(-1)[::97].numpy().copy()
                    rec["ema2/" + k] = dict(ema.named_parameters())[k].detach().reshape(-1)[::97].numpy().copy()
            np.savez_compressed(os.path.join(out, f"train_c{in_ch}_sc{int(sc)}.npz"), **rec)
    lam = [TR.make_warmup_cosine_scheduler(optim.SGD([torch.zeros(1, requires_grad=True)], lr=1.0), 10, 100, 0.1)
           .lr_lambdas[0](s) for s in [0, 5, 9, 10, 50, 99, 100, 150]]
    np.savez_compressed(os.path.join(out, "lr_lambda.npz"), lam=np.array(lam))


def gen_proxy_and_helpers(M, I, TR, out):
    """one_step_proxy_like_test_infer (inference.py:317-371), _match_batch and _sample_timesteps_stratified
    (train.py:133-172) outputs of the unmodified reference."""
    L = 512
    rec = {}
    for in_ch, cc in [(3, 1), (7, 5)]:
        model, _ = build_model(M, in_ch, cc, seed=3)
        diff = M.CustomDiffusion(T=1000)
        data = synthetic_chirps(1, L, snr=10.0, seed=88)
        cond = data["y_norm"]
        if cc == 5:
            cond = torch.cat([cond, gaussian((1, 4, 1), seed=8).expand(1, 4, L).contiguous() * 0.3], dim=1)
        z = gaussian((1, 1, L), seed=99)
        for cfg, snr in [(1.0, 2.0), (1.5, 10.0)]:
            with NoiseInjector([z]):
                x0 = I.one_step_proxy_like_test_infer(model, diff, data["clean_norm"], cond, 1.7, snr, torch.device("cpu"),
                                                      in_ch, cc, True, cfg, True, cond_scale=0.9, eps_scale=1.1)
            rec[f"proxy_c{in_ch}_cfg{cfg}_snr{snr}"] = x0.numpy()
    a = torch.arange(6).view(3, 2).float()
    rec["match_3_to_6"] = TR._match_batch(a, 6).numpy()
    rec["match_3_to_7"] = TR._match_batch(a, 7).numpy()
    rec["match_3_to_3"] = TR._match_batch(a, 3).numpy()
    torch.manual_seed(123)
    ts = TR._sample_timesteps_stratified(64, 500, 999, torch.device("cpu"), bins=8)
    rec["strat_sorted_64_8"] = np.sort(ts.numpy())           # the draw itself is RNG-order dependent; the strata are not
    edges = torch.linspace(500, 1000, 9).long().numpy()
    rec["strat_edges"] = edges
    rec["strat_counts"] = np.histogram(ts.numpy(), bins=edges)[0]
    np.savez_compressed(os.path.join(out, "proxy_helpers.npz"), **rec)


def gen_scores(I, out):
    """Scores of the unmodified reference helpers (inference.py:11-27, 247-279 and the window MAE of :303-314,
    sweep_infer.py:8-13, 232-237) on synthetic (clean, reconstruction) pairs."""
    sys.path.insert(0, REF)
    import sweep_infer as SW
    L, B, fs = 2048, 6, 4096.0
    d = synthetic_chirps(B, L, snr=12.0, seed=55)
    clean = d["clean_norm"][:, 0].numpy()
    rng = np.random.default_rng(9)
    rec = {"sigma": (0.5 + rng.uniform(size=B)).astype(np.float32)}
    xh = np.zeros_like(clean)
    for b in range(B):
        shift = int(rng.integers(-7, 8))
        xh[b] = np.roll(clean[b], shift) * (0.8 + 0.05 * b) + 0.02 * rng.standard_normal(L).astype(np.float32)
    rec["xhat"] = xh
    cols = {k: [] for k in ["corr_last", "mae_last", "nmae_sigma", "best_lag_full", "best_lag_64", "xc_mae", "xc_nmae_clean",
                            "xc_nmae_sigma", "objective"]}
    for b in range(B):
        m = I._score_last_window(xh[b], clean[b], fs, secs=0.2)
        w = int(fs * 0.2)
        nm = float(np.mean(np.abs(xh[b][L - w:] - clean[b][L - w:]))) / (float(rec["sigma"][b]) + 1e-12)
        cols["corr_last"].append(m["corr_last"]); cols["mae_last"].append(m["mae_last"]); cols["nmae_sigma"].append(nm)
        cols["best_lag_full"].append(I._best_lag_by_xcorr(clean[b], xh[b], 0))
        cols["best_lag_64"].append(I._best_lag_by_xcorr(clean[b], xh[b], 64))
        a_al, b_al, t_a = I._align_xcorr(clean[b], xh[b], 1.0 / fs, max_shift=64)
        mask = (t_a >= -0.080) & (t_a <= 0.040)
        mae = float(np.mean(np.abs(b_al[mask] - a_al[mask])))
        cols["xc_mae"].append(mae)
        cols["xc_nmae_clean"].append(mae / (float(np.mean(np.abs(a_al[mask]))) + 1e-12))
        cols["xc_nmae_sigma"].append(mae / (float(rec["sigma"][b]) + 1e-12))
        cols["objective"].append(SW._objective({"corr_last": m["corr_last"], "nmae_sigma": nm}, {"corr_last": m["corr_last"]}))
    for k, v in cols.items():
        rec[k] = np.array(v, dtype=np.float64)
    np.savez_compressed(os.path.join(out, "scores.npz"), **rec)


def gen_whitening(I, out):
    """Outputs of the unmodified reference whitening / sigma helpers (inference.py:36-38, 125-205) on coloured synthetic data."""
    rng = np.random.default_rng(21)
    fs = 4096.0
    rec = {}
    for L in (2048, 1000):
        n = rng.standard_normal(L + 64)
        col = np.convolve(n, np.hanning(33) / np.hanning(33).sum(), mode="valid")[:L]       # coloured noise
        clean = synthetic_chirps(1, L, snr=8.0, seed=70 + L)["clean_norm"][0, 0].numpy().astype(np.float64) * 0.3
        y = (col * 3.0 + clean + 0.7).astype(np.float32)
        x = clean.astype(np.float32)
        y_w, x_w, P = I._whiten_pair_train_like(y, x, fs)
        back = I._dewhiten_train_like(y_w, P)
        P_model = 1e-3 + np.abs(np.sin(np.linspace(0, 3, 513))) ** 2 + np.linspace(0, 1, 513)
        ym, xm, Pm = I._whiten_pair_model(y, x, P_model, fs)
        backm = I._dewhiten_model(ym, Pm)
        rec.update({f"y_{L}": y, f"x_{L}": x, f"yw_{L}": y_w, f"xw_{L}": x_w, f"P_{L}": P, f"back_{L}": back,
                    f"Pmodel_{L}": P_model, f"ym_{L}": ym, f"xm_{L}": xm, f"Pm_{L}": Pm, f"backm_{L}": backm,
                    f"sig_std_{L}": np.array(I._pick_sigma(y, "std", 1.0)), f"sig_mad_{L}": np.array(I._pick_sigma(y, "mad", 1.0)),
                    f"sig_fixed_{L}": np.array(I._pick_sigma(y, "fixed", 2.5))})
    np.savez_compressed(os.path.join(out, "whitening.npz"), **rec)


def gen_whitening_welch(I, out):
    """The Welch variant (inference.py:161-179: scipy.signal.welch inside `_whiten_pair_welch`, `_dewhiten_welch`) of the
    unmodified reference on the inputs of gen_whitening, plus a multi-segment length; and scipy's welch itself on a batch."""
    from scipy.signal import welch
    rng = np.random.default_rng(22)
    fs = 4096.0
    rec = {}
    for L in (2048, 1000, 10000):
        n = rng.standard_normal(L + 64)
        col = np.convolve(n, np.hanning(33) / np.hanning(33).sum(), mode="valid")[:L]
        clean = synthetic_chirps(1, L, snr=8.0, seed=70 + L)["clean_norm"][0, 0].numpy().astype(np.float64) * 0.3
        y = (col * 3.0 + clean + 0.7).astype(np.float32)
        x = clean.astype(np.float32)
        y_w, x_w, (freqs, P) = I._whiten_pair_welch(y, x, fs)
        back = I._dewhiten_welch(y_w, (freqs, P), fs)
        f, Pxx = welch(y, fs=fs, nperseg=min(4096, L))
        rec.update({f"y_{L}": y, f"x_{L}": x, f"yw_{L}": y_w, f"xw_{L}": x_w, f"P_{L}": P, f"freqs_{L}": freqs, f"back_{L}": back,
                    f"Pxx_{L}": Pxx.astype(np.float64), f"f_{L}": f})
    np.savez_compressed(os.path.join(out, "whitening_welch.npz"), **rec)


def gen_ingest(out):
    """The unmodified reference ingest path (dataloader.NoisyWaveDataset + pad_collate, dataloader.py:26-268) on a small HDF5
    file with gen.py's layout (variable-length float32 rows of ragged lengths, per-sample masses / spins, model PSD rows, root
    attributes).  This image has no h5py: the fixture is written by the package's `_hdf5.write_file`, and the reference reads
    it through the package's minimal reader registered as `h5py` (only `h5py.File(path, 'r')` indexing is used by the
    reference); everything after the read -- NaN scrub, whitening, sigma, metadata broadcast, left-pad collate -- is the
    reference's own code."""
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _hdf5
    rng = np.random.default_rng(33)
    lens = [1024, 768, 1024, 512, 1000, 768]
    sig, noisy = [], []
    for i, L in enumerate(lens):
        c = synthetic_chirps(1, L, snr=9.0, seed=300 + i)["clean_norm"][0, 0].numpy().astype(np.float64) * 2e-21
        n = np.convolve(rng.standard_normal(L + 32), np.hanning(33) / np.hanning(33).sum(), mode="valid")[:L] * 1e-21
        sig.append(c.astype(np.float32))
        noisy.append((c + n).astype(np.float32))
    noisy[3][5] = np.nan                                          # exercised by the NaN scrub (dataloader.py:160-164)
    psd_model = [1e-42 * (1.0 + np.abs(np.sin(np.linspace(0, 3, 513))) ** 2 + np.linspace(0, 1, 513) * (1 + 0.1 * i)) for i in range(6)]
    fixture = os.path.join(out, "ingest_fixture.h5")
    _hdf5.write_file(fixture, {"signal": sig, "noisy": noisy, "noise": [b - a for a, b in zip(sig, noisy)],
                               "lengths": np.array(lens, dtype=np.int64), "mass1": rng.uniform(20, 70, 6), "mass2": rng.uniform(10, 40, 6),
                               "spin1z": rng.uniform(-0.5, 0.5, 6), "spin2z": rng.uniform(-0.5, 0.5, 6),
                               "psd_model": np.stack(psd_model)},
                     {"sampling_rate": 4096.0, "delta_t": 1.0 / 4096.0, "time_axis": "seconds-rel-peak", "padding": "none"})
    h5stub = types.ModuleType("h5py")
    h5stub.File = _hdf5.File
    sys.modules["h5py"] = h5stub
    sys.path.insert(0, REF)
    import importlib
    D = importlib.import_module("dataloader")
    rec = {}
    for tag, kw in {"raw_std": dict(whiten=False, sigma_mode="std"), "train_std": dict(whiten=True, whiten_mode="train", sigma_mode="std"),
                    "model_mad": dict(whiten=True, whiten_mode="auto", sigma_mode="mad")}.items():
        ds = D.NoisyWaveDataset(fixture, include_metadata=True, mass_scale=65.0, **kw)
        batch = D.pad_collate([ds[i] for i in range(6)])
        for name, t in zip(("clean", "noisy", "sigma", "mask", "meta"), batch):
            rec[f"{tag}/{name}"] = t.numpy()
        one = ds[3]
        rec[f"{tag}/item3_noisy"] = one[1].numpy()
        ds.close()
    np.savez_compressed(os.path.join(out, "ingest.npz"), **rec)


ARCH_CASES = [("b16_k5_d3", 16, 5, 3, 3, 1, 256), ("b8_k7_d2", 8, 7, 2, 7, 5, 250), ("b32_k1_d3", 32, 1, 3, 3, 1, 256)]


def gen_arch(M, TR, out):
    """Non-default UNet1D(base_ch, kernel, depth) (models.py:78-88) through the unmodified reference: eps_hat and every parameter
    gradient of the reference loss (train.py:53-58), for pinning the oracle (and through it the generic CUDA kernels) there."""
    rec = {}
    for tag, base_ch, kernel, depth, in_ch, cc, L in ARCH_CASES:
        model = M.UNet1D(in_ch=in_ch, base_ch=base_ch, depth=depth, kernel=kernel, cond_in_ch=cc, use_selfcond=True)
        model.load_state_dict(make_state_dict(in_ch=in_ch, cond_in_ch=cc, base_ch=base_ch, depth=depth, kernel=kernel, seed=21),
                              strict=True)
        x = gaussian((2, in_ch, L), seed=200 + L + base_ch)
        t = torch.tensor([24, 731])
        target = gaussian((2, 1, L), seed=300 + base_ch)
        eps = model(x, t)
        loss = TR._element_loss(eps, target, torch.ones_like(target), "huber", 0.5).mean()
        loss.backward()
        rec[f"{tag}/eps"] = eps.detach().numpy()
        rec[f"{tag}/loss"] = np.float64(loss.item())
        for k, p_ in model.named_parameters():
            rec[f"{tag}/grad/{k}"] = p_.grad.numpy()
    np.savez_compressed(os.path.join(out, "arch.npz"), **rec)


def gen_checkpoint(M, I, TR, out):
    """A reference-format checkpoint (payload of train.py:606-630) written from the unmodified reference classes: a small
    non-default UNet1D (base_ch=16, depth=2, time_dim=32, 7 input channels) after two AdamW steps on the reference loss with an
    EMA copy (train.py:73-81), plus what the reference computes from it -- eps_hat of the raw and of the EMA weights -- and the
    outputs of the reference's measurement loaders (`_load_measurement_from_h5`, `_meta_to_stack`, inference.py:59-122) on the
    ingest fixture."""
    import copy
    import types
    torch.manual_seed(11)
    in_ch, cc, L = 7, 5, 256
    model = M.UNet1D(in_ch=in_ch, base_ch=16, time_dim=32, depth=2, t_embed_max_time=999, cond_in_ch=cc, use_selfcond=True)
    with torch.no_grad():
        model.final.weight.normal_(0.0, 0.05)
        model.final.bias.normal_(0.0, 0.05)
    ema = copy.deepcopy(model)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=1e-4)
    diff = M.CustomDiffusion(T=1000, device="cpu")
    data = synthetic_chirps(3, L, snr=10.0, seed=90)
    cond = torch.cat([data["y_norm"], 0.3 * gaussian((3, 4, 1), seed=91).expand(3, 4, L)], dim=1)
    for it in range(2):
        t = torch.tensor([600, 800, 999]) - it
        x_t, eps = diff.q_sample(data["clean_norm"], t)
        net = torch.cat([x_t, cond, torch.zeros_like(x_t)], dim=1)
        loss = TR._element_loss(model(net, t), eps, torch.ones_like(eps), "huber", 0.5).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        TR.update_ema(ema, model, 0.9)
    args = dict(data="train.h5", model_dir="model", epochs=2, batch_size=3, lr=2e-3, weight_decay=1e-4, T=1000, base_ch=16, time_dim=32,
                depth=2, ema=True, ema_decay=0.9, p_uncond=0.2, p_selfcond=0.5, loss="huber", huber_beta=0.5)
    payload = {"model_state": model.state_dict(), "optimizer_state": opt.state_dict(),
               "args": {**args, "conditional": True, "in_ch": in_ch, "cond_in_ch": cc, "meta_enabled": True, "meta_channels": 4,
                        "conditioning": "concat[y + meta]+selfcond", "whiten": True, "whiten_mode": "auto", "sigma_mode": "std",
                        "dropout_y_only": True, "meta_scale": {"M": 65.0, "q": 10.0}},
               "epoch": 2, "model_ema_state": ema.state_dict()}
    torch.save(payload, os.path.join(out, "ref_checkpoint.pth"))
    rec = {}
    x = gaussian((2, in_ch, L), seed=92)
    tt = torch.tensor([24, 731])
    with torch.no_grad():
        rec["x"], rec["t"] = x.numpy(), tt.numpy()
        rec["eps_raw"] = model.eval()(x, tt).numpy()
        rec["eps_ema"] = ema.eval()(x, tt).numpy()
    # measurement loaders on the ingest fixture (read through the package's reader registered as h5py, see gen_ingest)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _hdf5
    h5stub = types.ModuleType("h5py")
    h5stub.File = _hdf5.File
    I.h5py = h5stub
    y, clean, fs, P_model, (fw, Pw), meta = I._load_measurement_from_h5(os.path.join(out, "ingest_fixture.h5"), 4)
    rec["meas/y"], rec["meas/clean"], rec["meas/fs"], rec["meas/P_model"] = y, clean, np.float64(fs), P_model
    rec["meas/meta_keys"] = np.array(sorted(meta))
    rec["meas/meta_vals"] = np.array([meta[k] for k in sorted(meta)], dtype=np.float64)
    meta["q"] = meta["mass1"] / meta["mass2"]
    meta["chirp_mass"] = 21.5
    for need in (1, 3, 5, 7, 9):
        st = I._meta_to_stack(meta, 64, need, 65.0, 10.0)
        rec[f"meta_stack/{need}"] = np.zeros((0, 64), np.float32) if st is None else st
    np.savez_compressed(os.path.join(out, "checkpoint.npz"), **rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--only", default=None, help="run one generator only (whitening_welch | ingest | checkpoint | arch)")
    args = ap.parse_args()
    if args.only == "whitening_welch":
        M, I, TR = import_reference()
        gen_whitening_welch(I, args.out)
        return
    if args.only == "ingest":
        gen_ingest(args.out)
        return
    if args.only == "arch":
        M, I, TR = import_reference()
        gen_arch(M, TR, args.out)
        return
    if args.only == "checkpoint":
        M, I, TR = import_reference()
        gen_checkpoint(M, I, TR, args.out)
        return
    torch.set_num_threads(8)
    M, I, TR = import_reference()
    gen_schedule(M, I, args.out)
    gen_forward(M, args.out)
    gen_chains(M, I, args.out)
    gen_train(M, TR, args.out)
    gen_proxy_and_helpers(M, I, TR, args.out)
    gen_scores(I, args.out)
    gen_whitening(I, args.out)
    gen_whitening_welch(I, args.out)
    gen_ingest(args.out)
    gen_checkpoint(M, I, TR, args.out)
    gen_arch(M, TR, args.out)
    tot = sum(os.path.getsize(os.path.join(args.out, f)) for f in os.listdir(args.out) if f.endswith(".npz"))
    print(f"golden fixtures written to {args.out}: {tot / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
