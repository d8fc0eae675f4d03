"""GPU whitening / de-whitening / sigma against outputs of the unmodified reference helpers
(tests/golden/whitening.npz, make_golden.py:gen_whitening).

Tolerances: the whitened float32 outputs and the float64 PSD come from fp64 FFTs on both sides (pocketfft vs cuFFT): rel-L2
<= 1e-6 (float32 rounding of the output dominates), PSD <= 1e-10.  De-whitening: numpy >= 2 transforms a float32 input in
single precision (inference.py:157 gets the float32 y_w), cuFFT here runs in fp64: rel-L2 <= 1e-5."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("L", [2048, 1000])
def test_whitening_matches_reference_golden(golden_dir, L):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening as W
    g = np.load(os.path.join(golden_dir, "whitening.npz"))
    y, x = g[f"y_{L}"], g[f"x_{L}"]
    y_w, x_w, P = W._whiten_pair_train_like(y, x, 4096.0)
    assert y_w.dtype == np.float32 and P.dtype == np.float64 and P.shape == (L // 2 + 1,)
    assert rel(P, g[f"P_{L}"]) <= 1e-10
    assert rel(y_w, g[f"yw_{L}"]) <= 1e-6 and rel(x_w, g[f"xw_{L}"]) <= 1e-6
    assert W._whiten_pair_train_like(y, None, 4096.0)[1] is None
    assert rel(W._dewhiten_train_like(g[f"yw_{L}"], g[f"P_{L}"]), g[f"back_{L}"]) <= 1e-5
    ym, xm, Pm = W._whiten_pair_model(y, x, g[f"Pmodel_{L}"], 4096.0)
    assert rel(Pm, g[f"Pm_{L}"]) <= 1e-12
    assert rel(ym, g[f"ym_{L}"]) <= 1e-6 and rel(xm, g[f"xm_{L}"]) <= 1e-6
    assert rel(W._dewhiten_model(g[f"ym_{L}"], g[f"Pm_{L}"]), g[f"backm_{L}"]) <= 1e-5
    assert abs(W._pick_sigma(y, "std", 1.0) - float(g[f"sig_std_{L}"])) <= 1e-12 * float(g[f"sig_std_{L}"])
    assert abs(W._pick_sigma(y, "mad", 1.0) - float(g[f"sig_mad_{L}"])) <= 1e-12 * float(g[f"sig_mad_{L}"])
    assert W._pick_sigma(y, "fixed", 2.5) == float(g[f"sig_fixed_{L}"])
    with pytest.raises(ValueError):
        W._pick_sigma(y, "bogus", 1.0)


def test_batched_whitening_equals_per_sample_and_round_trips(golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening as W
    g = np.load(os.path.join(golden_dir, "whitening.npz"))
    y = torch.from_numpy(np.stack([g["y_2048"], g["y_2048"][::-1].copy(), 2.0 * g["y_2048"]])).cuda()
    y_w, _, P = W.whiten_train_like(y)
    assert rel(y_w[0].cpu().numpy(), g["yw_2048"]) <= 1e-6
    # whiten -> de-whiten returns the mean-removed input (spectral floor aside)
    back = W.apply_psd(y_w, P, dewhiten=True)
    ref = (y.double() - y.double().mean(dim=1, keepdim=True)).cpu().numpy()
    assert rel(back.cpu().numpy(), ref) <= 1e-5
    s = W.sigma(y, "std")
    assert torch.allclose(s, y.double().std(dim=1, unbiased=False), rtol=1e-12)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        W.whiten_train_like(torch.zeros(1, 64))


def test_reconstruct_batch_pipeline_equals_its_stages_and_the_oracle():
    """pipeline.reconstruct_batch (inference.py:655-826 on the device) == whiten -> sigma -> oracle DDIM -> de-whiten by hand."""
    import oracle
    from weights import make_state_dict, synthetic_chirps
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import pipeline, whitening as W
    B, L, fs = 3, 1024, 4096.0
    d = synthetic_chirps(B, L, snr=10.0, seed=17)
    y_raw = (d["y_norm"][:, 0] * 3e-3 + 1e-3).cuda()            # "strain-like" scale and a DC offset
    clean_raw = (d["clean_norm"][:, 0] * 3e-3).cuda()
    sd = make_state_dict(3, 1, seed=0)
    m = UNet1D(in_ch=3, cond_in_ch=1, use_selfcond=True)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    diff = CustomDiffusion(T=1000, device="cuda")
    r = pipeline.reconstruct_batch(m, diff, y_raw, fs=fs, clean_raw=clean_raw, whiten=True, whiten_mode="train", sigma_mode="std",
                                   start_snr=2.0, steps=6, eta=0.0, seed=5, compute_dtype="fp32", score_secs=0.2)
    assert r["whiten_kind"] == "train" and r["start_t"] == 289
    # stage by stage
    y_w, c_w, P = W.whiten_train_like(y_raw, clean_raw)
    sig = W.sigma(y_w, "std")
    assert torch.allclose(sig, r["sigma"], rtol=1e-12)
    y_norm = (y_w / sig.float()[:, None])[:, None, :]
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    xT = inf.philox_normal(B, L, 5, 0, 0, "cuda").cpu()
    ref = oracle.ddim_sample(sd, cfg, ab, y_norm.cpu(), T=1000, steps=6, eta=0.0, start_t=289, noise=[xT])
    assert rel(r["x0_hat_norm"].cpu().numpy(), ref.numpy()) <= 1e-4
    back = W.apply_psd((ref.cuda() * sig.float().view(B, 1, 1)).view(B, L), P, dewhiten=True)
    assert rel(r["x0_hat_strain"].cpu().numpy(), back.cpu().numpy()) <= 1e-4
    assert set(r["strain"]) >= {"corr_last", "mae_last", "nmae_sigma", "overlap"} and r["objective"].shape == (B,)
    assert torch.isfinite(r["objective"]).all()
    # raw (un-whitened) flavour and the bf16 engine
    r2 = pipeline.reconstruct_batch(m, diff, y_raw, fs=fs, whiten=False, steps=4, start_t=200, seed=5, compute_dtype="bf16")
    assert r2["whiten_kind"] == "raw" and r2["x0_hat_strain"].shape == (B, L) and torch.isfinite(r2["x0_hat_strain"]).all()


@pytest.mark.parametrize("L", [2048, 1000, 10000])
def test_welch_whitening_matches_reference_golden(golden_dir, L):
    """The Welch variant (inference.py:161-179): scipy.signal.welch on the device (one segment at L <= 4096, four at L = 10000),
    interpolation onto the rfft grid, whitening and de-whitening -- against outputs of the unmodified reference helper.  scipy
    transforms the float32 input in single precision; cuFFT here runs in fp64: PSD rel-L2 <= 2e-6, whitened signals <= 1e-4."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening as W
    g = np.load(os.path.join(golden_dir, "whitening_welch.npz"))
    y, x = g[f"y_{L}"], g[f"x_{L}"]
    nper = min(4096, L)
    Pxx = W.welch_psd(torch.from_numpy(np.stack([y, 2.0 * y])).cuda(), 4096.0, nper).cpu().numpy()
    assert Pxx.shape == (2, nper // 2 + 1)
    assert rel(Pxx[0], g[f"Pxx_{L}"]) <= 2e-6 and rel(Pxx[1], 4.0 * g[f"Pxx_{L}"]) <= 2e-6
    y_w, x_w, (freqs, P) = W._whiten_pair_welch(y, x, 4096.0)
    assert np.array_equal(freqs, g[f"freqs_{L}"])
    assert rel(P, g[f"P_{L}"]) <= 2e-6
    # the reference's rfft(y) runs in single precision here (float32 input, numpy >= 2; no float64 cast at inference.py:168-171)
    assert rel(y_w, g[f"yw_{L}"]) <= 1e-4 and rel(x_w, g[f"xw_{L}"]) <= 1e-4
    assert rel(W._dewhiten_welch(g[f"yw_{L}"], (freqs, g[f"P_{L}"]), 4096.0), g[f"back_{L}"]) <= 1e-5
    # np.interp on an arbitrary grid (a saved Welch PSD with its own frequency array, dataloader.py:136-139)
    fw = np.sort(np.random.default_rng(1).uniform(0.0, 2100.0, size=97))
    Pw = np.random.default_rng(2).uniform(0.5, 2.0, size=97)
    got = W.interp_grid(torch.from_numpy(fw), torch.from_numpy(Pw), L, 4096.0)[0].cpu().numpy()
    ref = np.interp(np.fft.rfftfreq(L, 1 / 4096.0), fw, Pw, left=Pw[0], right=Pw[-1])
    assert rel(got, ref) <= 1e-13


@pytest.mark.parametrize("L", [64, 512, 2048, 4096, 16384])
def test_fused_whitening_kernel_equals_cufft_path(L):
    """gwf_whiten_train_like for power-of-two lengths runs in ONE kernel (shared-memory fp64 FFT, whiten.cu::whiten_fused_kernel);
    against the cuFFT + spectral-kernel path (option fused = 0) and against numpy's float64 recipe (inference.py:137-153)."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening as W
    rng = np.random.default_rng(L)
    B = 5
    t = np.arange(L) / 4096.0
    y = (rng.standard_normal((B, L)) * 3e-3 + 1e-3 + 2e-3 * np.sin(2 * np.pi * 60.0 * t)).astype(np.float32)
    x = (rng.standard_normal((B, L)) * 1e-3).astype(np.float32)
    lib = W.load()
    outs = {}
    for fused in (1, 0):
        assert lib.gwf_set_option(b"fused", fused) == 0
        try:
            y_w, x_w, P = W.whiten_train_like(torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda())
            only_y = W.whiten_train_like(torch.from_numpy(y).cuda())
        finally:
            lib.gwf_set_option(b"fused", 1)
        assert only_y[1] is None and torch.equal(only_y[0], y_w)
        outs[fused] = (y_w.cpu().numpy(), x_w.cpu().numpy(), P.cpu().numpy())
    # numpy float64 recipe
    for b in range(B):
        y64 = y[b].astype(np.float64) - y[b].astype(np.float64).mean()
        Y = np.fft.rfft(y64)
        Pn = np.abs(Y) ** 2
        if Pn.size > 9:
            Pn = np.convolve(Pn, np.ones(9) / 9.0, mode="same")
        Pn = np.maximum(Pn, 1e-20)
        yw = np.fft.irfft(Y / np.sqrt(Pn), n=L).astype(np.float32)
        x64 = x[b].astype(np.float64) - x[b].astype(np.float64).mean()
        xw = np.fft.irfft(np.fft.rfft(x64) / np.sqrt(Pn), n=L).astype(np.float32)
        for fused in (1, 0):
            assert rel(outs[fused][2][b], Pn) <= 1e-10, (fused, b, "P")
            assert rel(outs[fused][0][b], yw) <= 1e-6 and rel(outs[fused][1][b], xw) <= 1e-6, (fused, b)
    assert rel(outs[1][2], outs[0][2]) <= 1e-12 and rel(outs[1][0], outs[0][0]) <= 1e-6


@pytest.mark.parametrize("L", [256, 4096, 8192])
@pytest.mark.parametrize("shared", [False, True])
def test_fused_apply_psd_equals_cufft_path(L, shared):
    """gwf_apply_psd (model-PSD whitening, the loader's 1e-20-floor variant, de-whitening; inference.py:155-159, 190-203,
    dataloader.py:127-143) in one kernel for power-of-two lengths, against the cuFFT path and numpy."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening as W
    rng = np.random.default_rng(L + shared)
    B = 4
    sig = (rng.standard_normal((B, L)) * 2e-3 + 5e-4).astype(np.float32)
    P = np.abs(rng.standard_normal((L // 2 + 1,) if shared else (B, L // 2 + 1))) * 1e-6 + 1e-9
    lib = W.load()
    for dewhiten, floor in [(False, False), (False, True), (True, False)]:
        outs = {}
        for fused in (1, 0):
            assert lib.gwf_set_option(b"fused", fused) == 0
            try:
                o64 = W.apply_psd(torch.from_numpy(sig).cuda(), torch.from_numpy(P).cuda(), dewhiten, torch.float64, loader_floor=floor)
                o32 = W.apply_psd(torch.from_numpy(sig).cuda(), torch.from_numpy(P).cuda(), dewhiten, torch.float32, loader_floor=floor)
            finally:
                lib.gwf_set_option(b"fused", 1)
            outs[fused] = (o64.cpu().numpy(), o32.cpu().numpy())
        Y = np.fft.rfft(sig.astype(np.float64), axis=1)
        gain = np.sqrt(P + 1e-12) if dewhiten else 1.0 / np.sqrt(P + (1e-20 if floor else 1e-12))
        ref = np.fft.irfft(Y * gain, n=L, axis=1)
        for fused in (1, 0):
            assert rel(outs[fused][0], ref) <= 1e-12, (fused, dewhiten, floor)
            assert rel(outs[fused][1], ref) <= 1e-6
