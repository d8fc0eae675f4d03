"""GPU half of the ingest path: `dataloader.BatchLoader` (raw rows -> pinned left-padded staging -> H2D ->
batched whitening / sigma kernels) against what the UNMODIFIED reference data loader returned for the same HDF5 file
(tests/golden/ingest.npz, make_golden.py:gen_ingest: NoisyWaveDataset + pad_collate, dataloader.py:26-268), and
`train_diffusion(args)` with the reference signature on that file.

Tolerances: whitened float32 rows rel-L2 <= 1e-5 (fp64 FFTs on both sides, float32 outputs), sigma 1e-6 relative, masks and
metadata exact."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {"raw_std": dict(whiten=False, sigma_mode="std"), "train_std": dict(whiten=True, whiten_mode="train", sigma_mode="std"),
         "model_mad": dict(whiten=True, whiten_mode="auto", sigma_mode="mad")}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("tag", list(CASES))
def test_batch_loader_matches_reference_dataloader(golden_dir, tag):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import dataloader as D
    g = np.load(os.path.join(golden_dir, "ingest.npz"))
    fx = os.path.join(golden_dir, "ingest_fixture.h5")
    loader = D.make_dataloader(fx, batch_size=6, shuffle=False, mass_scale=65.0, **CASES[tag])
    assert len(loader) == 1 and len(loader.dataset) == 6 and loader.dataset.fs == 4096.0
    clean, noisy, sigma, mask, meta = next(iter(loader))
    assert noisy.is_cuda and noisy.shape == (6, 1, 1024) and meta.shape == (6, 4, 1024) and sigma.shape == (6,)
    assert np.array_equal(mask.cpu().numpy(), g[f"{tag}/mask"])
    assert rel(noisy.cpu().numpy(), g[f"{tag}/noisy"]) <= 1e-5 and rel(clean.cpu().numpy(), g[f"{tag}/clean"]) <= 1e-5
    assert np.allclose(sigma.cpu().numpy(), g[f"{tag}/sigma"], rtol=1e-6, atol=0)
    assert np.allclose(meta.cpu().numpy(), g[f"{tag}/meta"], rtol=1e-7, atol=0)
    # per-sample contract of the dataset (CPU tensors, dataloader.py:153-228)
    c3, n3, s3, m3, meta3 = loader.dataset[3]
    assert not n3.is_cuda and n3.shape == (1, 512) and meta3.shape == (4, 512) and float(m3.sum()) == 512.0
    assert rel(n3.numpy(), g[f"{tag}/item3_noisy"]) <= 1e-5
    # two batches of three through the double-buffered staging == the one batch of six, per sample (its own padding)
    l3 = D.make_dataloader(fx, batch_size=3, shuffle=False, mass_scale=65.0, **CASES[tag])
    got = list(l3)
    assert len(got) == 2 and got[1][1].shape == (3, 1, 1000)
    for k, (c, n, s, m, me) in enumerate(got):
        Lk = n.shape[-1]
        for j in range(3):
            b = 3 * k + j
            valid = int(g[f"{tag}/mask"][b].sum())
            assert rel(n[j, 0, Lk - valid:].cpu().numpy(), g[f"{tag}/noisy"][b, 0, 1024 - valid:]) <= 1e-5
            assert float(m[j].sum()) == valid


def test_shuffle_order_is_the_random_samplers(golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import dataloader as D
    fx = os.path.join(golden_dir, "ingest_fixture.h5")
    loader = D.make_dataloader(fx, batch_size=2, shuffle=True)
    torch.manual_seed(123)
    mine = loader._order()
    torch.manual_seed(123)
    ref = list(torch.utils.data.RandomSampler(range(6)))
    assert mine == ref


def test_train_diffusion_with_the_reference_signature(golden_dir, tmp_path):
    """train_diffusion(args): args.data -> meta scale -> make_dataloader -> ragged batches (three different (B, L) shapes sharing
    one optimiser run) -> checkpoint payload of train.py:606-630."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import train as TR
    args = argparse.Namespace(data=golden_dir, model_dir=str(tmp_path), epochs=2, batch_size=2, lr=1e-3, weight_decay=1e-4, T=1000,
                              base_ch=64, time_dim=128, depth=3, device="cuda", num_workers=0, seed=7, p_uncond=0.2, p_selfcond=0.5,
                              t_min_frac=0.5, force_cond_epochs=0, t_cover="rand", t_bins=0, t_multi=1, loss="huber", huber_beta=0.5,
                              clip_grad=1.0, clamp_inputs=10.0, skip_bad_batches=True, skip_loss_threshold=50.0, amp=False, ema=True,
                              ema_decay=0.99, warmup_steps=2, cosine_decay=True, min_lr_scale=0.1, loss_weight_power=0.0,
                              whiten=True, whiten_mode="train", sigma_mode="std", sigma_fixed=1.0, init_from=None,
                              dropout_y_only=True)
    out = TR.train_diffusion(args)
    assert len(out["losses"]) == 6 and all(np.isfinite(out["losses"]))
    ck = out["checkpoint"]
    assert ck["args"]["in_ch"] == 7 and ck["args"]["cond_in_ch"] == 5 and ck["args"]["meta_channels"] == 4
    assert 20.0 < ck["args"]["meta_scale"]["M"] <= 70.0                 # 95th percentile of the fixture's masses
    st = out["stepper"]
    assert st.applied_steps() + st.skipped_batches() == 6
    assert os.path.exists(os.path.join(str(tmp_path), "latest_model", "model_diffusion.pth"))
