"""On-device scoring against the scores computed by the unmodified reference helpers
(tests/golden/scores.npz, make_golden.py:gen_scores).  fp64 accumulation on both sides except where the reference sums in
float32 (np.mean / np.dot of float32 arrays): tolerance 2e-6 relative there, 1e-9 for the float64 quantities."""
import os

import numpy as np
import pytest
import torch

from weights import synthetic_chirps

pytestmark = pytest.mark.gpu


def test_scores_match_reference_golden(golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import scoring
    g = np.load(os.path.join(golden_dir, "scores.npz"))
    L, B, fs = 2048, 6, 4096.0
    clean = synthetic_chirps(B, L, snr=12.0, seed=55)["clean_norm"].cuda()
    xhat = torch.from_numpy(g["xhat"]).cuda()
    sigma = torch.from_numpy(g["sigma"]).cuda()
    r = scoring.score_batch(xhat, clean, fs, sigma=sigma, secs=0.2, max_shift=64)
    np.testing.assert_allclose(r["corr_last"].cpu().numpy(), g["corr_last"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["mae_last"].cpu().numpy(), g["mae_last"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["nmae_sigma"].cpu().numpy(), g["nmae_sigma"], rtol=2e-6)
    assert np.array_equal(r["best_lag"].cpu().numpy(), g["best_lag_64"])
    np.testing.assert_allclose(r["xc_mae"].cpu().numpy(), g["xc_mae"], rtol=2e-6)
    np.testing.assert_allclose(r["xc_nmae_clean"].cpu().numpy(), g["xc_nmae_clean"], rtol=2e-6)
    np.testing.assert_allclose(r["xc_nmae_sigma"].cpu().numpy(), g["xc_nmae_sigma"], rtol=2e-6)
    full = scoring.best_lag_by_xcorr(clean, xhat, 0)                      # max_shift <= 0: every lag (inference.py:248-249)
    assert np.array_equal(full.cpu().numpy().astype(np.float64), g["best_lag_full"])
    m = {"corr_last": r["corr_last"], "nmae_sigma": r["nmae_sigma"]}
    J = scoring.objective(m, {"corr_last": r["corr_last"]})
    np.testing.assert_allclose(J.cpu().numpy(), g["objective"], rtol=1e-6)
    lw = scoring.score_last_window(xhat, clean, fs, secs=0.2)
    assert torch.equal(lw["corr_last"], r["corr_last"])
    # overlap of a signal with itself / its negative
    o = scoring.score_batch(clean, clean, fs)["overlap"]
    assert float((o - 1.0).abs().max()) < 1e-12
    assert float((scoring.score_batch(-clean, clean, fs)["overlap"] + 1.0).abs().max()) < 1e-12


def test_scoring_rejects_cpu_tensors():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import scoring
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        scoring.score_batch(torch.zeros(1, 16), torch.zeros(1, 16), 4096.0)
