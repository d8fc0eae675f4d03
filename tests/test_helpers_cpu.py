"""Host-side helpers of the training mirror against outputs of the unmodified reference (tests/golden/proxy_helpers.npz)."""
import os

import numpy as np
import torch


def test_match_batch_and_lr_lambda(golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import train as TR
    g = np.load(os.path.join(golden_dir, "proxy_helpers.npz"))
    a = torch.arange(6).view(3, 2).float()
    assert np.array_equal(TR._match_batch(a, 6).numpy(), g["match_3_to_6"])
    assert np.array_equal(TR._match_batch(a, 7).numpy(), g["match_3_to_7"])
    assert np.array_equal(TR._match_batch(a, 3).numpy(), g["match_3_to_3"])
    lam = np.load(os.path.join(golden_dir, "lr_lambda.npz"))["lam"]
    mine = [TR.warmup_cosine_lambda(s, 10, 100, 0.1) for s in [0, 5, 9, 10, 50, 99, 100, 150]]
    assert np.allclose(mine, lam, rtol=0, atol=1e-15)
    opt = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=1.0)
    sched = TR.make_warmup_cosine_scheduler(opt, 10, 100, 0.1)
    assert abs(sched.lr_lambdas[0](50) - lam[4]) < 1e-15


def test_stratified_timesteps_cover_the_reference_strata(golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import train as TR
    g = np.load(os.path.join(golden_dir, "proxy_helpers.npz"))
    torch.manual_seed(7)
    t = TR._sample_timesteps_stratified(64, 500, 999, torch.device("cpu"), bins=8)
    assert t.dtype == torch.int64 and t.shape == (64,)
    assert np.array_equal(np.histogram(t.numpy(), bins=g["strat_edges"])[0], g["strat_counts"])
    assert int(t.min()) >= 500 and int(t.max()) <= 999
    t = TR._sample_timesteps_stratified(10, 0, 999, torch.device("cpu"), bins=0)       # bins=0 -> one stratum per sample
    assert len(set((t // 100).tolist())) == 10
    t = TR._sample_timesteps_stratified(5, 7, 7, torch.device("cpu"), bins=3)          # degenerate range
    assert t.tolist() == [7] * 5


def test_element_loss_matches_reference_formula():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import train as TR
    import oracle
    torch.manual_seed(0)
    a, b = torch.randn(2, 1, 33), torch.randn(2, 1, 33)
    m = (torch.rand(2, 1, 33) > 0.3).float()
    for lt in ("huber", "mse"):
        assert torch.equal(TR._element_loss(a, b, m, lt, 0.5), oracle.element_loss(a, b, m, lt, 0.5))
