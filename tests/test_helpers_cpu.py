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


def test_stratified_timesteps_cover_every_stratum():
    """`_sample_timesteps_stratified` (train.py:147-172): every stratum [edge_i, edge_{i+1}) of [t_min, t_max] receives q or q + 1
    of the bsz draws, all draws lie in range, and the result is a permutation of the per-stratum draws."""
    import torch
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import _sample_timesteps_stratified
    torch.manual_seed(0)
    for bsz, t_min, t_max, bins in [(256, 500, 999, 0), (37, 500, 999, 8), (16, 0, 999, 64), (5, 990, 999, 0), (12, 7, 7, 4)]:
        t = _sample_timesteps_stratified(bsz, t_min, t_max, "cpu", bins=bins)
        assert t.dtype == torch.int64 and t.shape == (bsz,)
        assert int(t.min()) >= t_min and int(t.max()) <= t_max
        b = max(1, min(bins if bins > 0 else bsz, bsz))
        edges = torch.linspace(t_min, t_max + 1, b + 1).long().tolist()
        q, r = divmod(bsz, b)
        left = sorted(t.tolist())
        for i in range(b):
            lo, hi = edges[i], max(edges[i + 1] - 1, edges[i])
            want = q + 1 if i < r else q
            got = [v for v in left if lo <= v <= hi]
            assert len(got) >= want, (bsz, bins, i, lo, hi, want, len(got))      # >=: a degenerate stratum may overlap its neighbour
        # different calls give different draws (it is random), same seed gives the same
        torch.manual_seed(3)
        a1 = _sample_timesteps_stratified(bsz, t_min, t_max, "cpu", bins=bins)
        torch.manual_seed(3)
        a2 = _sample_timesteps_stratified(bsz, t_min, t_max, "cpu", bins=bins)
        assert torch.equal(a1, a2)


def test_sweep_sample_combo_consumes_the_rngs_in_reference_order():
    """sweep.sample_combo mirrors the nested `sample_combo` of sweep_infer.py:291-303, which draws from the GLOBAL `random` and
    `np.random` streams in a fixed order (cfg-mode coin, start_snr, cfg_scale, cfg_center, cfg_width from numpy; dc_weight,
    init_mode, eta choices from `random`): with the same seeds a reference run and this one must visit the same combinations."""
    import math
    import random
    import types
    import numpy as np
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import sweep as S
    a = types.SimpleNamespace(cfg_mode="auto", start_snr_min=5.0, start_snr_max=40.0, cfg_min=1.0, cfg_max=3.0, cfg_center_min=0.5,
                              cfg_center_max=0.9, cfg_width_min=0.05, cfg_width_max=0.3, dc_choices=[0.0, 0.05, 0.1],
                              init_choices=["noise", "y-blend"], eta_choices=[0.0, 0.5, 1.0])
    random.seed(7)
    np.random.seed(7)
    got = [S.sample_combo(a) for _ in range(5)]
    random.seed(7)
    np.random.seed(7)
    for c in got:
        coin = random.random()
        u = [np.random.uniform(lo, hi) for lo, hi in [(math.log10(5.0), math.log10(40.0)), (1.0, 3.0), (0.5, 0.9), (0.05, 0.3)]]
        dc, init, eta = random.choice(a.dc_choices), random.choice(a.init_choices), random.choice(a.eta_choices)
        assert c["cfg_mode"] == ("gauss" if coin < 0.7 else "const")
        assert c["start_snr"] == 10 ** u[0] and c["cfg_scale"] == u[1] and c["cfg_center"] == u[2] and c["cfg_width"] == u[3]
        assert (c["dc_weight"], c["init_mode"], c["eta"]) == (float(dc), init, float(eta))
    a.cfg_mode = "const"                                       # a fixed mode draws no coin
    random.seed(1)
    S.sample_combo(a)
    r_after = random.random()
    random.seed(1)
    random.choice(a.dc_choices), random.choice(a.init_choices), random.choice(a.eta_choices)
    assert random.random() == r_after
