"""Pin the CPU oracle against vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import ModelCfg
from weights import make_state_dict, synthetic_chirps, gaussian
from make_golden import CHAIN_CASES


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _sub(x):
    return x[:, ::4, ::4].numpy()


def test_schedule_tables(golden_dir):
    g = _load(golden_dir, "schedule.npz")
    betas = oracle.cosine_beta_schedule(1000)
    ab = oracle.alpha_bar_from_betas(betas)
    assert np.array_equal(betas.numpy(), g["betas"])
    assert np.array_equal(ab.numpy(), g["alpha_bar"])
    assert np.array_equal(oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(50)).numpy(), g["alpha_bar_T50"])
    for key in [k for k in g if k.startswith("sched_")]:
        _, T, steps, st = key.split("_")
        st = None if st == "None" else int(st)
        assert np.array_equal(oracle.build_t_schedule(int(T), int(steps), st).numpy(), g[key]), key
    tt = torch.from_numpy(g["temb_t"])
    assert np.array_equal(oracle.time_embedding(tt, 128, 999.0).numpy(), g["temb_128"])
    assert np.array_equal(oracle.time_embedding(tt, 33, 49.0).numpy(), g["temb_33"])
    w = [oracle.cfg_weight(i, 10, mode, 1.5, 0.5, 0.3) for mode in ["const", "tophat", "gauss"] for i in [0, 3, 5, 9]]
    assert np.array_equal(np.array(w), g["cfg_w"])
    assert [oracle.t_for_target_snr(ab, s) for s in [0.9, 2.0, 10.0, 20.0]] == list(g["t_for_snr"])
    assert np.array_equal(oracle.snr_from_alpha_bar(ab), g["snr_tab"])
    with pytest.raises(ValueError):
        oracle.cfg_weight(0, 10, "nope", 1.0, 0.5, 0.3)


@pytest.mark.parametrize("tag,in_ch,cc,L,B", [("c3_L256", 3, 1, 256, 2), ("c7_L256", 7, 5, 256, 2),
                                               ("c3_L500", 3, 1, 500, 1), ("c7_L1024", 7, 5, 1024, 1)])
def test_forward_per_layer(golden_dir, tag, in_ch, cc, L, B):
    g = _load(golden_dir, f"forward_{tag}.npz")
    sd = make_state_dict(in_ch=in_ch, cond_in_ch=cc, seed=0)
    cfg = ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    x = gaussian((B, in_ch, L), seed=100 + L + in_ch)
    if cc == 5:
        x[:, 2:6, :] = x[:, 2:6, :1].clone()
    t = torch.from_numpy(g["t"])
    with torch.no_grad():
        taps = oracle.unet_forward_taps(sd, cfg, x, t)
    assert np.array_equal(taps["eps"].numpy(), g["eps"])
    for name in ["enc0", "enc1", "enc2", "mid", "dec0", "dec1", "dec2"]:
        assert np.array_equal(_sub(taps[name + ".raw"]), g[name + ".raw"]), name
    # the conv inputs captured by hooks pin the post-FiLM activations, the pooling and the concat order
    assert np.array_equal(_sub(torch.nn.functional.avg_pool1d(taps["enc0.out"], 2, 2)), g["enc1.in"])
    fin = torch.cat([taps["dec2.out"] if taps["dec2.out"].size(-1) == L else
                     torch.nn.functional.pad(taps["dec2.out"], (0, L - taps["dec2.out"].size(-1))), x[:, :1]], dim=1)
    assert np.array_equal(_sub(fin), g["final.in"])
    assert abs(oracle.conv_flops_per_sample(cfg, 4096) / 1e9 - (2.2243 if in_ch == 3 else 2.2442)) < 2e-4


@pytest.mark.parametrize("in_ch,cc", [(3, 1), (7, 5)])
def test_chains_match_reference(golden_dir, in_ch, cc):
    L = 256
    sd = make_state_dict(in_ch=in_ch, cond_in_ch=cc, seed=1)
    cfg = ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    y = synthetic_chirps(2, L, snr=10.0, seed=77)["y_norm"]
    cond = y if cc == 1 else torch.cat([y, gaussian((2, 4, 1), seed=5).expand(2, 4, L).contiguous() * 0.3], dim=1)
    for tag, kw in CHAIN_CASES.items():
        if cc == 5 and tag not in ("cfg15_dc", "ddim10_s289"):
            continue
        g = _load(golden_dir, f"chain_c{in_ch}_{tag}.npz")
        draws = [torch.cat([gaussian((1, 1, L), seed=9000 + 100 * b + k) for b in range(2)], 0) for k in range(64)]
        used = []

        def gen():
            for d in draws:
                used.append(1)
                yield d
        kws = dict(kw)
        out = oracle.ddim_sample(sd, cfg, ab, cond, T=1000, noise=gen(), **kws)
        assert len(used) == int(g["n_draws"]), tag
        err = (out - torch.from_numpy(g["x_final"])).abs().max().item()
        ref = np.abs(g["x_final"]).max()
        assert err <= 2e-6 * max(ref, 1.0), (tag, err, ref)


@pytest.mark.parametrize("in_ch,cc", [(7, 5), (3, 1)])
@pytest.mark.parametrize("sc", [False, True])
def test_train_step_matches_reference(golden_dir, in_ch, cc, sc):
    L, B = 256, 4
    g = _load(golden_dir, f"train_c{in_ch}_sc{int(sc)}.npz")
    sd = make_state_dict(in_ch=in_ch, cond_in_ch=cc, seed=2)
    cfg = ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    data = synthetic_chirps(B, L, snr=12.0, seed=31)
    clean, y = data["clean_norm"], data["y_norm"]
    mask = torch.ones(B, 1, L)
    mask[1, :, :37] = 0.0
    cond = y if cc == 1 else torch.cat([y, gaussian((B, 4, 1), seed=6).expand(B, 4, L).contiguous() * 0.3], dim=1)
    t = torch.tensor([500, 731, 999, 612])
    eps = gaussian((B, 1, L), seed=41)
    drop = torch.tensor([0.0, 1.0, 0.0, 0.0]).view(B, 1, 1)
    params = {k: v.clone() for k, v in sd.items()}
    ema = {k: v.clone() for k, v in sd.items()}
    mom = {k: torch.zeros_like(v) for k, v in sd.items()}
    var = {k: torch.zeros_like(v) for k, v in sd.items()}
    for step in range(2):
        loss, grads, eps_hat = oracle.train_step(params, cfg, ab, clean_norm=clean, cond_stack=cond, mask=mask, t=t,
                                                 eps=eps, drop=drop, selfcond=sc)
        assert abs(float(loss) - float(g[f"loss{step}"])) <= 1e-6 * max(1.0, abs(float(loss)))
        if step == 0:
            np.testing.assert_allclose(eps_hat.numpy(), g["eps_hat"], rtol=0, atol=2e-6)
            for k, gr in grads.items():
                ref = g["grad/" + k]
                mine = gr.numpy() if gr.numel() <= 4096 else gr.reshape(-1)[::97].numpy()
                np.testing.assert_allclose(mine, ref, rtol=0, atol=2e-6 * max(1e-3, np.abs(ref).max()), err_msg=k)
        clipped, gn = oracle.clip_grad_norm(grads, 1.0)
        assert abs(gn - float(g[f"grad_norm{step}"])) <= 1e-5 * gn
        lr = 2e-4 * oracle.warmup_cosine_lambda(step, 10, 100, 0.1)
        assert abs(lr - float(g[f"lr{step}"])) <= 1e-12
        for k in params:
            params[k], mom[k], var[k] = oracle.adamw_step(params[k], clipped[k], mom[k], var[k], step + 1, lr)
            ema[k] = oracle.ema_step(ema[k], params[k], 0.999)
    for k in params:
        ref = g["p2/" + k]
        mine = params[k].numpy() if params[k].numel() <= 4096 else params[k].reshape(-1)[::97].numpy()
        np.testing.assert_allclose(mine, ref, rtol=0, atol=3e-7, err_msg=k)
        refe = g["ema2/" + k]
        minee = ema[k].numpy() if ema[k].numel() <= 4096 else ema[k].reshape(-1)[::97].numpy()
        np.testing.assert_allclose(minee, refe, rtol=0, atol=3e-7, err_msg=k)
    lam = np.load(os.path.join(golden_dir, "lr_lambda.npz"))["lam"]
    mine = [oracle.warmup_cosine_lambda(s, 10, 100, 0.1) for s in [0, 5, 9, 10, 50, 99, 100, 150]]
    assert np.allclose(mine, lam, rtol=0, atol=1e-15)


def test_non_default_architectures_match_reference(golden_dir):
    """UNet1D(base_ch, kernel, depth) outside the CLI defaults (models.py:78-88): the oracle's eps_hat and autograd gradients
    against the unmodified reference's (tests/golden/make_golden.py::gen_arch) -- this pins the oracle where the shape-generic
    CUDA kernels are compared with it (tests/test_gpu_generic_arch.py)."""
    from make_golden import ARCH_CASES
    g = _load(golden_dir, "arch.npz")
    for tag, base_ch, kernel, depth, in_ch, cc, L in ARCH_CASES:
        sd = {k: v.clone().requires_grad_(True) for k, v in
              make_state_dict(in_ch=in_ch, cond_in_ch=cc, base_ch=base_ch, depth=depth, kernel=kernel, seed=21).items()}
        cfg = ModelCfg(in_ch=in_ch, base_ch=base_ch, depth=depth, kernel=kernel, cond_in_ch=cc, use_selfcond=True)
        x = gaussian((2, in_ch, L), seed=200 + L + base_ch)
        target = gaussian((2, 1, L), seed=300 + base_ch)
        eps = oracle.unet_forward(sd, cfg, x, torch.tensor([24, 731]))
        assert float((eps.detach() - torch.from_numpy(g[f"{tag}/eps"])).norm() / torch.from_numpy(g[f"{tag}/eps"]).norm()) <= 2e-6, tag
        loss = torch.nn.functional.smooth_l1_loss(eps, target, beta=0.5, reduction="none").mean()
        assert abs(float(loss) - float(g[f"{tag}/loss"])) <= 2e-6 * abs(float(g[f"{tag}/loss"]))
        loss.backward()
        for k, v in sd.items():
            ref = torch.from_numpy(g[f"{tag}/grad/{k}"])
            assert float((v.grad - ref).norm()) <= 5e-6 * max(float(ref.norm()), 1e-6), (tag, k)
