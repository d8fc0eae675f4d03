"""GPU parity on the configurations bench.py measures (BASELINE.json configs 1-4), directly against the CPU oracle.

  * the fused bf16 / tcgen05 inference kernels (conv_in_gn, conv_gn, conv_gn2 + head dots) at L = 4096 with more samples than
    CTA groups (B >= 19), in_ch 3 and 7: every block's output and eps_hat vs the oracle, <= 1e-2 (north_star bf16 tolerance);
  * the T = 1000 DDPM chain at L = 4096 (config 1 / the `sampling` bench object) in bf16 with injected noise: teacher-forced
    eps_hat at t in {999, 529, 289, 55, 0} <= 1e-2 and the end-to-end reconstruction against the oracle chain;
  * a 50-step chain at L = 16384 (config 4 shape);
  * world-size independence of the on-device Philox chain (graph replay AND eager): full batch == concatenated shards;
  * a cached plan follows a changed Philox key and changed weights (frozen graph arguments, stale FiLM table);
  * bf16 backward vs the oracle at L = 4096 with self-conditioning and in_ch = 7 (config 2 shape, small batch).
"""
import pytest
import torch

import oracle
from weights import gaussian, make_state_dict, synthetic_chirps

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2
NAMES = ["enc0", "enc1", "enc2", "mid", "dec0", "dec1", "dec2"]


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def overlap(a, b):
    a, b = a.double().cpu().reshape(a.shape[0], -1), b.double().cpu().reshape(b.shape[0], -1)
    return float(((a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1) + 1e-30)).min())


def _model(in_ch, cc, seed, dtype):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import UNet1D
    m = UNet1D(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True, compute_dtype=dtype)
    m.load_state_dict(make_state_dict(in_ch, cc, seed=seed), strict=True)
    return m.cuda().eval()


def _sample(model, diff, cond, kw, noise=None, **extra):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf
    full = dict(T=1000, device="cuda", length=cond.shape[-1], debug=False, x0_std_est=0.14, cond_scale=1.0, eps_scale=1.0,
                pred_type="eps", in_ch=model.in_ch, cond_in_ch=model.cond_in_ch, use_selfcond=True, cfg_mode="const",
                cfg_center=0.5, cfg_width=0.3, cfg_u_only_thresh=0.0, dc_weight=0.0, cfg_scale=1.0, start_t=None,
                init_mode="noise")
    full.update(kw)
    return inf.ddim_sample(model, diff, cond.cuda(), noise=noise, **full, **extra)


@pytest.mark.parametrize("in_ch,cc,B", [(3, 1, 21), (7, 5, 19)])
def test_fused_inference_kernels_vs_oracle_L4096(in_ch, cc, B):
    """Every fused block output and eps_hat of the benchmarked bf16 path directly against the oracle (no CUDA-vs-CUDA hop)."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.engine import ModelSpec, UNetEngine
    L = 4096
    sd = make_state_dict(in_ch, cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    x = gaussian((B, in_ch, L), seed=17 + in_ch)
    t = torch.tensor(([24, 999, 500, 3, 250, 55, 289, 0] * B)[:B])
    with torch.no_grad():
        taps = oracle.unet_forward_taps(sd, cfg, x, t)
    spec = ModelSpec(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    eng = UNetEngine({k: v.cuda() for k, v in sd.items()}, spec, dtype="bf16", conv_impl="tc")
    # (a) the product configuration: first block fused, six conv_gn blocks, head dots
    for _ in range(2):
        eps = eng.forward(x.cuda(), t.cuda())
    ws = eng.workspace(B, L, False)
    assert ws.head_fused and sum(bool(v) for v in eng._fuse_ok.values()) == 6
    assert eng.lib.gw_conv_in_gn_group(in_ch, L, 64, cc) > 0
    for li, n in enumerate(NAMES[:-1]):                       # dec2's activation is not materialised with the head dots
        assert rel_l2(ws.out[li].float().transpose(1, 2), taps[n + ".out"]) <= BF16_TOL, (n, "out")
    err = rel_l2(eps, taps["eps"])
    assert err <= BF16_TOL, err
    per_sample = ((eps.cpu() - taps["eps"]).flatten(1).norm(dim=1) / taps["eps"].flatten(1).norm(dim=1))
    assert float(per_sample.max()) <= BF16_TOL, per_sample            # no single sample (e.g. of a partial round) is off
    # (b) the same kernels writing the last activation (head dots off)
    eng.fuse_head = False
    eps_b = eng.forward(x.cuda(), t.cuda())
    assert rel_l2(ws.out[6].float().transpose(1, 2), taps["dec2.out"]) <= BF16_TOL
    assert rel_l2(eps_b, taps["eps"]) <= BF16_TOL


@pytest.mark.slow
def test_ddpm1000_bf16_chain_vs_oracle_L4096():
    """BASELINE config 1 / the `sampling` bench object: steps = 1000, eta = 1, L = 4096, bf16 fused kernels in one CUDA graph,
    identical injected noise.  Per-step eps_hat is checked teacher-forced on the oracle's own x_t (the free-running chain
    amplifies any rounding by 1/sqrt(alpha_bar_t) ~ 2e4 at t ~ 999)."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion
    L, B = 4096, 4
    sd = make_state_dict(3, 1, seed=1)
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    y = synthetic_chirps(B, L, snr=10.0, seed=78)["y_norm"]
    gen = torch.Generator().manual_seed(4242)
    noise = torch.randn(1000, B, 1, L, generator=gen)
    trace = []
    ref = oracle.ddim_sample(sd, cfg, ab, y, T=1000, steps=1000, eta=1.0, noise=list(noise), trace=trace)
    assert len(trace) == 1000
    model = _model(3, 1, seed=1, dtype="bf16")
    diff = CustomDiffusion(T=1000, device="cuda")
    eng = model.engine("bf16")
    for tq in (999, 529, 289, 55, 0):
        i = 999 - tq
        assert int(trace[i]["t"]) == tq
        sc = trace[i - 1]["x0"] if i > 0 else torch.zeros(B, 1, L)
        net = torch.cat([trace[i]["x_in"], y, sc], dim=1).cuda()
        eps = eng.forward(net, torch.full((B,), tq, dtype=torch.long, device="cuda"))
        err = rel_l2(eps, trace[i]["eps"])
        assert err <= BF16_TOL, (tq, err)
    out = _sample(model, diff, y, dict(steps=1000, eta=1.0), noise=noise, use_graph=True)
    ov, rl = overlap(out, ref), rel_l2(out, ref)
    print(f"ddpm1000 bf16 vs oracle: overlap {ov:.6f} rel-L2 {rl:.3e}")
    assert torch.isfinite(out).all()
    assert ov >= 0.999, (ov, rl)
    assert rl <= 3e-2, (ov, rl)
    out32 = _sample(model, diff, y, dict(steps=1000, eta=1.0), noise=noise, use_graph=True, compute_dtype="fp32")
    r32 = rel_l2(out32, ref)
    print(f"ddpm1000 fp32 vs oracle: rel-L2 {r32:.3e}")
    assert r32 <= 1e-3, r32             # 1000 steps of fp32 rounding in a different summation order than ATen


def test_chain_L16384_bf16_vs_oracle():
    """BASELINE config 4 shape: 16384-sample segments, 50 stochastic steps, bf16 fused kernels vs the oracle chain."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion
    L, B = 16384, 2
    sd = make_state_dict(3, 1, seed=1)
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    y = synthetic_chirps(B, L, snr=10.0, seed=81)["y_norm"]
    noise = torch.randn(51, B, 1, L, generator=torch.Generator().manual_seed(7))
    trace = []
    ref = oracle.ddim_sample(sd, cfg, ab, y, T=1000, steps=50, eta=1.0, noise=list(noise), trace=trace)
    model = _model(3, 1, seed=1, dtype="bf16")
    diff = CustomDiffusion(T=1000, device="cuda")
    eng = model.engine("bf16")
    for i in (0, 25, len(trace) - 1):
        sc = trace[i - 1]["x0"] if i > 0 else torch.zeros(B, 1, L)
        net = torch.cat([trace[i]["x_in"], y, sc], dim=1).cuda()
        tq = int(trace[i]["t"])
        eps = eng.forward(net, torch.full((B,), tq, dtype=torch.long, device="cuda"))
        assert rel_l2(eps, trace[i]["eps"]) <= BF16_TOL, (tq, rel_l2(eps, trace[i]["eps"]))
    out = _sample(model, diff, y, dict(steps=50, eta=1.0), noise=noise, use_graph=True)
    ov, rl = overlap(out, ref), rel_l2(out, ref)
    print(f"L16384 ddpm-50 bf16 vs oracle: overlap {ov:.6f} rel-L2 {rl:.3e}")
    assert ov >= 0.999 and rl <= 3e-2, (ov, rl)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_philox_chain_is_world_size_independent(dtype, use_graph):
    """x_T and every step's noise come from Philox(seed, global sample index, step): a batch sharded over ranks (or chunked
    through a cached plan / captured graph) reproduces the unsharded run bit for bit."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion
    L, B = 512, 8
    model = _model(3, 1, seed=1, dtype=dtype)
    diff = CustomDiffusion(T=1000, device="cuda")
    y = synthetic_chirps(B, L, snr=10.0, seed=79)["y_norm"]
    kw = dict(steps=6, eta=1.0, start_t=200)
    full = _sample(model, diff, y, kw, seed=1234, sample0=0, use_graph=use_graph)
    # two "ranks": same plan shape (B/2), so the second call replays the graph captured by the first with another key
    half_a = _sample(model, diff, y[:4], kw, seed=1234, sample0=0, use_graph=use_graph)
    half_b = _sample(model, diff, y[4:], kw, seed=1234, sample0=4, use_graph=use_graph)
    assert torch.equal(full, torch.cat([half_a, half_b], 0))
    assert not torch.equal(half_a, _sample(model, diff, y[:4], kw, seed=1235, sample0=0, use_graph=use_graph))
    again = _sample(model, diff, y[4:], kw, seed=1234, sample0=4, use_graph=use_graph)
    assert torch.equal(again, half_b)


def test_cached_plan_follows_weight_updates():
    """A cached SamplerPlan (FiLM table computed at construction) and the module's engines must see new weights: after
    load_state_dict, and after FusedTrainStep updated the flat parameter buffer through raw pointers."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep
    L, B = 256, 2
    diff = CustomDiffusion(T=1000, device="cuda")
    y = synthetic_chirps(B, L, snr=10.0, seed=5)["y_norm"]
    noise = torch.stack([gaussian((B, 1, L), seed=300 + k) for k in range(8)], 0)
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    kw = dict(steps=5, eta=1.0, start_t=300)
    model = _model(3, 1, seed=1, dtype="fp32")
    out1 = _sample(model, diff, y, kw, noise=noise)
    sd2 = make_state_dict(3, 1, seed=9)
    model.load_state_dict(sd2)
    out2 = _sample(model, diff, y, kw, noise=noise)                       # same cached plan, new weights
    ref2 = oracle.ddim_sample(sd2, cfg, ab, y, T=1000, noise=list(noise), **kw)
    assert rel_l2(out2, ref2) <= 1e-4 and rel_l2(out1, ref2) > 1e-2
    # training updates through the flat buffer; model(x, t) and ddim_sample(model, ...) afterwards use the new weights
    d = synthetic_chirps(B, L, snr=12.0, seed=31)
    st = FusedTrainStep(model, diff, B, L, lr=5e-2, p_uncond=0.0, ema_decay=None)
    st.load_batch(d["clean_norm"].cuda(), d["y_norm"].cuda(), None)
    st.step(use_graph=False)
    torch.cuda.synchronize()
    sd3 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    assert float((sd3["mid.0.weight"] - sd2["mid.0.weight"]).abs().max()) > 1e-3
    out3 = _sample(model, diff, y, kw, noise=noise)
    ref3 = oracle.ddim_sample(sd3, cfg, ab, y, T=1000, noise=list(noise), **kw)
    assert rel_l2(out3, ref3) <= 1e-4, rel_l2(out3, ref3)
    x = gaussian((B, 3, L), seed=2)
    t = torch.tensor([10, 900])
    with torch.no_grad():
        assert rel_l2(model(x.cuda(), t.cuda()), oracle.unet_forward(sd3, cfg, x, t)) <= 1e-5


def test_backward_bf16_selfcond_c7_L4096_vs_oracle():
    """Config-2 shape (L = 4096, in_ch = 7, self-conditioning forward on) at a batch the oracle finishes in seconds."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import UNet1D
    in_ch, cc, B, L = 7, 5, 2, 4096
    sd = make_state_dict(in_ch=in_ch, cond_in_ch=cc, seed=2)
    data = synthetic_chirps(B, L, snr=12.0, seed=31)
    clean, y = data["clean_norm"], data["y_norm"]
    mask = torch.ones(B, 1, L)
    mask[1, :, :301] = 0.0
    cond = torch.cat([y, gaussian((B, 4, 1), seed=6).expand(B, 4, L).contiguous() * 0.3], dim=1)
    t = torch.tensor([731, 999])
    eps = gaussian((B, 1, L), seed=41)
    drop = torch.tensor([0.0, 1.0]).view(B, 1, 1)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    loss_o, grads_o, eps_o = oracle.train_step(sd, cfg, ab, clean_norm=clean, cond_stack=cond, mask=mask, t=t, eps=eps, drop=drop,
                                               selfcond=True)
    m = UNet1D(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True, compute_dtype="bf16")
    m.load_state_dict(sd, strict=True)
    st = FusedTrainStep(m.cuda(), CustomDiffusion(T=1000, device="cuda"), B, L, p_uncond=0.2, warmup_steps=10, total_steps=100)
    st.load_batch(clean.cuda(), cond.cuda(), mask.cuda())
    st.step(selfcond=True, t=t.cuda(), eps=eps.cuda(), drop=drop.cuda(), use_graph=False)
    torch.cuda.synchronize()
    assert rel_l2(st.eps_hat, eps_o) <= BF16_TOL, rel_l2(st.eps_hat, eps_o)
    assert abs(float(st.loss) - float(loss_o)) <= 1e-2 * abs(float(loss_o))
    grads = st.layout.views(st.flat_g)
    flat_o = torch.cat([grads_o[k].reshape(-1) for k in grads_o]).double()
    flat_m = torch.cat([grads[k].cpu().reshape(-1) for k in grads_o]).double()
    cos = float((flat_o * flat_m).sum() / (flat_o.norm() * flat_m.norm()))
    assert cos >= 0.999, cos
    tot = float(flat_o.norm())
    for k, go in grads_o.items():
        err = float((grads[k].cpu().double() - go.double()).norm())
        assert err <= 5e-2 * max(float(go.norm()), 2e-2 * tot), (k, err, float(go.norm()))


@pytest.mark.parametrize("use_graph", [False, True])
def test_layer_chaining_is_bit_identical(use_graph):
    """gw_conv_gn3: the fused layer kernels launched with programmatic stream serialization, each ordering itself per SAMPLE
    through the producer's completion flags instead of a grid-wide dependency (option GWB200_CHAIN=1).  Same arithmetic, so the
    chain result is bit-identical to the plain launch order -- with more samples than CTA groups, twice through the same plan
    (the second chain must not be satisfied by the first chain's flags)."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion
    L, B = 4096, 40
    diff = CustomDiffusion(T=1000, device="cuda")
    y = synthetic_chirps(B, L, snr=10.0, seed=80)["y_norm"]
    noise = torch.stack([gaussian((B, 1, L), seed=700 + k) for k in range(14)], 0)
    kw = dict(steps=12, eta=1.0, start_t=529)
    outs = {}
    for chain in (False, True):
        model = _model(3, 1, seed=1, dtype="bf16")
        model.engine("bf16").chain_layers = chain
        a = _sample(model, diff, y, kw, noise=noise, use_graph=use_graph)
        b = _sample(model, diff, y, kw, noise=noise, use_graph=use_graph)
        assert torch.equal(a, b)
        outs[chain] = a
    assert torch.equal(outs[False], outs[True])


@pytest.mark.parametrize("in_ch,cc", [(3, 1), (7, 5)])
def test_single_group_epilogue_flavour_is_bit_identical(in_ch, cc):
    """conv_gn_kernel<..., ONE_T>: launches that carry at most one sample per CTA group let all 16 epilogue warps work on that sample
    (the second warpgroup would idle).  The statistics are summed in the same order in both flavours, so a sample's result does
    not depend on the batch it travels in: small batch (single-group flavour) == the same samples inside a large batch (two
    warpgroups) == the small batch with the flavour switched off."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, _cabi
    L, B = 4096, 40
    lib = _cabi.load()
    diff = CustomDiffusion(T=1000, device="cuda")
    data = synthetic_chirps(B, L, snr=10.0, seed=81)["y_norm"]
    y = data if cc == 1 else torch.cat([data, gaussian((B, 4, 1), seed=6).expand(B, 4, L).contiguous() * 0.3], dim=1)
    kw = dict(steps=6, eta=1.0, start_t=289)
    big = _sample(_model(in_ch, cc, seed=1, dtype="bf16"), diff, y, kw, seed=5, sample0=0)
    small = _sample(_model(in_ch, cc, seed=1, dtype="bf16"), diff, y[:6], kw, seed=5, sample0=0)
    assert lib.gw_set_option(b"one_group", 0) == 0
    try:
        small_two = _sample(_model(in_ch, cc, seed=1, dtype="bf16"), diff, y[:6], kw, seed=5, sample0=0)
    finally:
        lib.gw_set_option(b"one_group", 1)
    assert torch.equal(small, small_two)
    assert torch.equal(small, big[:6])
