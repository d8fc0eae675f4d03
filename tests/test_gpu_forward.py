"""GPU parity: CUDA engine (through the C-ABI) vs the CPU oracle and the reference-generated golden vectors.

Tolerances (BASELINE.json north_star): per-layer and eps_hat rel-L2 <= 1e-5 in fp32-exact mode, <= 1e-2 in bf16 mode.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from weights import make_state_dict, gaussian

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 1e-2
NAMES = ["enc0", "enc1", "enc2", "mid", "dec0", "dec1", "dec2"]


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _engine(sd, in_ch, cc, dtype, impl):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.engine import ModelSpec, UNetEngine
    spec = ModelSpec(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    return UNetEngine({k: v.cuda() for k, v in sd.items()}, spec, dtype=dtype, conv_impl=impl)


@pytest.mark.parametrize("mode,tol", [("fp32", FP32_TOL), ("bf16_simt", BF16_TOL), ("bf16_tc", BF16_TOL)])
@pytest.mark.parametrize("in_ch,cc,L,B", [(3, 1, 1024, 2), (7, 5, 2048, 3), (3, 1, 4096, 1)])
def test_forward_per_layer_vs_oracle(mode, tol, in_ch, cc, L, B):
    sd = make_state_dict(in_ch, cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    x = gaussian((B, in_ch, L), seed=7 + L)
    t = torch.tensor(([24, 999, 500] * B)[:B])
    with torch.no_grad():
        taps = oracle.unet_forward_taps(sd, cfg, x, t)
    dtype, impl = {"fp32": ("fp32", "simt"), "bf16_simt": ("bf16", "simt"), "bf16_tc": ("bf16", "tc")}[mode]
    eng = _engine(sd, in_ch, cc, dtype, impl)
    eps = eng.forward(x.cuda(), t.cuda(), keep_raw=True)
    ws = eng.workspace(B, L, True)
    for li, n in enumerate(NAMES):
        assert rel_l2(ws.raw[li].float().transpose(1, 2), taps[n + ".raw"]) <= tol, (n, "raw")
        assert rel_l2(ws.out[li].float().transpose(1, 2), taps[n + ".out"]) <= tol, (n, "out")
    assert rel_l2(eps, taps["eps"]) <= tol
    if mode == "bf16_tc":
        assert all(eng.tc_supported(li, ws.lay_len[li], ws.lay_len[li] if li <= 3 else ws.lay_len[li] // 2)
                   for li in range(1, 7)), "tcgen05 path was not used"


@pytest.mark.parametrize("L", [500, 1000])
def test_forward_odd_lengths_fp32(L):
    """pad/trim semantics when L is not divisible by 8 (models.py:218-220, 227-229)."""
    sd = make_state_dict(3, 1, seed=0)
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    x = gaussian((2, 3, L), seed=11)
    t = torch.tensor([529, 3])
    with torch.no_grad():
        ref = oracle.unet_forward(sd, cfg, x, t)
    for dtype, impl, tol in [("fp32", "simt", FP32_TOL), ("bf16", "tc", BF16_TOL)]:
        eng = _engine(sd, 3, 1, dtype, impl)
        assert rel_l2(eng.forward(x.cuda(), t.cuda()), ref) <= tol


def test_forward_matches_reference_golden(golden_dir):
    """eps_hat against the vectors written by the unmodified reference (tests/golden/make_golden.py)."""
    for tag, in_ch, cc, L, B in [("c3_L256", 3, 1, 256, 2), ("c7_L256", 7, 5, 256, 2), ("c3_L500", 3, 1, 500, 1),
                                 ("c7_L1024", 7, 5, 1024, 1)]:
        g = dict(np.load(os.path.join(golden_dir, f"forward_{tag}.npz")))
        sd = make_state_dict(in_ch, cc, seed=0)
        x = gaussian((B, in_ch, L), seed=100 + L + in_ch)
        if cc == 5:
            x[:, 2:6, :] = x[:, 2:6, :1].clone()
        t = torch.from_numpy(g["t"])
        ref = torch.from_numpy(g["eps"])
        eng = _engine(sd, in_ch, cc, "fp32", "simt")
        eps = eng.forward(x.cuda(), t.cuda(), keep_raw=True)
        assert rel_l2(eps, ref) <= FP32_TOL, tag
        ws = eng.workspace(B, L, True)
        for li, n in enumerate(NAMES):
            sub = ws.raw[li].float().transpose(1, 2)[:, ::4, ::4]
            assert rel_l2(sub, torch.from_numpy(g[n + ".raw"])) <= FP32_TOL, (tag, n)
        engb = _engine(sd, in_ch, cc, "bf16", "tc")
        assert rel_l2(engb.forward(x.cuda(), t.cuda()), ref) <= BF16_TOL, tag


def test_module_api_forward_and_errors():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import UNet1D
    sd = make_state_dict(3, 1, seed=0)
    m = UNet1D(in_ch=3)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x = gaussian((2, 3, 512), seed=1)
    t = torch.tensor([10, 700])
    cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
    with torch.no_grad():
        ref = oracle.unet_forward(sd, cfg, x, t)
        out = m(x.cuda(), t.cuda())
        assert out.shape == (2, 1, 512) and out.dtype == torch.float32
        assert rel_l2(out, ref) <= FP32_TOL
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outb = m(x.cuda(), t.cuda())
        assert rel_l2(outb, ref) <= BF16_TOL
        with pytest.raises(ValueError):
            m(torch.zeros(2, 4, 512, device="cuda"), t.cuda())
    # a weight update must be picked up (packed bf16 copies are refreshed)
    with torch.no_grad():
        m.final.weight.mul_(2.0)
        m.final.bias.mul_(2.0)
        out2 = m(x.cuda(), t.cuda())
    assert rel_l2(out2, 2.0 * ref) <= FP32_TOL


def test_q_sample_kernel():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion
    d = CustomDiffusion(T=1000, device="cuda")
    x0 = gaussian((3, 1, 777), seed=2)
    eps = gaussian((3, 1, 777), seed=3)
    t = torch.tensor([0, 500, 999])
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    ref = oracle.q_sample(ab, x0, t, eps)
    xt, e = d.q_sample(x0.cuda(), t.cuda(), noise=eps.cuda())
    assert torch.equal(e.cpu(), eps)
    assert rel_l2(xt, ref) <= 1e-6
    xt2, e2 = d.q_sample(x0.cuda(), t.cuda())
    assert abs(float(e2.std()) - 1.0) < 0.1


@pytest.mark.parametrize("in_ch,cc,L,B", [(3, 1, 1024, 2), (7, 5, 4096, 2), (3, 1, 500, 3), (7, 5, 250, 1)])
@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_fused_first_block(in_ch, cc, L, B, dtype, tol):
    """Inference path of the first block (gw_conv_in_block: stats pass + recompute/apply pass, no raw tensor)."""
    import torch.nn.functional as F
    sd = make_state_dict(in_ch, cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    x = gaussian((B, in_ch, L), seed=3 + L)
    t = torch.tensor(([24, 999, 500] * B)[:B])
    with torch.no_grad():
        taps = oracle.unet_forward_taps(sd, cfg, x, t)
    eng = _engine(sd, in_ch, cc, dtype, "auto")
    eng.fuse_first_block = True
    eps = eng.forward(x.cuda(), t.cuda())
    ws = eng.workspace(B, L, False)
    assert rel_l2(ws.out[0].float().transpose(1, 2), taps["enc0.out"]) <= tol
    assert rel_l2(ws.pooled[0].float().transpose(1, 2), F.avg_pool1d(taps["enc0.out"], 2, 2)) <= tol
    assert rel_l2(eps, taps["eps"]) <= tol
    eng.fuse_first_block = False
    eps2 = eng.forward(x.cuda(), t.cuda())
    assert rel_l2(eps2, taps["eps"]) <= tol


@pytest.mark.parametrize("in_ch,cc,L,B", [(3, 1, 4096, 21), (7, 5, 4096, 5), (3, 1, 512, 5), (7, 5, 256, 3), (3, 1, 16384, 2),
                                          (7, 5, 1536, 4)])
@pytest.mark.parametrize("keep_raw", [False, True])
def test_fused_conv_gn_block(in_ch, cc, L, B, keep_raw):
    """gw_conv_gn (conv + GroupNorm + SiLU + cond + FiLM + pool in one kernel, statistics exchanged between the CTAs of a
    sample) against gw_conv_tc + gw_gn_apply and against the oracle."""
    sd = make_state_dict(in_ch, cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    x = gaussian((B, in_ch, L), seed=3 + L)
    t = torch.tensor(([24, 999, 500, 3, 250] * B)[:B])
    fused, plain = _engine(sd, in_ch, cc, "bf16", "tc"), _engine(sd, in_ch, cc, "bf16", "tc")
    fused.fuse_gn_train = True
    fused.fuse_head = False                      # this test reads the last decoder's activation (see test_fused_head_dots)
    plain.fuse_gn = False
    # three forwards through the same workspace: the exchange epoch must advance between launches
    for _ in range(3):
        eps_f = fused.forward(x.cuda(), t.cuda(), keep_raw=keep_raw)
    eps_p = plain.forward(x.cuda(), t.cuda(), keep_raw=keep_raw)
    wf, wp = fused.workspace(B, L, keep_raw), plain.workspace(B, L, keep_raw)
    n_fused = sum(bool(v) for v in fused._fuse_ok.values())
    assert n_fused == 6, fused._fuse_ok
    for li, n in enumerate(NAMES):
        assert rel_l2(wf.out[li].float(), wp.out[li].float()) <= 5e-3, (n, "out")
        if li < 3:
            assert rel_l2(wf.pooled[li].float(), wp.pooled[li].float()) <= 5e-3, (n, "pooled")
        if keep_raw:
            # layer 1 sees identical inputs on both paths; deeper layers inherit the rounding differences of `out`
            assert rel_l2(wf.raw[li].float(), wp.raw[li].float()) <= (1e-6 if li <= 1 else 5e-3), (n, "raw")
            assert torch.allclose(wf.stats[li], wp.stats[li], rtol=1e-4 if li <= 1 else 2e-2, atol=1e-3), (n, "stats")
    assert rel_l2(eps_f, eps_p) <= BF16_TOL      # two bf16 roundings of the same network
    if B * L <= 3 * 4096:
        with torch.no_grad():
            ref = oracle.unet_forward(sd, cfg, x, t)
        assert rel_l2(eps_f, ref) <= BF16_TOL


@pytest.mark.parametrize("in_ch,cc,L,B", [(3, 1, 1000, 3), (7, 5, 2050, 2), (3, 1, 9000, 2), (1, 0, 768, 2)])
def test_fused_first_block_partial_tiles(in_ch, cc, L, B):
    """gw_conv_in_gn on lengths that leave a partial CTA slice (256 rows up to L = 8192, 512 beyond) and without
    conditioning; the deeper blocks of these lengths take the unfused kernels."""
    sd = make_state_dict(in_ch, cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=(in_ch > 1))
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.engine import ModelSpec, UNetEngine
    spec = ModelSpec(in_ch=in_ch, cond_in_ch=cc, use_selfcond=(in_ch > 1))
    fused = UNetEngine({k: v.cuda() for k, v in sd.items()}, spec, dtype="bf16", conv_impl="tc")
    plain = UNetEngine({k: v.cuda() for k, v in sd.items()}, spec, dtype="bf16", conv_impl="tc")
    plain.fuse_gn = False
    fused.direct_first = False                   # this test covers the exchange-based kernel (see test_direct_first_block)
    x = gaussian((B, in_ch, L), seed=5 + L)
    t = torch.tensor(([700, 12, 333] * B)[:B])
    assert fused.lib.gw_conv_in_gn_group(in_ch, L, 64, cc) > 0
    for _ in range(2):
        eps_f = fused.forward(x.cuda(), t.cuda())
    eps_p = plain.forward(x.cuda(), t.cuda())
    wf, wp = fused.workspace(B, L, False), plain.workspace(B, L, False)
    assert rel_l2(wf.out[0].float(), wp.out[0].float()) <= 2e-3
    assert rel_l2(wf.pooled[0].float(), wp.pooled[0].float()) <= 2e-3
    assert rel_l2(eps_f, eps_p) <= BF16_TOL
    if B * L <= 3 * 4096:
        with torch.no_grad():
            ref = oracle.unet_forward(sd, cfg, x, t)
        assert rel_l2(eps_f, ref) <= BF16_TOL


@pytest.mark.parametrize("in_ch,cc,L,B", [(3, 1, 4096, 21), (7, 5, 4096, 5), (3, 1, 512, 5), (7, 5, 1536, 4), (3, 1, 16384, 2)])
def test_fused_head_dots(in_ch, cc, L, B):
    """gw_conv_gn2: the last decoder's fused kernel leaves the three head-conv dot products per position (formed from its fp32
    epilogue values) and gw_final_step(dtype = GW_DOTS) finishes eps_hat from them -- against the path that writes the bf16
    activation and streams it back, and against the oracle."""
    sd = make_state_dict(in_ch, cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    x = gaussian((B, in_ch, L), seed=5 + L)
    t = torch.tensor(([24, 999, 500, 3, 250] * B)[:B])
    a, b = _engine(sd, in_ch, cc, "bf16", "tc"), _engine(sd, in_ch, cc, "bf16", "tc")
    assert a.fuse_head
    b.fuse_head = False
    for _ in range(2):
        eps_a = a.forward(x.cuda(), t.cuda())
    eps_b = b.forward(x.cuda(), t.cuda())
    assert a.workspace(B, L, False).head_fused and not b.workspace(B, L, False).head_fused
    assert rel_l2(eps_a, eps_b) <= 5e-3          # fp32 vs bf16-rounded activations under the head conv
    if B * L <= 5 * 4096:
        with torch.no_grad():
            eps_o = oracle.unet_forward(sd, cfg, x, t)
        assert rel_l2(eps_a.cpu(), eps_o) <= BF16_TOL


@pytest.mark.parametrize("in_ch,cc,L,B", [(3, 1, 4096, 21), (7, 5, 4096, 5), (3, 1, 16384, 2), (7, 5, 512, 3)])
def test_cta_pair_mma_equals_single_cta(in_ch, cc, L, B):
    """conv_gn_kernel with tcgen05 CTA pairs (cta_group::2: one M = 256 MMA per pair, each CTA staging half of every weight
    tile) against the single-CTA kernel.  Every output row is accumulated in the same K order; only the order in which the
    CTAs' GroupNorm partial sums are added differs (slice -> CTA mapping), i.e. fp32 rounding of the statistics."""
    sd = make_state_dict(in_ch, cc, seed=0)
    x = gaussian((B, in_ch, L), seed=9 + L)
    t = torch.tensor(([24, 999, 500, 3, 250] * B)[:B])
    eng = _engine(sd, in_ch, cc, "bf16", "tc")
    outs = []
    try:
        for pair2 in (0, 1):
            assert eng.lib.gw_set_option(b"pair2", pair2) == 0
            for _ in range(2):
                eps = eng.forward(x.cuda(), t.cuda())
            ws = eng.workspace(B, L, False)
            outs.append((eps.clone(), [o.clone() for o in ws.out[:6]], [p.clone() for p in ws.pooled]))
    finally:
        eng.lib.gw_set_option(b"pair2", 0)
    assert rel_l2(outs[1][0], outs[0][0]) <= 1e-3
    for a, b in zip(outs[0][1] + outs[0][2], outs[1][1] + outs[1][2]):
        assert rel_l2(b.float(), a.float()) <= 1e-3               # bf16 tensors: a last-bit flip here and there
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    if B * L <= 5 * 4096:
        with torch.no_grad():
            assert rel_l2(outs[1][0], oracle.unet_forward(sd, cfg, x, t)) <= BF16_TOL


@pytest.mark.parametrize("in_ch,cc,L,B", [(3, 1, 4096, 21), (7, 5, 4096, 5), (3, 1, 1000, 3), (7, 5, 2050, 2), (1, 0, 768, 2),
                                          (3, 1, 16384, 2), (7, 5, 250, 3)])
def test_direct_first_block(in_ch, cc, L, B):
    """gw_conv_in_direct: the first block in one pass, GroupNorm statistics from the analytic moments of the conv output (a
    quadratic form of the input's lagged cross products) -- block output, pooled output and eps_hat against the oracle, the
    statistics themselves against the oracle's raw conv output, inputs with a large self-conditioning channel included."""
    import torch.nn.functional as F
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.engine import ModelSpec, UNetEngine
    sd = make_state_dict(in_ch, cc, seed=0)
    sc = in_ch > 1
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=sc)
    spec = ModelSpec(in_ch=in_ch, cond_in_ch=cc, use_selfcond=sc)
    eng = UNetEngine({k: v.cuda() for k, v in sd.items()}, spec, dtype="bf16", conv_impl="tc")
    eng.direct_first = True
    assert eng.lib.gw_conv_in_direct_ws_floats(B, in_ch, L, 64, cc) > 0
    x = gaussian((B, in_ch, L), seed=13 + L)
    x[:, 0] += 0.7                                # a DC offset: the variance is a difference of large moments
    if sc:
        x[0, -1] *= 3000.0                        # x0_hat of the first reverse steps is ~1e4
    t = torch.tensor(([700, 12, 333] * B)[:B])
    with torch.no_grad():
        taps = oracle.unet_forward_taps(sd, cfg, x, t)
    for _ in range(2):
        eps = eng.forward(x.cuda(), t.cuda())
    ws = eng.workspace(B, L, False)
    assert ws.coef0 is not None
    assert rel_l2(ws.out[0].float().transpose(1, 2), taps["enc0.out"]) <= 5e-3
    assert rel_l2(ws.pooled[0].float().transpose(1, 2), F.avg_pool1d(taps["enc0.out"], 2, 2)) <= 5e-3
    assert rel_l2(eps, taps["eps"]) <= BF16_TOL
    # the analytic statistics: A = 0.5 rstd gn_w of channel pair 0 / 4 .. against the oracle's raw conv output
    raw = taps["enc0.raw"].double().reshape(B, 8, 8 * L)
    rstd = 1.0 / torch.sqrt(raw.var(dim=2, unbiased=False) + 1e-5)
    nca = cc if cc in (0, 1, 5) else 8
    cf = (8 + 2 * nca + 3) // 4 * 4
    coef = ws.coef0[: B * 32 * cf].view(B, 32, cf).cpu().double()
    a_ref = 0.5 * rstd[:, :, None] * sd["encoders.0.1.weight"].double().view(1, 8, 8)
    a_got = coef[:, :, 0:2].reshape(B, 64).view(B, 8, 8)
    assert float(((a_got - a_ref).abs() / a_ref.abs().clamp_min(1e-12)).max()) <= 2e-4
