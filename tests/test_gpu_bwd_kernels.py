"""tcgen05 backward GEMMs (dgrad through conv_tc, wgrad_tc) against the CUDA-core fp32 kernels on the same bf16 tensors.

Both families read identical bf16 activations / gradients; they differ by bf16 weights (dgrad) and accumulation order, so
per-tensor rel-L2 <= 1e-2 (north_star bf16 bound)."""
import pytest
import torch

from weights import gaussian, make_state_dict, synthetic_chirps

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _run(in_ch, cc, B, L, dgrad, wgrad, wvariant=0, fuse_head=False):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep
    sd = make_state_dict(in_ch, cc, seed=4)
    m = UNet1D(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True, compute_dtype="bf16")
    m.load_state_dict(sd)
    m = m.cuda()
    st = FusedTrainStep(m, CustomDiffusion(T=1000, device="cuda"), B, L, compute_dtype="bf16", seed=5)
    st.bwd.dgrad_impl, st.bwd.wgrad_impl, st.bwd.wgrad_variant = dgrad, wgrad, wvariant
    st.bwd.fuse_head = fuse_head
    d = synthetic_chirps(B, L, seed=9)
    cond = d["y_norm"] if cc == 1 else torch.cat([d["y_norm"], 0.2 * gaussian((B, 4, 1), 3).expand(B, 4, L)], 1)
    st.load_batch(d["clean_norm"].cuda(), cond.contiguous().cuda(), None)
    st.step(use_graph=False)
    torch.cuda.synchronize()
    return st, {k: v.clone() for k, v in st.layout.views(st.flat_g).items()}


@pytest.mark.parametrize("in_ch,cc,B,L", [(3, 1, 3, 1024), (7, 5, 2, 4096), (3, 1, 5, 192)])
def test_tc_backward_matches_simt(in_ch, cc, B, L):
    _, ref = _run(in_ch, cc, B, L, "simt", "simt")
    for dg, wg, wv, what in [("simt", "tc", 1, "wgrad_tc (one box per tap)"), ("simt", "tc", 0, "wgrad_tc (shifted descriptors)"),
                             ("tc", "simt", 0, "dgrad_tc"), ("tc", "tc", 0, "both"), ("tc", "tc", 2, "both + fused head gradient"),
                             ("simt", "tc", 4, "wgrad_tc with atomic split-K")]:
        _, got = _run(in_ch, cc, B, L, dg, wg, (wv & 1) | (2 if wv & 4 else 0), fuse_head=bool(wv & 2))
        tot = float(torch.cat([v.reshape(-1) for v in ref.values()]).norm())
        for k in ref:
            err = float((got[k].double() - ref[k].double()).norm())
            assert err <= 1e-2 * max(float(ref[k].norm()), 1e-2 * tot), (what, k, err, float(ref[k].norm()))


def _run_gn(in_ch, cc, B, L, fused, slice_bytes=None):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _cabi
    lib = _cabi.load()
    if slice_bytes is not None:
        assert lib.gw_set_option(b"gn_bwd_fused_slice", slice_bytes) == 0
    try:
        from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
        from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep
        sd = make_state_dict(in_ch, cc, seed=4)
        m = UNet1D(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True, compute_dtype="bf16")
        m.load_state_dict(sd)
        m = m.cuda()
        st = FusedTrainStep(m, CustomDiffusion(T=1000, device="cuda"), B, L, compute_dtype="bf16", seed=5)
        st.bwd.fuse_gn_bwd = fused
        d = synthetic_chirps(B, L, seed=9)
        cond = d["y_norm"] if cc == 1 else torch.cat([d["y_norm"], 0.2 * gaussian((B, 4, 1), 3).expand(B, 4, L)], 1)
        st.load_batch(d["clean_norm"].cuda(), cond.contiguous().cuda(), None)
        st.step(use_graph=False)
        st.step(use_graph=False)           # second launch: the exchange epochs advance, no reset of the sync buffer
        torch.cuda.synchronize()
        return {k: v.clone() for k, v in st.layout.views(st.flat_g).items()}, float(st.loss)
    finally:
        if slice_bytes is not None:
            lib.gw_set_option(b"gn_bwd_fused_slice", 45056)


@pytest.mark.parametrize("in_ch,cc,B,L,slice_bytes", [(7, 5, 2, 4096, None), (3, 1, 3, 1024, None), (3, 1, 5, 192, None),
                                                       (7, 5, 48, 512, 12000)])
def test_one_pass_gn_backward_matches_two_pass(in_ch, cc, B, L, slice_bytes):
    """gn_bwd_fused.cu (operands read once, group sums exchanged between the CTAs of a sample) against the two-pass streaming
    kernels: same math, different summation order of the fp32 sums -> per-tensor rel-L2 <= 5e-3.  The last case forces 16 CTAs
    per sample and more samples than CTA groups (every group loops over several samples)."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _cabi
    lib = _cabi.load()
    if slice_bytes is not None:
        lib.gw_set_option(b"gn_bwd_fused_slice", slice_bytes)
    g0 = lib.gw_gn_bwd_fused_group(L, 64, cc, 1, 1)
    if slice_bytes is not None:
        lib.gw_set_option(b"gn_bwd_fused_slice", 45056)
    assert g0 > 0, "the first block must take the one-pass kernel in this test"
    ref, loss_ref = _run_gn(in_ch, cc, B, L, False)
    got, loss_got = _run_gn(in_ch, cc, B, L, True, slice_bytes)
    assert abs(loss_ref - loss_got) <= 1e-4 * abs(loss_ref)
    tot = float(torch.cat([v.reshape(-1) for v in ref.values()]).norm())
    for k in ref:
        err = float((got[k].double() - ref[k].double()).norm())
        assert err <= 5e-3 * max(float(ref[k].norm()), 1e-2 * tot), (k, err, float(ref[k].norm()))


@pytest.mark.parametrize("in_ch,cc,B,L", [(7, 5, 2, 4096), (3, 1, 3, 1024), (3, 1, 5, 200)])
def test_specialised_gn_backward_matches_generic(in_ch, cc, B, L):
    """Compile-time-specialised streaming GroupNorm-backward kernels (+ the conv-bias gradient formed analytically from
    sum_l xhat instead of summing the bf16-rounded d_raw) against the generic streaming kernels."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _cabi
    lib = _cabi.load()
    assert lib.gw_set_option(b"gn_bwd_stats_fast", 0) == 0
    try:
        ref, loss_ref = _run_gn(in_ch, cc, B, L, False)
    finally:
        lib.gw_set_option(b"gn_bwd_stats_fast", 1)
    got, loss_got = _run_gn(in_ch, cc, B, L, False)
    assert abs(loss_ref - loss_got) <= 1e-4 * abs(loss_ref)
    tot = float(torch.cat([v.reshape(-1) for v in ref.values()]).norm())
    for k in ref:
        err = float((got[k].double() - ref[k].double()).norm())
        assert err <= 5e-3 * max(float(ref[k].norm()), 1e-2 * tot), (k, err, float(ref[k].norm()))
