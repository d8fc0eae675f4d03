"""CPU-side checks: the C-ABI library loads and exports everything include/gwb200.h declares; host-side logic
(schedule, cfg weights, state_dict contract, error behaviour) matches the reference-generated golden vectors."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _cabi
    return _cabi.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "gwb200.h")).read()
    names = set(re.findall(r"^\s*(?:int|long|const char\*)\s+(gw_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/gwb200.h but not exported"
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _cabi
    assert set(_cabi.exported_symbols()) <= names | {"gw_version", "gw_last_error"}
    assert lib.gw_version() >= 100


def test_fft_library_exports_every_declared_symbol(lib):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening
    hdr = open(os.path.join(ROOT, "include", "gwb200_fft.h")).read()
    names = set(re.findall(r"^\s*(?:int|long|const char\*)\s+(gwf_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert names == set(whitening.exported_symbols()) and len(names) == 11
    flib = whitening.load()
    for n in names:
        assert hasattr(flib, n)
    assert flib.gwf_workspace_bytes(2, 1024) == 2 * 1024 * 8 + 2 * 2 * 513 * 16


def test_conv_tc_shape_helpers(lib):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200._cabi import ConvTcShape
    # dec0 at L=1024: pair space, two phase tiles of 256 columns, 20 segments of 64
    s = ConvTcShape(2, 1, 4, 1024, 256, 512, 256, 256)
    assert lib.gw_conv_tc_packed_elems(ctypes.byref(s)) == 2 * 256 * 20 * 64
    assert lib.gw_conv_tc_n_part(ctypes.byref(s)) == 4 * 2
    s = ConvTcShape(1, 0, 4, 2048, 64, 2048, 0, 128)
    assert lib.gw_conv_tc_packed_elems(ctypes.byref(s)) == 128 * 3 * 64
    bad = ConvTcShape(1, 0, 4, 2048, 48, 2048, 0, 128)
    assert lib.gw_conv_tc_packed_elems(ctypes.byref(bad)) < 0
    assert b"C0" in lib.gw_last_error()


def test_state_dict_contract_matches_reference_layout():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import UNet1D
    from weights import make_state_dict
    for in_ch, cc, n in [(3, 1, 1061572), (7, 5, 1066948)]:
        m = UNet1D(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
        sd = make_state_dict(in_ch, cc)          # keys/shapes as loaded into the reference by make_golden.py
        assert list(m.state_dict().keys()).sort() == list(sd.keys()).sort()
        m.load_state_dict(sd, strict=True)
        assert len(m.state_dict()) == 60
        assert sum(p.numel() for p in m.parameters()) == n
        assert m.cond_in_ch == cc and m.use_selfcond and m.in_ch_total == in_ch
    m = UNet1D(in_ch=3)
    assert float(m.final.weight.abs().max()) == 0.0      # reference zero-inits the head (models.py:132-134)
    assert UNet1D(in_ch=1).cond_in_ch == 0 and not UNet1D(in_ch=1).use_selfcond
    assert UNet1D(in_ch=2).cond_in_ch == 1


def test_cpu_tensor_raises_not_falls_back():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import UNet1D
    m = UNet1D(in_ch=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            m(torch.zeros(1, 3, 256), torch.zeros(1, dtype=torch.long))


def test_host_schedule_matches_reference(golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, cosine_beta_schedule
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf
    g = dict(np.load(os.path.join(golden_dir, "schedule.npz")))
    d = CustomDiffusion(T=1000)
    assert np.array_equal(cosine_beta_schedule(1000).numpy(), g["betas"])
    assert np.array_equal(d.alpha_bar.numpy(), g["alpha_bar"])
    assert d.T == 1000 and d.device == "cpu"
    for key in [k for k in g if k.startswith("sched_")]:
        _, T, steps, st = key.split("_")
        st = None if st == "None" else int(st)
        assert np.array_equal(inf._build_t_schedule(int(T), int(steps), "cpu", st).numpy(), g[key]), key
    w = [inf._cfg_weight(i, 10, mode, 1.5, 0.5, 0.3) for mode in ["const", "tophat", "gauss"] for i in [0, 3, 5, 9]]
    assert np.array_equal(np.array(w), g["cfg_w"])
    assert [inf.t_for_target_snr(d, s) for s in [0.9, 2.0, 10.0, 20.0]] == list(g["t_for_snr"])
    with pytest.raises(ValueError):
        inf._cfg_weight(0, 10, "bogus", 1.0, 0.5, 0.3)
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.models import TimeEmbedding
    assert np.array_equal(TimeEmbedding(128, 999.0)(torch.from_numpy(g["temb_t"])).numpy(), g["temb_128"])


def test_flat_import_shims():
    import importlib
    import sys
    p = os.path.join(ROOT, "src", "snr_denoising")
    sys.path.insert(0, p)
    try:
        for name in ("models", "inference"):
            sys.modules.pop(name, None)
        models = importlib.import_module("models")
        inference = importlib.import_module("inference")
        assert hasattr(models, "UNet1D") and hasattr(models, "CustomDiffusion") and hasattr(models, "cosine_beta_schedule")
        assert hasattr(inference, "ddim_sample") and hasattr(inference, "_build_t_schedule")
    finally:
        sys.path.remove(p)
        for name in ("models", "inference"):
            sys.modules.pop(name, None)
