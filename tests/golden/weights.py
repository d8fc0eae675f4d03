"""Deterministic, torch-version-independent weights and inputs for the parity fixtures.

Everything comes from numpy's PCG64 (stable across numpy versions), so the same tensors can be
regenerated on the GPU box without shipping 4 MB of weights.  Shapes follow the reference's
`UNet1D.state_dict()` contract (models.py:105-152).  `final.*` is drawn
from N(0, 0.05^2) because the reference zero-initialises it (models.py:132-134), which would
make every comparison pass trivially.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def make_state_dict(in_ch: int = 3, cond_in_ch: int = 1, base_ch: int = 64, depth: int = 3,
                    time_dim: int = 128, kernel: int = 3, seed: int = 0) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(max(fan_in, 1))
        return torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = uni((cout, cin, k), cin * k)
        sd[name + ".bias"] = uni((cout,), cin * k)

    def lin(name, cout, cin):
        sd[name + ".weight"] = uni((cout, cin), cin)
        sd[name + ".bias"] = uni((cout,), cin)

    def gn(name, c):
        sd[name + ".weight"] = torch.from_numpy((1.0 + 0.2 * rng.uniform(-1, 1, size=(c,))).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy((0.2 * rng.uniform(-1, 1, size=(c,))).astype(np.float32))

    chs = [base_ch * (2 ** i) for i in range(depth)]
    lin("time_mlp.1", base_ch, time_dim)
    cin = in_ch
    for i, c in enumerate(chs):
        conv(f"encoders.{i}.0", c, cin, kernel)
        gn(f"encoders.{i}.1", c)
        cin = c
    conv("mid.0", cin, cin, kernel)
    gn("mid.1", cin)
    prev = chs[-1]
    for i, c in enumerate(reversed(chs)):
        conv(f"decoders.{i}.0", c, prev + c, kernel)
        gn(f"decoders.{i}.1", c)
        prev = c
    sd["final.weight"] = torch.from_numpy((0.05 * rng.standard_normal((1, prev + 1, kernel))).astype(np.float32))
    sd["final.bias"] = torch.from_numpy((0.05 * rng.standard_normal((1,))).astype(np.float32))
    for i, c in enumerate(chs):
        lin(f"tproj_enc.{i}.1", 2 * c, base_ch)
    lin("tproj_mid.1", 2 * chs[-1], base_ch)
    for i, c in enumerate(reversed(chs)):
        lin(f"tproj_dec.{i}.1", 2 * c, base_ch)
    if cond_in_ch > 0:
        for i, c in enumerate(chs):
            conv(f"cond_enc.{i}", c, cond_in_ch, 1)
        conv("cond_mid", chs[-1], cond_in_ch, 1)
        for i, c in enumerate(reversed(chs)):
            conv(f"cond_dec.{i}", c, cond_in_ch, 1)
    return sd


def synthetic_chirps(B: int, L: int, snr: float = 10.0, seed: int = 1234, fs: float = 4096.0,
                     snr_hi: float | None = None):
    """Whitened-domain Newtonian chirp + unit white noise.

    Returns dict(clean_norm [B,1,L], y_norm [B,1,L], sigma [B], snr [B]) as float32 torch tensors.
    If `snr_hi` is given, per-sample SNR ~ U[snr, snr_hi].
    """
    rng = np.random.default_rng(seed)
    n = np.arange(L, dtype=np.float64)
    clean = np.zeros((B, L), dtype=np.float64)
    snrs = np.full((B,), float(snr))
    if snr_hi is not None:
        snrs = rng.uniform(snr, snr_hi, size=(B,))
    for b in range(B):
        f0, fmax = 30.0, 400.0
        tc = 0.9 * L / fs                               # coalescence at 90 % of the segment
        dur = rng.uniform(0.35, 0.8) * tc               # chirp-mass-like knob: how long the inspiral is in band
        phi0 = rng.uniform(0, 2 * np.pi)
        tsec = n / fs
        tau = np.clip(tc - tsec, 1e-6, None)
        # f(tau) = f0 * (tau/dur)^(-3/8), capped
        f = np.minimum(f0 * (tau / dur) ** (-3.0 / 8.0), fmax)
        phase = 2 * np.pi * np.cumsum(f) / fs + phi0
        h = f ** (2.0 / 3.0) * np.cos(phase)
        h[tsec < (tc - dur)] = 0.0
        post = tsec >= tc
        h[post] = (fmax ** (2.0 / 3.0)) * np.exp(-(tsec[post] - tc) / 0.004) * np.cos(
            2 * np.pi * 250.0 * (tsec[post] - tc) + phase[np.argmax(post) - 1])
        on = np.nonzero(h)[0]
        if len(on) > 8:                                 # Tukey(0.1) taper over the support
            m = len(on)
            k = max(1, int(0.05 * m))
            win = np.ones(m)
            ramp = 0.5 * (1 - np.cos(np.pi * np.arange(k) / k))
            win[:k] = ramp
            win[-k:] = ramp[::-1]
            h[on] *= win
        h *= snrs[b] / (np.linalg.norm(h) + 1e-30)
        clean[b] = h
    noise = rng.standard_normal((B, L))
    y = clean + noise
    sigma = y.std(axis=1)
    out = {
        "clean_norm": torch.from_numpy((clean / sigma[:, None]).astype(np.float32))[:, None, :],
        "y_norm": torch.from_numpy((y / sigma[:, None]).astype(np.float32))[:, None, :],
        "sigma": torch.from_numpy(sigma.astype(np.float32)),
        "snr": torch.from_numpy(snrs.astype(np.float32)),
    }
    return out


def gaussian(shape, seed: int) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
