"""Batched on-device scoring of reconstructions.

The reference scores one sample at a time on the host in numpy: tail-window Pearson correlation + MAE
(`inference._score_last_window`, inference.py:11-27), the cross-correlation lag search and aligned-window MAE
(`_best_lag_by_xcorr`, `_align_xcorr`, inference.py:247-279, 303-314) and the sweep objective
J = r_strain + 0.5 r_white - 0.1 NMAE_sigma (`sweep_infer._objective`, sweep_infer.py:8-13).  Here a whole batch is scored by
one kernel (`gw_score_batch`, fp64 accumulation), so an SNR sweep never leaves the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _cabi

__all__ = ["score_batch", "score_last_window", "best_lag_by_xcorr", "objective"]

_COLS = ["corr_last", "mae_last", "nmae_sigma", "overlap", "best_lag", "xc_mae", "xc_nmae_clean", "xc_nmae_sigma",
         "peak_index", "aligned_len", "xc_count", "tail_count"]


def score_batch(xhat: torch.Tensor, clean: torch.Tensor, fs: float, sigma: Optional[torch.Tensor] = None, secs: float = 0.8,
                max_shift: int = 0, delta_t: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """xhat, clean: [B, L] / [B, 1, L] CUDA tensors -> dict of fp64 [B] tensors (see include/gwb200.h: gw_score_batch)."""
    if xhat.device.type != "cuda":
        raise RuntimeError("gwb200 score_batch runs on CUDA (sm_100a) only: no CPU fallback")
    B, L = xhat.shape[0], xhat.shape[-1]
    x = xhat.reshape(B, L).float().contiguous()
    c = clean.to(xhat.device).reshape(B, L).float().contiguous()
    s = sigma.to(xhat.device).reshape(B).float().contiguous() if sigma is not None else None
    out = torch.empty(B, len(_COLS), device=xhat.device, dtype=torch.float64)
    _cabi.check(_cabi.load().gw_score_batch(_cabi.ptr(x), _cabi.ptr(c), _cabi.ptr(s), B, L, float(fs), float(secs), int(max_shift),
                                            float(delta_t if delta_t is not None else 1.0 / fs), _cabi.ptr(out),
                                            _cabi.stream_ptr()), "score_batch")
    return {k: out[:, i] for i, k in enumerate(_COLS)}


def score_last_window(x: torch.Tensor, c: torch.Tensor, fs: float, secs: float = 0.8) -> Dict[str, torch.Tensor]:
    """Batched `inference._score_last_window` (inference.py:20-27)."""
    r = score_batch(x, c, fs, secs=secs, max_shift=1)
    return {"corr_last": r["corr_last"], "mae_last": r["mae_last"]}


def best_lag_by_xcorr(a: torch.Tensor, b: torch.Tensor, max_shift: int = 0) -> torch.Tensor:
    """Batched `inference._best_lag_by_xcorr(a, b, max_shift)` (inference.py:247-262): int64 [B]."""
    return score_batch(b, a, 1.0, max_shift=max_shift)["best_lag"].long()


def objective(m_strain: Optional[Dict[str, torch.Tensor]], m_white: Optional[Dict[str, torch.Tensor]]) -> torch.Tensor:
    """`sweep_infer._objective` (sweep_infer.py:8-13) on tensors: r_s + 0.5 r_w - 0.1 nmae_sigma (missing terms count 0)."""
    r_s = m_strain["corr_last"] if m_strain else 0.0
    r_w = m_white["corr_last"] if m_white else 0.0
    nm = m_strain.get("nmae_sigma", 0.0) if m_strain else 0.0
    return r_s + 0.5 * r_w - 0.1 * nm
