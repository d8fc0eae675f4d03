"""Batched, device-resident version of the core of the reference deployment CLI (`inference.main`, inference.py:655-826):

    raw strain -> whiten -> sigma -> normalise -> conditioning stack -> start_t -> ddim_sample -> x sigma -> de-whiten -> scores

The reference does this for ONE sample with numpy on the host around a batch-1 `ddim_sample`; here every stage is a kernel
(whitening.py, inference.ddim_sample, scoring.py) and a whole batch stays on the GPU.  HDF5 / npy IO, plotting and argparse
stay out (out of scope).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import scoring, whitening
from .inference import ddim_sample, t_for_target_snr

# inference.py:706 fallback sigmas when the estimate is degenerate
_FALLBACK_SIGMA = {"train": 2.914e-12, "welch": 2.914e-16, "model": 2.914e-16, "raw": 2.914e-12}


@torch.no_grad()
def reconstruct_batch(model, diffusion, y_raw: torch.Tensor, *, fs: float, clean_raw: Optional[torch.Tensor] = None,
                      meta: Optional[torch.Tensor] = None, whiten: bool = True, whiten_mode: str = "train",
                      P_model: Optional[torch.Tensor] = None, sigma_mode: str = "std", sigma_fixed: float = 1.0,
                      start_snr: Optional[float] = None, start_t: Optional[int] = None, steps: int = 50, eta: float = 0.0,
                      init_mode: str = "noise", x0_std_est: float = 0.14, dc_weight: float = 0.0, cond_scale: float = 1.0,
                      eps_scale: float = 1.0, pred_type: str = "eps", cfg_scale: float = 1.0, cfg_mode: str = "const",
                      cfg_center: float = 0.5, cfg_width: float = 0.3, cfg_u_only_thresh: float = 0.0, drop_y_only: bool = True,
                      score_secs: float = 0.8, xcorr_window_samp: int = 0, seed: int = 0, sample0: int = 0,
                      noise: Optional[torch.Tensor] = None, compute_dtype: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """y_raw [B, L] (CUDA) -> dict(x0_hat_norm, x0_hat_white, x0_hat_strain [B, L], sigma [B], start_t, scores...)."""
    if y_raw.device.type != "cuda":
        raise RuntimeError("gwb200 reconstruct_batch runs on CUDA (sm_100a) only: no CPU fallback")
    dev = y_raw.device
    B, L = y_raw.shape[0], y_raw.shape[-1]
    y_raw = y_raw.reshape(B, L).float()
    clean = clean_raw.to(dev).reshape(B, L).float() if clean_raw is not None else None
    # ---- whitening (inference.py:655-700); 'auto' resolves model -> train (a saved Welch PSD is the data loader's business)
    P = None
    kind = "raw"
    if whiten:
        mode = whiten_mode
        if mode == "auto":
            mode = "model" if P_model is not None else "train"
        if mode == "welch":                                   # inference.py:676-679: Welch PSD estimated from y itself
            y_c, c_c, P = whitening.whiten_welch(y_raw, clean, fs)
            kind = "welch"
        elif mode == "model" and P_model is not None:
            P = whitening.interp_psd_for_length(P_model, L, fs)
            y_c = whitening.apply_psd(y_raw, P, dewhiten=False, out_dtype=torch.float32)
            c_c = whitening.apply_psd(clean, P, dewhiten=False, out_dtype=torch.float32) if clean is not None else None
            kind = "model"
        else:
            y_c, c_c, P = whitening.whiten_train_like(y_raw, clean)
            kind = "train"
    else:
        y_c, c_c = y_raw, clean
    # ---- sigma in the conditioning domain, with the reference's fallback (inference.py:703-717)
    sig = whitening.sigma(y_c, sigma_mode, sigma_fixed)
    bad = ~torch.isfinite(sig) | (sig < 1e-20)
    sig = torch.where(bad, torch.full_like(sig, _FALLBACK_SIGMA[kind]), sig)
    sig32 = sig.float().view(B, 1, 1)
    y_norm = (y_c.view(B, 1, L) / sig32)
    clean_norm = (c_c.view(B, 1, L) / sig32) if c_c is not None else None
    # ---- conditioning stack (inference.py:729-746): y + metadata channels (zeros when absent)
    Cc = model.cond_in_ch
    if Cc <= 1:
        cond = y_norm
    else:
        m = meta.to(dev).float() if meta is not None else torch.zeros(B, Cc - 1, L, device=dev)
        if m.ndim == 2:
            m = m[:, :, None].expand(B, Cc - 1, L)
        cond = torch.cat([y_norm, m], dim=1)
    # ---- start_t (inference.py:749-751)
    T = diffusion.T
    st = t_for_target_snr(diffusion, start_snr) if start_snr is not None else start_t
    x0n = ddim_sample(model, diffusion, cond.contiguous(), T, steps, eta, dev, L, False, st, init_mode, x0_std_est, dc_weight,
                      cond_scale, eps_scale, pred_type, model.in_ch, Cc, model.use_selfcond, cfg_scale, cfg_mode, cfg_center,
                      cfg_width, cfg_u_only_thresh, drop_y_only=drop_y_only, seed=seed, sample0=sample0, noise=noise,
                      compute_dtype=compute_dtype)
    x0_white = (x0n * sig32).view(B, L)                                     # inference.py:815
    if kind in ("train", "model", "welch"):                                 # inference.py:818-826
        x0_strain = whitening.apply_psd(x0_white, P, dewhiten=True)
    else:
        x0_strain = x0_white.double()
    out = {"x0_hat_norm": x0n, "x0_hat_white": x0_white, "x0_hat_strain": x0_strain, "sigma": sig, "whiten_kind": kind,
           "start_t": (T - 1) if st is None else int(st)}
    if clean is not None:                                                   # scores (inference.py:836-865; sweep_infer.py:225-241)
        ms = scoring.score_batch(x0_strain.float(), clean, fs, sigma=sig.float(), secs=score_secs, max_shift=max(1, xcorr_window_samp))
        out["strain"] = ms
        if c_c is not None:
            out["white"] = scoring.score_batch(x0_white, c_c, fs, sigma=sig.float(), secs=score_secs, max_shift=1)
            out["objective"] = scoring.objective(ms, out["white"])
    return out
