"""GPU whitening / de-whitening / sigma estimation around the reverse chain.

Batched fp64 equivalents of the reference's per-sample numpy helpers (inference.py:36-38, 125-205; the training data loader
uses the same train-like recipe, dataloader.py:110-151), bound from libgwb200_fft.so (include/gwb200_fft.h).  The functions
with the reference's names at the bottom take / return numpy arrays exactly like the reference (one sample), so
`inference.main`-style code can call them unchanged; the batched functions keep everything on the device.
The Welch variant (`_whiten_pair_welch`, scipy.signal.welch) is not implemented.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np
import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_PKG, "libgwb200_fft.so")
_lib = None
_P, _I, _L, _D = C.c_void_p, C.c_int, C.c_long, C.c_double
_SIGS = {
    "gwf_last_error": ([], C.c_char_p),
    "gwf_workspace_bytes": ([_I, _I], _L),
    "gwf_whiten_train_like": ([_P, _P, _I, _I, _P, _P, _P, _P, _P], _I),
    "gwf_apply_psd": ([_P, _I, _I, _P, _I, _I, _P, _P, _P, _P], _I),
    "gwf_interp_psd": ([_P, _I, _I, _D, _P, _P], _I),
    "gwf_sigma": ([_P, _I, _I, _I, _P, _P], _I),
}


def exported_symbols():
    return list(_SIGS)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for this path)")
        lib = C.CDLL(_LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, res
        _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"gwb200_fft {what} failed ({rc}): {load().gwf_last_error().decode()}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _prep(a: torch.Tensor) -> torch.Tensor:
    if a.device.type != "cuda":
        raise RuntimeError("gwb200 whitening runs on CUDA (sm_100a) only: no CPU fallback")
    return a.reshape(a.shape[0], a.shape[-1]).float().contiguous()


def _work(B, L, device):
    return torch.empty(load().gwf_workspace_bytes(B, L), device=device, dtype=torch.uint8)


def whiten_train_like(y: torch.Tensor, x: Optional[torch.Tensor] = None):
    """Batched `_whiten_pair_train_like` (inference.py:137-153): mean removal, rfft, 9-tap smoothed periodogram floored at
    1e-20, divide by sqrt(P), irfft.  y, x: [B, L] CUDA -> (y_w fp32, x_w fp32 or None, P fp64 [B, L//2+1])."""
    y = _prep(y)
    B, L = y.shape
    xx = _prep(x) if x is not None else None
    y_w = torch.empty_like(y)
    x_w = torch.empty_like(y) if xx is not None else None
    P = torch.empty(B, L // 2 + 1, device=y.device, dtype=torch.float64)
    w = _work(B, L, y.device)
    _check(load().gwf_whiten_train_like(_ptr(y), _ptr(xx), B, L, _ptr(y_w), _ptr(x_w), _ptr(P), _ptr(w), _stream()), "whiten_train_like")
    return y_w, x_w, P


def apply_psd(sig: torch.Tensor, P: torch.Tensor, dewhiten: bool, out_dtype=torch.float64) -> torch.Tensor:
    """rfft(sig) * sqrt(P + 1e-12) (dewhiten; `_dewhiten_train_like` / `_dewhiten_model`) or / sqrt(P + 1e-12) (`_whiten_pair_model`),
    then irfft.  P: [B, F] or one shared row [F]."""
    s = _prep(sig)
    B, L = s.shape
    P = P.to(s.device, torch.float64).contiguous()
    shared = 1 if P.ndim == 1 else 0
    o32 = torch.empty_like(s) if out_dtype == torch.float32 else None
    o64 = torch.empty(B, L, device=s.device, dtype=torch.float64) if out_dtype == torch.float64 else None
    w = _work(B, L, s.device)
    _check(load().gwf_apply_psd(_ptr(s), B, L, _ptr(P), shared, 2 if dewhiten else 1, _ptr(o32), _ptr(o64), _ptr(w), _stream()), "apply_psd")
    return o32 if o32 is not None else o64


def interp_psd_for_length(P_model: torch.Tensor, L_tgt: int, fs: float) -> torch.Tensor:
    """`_interp_psd_for_length` (inference.py:181-188) on the device: fp64 [L_tgt//2+1]."""
    Ps = P_model.to("cuda", torch.float64).contiguous()
    out = torch.empty(L_tgt // 2 + 1, device=Ps.device, dtype=torch.float64)
    _check(load().gwf_interp_psd(_ptr(Ps), Ps.numel(), L_tgt, float(fs), _ptr(out), _stream()), "interp_psd")
    return out


def sigma(y: torch.Tensor, mode: str = "std", fixed: float = 1.0) -> torch.Tensor:
    """Batched `_pick_sigma` (inference.py:125-135): "std" (population, fp64), "mad" (1.4826 * MAD + 1e-24) or "fixed"."""
    if mode == "fixed":
        return torch.full((y.shape[0],), float(fixed), device=y.device, dtype=torch.float64)
    if mode not in ("std", "mad"):
        raise ValueError(f"unknown sigma-mode: {mode}")
    yy = _prep(y)
    out = torch.empty(yy.shape[0], device=yy.device, dtype=torch.float64)
    _check(load().gwf_sigma(_ptr(yy), yy.shape[0], yy.shape[1], 0 if mode == "std" else 1, _ptr(out), _stream()), "sigma")
    return out


# ---------------------------------------------------------------------------------------------- reference-named, per sample
def _dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()[None]


def _whiten_pair_train_like(y: np.ndarray, x: Optional[np.ndarray], fs: float) -> Tuple[np.ndarray, Optional[np.ndarray], np.ndarray]:
    y_w, x_w, P = whiten_train_like(_dev(y), _dev(x) if x is not None else None)
    return y_w[0].cpu().numpy(), (x_w[0].cpu().numpy() if x_w is not None else None), P[0].cpu().numpy()


def _dewhiten_train_like(sig: np.ndarray, P: np.ndarray) -> np.ndarray:
    return apply_psd(_dev(sig), torch.from_numpy(np.asarray(P, dtype=np.float64)), True)[0].cpu().numpy()


def _interp_psd_for_length(P: np.ndarray, L_src: int, L_tgt: int, fs: float) -> np.ndarray:
    return interp_psd_for_length(torch.from_numpy(np.asarray(P, dtype=np.float64)), L_tgt, fs).cpu().numpy()


def _whiten_pair_model(y: np.ndarray, x: Optional[np.ndarray], P_model: np.ndarray, fs: float):
    P = interp_psd_for_length(torch.from_numpy(np.asarray(P_model, dtype=np.float64)), len(y), fs)
    y_w = apply_psd(_dev(y), P, False, torch.float32)[0].cpu().numpy()
    x_w = apply_psd(_dev(x), P, False, torch.float32)[0].cpu().numpy() if x is not None else None
    return y_w, x_w, P.cpu().numpy()


def _dewhiten_model(sig: np.ndarray, P: np.ndarray) -> np.ndarray:
    return _dewhiten_train_like(sig, P)


def _mad_std(x: np.ndarray) -> float:
    return float(sigma(_dev(x), "mad")[0])


def _pick_sigma(y: np.ndarray, mode: str, fixed: float) -> float:
    if mode not in ("std", "mad", "fixed"):
        raise ValueError(f"unknown sigma-mode: {mode}")
    return float(fixed) if mode == "fixed" else float(sigma(_dev(y), mode)[0])
