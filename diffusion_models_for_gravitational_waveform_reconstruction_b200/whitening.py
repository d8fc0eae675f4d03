"""GPU whitening / de-whitening / sigma estimation around the reverse chain.

Batched fp64 equivalents of the reference's per-sample numpy helpers (inference.py:36-38, 125-205; the training data loader
uses the same train-like recipe, dataloader.py:110-151), bound from libgwb200_fft.so (include/gwb200_fft.h).  The functions
with the reference's names at the bottom take / return numpy arrays exactly like the reference (one sample), so
`inference.main`-style code can call them unchanged; the batched functions keep everything on the device.
`welch_psd` is scipy.signal.welch (defaults) on the device, `whiten_welch` the reference's Welch variant around it.
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
    "gwf_set_option": ([C.c_char_p, _I], _I),
    "gwf_workspace_bytes": ([_I, _I], _L),
    "gwf_whiten_train_like": ([_P, _P, _I, _I, _P, _P, _P, _P, _P], _I),
    "gwf_apply_psd": ([_P, _I, _I, _P, _I, _I, _P, _P, _P, _P], _I),
    "gwf_interp_psd": ([_P, _I, _I, _D, _P, _P], _I),
    "gwf_interp_psd_batch": ([_P, _I, _I, _I, _D, _P, _P], _I),
    "gwf_interp_grid": ([_P, _P, _I, _I, _I, _D, _P, _P], _I),
    "gwf_welch_workspace_bytes": ([_I, _I, _I], _L),
    "gwf_welch_psd": ([_P, _I, _I, _D, _I, _P, _P, _P], _I),
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


def apply_psd(sig: torch.Tensor, P: torch.Tensor, dewhiten: bool, out_dtype=torch.float64, loader_floor: bool = False) -> torch.Tensor:
    """rfft(sig) * sqrt(P + 1e-12) (dewhiten; `_dewhiten_train_like` / `_dewhiten_model`) or / sqrt(P + 1e-12) (`_whiten_pair_model`),
    then irfft.  P: [B, F] or one shared row [F].  `loader_floor`: whiten with the data loader's 1e-20 floor (dataloader.py:127-143)."""
    s = _prep(sig)
    B, L = s.shape
    P = P.to(s.device, torch.float64).contiguous()
    shared = 1 if P.ndim == 1 else 0
    o32 = torch.empty_like(s) if out_dtype == torch.float32 else None
    o64 = torch.empty(B, L, device=s.device, dtype=torch.float64) if out_dtype == torch.float64 else None
    w = _work(B, L, s.device)
    mode = 2 if dewhiten else (3 if loader_floor else 1)
    _check(load().gwf_apply_psd(_ptr(s), B, L, _ptr(P), shared, mode, _ptr(o32), _ptr(o64), _ptr(w), _stream()), "apply_psd")
    return o32 if o32 is not None else o64


def interp_psd_for_length(P_model: torch.Tensor, L_tgt: int, fs: float) -> torch.Tensor:
    """`_interp_psd_for_length` (inference.py:181-188) on the device: fp64 [L_tgt//2+1]."""
    Ps = P_model.to("cuda", torch.float64).contiguous()
    out = torch.empty(L_tgt // 2 + 1, device=Ps.device, dtype=torch.float64)
    _check(load().gwf_interp_psd(_ptr(Ps), Ps.numel(), L_tgt, float(fs), _ptr(out), _stream()), "interp_psd")
    return out


def interp_psd_batch(P_src: torch.Tensor, L_tgt: int, fs: float) -> torch.Tensor:
    """Per-sample `_interp_psd_for_length`: P_src fp64 [B, n_src] -> [B, L_tgt//2+1]."""
    Ps = P_src.to("cuda", torch.float64).contiguous()
    B, n_src = Ps.shape
    out = torch.empty(B, L_tgt // 2 + 1, device=Ps.device, dtype=torch.float64)
    _check(load().gwf_interp_psd_batch(_ptr(Ps), B, n_src, L_tgt, float(fs), _ptr(out), _stream()), "interp_psd_batch")
    return out


def interp_grid(xp: torch.Tensor, fp: torch.Tensor, L_tgt: int, fs: float) -> torch.Tensor:
    """np.interp(rfftfreq(L_tgt, 1/fs), xp, fp, left=fp[0], right=fp[-1]) per row (dataloader.py:136-139): [B, n] -> [B, L_tgt//2+1]."""
    xs = xp.to("cuda", torch.float64).contiguous()
    fs_ = fp.to("cuda", torch.float64).contiguous()
    if xs.ndim == 1:
        xs, fs_ = xs[None], fs_[None]
    B, n = xs.shape
    out = torch.empty(B, L_tgt // 2 + 1, device=xs.device, dtype=torch.float64)
    _check(load().gwf_interp_grid(_ptr(xs), _ptr(fs_), B, n, L_tgt, float(fs), _ptr(out), _stream()), "interp_grid")
    return out


def welch_psd(y: torch.Tensor, fs: float, nperseg: int) -> torch.Tensor:
    """scipy.signal.welch(y, fs=fs, nperseg=nperseg) with scipy's defaults, batched: y [B, L] CUDA -> Pxx fp64 [B, nperseg//2+1]
    (frequencies: rfftfreq(nperseg, 1/fs))."""
    yy = _prep(y)
    B, L = yy.shape
    nb = load().gwf_welch_workspace_bytes(B, L, int(nperseg))
    if nb <= 0:
        raise ValueError(f"welch_psd: nperseg={nperseg} for L={L}")
    w = torch.empty(nb, device=yy.device, dtype=torch.uint8)
    out = torch.empty(B, nperseg // 2 + 1, device=yy.device, dtype=torch.float64)
    _check(load().gwf_welch_psd(_ptr(yy), B, L, float(fs), int(nperseg), _ptr(out), _ptr(w), _stream()), "welch_psd")
    return out


def whiten_welch(y: torch.Tensor, x: Optional[torch.Tensor], fs: float):
    """Batched `_whiten_pair_welch` (inference.py:161-173): Welch PSD of y (nperseg = min(4096, L)), interpolated onto the rfft
    grid, divide by sqrt(P + 1e-12), no mean removal.  -> (y_w fp32, x_w fp32 or None, P fp64 [B, L//2+1])."""
    yy = _prep(y)
    L = yy.shape[1]
    P = interp_psd_batch(welch_psd(yy, fs, min(4096, L)), L, fs)
    y_w = apply_psd(yy, P, False, torch.float32)
    x_w = apply_psd(_prep(x), P, False, torch.float32) if x is not None else None
    return y_w, x_w, P


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


def _whiten_pair_welch(y: np.ndarray, x: Optional[np.ndarray], fs: float):
    """inference.py:161-173: -> (y_w, x_w, (freqs, P))."""
    y_w, x_w, P = whiten_welch(_dev(y), _dev(x) if x is not None else None, fs)
    freqs = np.fft.rfftfreq(len(y), 1 / fs)
    return y_w[0].cpu().numpy(), (x_w[0].cpu().numpy() if x_w is not None else None), (freqs, P[0].cpu().numpy())


def _dewhiten_welch(sig: np.ndarray, freqs_P, fs: float) -> np.ndarray:
    """inference.py:175-179."""
    return _dewhiten_train_like(sig, freqs_P[1])


def _mad_std(x: np.ndarray) -> float:
    return float(sigma(_dev(x), "mad")[0])


def _pick_sigma(y: np.ndarray, mode: str, fixed: float) -> float:
    if mode not in ("std", "mad", "fixed"):
        raise ValueError(f"unknown sigma-mode: {mode}")
    return float(fixed) if mode == "fixed" else float(sigma(_dev(y), mode)[0])
