"""Drop-in mirror of the reference `models` module (src/snr_denoising/models.py) backed by the sm_100a kernels.

Same public names, constructor arguments, attributes and `state_dict()` keys/shapes as the reference
(SURVEY.md section 8b / Appendix A), so `load_state_dict(ckpt['model_state'], strict=True)` works both ways.
The modules below only *hold parameters*; `UNet1D.forward` hands them to `UNetEngine` (engine.py), which runs the
hand-written CUDA kernels.  There is no PyTorch / CPU fallback: a CPU tensor raises.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .engine import ModelSpec, UNetEngine

__all__ = ["TimeEmbedding", "cosine_beta_schedule", "CustomDiffusion", "UNet1D"]


class TimeEmbedding(nn.Module):
    """Parameter-free placeholder for the sinusoidal embedding (models.py:9-31); evaluated inside gw_film_vectors."""

    def __init__(self, dim: int, max_time: float = 999.0):
        super().__init__()
        self.dim = dim
        self.max_time = float(max_time)

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        # host-side convenience only (schedule tooling); the network path never calls this
        ts = t.float() / max(self.max_time, 1.0)
        half = self.dim // 2
        k = torch.arange(half, dtype=torch.float32, device=t.device)
        arg = ts[:, None] * torch.exp(k * -(math.log(10000) / max(half - 1, 1)))[None, :]
        out = torch.cat([arg.sin(), arg.cos()], dim=1)
        if self.dim % 2 == 1:
            out = torch.cat([out, torch.zeros(t.size(0), 1, device=t.device)], dim=1)
        return out


def cosine_beta_schedule(T: int, s: float = 0.008) -> torch.Tensor:
    """models.py:34-40: fp32 cosine alpha-bar ratio schedule, clamped to [0, 0.999]."""
    grid = torch.linspace(0, T, T + 1, dtype=torch.float32)
    f = torch.cos(((grid / T) + s) / (1 + s) * (math.pi / 2)) ** 2
    f = f / f[0]
    return (1 - (f[1:] / f[:-1])).clamp(min=0.0, max=0.999)


class CustomDiffusion:
    """models.py:43-59.  `q_sample` runs the fused gw_q_sample kernel; `noise=`/`generator=` are additive kwargs."""

    def __init__(self, T: int = 1000, device="cpu"):
        self.device = device
        self.T = T
        betas = cosine_beta_schedule(T).to(device)
        self.alpha_bar = torch.cumprod(1.0 - betas, dim=0)
        self.betas = betas
        self._sq = None

    def _tables(self, device):
        if self._sq is None or self._sq[0].device != device:
            ab = self.alpha_bar.to(device)
            self._sq = (ab.sqrt().contiguous(), (1 - ab).sqrt().contiguous())
        return self._sq

    def q_sample(self, x0: torch.Tensor, t: torch.Tensor, noise: Optional[torch.Tensor] = None,
                 generator: Optional[torch.Generator] = None):
        from . import _cabi
        if x0.device.type != "cuda":
            raise RuntimeError("gwb200 q_sample: CUDA tensors only (no CPU fallback)")
        lib = _cabi.load()
        B, C, L = x0.shape
        x0c = x0.contiguous().float().reshape(B * C, L)
        tt = t.long().to(x0.device).reshape(-1).repeat_interleave(C).contiguous()
        sa, sm = self._tables(x0.device)
        if noise is None:
            noise = torch.randn(x0.shape, device=x0.device, dtype=torch.float32, generator=generator)
        eps = noise.contiguous().float().reshape(B * C, L)
        out = torch.empty(B * C, 1, L, device=x0.device, dtype=torch.float32)
        _cabi.check(lib.gw_q_sample(_cabi.ptr(x0c), _cabi.ptr(tt), _cabi.ptr(sa), _cabi.ptr(sm), _cabi.ptr(eps), 0, 0, 0, 0,
                                    0.0, _cabi.ptr(out), B * C, 1, L, _cabi.stream_ptr()), "q_sample")
        return out.view(B, C, L), eps.view(B, C, L)


class _Holder(nn.Module):
    """weight/bias container with the reference layer's default initialisation."""

    def __init__(self, wshape, fan_in: Optional[int], kind: str):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*wshape))
        self.bias = nn.Parameter(torch.empty(wshape[0]))
        with torch.no_grad():
            if kind == "norm":
                self.weight.fill_(1.0)
                self.bias.zero_()
            elif kind == "zero":                       # models.py:133-134
                self.weight.zero_()
                self.bias.zero_()
            else:                                      # torch Conv1d / Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                b = 1.0 / math.sqrt(fan_in)
                self.weight.uniform_(-b, b)
                self.bias.uniform_(-b, b)


class _Empty(nn.Module):
    pass


def _pair(a: nn.Module, b: nn.Module) -> nn.ModuleList:
    return nn.ModuleList([a, b])


class UNet1D(nn.Module):
    """models.py:62-231: 1-D U-Net with FiLM time conditioning and per-stage 1x1-conv conditioning bias.

    Input channels: [x_t | cond_0..cond_{K-1} | optional x0 self-conditioning]; output eps_hat [B,1,L].
    Extra (additive) kwargs: `compute_dtype` in {"fp32", "bf16"} selects the kernel family
    (fp32 = exact CUDA-core mode, bf16 = tcgen05 tensor-core mode).
    """

    def __init__(self, in_ch: int = 1, base_ch: int = 64, time_dim: int = 128, depth: int = 3, kernel: int = 3,
                 t_embed_max_time: float = 999.0, cond_in_ch: Optional[int] = None, use_selfcond: Optional[bool] = None,
                 compute_dtype: str = "fp32"):
        super().__init__()
        if use_selfcond is None:
            use_selfcond = in_ch >= 3
        self.use_selfcond = bool(use_selfcond)
        if cond_in_ch is None:
            cond_in_ch = max(in_ch - 1 - (1 if self.use_selfcond else 0), 0)
        self.cond_in_ch = int(cond_in_ch)
        self.in_ch = in_ch
        self.in_ch_total = in_ch
        self.compute_dtype = compute_dtype
        self.spec = ModelSpec(in_ch=in_ch, base_ch=base_ch, time_dim=time_dim, depth=depth, kernel=kernel,
                              max_time=float(t_embed_max_time), cond_in_ch=self.cond_in_ch, use_selfcond=self.use_selfcond)
        chs: List[int] = self.spec.chs
        k = kernel

        def block(cin, cout):
            return _pair(_Holder((cout, cin, k), cin * k, "conv"), _Holder((cout,), None, "norm"))

        self.time_mlp = _pair(TimeEmbedding(time_dim, max_time=t_embed_max_time), _Holder((base_ch, time_dim), time_dim, "lin"))
        self.encoders = nn.ModuleList()
        cin = in_ch
        for c in chs:
            self.encoders.append(block(cin, c))
            cin = c
        self.mid = block(cin, cin)
        self.decoders = nn.ModuleList()
        prev = chs[-1]
        for c in reversed(chs):
            self.decoders.append(block(prev + c, c))
            prev = c
        self.final = _Holder((1, prev + 1, k), (prev + 1) * k, "zero")
        self.tproj_enc = nn.ModuleList([_pair(_Empty(), _Holder((2 * c, base_ch), base_ch, "lin")) for c in chs])
        self.tproj_mid = _pair(_Empty(), _Holder((2 * chs[-1], base_ch), base_ch, "lin"))
        self.tproj_dec = nn.ModuleList([_pair(_Empty(), _Holder((2 * c, base_ch), base_ch, "lin")) for c in reversed(chs)])
        if self.cond_in_ch > 0:
            self.cond_enc = nn.ModuleList([_Holder((c, self.cond_in_ch, 1), self.cond_in_ch, "conv") for c in chs])
            self.cond_mid = _Holder((chs[-1], self.cond_in_ch, 1), self.cond_in_ch, "conv")
            self.cond_dec = nn.ModuleList([_Holder((c, self.cond_in_ch, 1), self.cond_in_ch, "conv") for c in reversed(chs)])
        else:
            self.cond_enc = nn.ModuleList([_Empty() for _ in chs])
            self.cond_mid = _Empty()
            self.cond_dec = nn.ModuleList([_Empty() for _ in chs])
        self._engines: Dict[tuple, UNetEngine] = {}
        self._bwd: Dict[int, object] = {}
        self._versions = None

    # ------------------------------------------------------------------
    def _param_versions(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, compute_dtype: Optional[str] = None, conv_impl: str = "auto") -> UNetEngine:
        """The kernel engine bound to this module's (CUDA) parameters; re-packs weights when they change."""
        cd = compute_dtype or self.compute_dtype
        key = (cd, conv_impl)
        ver = self._param_versions()
        if self._versions != ver:
            for e in self._engines.values():
                e.p = {k: v.detach() for k, v in self.named_parameters()}
                e.refresh()
            self._versions = ver
        eng = self._engines.get(key)
        if eng is None or eng.device != next(self.parameters()).device:
            params = {k: v.detach() for k, v in self.named_parameters()}
            eng = UNetEngine(params, self.spec, dtype=cd, conv_impl=conv_impl)
            self._engines[key] = eng
        return eng

    def _backward_engine(self, compute_dtype: Optional[str] = None, conv_impl: str = "auto"):
        """BackwardEngine (train.py) for the autograd path; gradients come back in ParamLayout order."""
        from .engine import ParamLayout
        from .train import BackwardEngine
        eng = self.engine(compute_dtype, conv_impl)
        bw = self._bwd.get(id(eng))
        if bw is None:
            layout = ParamLayout(self.spec, {k: tuple(p.shape) for k, p in self.named_parameters()})
            bw = BackwardEngine(eng, layout)
            self._bwd = {id(eng): bw}
        return bw

    def _apply(self, fn, *a, **kw):
        self._engines = {}
        self._bwd = {}
        self._versions = None
        return super()._apply(fn, *a, **kw)

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        if x.device.type != "cuda":
            raise RuntimeError("gwb200 UNet1D runs on CUDA (sm_100a) only: no CPU fallback")
        cd = self.compute_dtype
        if torch.is_autocast_enabled():      # reference inference/train wrap the call in autocast (inference.py:444)
            cd = "bf16"
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import unet_autograd_forward
            return unet_autograd_forward(self, x, t, cd)
        return self.engine(cd).forward(x, t)
