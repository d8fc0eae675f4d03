"""CPU oracle: a functional restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

The reference (`/root/reference/src/snr_denoising/{models,inference,train}.py`) does all of
its arithmetic through PyTorch ATen ops (torch 2.11, unpinned by the reference), so this
restatement uses the same ATen CPU ops through `torch.nn.functional`, but is written as pure
functions over a `state_dict` instead of `nn.Module`s so it can travel to the GPU box, where
`/root/reference` does not exist.

Parity pin: `tests/golden/make_golden.py` imports the unmodified reference in the build
container and stores its outputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py`
checks every function here against those vectors (fp32, bit-for-bit or <= 2e-6).  The
reference itself ships no tests or golden vectors.

Each function cites the reference lines it follows (paths relative to
`/root/reference/src/snr_denoising/`).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

__all__ = [
    "ModelCfg", "time_embedding", "cosine_beta_schedule", "alpha_bar_from_betas",
    "q_sample", "film_vectors", "unet_forward", "unet_forward_taps", "build_t_schedule",
    "cfg_weight", "snr_from_alpha_bar", "t_for_target_snr", "ddim_sample", "predict_x0_norm",
    "element_loss", "train_loss", "warmup_cosine_lambda", "adamw_step", "ema_step",
    "clip_grad_norm", "train_step", "conv_flops_per_sample", "level_lengths",
]


class ModelCfg:
    """Channel bookkeeping of UNet1D.__init__ (models.py:78-152)."""

    def __init__(self, in_ch: int = 1, base_ch: int = 64, time_dim: int = 128, depth: int = 3,
                 kernel: int = 3, t_embed_max_time: float = 999.0,
                 cond_in_ch: Optional[int] = None, use_selfcond: Optional[bool] = None):
        if use_selfcond is None:                       # models.py:91-94
            use_selfcond = in_ch >= 3
        if cond_in_ch is None:                         # models.py:96-98
            cond_in_ch = max(in_ch - 1 - (1 if use_selfcond else 0), 0)
        self.in_ch, self.base_ch, self.time_dim = int(in_ch), int(base_ch), int(time_dim)
        self.depth, self.kernel = int(depth), int(kernel)
        self.max_time = float(t_embed_max_time)
        self.cond_in_ch, self.use_selfcond = int(cond_in_ch), bool(use_selfcond)
        self.chs = [self.base_ch * (2 ** i) for i in range(self.depth)]   # models.py:113

    @classmethod
    def from_state_dict(cls, sd: StateDict, t_embed_max_time: float = 999.0,
                        use_selfcond: Optional[bool] = None) -> "ModelCfg":
        in_ch = sd["encoders.0.0.weight"].shape[1]
        base = sd["encoders.0.0.weight"].shape[0]
        depth = len([k for k in sd if k.startswith("encoders.") and k.endswith(".0.weight")])
        cond = sd["cond_mid.weight"].shape[1] if "cond_mid.weight" in sd else 0
        if use_selfcond is None:
            use_selfcond = (in_ch - 1 - cond) >= 1
        return cls(in_ch=in_ch, base_ch=base, time_dim=sd["time_mlp.1.weight"].shape[1],
                   depth=depth, kernel=sd["encoders.0.0.weight"].shape[2],
                   t_embed_max_time=t_embed_max_time, cond_in_ch=cond, use_selfcond=use_selfcond)


# --------------------------------------------------------------------------------------
# schedule / embedding
# --------------------------------------------------------------------------------------
def time_embedding(t: Tensor, dim: int, max_time: float = 999.0) -> Tensor:
    """models.py:19-31 (TimeEmbedding.forward)."""
    ts = t.float() / max(float(max_time), 1.0)
    half = dim // 2
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / max(half - 1, 1)))
    arg = ts[:, None] * freqs[None, :]
    emb = torch.cat([arg.sin(), arg.cos()], dim=1)
    if dim % 2 == 1:
        emb = torch.cat([emb, torch.zeros(t.size(0), 1)], dim=1)
    return emb


def cosine_beta_schedule(T: int, s: float = 0.008) -> Tensor:
    """models.py:34-40."""
    t = torch.linspace(0, T, T + 1, dtype=torch.float32)
    ac = torch.cos(((t / T) + s) / (1 + s) * (math.pi / 2)) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return betas.clamp(min=0.0, max=0.999)


def alpha_bar_from_betas(betas: Tensor) -> Tensor:
    """models.py:48-49 (fp32 cumprod)."""
    return torch.cumprod(1.0 - betas, dim=0)


def q_sample(alpha_bar: Tensor, x0: Tensor, t: Tensor, eps: Tensor) -> Tensor:
    """models.py:52-59 with the noise injected (the reference draws randn_like inside)."""
    t = t.long()
    a = alpha_bar.sqrt()[t].view(-1, 1, 1)
    m = (1 - alpha_bar).sqrt()[t].view(-1, 1, 1)
    return a * x0 + m * eps


def film_vectors(sd: StateDict, cfg: ModelCfg, t: Tensor) -> List[Tensor]:
    """time_mlp + every tproj_* (models.py:105-109, 137-142, 197, 206, 213, 224).

    Returns [enc0..enc{d-1}, mid, dec0..dec{d-1}] each [B, 2C] with (gamma | beta) halves.
    """
    emb = time_embedding(t, cfg.time_dim, cfg.max_time)
    ctx = F.silu(F.linear(emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    a = F.silu(ctx)
    names = [f"tproj_enc.{i}.1" for i in range(cfg.depth)] + ["tproj_mid.1"] + \
            [f"tproj_dec.{i}.1" for i in range(cfg.depth)]
    return [F.linear(a, sd[n + ".weight"], sd[n + ".bias"]) for n in names]


def level_lengths(L: int, depth: int = 3) -> List[int]:
    """avg_pool1d(2,2) floors (models.py:208)."""
    out = [L]
    for _ in range(depth):
        out.append(out[-1] // 2)
    return out


# --------------------------------------------------------------------------------------
# network forward
# --------------------------------------------------------------------------------------
def _block(sd: StateDict, prefix: str, h: Tensor) -> Tuple[Tensor, Tensor]:
    """Conv1d(k, pad k//2) -> GroupNorm(gcd(8,C)) -> SiLU (models.py:154-167). Returns (raw conv, activated)."""
    w = sd[prefix + ".0.weight"]
    raw = F.conv1d(h, w, sd[prefix + ".0.bias"], padding=w.shape[2] // 2)
    g = max(1, math.gcd(8, w.shape[0]))
    act = F.silu(F.group_norm(raw, g, sd[prefix + ".1.weight"], sd[prefix + ".1.bias"], eps=1e-5))
    return raw, act


def _cond_bias(sd: StateDict, name: str, cond: Optional[Tensor], L: int):
    """models.py:188-193."""
    if cond is None or (name + ".weight") not in sd:
        return 0.0
    c = F.interpolate(cond, size=L, mode="linear", align_corners=False)
    return F.conv1d(c, sd[name + ".weight"], sd[name + ".bias"])


def _film(h: Tensor, v: Tensor) -> Tensor:
    """models.py:169-173."""
    C = h.shape[1]
    return h * (1 + v[:, :C, None]) + v[:, C:, None]


def unet_forward_taps(sd: StateDict, cfg: ModelCfg, x: Tensor, t: Tensor) -> Dict[str, Tensor]:
    """models.py:195-231 with every intermediate kept (for per-layer parity)."""
    taps: Dict[str, Tensor] = {}
    films = film_vectors(sd, cfg, t)
    x_t = x[:, :1]
    cond = x[:, 1:1 + cfg.cond_in_ch] if cfg.cond_in_ch > 0 else None
    d = cfg.depth
    skips = []
    h = x
    for i in range(d):
        raw, h = _block(sd, f"encoders.{i}", h)
        taps[f"enc{i}.raw"] = raw
        h = h + _cond_bias(sd, f"cond_enc.{i}", cond, h.size(-1))
        h = _film(h, films[i])
        taps[f"enc{i}.out"] = h
        skips.append(h)
        h = F.avg_pool1d(h, 2, 2)
    raw, h = _block(sd, "mid", h)
    taps["mid.raw"] = raw
    h = h + _cond_bias(sd, "cond_mid", cond, h.size(-1))
    h = _film(h, films[d])
    taps["mid.out"] = h
    for i in range(d):
        skip = skips[d - 1 - i]
        h = F.interpolate(h, scale_factor=2, mode="nearest")           # nn.Upsample, models.py:127
        if h.size(-1) != skip.size(-1):                                 # models.py:218-220
            diff = skip.size(-1) - h.size(-1)
            h = F.pad(h, (0, diff)) if diff > 0 else h[..., :skip.size(-1)]
        h = torch.cat([h, skip], dim=1)
        raw, h = _block(sd, f"decoders.{i}", h)
        taps[f"dec{i}.raw"] = raw
        h = h + _cond_bias(sd, f"cond_dec.{i}", cond, h.size(-1))
        h = _film(h, films[d + 1 + i])
        taps[f"dec{i}.out"] = h
    if h.size(-1) != x.size(-1):                                        # models.py:227-229
        diff = x.size(-1) - h.size(-1)
        h = F.pad(h, (0, diff)) if diff > 0 else h[..., :x.size(-1)]
    wf = sd["final.weight"]
    taps["eps"] = F.conv1d(torch.cat([h, x_t], dim=1), wf, sd["final.bias"], padding=wf.shape[2] // 2)
    return taps


def unet_forward(sd: StateDict, cfg: ModelCfg, x: Tensor, t: Tensor) -> Tensor:
    return unet_forward_taps(sd, cfg, x, t)["eps"]


def conv_flops_per_sample(cfg: ModelCfg, L: int) -> float:
    """SURVEY 2.3 / BASELINE.md section 4: sum over every Conv1d of 2*Cin*Cout*K*L."""
    Ls = level_lengths(L, cfg.depth)
    fl = 0.0
    cin = cfg.in_ch
    for i, c in enumerate(cfg.chs):
        fl += 2.0 * cin * c * cfg.kernel * Ls[i] + 2.0 * cfg.cond_in_ch * c * Ls[i]
        cin = c
    fl += 2.0 * cin * cin * cfg.kernel * Ls[cfg.depth] + 2.0 * cfg.cond_in_ch * cin * Ls[cfg.depth]
    prev = cin
    for i, c in enumerate(reversed(cfg.chs)):
        Li = Ls[cfg.depth - 1 - i]
        fl += 2.0 * (prev + c) * c * cfg.kernel * Li + 2.0 * cfg.cond_in_ch * c * Li
        prev = c
    fl += 2.0 * (prev + 1) * 1 * cfg.kernel * L
    return fl


# --------------------------------------------------------------------------------------
# sampler
# --------------------------------------------------------------------------------------
def snr_from_alpha_bar(alpha_bar: Tensor):
    """inference.py:209-211."""
    ab = alpha_bar.detach().cpu().numpy().clip(1e-12, 1 - 1e-12)
    import numpy as np
    return np.sqrt(ab / (1.0 - ab))


def t_for_target_snr(alpha_bar: Tensor, target_snr: float) -> int:
    """inference.py:213-215."""
    import numpy as np
    snr = snr_from_alpha_bar(alpha_bar)
    return int(np.argmin(np.abs(snr - float(target_snr))))


def build_t_schedule(T: int, steps: int, start_t: Optional[int]) -> Tensor:
    """inference.py:217-228."""
    if start_t is None:
        start_t = T - 1
    start_t = int(max(0, min(start_t, T - 1)))
    steps = int(max(1, min(steps, start_t + 1)))
    ts = torch.linspace(start_t, 0, steps).round().long()
    ts = torch.unique_consecutive(ts)
    if ts[0].item() != start_t:
        ts = torch.cat([torch.tensor([start_t]), ts])
    if ts[-1].item() != 0:
        ts = torch.cat([ts, torch.tensor([0])])
    return ts


def cfg_weight(i: int, N: int, mode: str, wmax: float, center: float, width: float) -> float:
    """inference.py:230-244."""
    s = 1.0 if N <= 1 else i / (N - 1)
    mode = mode.lower()
    if mode == "const":
        return float(wmax)
    if mode == "tophat":
        lo, hi = center - width * 0.5, center + width * 0.5
        return float(wmax) if (lo <= s <= hi) else 1.0
    if mode == "gauss":
        sig = max(width, 1e-9)
        return float(wmax) * math.exp(-0.5 * ((s - center) / sig) ** 2)
    raise ValueError(f"unknown cfg-mode: {mode}")


@torch.no_grad()
def ddim_sample(sd: StateDict, cfg: ModelCfg, alpha_bar: Tensor, cond_stack: Tensor, *,
                T: int, steps: int, eta: float, start_t: Optional[int] = None,
                init_mode: str = "noise", x0_std_est: float = 0.14, dc_weight: float = 0.0,
                cond_scale: float = 1.0, eps_scale: float = 1.0, pred_type: str = "eps",
                cfg_scale: float = 1.0, cfg_mode: str = "const", cfg_center: float = 0.5,
                cfg_width: float = 0.3, cfg_u_only_thresh: float = 0.0, drop_y_only: bool = True,
                noise: Optional[Sequence[Tensor]] = None, oracle_init_clean: Optional[Tensor] = None,
                trace: Optional[List[Dict[str, Tensor]]] = None,
                forward_fn: Optional[Callable[[Tensor, Tensor], Tensor]] = None) -> Tensor:
    """inference.py:374-514 for a batch [B, cond_in_ch, L] with injected noise.

    `noise[k]` is the k-th draw the reference would take from the global RNG: draw 0 initialises
    x_T (inference.py:409-415, or q_sample for oracle_init :403-405), draw k>=1 is the
    `randn_like` of the k-th step whose sigma_t > 0 (inference.py:483).  Each is [B,1,L].
    Every op is per sample, so a batch equals the reference looped over B=1 calls.
    """
    B, _, L = cond_stack.shape
    y = cond_stack[:, :1]
    meta = cond_stack[:, 1:] if cond_stack.size(1) > 1 else None
    sched = build_t_schedule(T, steps, start_t)
    ab = alpha_bar.clamp(1e-12, 1.0)
    ab_start = ab[int(sched[0])]
    draws = iter(noise) if noise is not None else None

    def draw():
        return next(draws) if draws is not None else torch.randn(B, 1, L)

    if oracle_init_clean is not None:
        t0 = torch.full((B,), int(sched[0]), dtype=torch.long)
        x_t = q_sample(alpha_bar, oracle_init_clean, t0, draw())
    elif init_mode == "noise":
        x_t = draw()
    elif init_mode == "scaled-noise":
        x_t = torch.sqrt(ab_start * (x0_std_est ** 2) + (1 - ab_start)) * draw()
    elif init_mode == "y-blend":
        x_t = torch.sqrt(ab_start) * y + torch.sqrt(1 - ab_start) * draw()
    else:
        raise ValueError(f"unknown init_mode: {init_mode}")
    x0_sc = torch.zeros_like(x_t) if cfg.use_selfcond else None
    fwd = forward_fn if forward_fn is not None else (lambda xi, ti: unet_forward(sd, cfg, xi, ti))

    def pack(xt, c, sc):
        return torch.cat([xt, c, sc], dim=1) if cfg.use_selfcond else torch.cat([xt, c], dim=1)

    N = len(sched)
    for i in range(N):
        t_now = int(sched[i])
        ab_t = ab[t_now]
        ab_prev = ab[int(sched[i + 1])] if i + 1 < N else torch.tensor(1.0)
        y_used = cond_scale * y
        c_on = torch.cat([y_used, meta], dim=1) if meta is not None else y_used
        c_off = torch.cat([torch.zeros_like(y_used), meta], dim=1) if (drop_y_only and meta is not None) \
            else torch.zeros_like(c_on)
        w = cfg_weight(i, N, cfg_mode, cfg_scale, cfg_center, cfg_width)
        tt = torch.full((B,), t_now, dtype=torch.long)
        if w <= cfg_u_only_thresh:
            out = fwd(pack(x_t, c_off, x0_sc), tt)
        elif abs(w - 1.0) <= 1e-6:
            out = fwd(pack(x_t, c_on, x0_sc), tt)
        else:
            oc = fwd(pack(x_t, c_on, x0_sc), tt)
            ou = fwd(pack(x_t, c_off, x0_sc), tt)
            out = ou + w * (oc - ou)
        out = out[:, :1]
        if pred_type == "eps":
            eps = eps_scale * out
            x0 = (x_t - torch.sqrt(1 - ab_t) * eps) / torch.sqrt(ab_t)
        else:
            x0 = out
            eps = (x_t - torch.sqrt(ab_t) * x0) / torch.sqrt(torch.clamp(1 - ab_t, min=1e-12))
        if dc_weight > 0:
            x0 = (1 - dc_weight) * x0 + dc_weight * y
        if cfg.use_selfcond:
            x0_sc = x0
        x_in = x_t
        if t_now == 0:
            x_t = x0
        else:
            sig = eta * torch.sqrt((1 - ab_prev) / (1 - ab_t) * (1 - ab_t / ab_prev))
            dirx = torch.sqrt(torch.clamp(1 - ab_prev - sig ** 2, min=0.0)) * eps
            nz = sig * draw() if sig.item() > 0 else 0.0
            x_t = torch.sqrt(ab_prev) * x0 + dirx + nz
        if trace is not None:
            trace.append({"t": torch.tensor(t_now), "x_in": x_in, "eps": eps, "x0": x0, "x_out": x_t})
    return x_t


# --------------------------------------------------------------------------------------
# training step
# --------------------------------------------------------------------------------------
@torch.no_grad()
def predict_x0_norm(sd: StateDict, cfg: ModelCfg, alpha_bar: Tensor, x_t: Tensor, cond: Tensor, t: Tensor) -> Tensor:
    """train.py:40-51."""
    t = t.long()
    eps_hat = unet_forward(sd, cfg, torch.cat([x_t, cond, torch.zeros_like(x_t)], dim=1), t)
    ab = alpha_bar[t].view(-1, 1, 1)
    return (x_t - torch.sqrt(1 - ab) * eps_hat) / torch.sqrt(ab)


def element_loss(eps_hat: Tensor, eps: Tensor, mask: Tensor, loss_type: str, huber_beta: float) -> Tensor:
    """train.py:53-58."""
    if loss_type == "huber":
        el = F.smooth_l1_loss(eps_hat, eps, reduction="none", beta=huber_beta)
    else:
        el = (eps_hat - eps) ** 2
    return el * mask


def train_loss(eps_hat: Tensor, eps: Tensor, mask: Tensor, alpha_bar: Tensor, t: Tensor,
               loss_type: str = "huber", huber_beta: float = 0.5, loss_weight_power: float = 0.0) -> Tensor:
    """train.py:411-421."""
    el = element_loss(eps_hat, eps, mask, loss_type, huber_beta)
    if loss_weight_power != 0.0:
        el = el * (1.0 - alpha_bar[t].view(-1, 1, 1)).pow(loss_weight_power)
    denom = mask.sum(dim=[1, 2]).clamp_min(1.0)
    return (el.sum(dim=[1, 2]) / denom).mean()


def warmup_cosine_lambda(step: int, warmup_steps: int, total_steps: int, min_lr_scale: float = 0.1) -> float:
    """train.py:84-91."""
    if step < warmup_steps:
        return max(1e-8, float(step + 1) / max(1, warmup_steps))
    p = (step - warmup_steps) / max(1, (total_steps - warmup_steps))
    p = min(max(p, 0.0), 1.0)
    return min_lr_scale + 0.5 * (1 - min_lr_scale) * (1 + math.cos(math.pi * p))


def clip_grad_norm(grads: Dict[str, Tensor], max_norm: float) -> Tuple[Dict[str, Tensor], float]:
    """torch.nn.utils.clip_grad_norm_ semantics (train.py:445): coef = min(1, max/(norm+1e-6))."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}, float(total)


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
               beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, wd: float = 1e-4):
    """torch.optim.AdamW single-tensor update (train.py:264, 447). `step` is 1-based."""
    p = p * (1 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def ema_step(ema: Tensor, p: Tensor, decay: float) -> Tensor:
    """train.py:73-81."""
    return ema * decay + p * (1.0 - decay)


def train_step(sd: StateDict, cfg: ModelCfg, alpha_bar: Tensor, *, clean_norm: Tensor, cond_stack: Tensor,
               mask: Tensor, t: Tensor, eps: Tensor, drop: Optional[Tensor] = None, selfcond: bool = False,
               clamp_inputs: float = 10.0, loss_type: str = "huber", huber_beta: float = 0.5,
               loss_weight_power: float = 0.0, dropout_y_only: bool = True):
    """Forward+backward of one batch: train.py:349-421, 438-439 with RNG draws injected.

    `clean_norm`, `cond_stack` are already sigma-normalised (train.py:336-347); `drop` is the
    [B,1,1] 0/1 CFG-dropout mask (train.py:386); `selfcond` is the per-batch coin (train.py:401).
    Returns (loss, grads dict, eps_hat).
    """
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    y = cond_stack[:, :1]
    meta = cond_stack[:, 1:] if cond_stack.size(1) > 1 else None
    if clamp_inputs > 0:                                       # train.py:350-352 (cond_stack itself is not re-built)
        clean_norm = clean_norm.clamp(-clamp_inputs, clamp_inputs)
        y_cl = y.clamp(-clamp_inputs, clamp_inputs)
    else:
        y_cl = y
    x_t = q_sample(alpha_bar, clean_norm, t, eps)
    if clamp_inputs > 0:
        x_t = x_t.clamp(-clamp_inputs, clamp_inputs)
    if drop is not None:
        if meta is not None and dropout_y_only:               # train.py:387-394 uses the clamped y_norm
            cond_used = torch.cat([y_cl * (1.0 - drop), meta], dim=1)
        else:
            cond_used = cond_stack * (1.0 - drop)
    else:
        cond_used = cond_stack
    if selfcond:
        x0_sc = predict_x0_norm(sd, cfg, alpha_bar, x_t, cond_used, t)
    else:
        x0_sc = torch.zeros_like(x_t)
    net_in = torch.cat([x_t, cond_used, x0_sc], dim=1)
    eps_hat = unet_forward(params, cfg, net_in, t)
    loss = train_loss(eps_hat, eps, mask, alpha_bar, t, loss_type, huber_beta, loss_weight_power)
    loss.backward()
    grads = {k: p.grad.detach() for k, p in params.items()}
    return loss.detach(), grads, eps_hat.detach()
