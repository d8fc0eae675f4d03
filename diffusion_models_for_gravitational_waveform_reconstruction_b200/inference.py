"""Drop-in mirror of the reference sampler entry points (src/snr_denoising/inference.py:209-244, 317-514).

`ddim_sample` keeps the reference signature and argument meaning; additive behaviour only:
  * `cond_stack` may be [B, cond_in_ch, L] (the reference is hard-wired to B=1);
  * `noise=` injects the draws the reference would take from the global RNG ([n_draws, B, 1, L], draw 0 = x_T,
    draw k = k-th stochastic step), `seed=`/`sample0=` select the on-device Philox stream instead;
  * `use_graph=` runs the whole chain as one CUDA graph, `compute_dtype=` picks fp32-exact or bf16/tcgen05 kernels.
The data / checkpoint loaders either side of the chain are mirrored too: `load_checkpoint` (inference.py:614-650,
sweep_infer.py:170-190), `_load_measurement_from_h5` / `_load_measurement_from_npy` (inference.py:59-93), `_meta_to_stack`
(inference.py:96-122).  Plotting and the argparse `main` (inference.py:517-903) are outside the hot path and not re-implemented.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .engine import SamplerPlan, build_t_schedule, cfg_weight
from .models import CustomDiffusion, UNet1D
# CPU pre/post-processing of the reference CLI, on the GPU here; same names as inference.py:36-38, 125-205
from .whitening import (_dewhiten_model, _dewhiten_train_like, _dewhiten_welch, _interp_psd_for_length, _mad_std,  # noqa: F401
                        _pick_sigma, _whiten_pair_model, _whiten_pair_train_like, _whiten_pair_welch)

__all__ = ["philox_normal", "snr_from_alpha_bar", "t_for_target_snr", "_build_t_schedule", "_cfg_weight", "_reduce_to_one_channel",
           "one_step_proxy_like_test_infer", "ddim_sample", "make_sampler_plan", "load_checkpoint", "_load_measurement_from_h5",
           "_load_measurement_from_npy", "_meta_to_stack"]


# ------------------------------------------------------------------------------------------------ loaders around the chain
def load_checkpoint(path: str, device="cuda", use_ema: bool = True, compute_dtype: Optional[str] = None):
    """Model + diffusion from a reference-format checkpoint (payload of train.py:606-630), the way every reference entry point
    does it (inference.py:614-650, sweep_infer.py:170-190, grid_infer.py:288-307): architecture from `ckpt['args']` with the
    reference's fallbacks, self-conditioning inferred from the channel count, EMA weights when present and wanted,
    `load_state_dict(strict=True)`.  Returns (model.eval() on `device`, CustomDiffusion, ck_args)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    ck_args = ckpt.get("args", {}) or {}
    in_ch = int(ck_args.get("in_ch", 3))
    cond_in_ch = int(ck_args.get("cond_in_ch", 1))
    T = int(ck_args.get("T", 1000))
    kw = {} if compute_dtype is None else {"compute_dtype": compute_dtype}
    model = UNet1D(in_ch=in_ch, base_ch=int(ck_args.get("base_ch", 64)), time_dim=int(ck_args.get("time_dim", 128)),
                   depth=int(ck_args.get("depth", 3)), t_embed_max_time=max(0, T - 1), cond_in_ch=cond_in_ch,
                   use_selfcond=(in_ch == (1 + cond_in_ch + 1)), **kw)
    loaded = False
    if use_ema and "model_ema_state" in ckpt:
        try:
            model.load_state_dict(ckpt["model_ema_state"], strict=True)
            loaded = True
        except Exception as e:                                    # inference.py:645-646: fall back to the raw weights
            print(f"[warn] EMA load failed ({e}); falling back to raw weights]")
    if not loaded:
        model.load_state_dict(ckpt["model_state"], strict=True)
    model = model.to(device).eval()
    return model, CustomDiffusion(T=T, device=device), ck_args


def _load_measurement_from_h5(h5_path: str, index: int):
    """inference.py:59-89: (y, clean, fs, P_model, (fw, Pw), meta_dict) of one HDF5 row (`gen.py:406-413` layout), read with
    h5py when installed and with the bundled reader otherwise."""
    from .dataloader import open_h5
    meta: Dict[str, float] = {}
    f = open_h5(h5_path)
    try:
        y = np.array(f["noisy"][index], dtype=np.float32)
        clean = np.array(f["signal"][index], dtype=np.float32) if "signal" in f else None
        fs = float(f.attrs.get("sampling_rate", 0.0)) or float(1.0 / f.attrs.get("delta_t", 1.0 / 4096.0))
        P_model = None
        if "psd_model" in f:
            P_model = np.array(f["psd_model"][index], dtype=np.float64)
        elif "psd" in f:                                          # legacy name
            P_model = np.array(f["psd"][index], dtype=np.float64)
        fw = Pw = None
        if ("psd_welch" in f) and ("psd_welch_freqs" in f):
            Pw = np.array(f["psd_welch"][index], dtype=np.float64)
            fw = np.array(f["psd_welch_freqs"][index], dtype=np.float64)
        for k in ["mass1", "mass2", "spin1z", "spin2z", "q", "chirp_mass", "snr", "epoch", "label_m1", "label_m2", "label_s1",
                  "label_s2"]:
            if k in f:
                try:
                    meta[k] = float(np.array(f[k][index]).reshape(()))
                except Exception:
                    pass
    finally:
        if hasattr(f, "close"):
            f.close()
    return y, clean, fs, P_model, (fw, Pw), meta


def _load_measurement_from_npy(npy_path: str, fs: float):
    """inference.py:91-93."""
    y = np.load(npy_path).astype(np.float32).ravel()
    return y, None, fs, None, (None, None), {}


def _meta_to_stack(meta: dict, L: int, cond_in_ch: int, M_SCALE: float, Q_SCALE: float) -> Optional[np.ndarray]:
    """inference.py:96-122: [cond_in_ch - 1, L] constant channels in the order m1, m2, s1, s2, q, chirp_mass (masses / M_SCALE,
    q clipped to [0, Q_SCALE] / Q_SCALE, spins as they are), zero rows beyond the sixth."""
    need = max(0, cond_in_ch - 1)
    if need <= 0:
        return None
    qv = meta.get("q", 0.0)
    if not np.isfinite(qv):
        qv = 0.0
    vals = [meta.get("mass1", 0.0) / max(M_SCALE, 1e-9), meta.get("mass2", 0.0) / max(M_SCALE, 1e-9), meta.get("spin1z", 0.0),
            meta.get("spin2z", 0.0), min(max(qv, 0.0), Q_SCALE) / max(Q_SCALE, 1e-9), meta.get("chirp_mass", 0.0) / max(M_SCALE, 1e-9)]
    arr = np.zeros((need, L), dtype=np.float32)
    for i, v in enumerate(vals[:need]):
        arr[i, :] = np.float32(float(v))
    return arr


def philox_normal(B: int, L: int, seed: int, sample0: int = 0, step: int = 0, device="cuda") -> torch.Tensor:
    """[B, 1, L] standard normals of the on-device Philox stream (seed, sample0 + b, step): what `ddim_sample` starts from
    when no `noise=` is injected (step 0), independent of batch chunking and world size."""
    from . import _cabi
    out = torch.empty(B, 1, L, device=device, dtype=torch.float32)
    _cabi.check(_cabi.load().gw_philox_normal(int(seed) & (2 ** 64 - 1), int(sample0), int(step), B, L, _cabi.ptr(out),
                                              _cabi.stream_ptr()), "philox_normal")
    return out


def snr_from_alpha_bar(alpha_bar: torch.Tensor) -> np.ndarray:
    ab = alpha_bar.detach().cpu().numpy().clip(1e-12, 1 - 1e-12)
    return np.sqrt(ab / (1.0 - ab))


def t_for_target_snr(diffusion: CustomDiffusion, target_snr: float) -> int:
    return int(np.argmin(np.abs(snr_from_alpha_bar(diffusion.alpha_bar) - float(target_snr))))


def _build_t_schedule(T: int, steps: int, device, start_t: Optional[int]) -> torch.Tensor:
    return torch.tensor(build_t_schedule(T, steps, start_t), dtype=torch.long, device=device)


def _cfg_weight(i: int, N: int, mode: str, wmax: float, center: float, width: float) -> float:
    return cfg_weight(i, N, mode, wmax, center, width)


def _reduce_to_one_channel(x: torch.Tensor) -> torch.Tensor:
    return x[:, :1, :] if (x.ndim == 3 and x.size(1) > 1) else x


def _stats(name: str, arr) -> str:
    """inference.py `_stats`: the debug line format of the reference sampler."""
    a = arr.detach().double() if isinstance(arr, torch.Tensor) else torch.as_tensor(np.asarray(arr), dtype=torch.float64)
    return (f"{name}: shape={tuple(a.shape)}, min={a.min().item():.3e}, max={a.max().item():.3e}, "
            f"mean={a.mean().item():.3e}, std={a.std(unbiased=False).item():.3e})")


def _corr_lag(x_t: torch.Tensor, y: torch.Tensor, delta_t: float):
    """The `corr_lag` diagnostic of the per-step JSONL log (inference.py:491-506): Pearson correlation of x_t and y after the
    cross-correlation lag search over +-0.25 s, one value per sample; the lag search runs on the device (gw_score_batch)."""
    from . import scoring
    B, L = x_t.shape
    win = min(L - 1, int(max(1.0, 0.25 / delta_t)))
    lags = scoring.best_lag_by_xcorr(x_t, y, max_shift=win).tolist()
    out = []
    for b, k in enumerate(lags):
        a, c = x_t[b].double(), y[b].double()
        if k < 0:
            a, c = a[-k:], c[: L + k]
        elif k > 0:
            a, c = a[: L - k], c[k:]
        a, c = a - a.mean(), c - c.mean()
        out.append(float((a * c).sum() / (torch.sqrt((a * a).sum() * (c * c).sum()) + 1e-30)))
    return out


_PLANS: Dict[Tuple, SamplerPlan] = {}


def make_sampler_plan(model: UNet1D, diffusion: CustomDiffusion, B: int, L: int, *, T: int, steps: int, eta: float,
                      start_t: Optional[int] = None, dc_weight: float = 0.0, eps_scale: float = 1.0,
                      pred_type: str = "eps", cfg_scale: float = 1.0, cfg_mode: str = "const", cfg_center: float = 0.5,
                      cfg_width: float = 0.3, cfg_u_only_thresh: float = 0.0, seed: int = 0, sample0: int = 0,
                      compute_dtype: Optional[str] = None, conv_impl: str = "auto", cache: bool = True) -> SamplerPlan:
    eng = model.engine(compute_dtype, conv_impl)
    key = (id(eng), B, L, T, steps, float(eta), start_t, float(dc_weight), float(eps_scale), pred_type, float(cfg_scale),
           cfg_mode, float(cfg_center), float(cfg_width), float(cfg_u_only_thresh))
    plan = _PLANS.get(key) if cache else None
    if plan is None:
        plan = SamplerPlan(eng, diffusion.alpha_bar, B, L, T=T, steps=steps, eta=eta, start_t=start_t, dc_weight=dc_weight,
                           eps_scale=eps_scale, pred_type=pred_type, cfg_scale=cfg_scale, cfg_mode=cfg_mode,
                           cfg_center=cfg_center, cfg_width=cfg_width, cfg_u_only_thresh=cfg_u_only_thresh, seed=seed,
                           sample0=sample0)
        if cache:
            if len(_PLANS) > 16:
                _PLANS.clear()
            _PLANS[key] = plan
    plan.set_rng(seed, sample0)
    return plan


@torch.no_grad()
def ddim_sample(model, diffusion, cond_stack: torch.Tensor,
                T: int, steps: int, eta: float,
                device, length: int, debug: bool,
                start_t: Optional[int], init_mode: str, x0_std_est: float,
                dc_weight: float, cond_scale: float, eps_scale: float, pred_type: str,
                in_ch: int, cond_in_ch: int, use_selfcond: bool, cfg_scale: float,
                cfg_mode: str, cfg_center: float, cfg_width: float, cfg_u_only_thresh: float,
                oracle_init: bool = False, clean_norm_311: Optional[torch.Tensor] = None,
                log_jsonl_path: Optional[str] = None, log_interval: int = 0,
                xcorr_window_samp: int = 0, delta_t: float = 1.0,
                amp: bool = False, drop_y_only: bool = True, *,
                noise: Optional[torch.Tensor] = None, seed: Optional[int] = None, sample0: int = 0,
                use_graph: bool = True, compute_dtype: Optional[str] = None, conv_impl: str = "auto",
                return_trace: bool = False) -> torch.Tensor:
    """eta-parameterised DDIM sampler (eta=0 deterministic, eta=1 with all steps = DDPM ancestral); inference.py:374-514."""
    if init_mode not in ("noise", "scaled-noise", "y-blend"):
        raise ValueError(f"unknown init_mode: {init_mode}")
    cfg_weight(0, 2, cfg_mode, cfg_scale, cfg_center, cfg_width)           # raises ValueError on an unknown cfg-mode
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("gwb200 ddim_sample runs on CUDA (sm_100a) only: no CPU fallback")
    cond_stack = cond_stack.to(dev).float()
    if cond_stack.ndim == 2:
        cond_stack = cond_stack.unsqueeze(0)
    B, _, L = cond_stack.shape
    if length != L:
        raise ValueError(f"length={length} does not match cond_stack length {L}")
    if model.in_ch != in_ch or model.cond_in_ch != cond_in_ch or bool(model.use_selfcond) != bool(use_selfcond):
        raise ValueError("in_ch / cond_in_ch / use_selfcond do not match the model")
    if compute_dtype is None and amp:
        compute_dtype = "bf16"            # reference AMP is fp16 autocast (inference.py:444); bf16 is the B200 analogue
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    plan = make_sampler_plan(model, diffusion, B, L, T=T, steps=steps, eta=eta, start_t=start_t, dc_weight=dc_weight,
                             eps_scale=eps_scale, pred_type=pred_type, cfg_scale=cfg_scale, cfg_mode=cfg_mode,
                             cfg_center=cfg_center, cfg_width=cfg_width, cfg_u_only_thresh=cfg_u_only_thresh, seed=seed,
                             sample0=sample0, compute_dtype=compute_dtype, conv_impl=conv_impl)
    y_chan = cond_stack[:, :1, :]
    meta = cond_stack[:, 1:, :] if cond_stack.size(1) > 1 else None
    ab_start = torch.tensor(plan.ab_start, device=dev)

    def draw0():
        if noise is not None:
            return noise[0].to(dev).float().reshape(B, 1, L)
        return philox_normal(B, L, seed, sample0, 0, dev)

    if oracle_init and (clean_norm_311 is not None):                        # inference.py:403-406
        t0 = torch.full((B,), plan.sched[0], dtype=torch.long, device=dev)
        x_t, _ = diffusion.q_sample(clean_norm_311.to(dev).float().reshape(B, 1, L), t0, noise=draw0())
        if debug:
            print(f"[debug] oracle-init enabled (t0={plan.sched[0]})")
    elif init_mode == "noise":
        x_t = draw0()
    elif init_mode == "scaled-noise":
        x_t = torch.sqrt(ab_start * (x0_std_est ** 2) + (1 - ab_start)) * draw0()
    else:
        x_t = torch.sqrt(ab_start) * y_chan + torch.sqrt(1 - ab_start) * draw0()

    y_used = cond_scale * y_chan                                           # inference.py:434-435
    cond_on = torch.cat([y_used, meta], dim=1) if meta is not None else y_used
    if drop_y_only and meta is not None:                                    # inference.py:446
        cond_off = torch.cat([torch.zeros_like(y_used), meta], dim=1)
    else:
        cond_off = torch.zeros_like(cond_on)

    want_noise = None
    if noise is not None:
        if noise.shape[0] < plan.n_draws:
            raise ValueError(f"noise has {noise.shape[0]} draws, the chain needs {plan.n_draws}")
        want_noise = noise[: plan.n_draws].to(dev).float().reshape(plan.n_draws, B, L).contiguous()
    trace = return_trace or bool(log_jsonl_path) or debug
    rebuild = False
    if (want_noise is None) != (plan.noise is None) or (want_noise is not None and plan.noise.shape != want_noise.shape):
        plan.noise = None if want_noise is None else torch.empty_like(want_noise)
        rebuild = True
    if want_noise is not None:
        plan.noise.copy_(want_noise)
    if trace and plan.trace_eps is None:
        plan.trace_eps = torch.empty(B, L, device=dev)
        plan.trace_x0 = torch.empty(B, L, device=dev)
        rebuild = True
    if rebuild:
        plan._graph = None
    plan.load_inputs(x_t, cond_on, cond_off, y_chan if dc_weight > 0 else None)

    if debug:
        print(f"[debug] init_mode={init_mode}, x0_std_est={x0_std_est:.3f}, ab_start={plan.ab_start:.6f}")
        print(f"[debug] schedule length={plan.N}, first={plan.sched[0]}, last={plan.sched[-1]}")

    if trace:
        # per-step host visibility (debug prints / JSONL / parity traces): run step by step
        if log_jsonl_path:
            os.makedirs(os.path.dirname(log_jsonl_path) or ".", exist_ok=True)
        steps_out = []
        for i in range(plan.N):
            x_in = plan.net[i & 1][:B, 0:1].clone() if return_trace else None
            plan.enqueue_step()
            x_now = plan.net[(i + 1) & 1][:B, 0:1]
            if return_trace:
                steps_out.append({"t": plan.sched[i], "x_in": x_in, "eps": plan.trace_eps.clone().view(B, 1, L),
                                  "x0": plan.trace_x0.clone().view(B, 1, L), "x_out": x_now.clone()})
            if debug and (i % max(1, plan.N // 5) == 0 or plan.sched[i] == 0):
                print(_stats(f"x_t (t={plan.sched[i]})", x_now))
                print(_stats(f"eps_hat (t={plan.sched[i]})", plan.trace_eps.view(B, 1, L)))
            if log_jsonl_path and ((i % max(1, log_interval) == 0) or (i == plan.N - 1)):
                corr_lag = _corr_lag(x_now.reshape(B, L), y_chan.reshape(B, L), delta_t)
                with open(log_jsonl_path, "a") as fh:
                    fh.write(json.dumps({"phase": "ddim_step", "i": i, "t": plan.sched[i],
                                         "i_norm": float(0.0 if plan.N <= 1 else i / (plan.N - 1)),
                                         "alpha_bar": float(diffusion.alpha_bar[plan.sched[i]]),
                                         "cfg_mode": cfg_mode, "cfg_w_t": float(plan.coef[i, 5]),
                                         "cfg_scale": float(cfg_scale),
                                         "corr_lag": corr_lag[0] if B == 1 else corr_lag}) + "\n")
        out = plan.net[plan.N & 1][:B, 0:1].clone()
        return (out, steps_out) if return_trace else out
    return plan.run(use_graph=use_graph).clone()


@torch.no_grad()
def one_step_proxy_like_test_infer(model, diffusion, clean_norm: torch.Tensor, cond_stack: torch.Tensor,
                                   sigma_scalar: float, target_snr: float, device, in_ch: int, cond_in_ch: int,
                                   use_selfcond: bool, cfg_scale: float, drop_y_only: bool, cond_scale: float = 1.0,
                                   eps_scale: float = 1.0, pred_type: str = "eps", amp: bool = False,
                                   noise: Optional[torch.Tensor] = None):
    """Single-forward x0 estimate at the t matching `target_snr` (inference.py:317-371)."""
    dev = torch.device(device)
    ab_np = diffusion.alpha_bar.detach().cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        t_pick = int(np.argmin(np.abs(np.sqrt(ab_np / (1 - ab_np)) - target_snr)))      # inference.py:329-331 (unclipped)
    B = clean_norm.shape[0]
    t = torch.full((B,), t_pick, dtype=torch.long, device=dev)
    x_t, _ = diffusion.q_sample(clean_norm.to(dev), t, noise=noise)
    sc = torch.zeros_like(x_t)
    y = cond_scale * cond_stack[:, :1, :].to(dev)
    meta = cond_stack[:, 1:, :].to(dev) if cond_stack.size(1) > 1 else None
    c_on = torch.cat([y, meta], dim=1) if meta is not None else y

    def net_in(c):
        return torch.cat([x_t, c, sc], dim=1) if use_selfcond else torch.cat([x_t, c], dim=1)

    eng = model.engine("bf16" if amp else None)
    out = eng.forward(net_in(c_on), t)
    if cfg_scale != 1.0:
        c_off = torch.cat([torch.zeros_like(y), meta], dim=1) if (drop_y_only and meta is not None) else torch.zeros_like(c_on)
        out_u = eng.forward(net_in(c_off), t)
        out = out_u + cfg_scale * (out - out_u)
    out = _reduce_to_one_channel(out)
    ab_t = diffusion.alpha_bar[t_pick].to(dev)
    if pred_type == "eps":
        x0 = (x_t - torch.sqrt(1 - ab_t) * (eps_scale * out)) / torch.sqrt(ab_t)
    else:
        x0 = out
    return x0 * torch.tensor(sigma_scalar, device=dev).view(1, 1, 1)
