#!/usr/bin/env python
"""Headline benchmark (driver contract): `python bench.py --gpus N --steps K --warmup W [--impl reference]`.

Default workload `train` = BASELINE.json configs[1]: one optimisation step of the reference training loop body
(train.py:320-456: q_sample, CFG dropout p=0.2, self-conditioning coin p=0.5, forward, masked Huber loss, backward,
clip 1.0, AdamW, EMA) on a batch of 256 x 4096-sample synthetic noisy chirps PER GPU, in_ch=7 (y + 4 metadata channels +
self-conditioning, what train.py builds), random-init U-Net, bf16 activations / tcgen05 GEMMs, fp32 master weights.
One bench "step" = one optimisation step.  Metric: training samples/sec (whole job, all ranks; weak scaling: the per-GPU
batch is fixed and the gradient bucket is all-reduced over NCCL).
  value : steps through `FusedTrainStep.step` with the batch resident in HBM (CUDA-graph replay);
  e2e   : the same through the public API with pinned HOST batches: H2D of (clean, cond, mask) every step (double-buffered,
          as the reference's pinned DataLoader does) and a D2H read of the loss every step.
The same JSON line carries a `sampling` object: the other half of BASELINE's metric, waveforms/sec of the full T=1000
DDPM reverse chain (`--workload ddpm1000` / `ddim50` make that the headline instead).
`roofline` = the tcgen05 conv GEMM family timed live with CUDA events (tensor bound); `kernels` lists every kernel family
with its share of the step and its own roofline.  `cpu_baseline` / `--impl reference` = the CPU oracle port of the
reference (torch CPU ops on all host cores; /root/reference is Python and cannot travel) on a bounded sample.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

WORKLOADS = {
    # name: (steps, eta, start_t, description)
    "ddpm1000": (1000, 1.0, None, "DDPM T=1000 full reverse chain (ddim_sample steps=1000 eta=1 cfg=1)"),
    "ddim50": (50, 0.0, None, "DDIM 50-step reconstruction (eta=0 cfg=1)"),
}
TRAIN_DESC = ("training step (train.py:320-456 defaults: huber 0.5, clip 1.0, AdamW 2e-4/1e-4, EMA 0.999, p_uncond 0.2, "
              "p_selfcond 0.5, t in [500, 999])")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s in sm if s >= 0.5 * max(sm)] if sm else []       # "under load" = upper half of what we saw
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class TimedLib:
    """Brackets every C-ABI call with CUDA events on the current stream (eager launches)."""

    def __init__(self, lib):
        self._lib, self.records, self.on = lib, [], False

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("gw_") or name in ("gw_last_error", "gw_conv_tc_packed_elems", "gw_conv_tc_n_part", "gw_conv_gn_group", "gw_conv_in_gn_group", "gw_conv_gn_sync_bytes", "gw_conv_in_direct_ws_floats", "gw_version",
                                                  "gw_gn_bwd_scratch_elems", "gw_wgrad_tc_scratch_elems", "gw_opt_scratch_doubles"):
            return fn

        def wrapped(*a):
            if not self.on:
                return fn(*a)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*a)
            e.record()
            self.records.append((name, s, e))
            return rc
        return wrapped


def measured_traffic(step_kind, kernel, B, L):
    """Per-launch DRAM traffic of `kernel` from the committed ncu capture (profiles/r01j_traffic.json); only valid for the
    shape it was captured at (B=256, L=4096), otherwise None."""
    p = os.path.join(ROOT, "profiles", "r01j_traffic.json")
    if not os.path.exists(p) or (B, L) != (256, 4096):
        return None
    d = json.load(open(p)).get(step_kind, {}).get(kernel)
    return None if d is None else d["bytes_per_launch"]


def make_inputs(B, L, cin, seed):
    from weights import synthetic_chirps
    d = synthetic_chirps(B, L, snr=10.0, seed=seed)
    y, clean = d["y_norm"], d["clean_norm"]
    cond = torch.cat([y, torch.zeros(B, 4, L)], dim=1) if cin == 7 else y      # metadata channels = 0
    return clean, cond


def layer_table(spec, B, L):
    """Per conv block: (name, Cin, Cout, L_lvl, is_encoder) for the algorithmic FLOP / byte counts."""
    d, lc = spec.depth, spec.layer_channels
    Ls = spec.level_lengths(L)
    rows = []
    for li in range(2 * d + 1):
        if li == 0:
            cin, Ll = spec.in_ch, Ls[0]
        elif li <= d:
            cin, Ll = lc[li - 1], Ls[li]
        else:
            cin, Ll = lc[li - 1] + spec.chs[2 * d - li], Ls[2 * d - li]
        rows.append({"name": spec.layer_names()[li], "cin": cin, "cout": lc[li], "L": Ll, "enc": li < d})
    return rows


def family_rooflines(records, spec, B, L, pk, n_steps, train):
    """Group the timed C-ABI calls of `n_steps` eager steps into kernel families with algorithmic work and rooflines."""
    t_by = collections.OrderedDict()
    for name, s, e in records:
        t_by[name] = t_by.get(name, 0.0) + s.elapsed_time(e) / n_steps        # ms per step
    tab = layer_table(spec, B, L)
    conv_fl = sum(2.0 * r["cin"] * r["cout"] * 3 * r["L"] * B for r in tab[1:])          # K2..K7
    act = [B * r["cout"] * r["L"] for r in tab]
    gn_fwd_bytes = sum((2.5 if r["enc"] else 2.0) * n * 2 for r, n in zip(tab, act))
    gn_bwd_bytes = sum((6.0 if r["enc"] else 5.0) * n * 2 for r, n in zip(tab, act))
    step_ms = sum(t_by.values())
    fams = []

    def add(name, keys, bound, work, note):
        ms = sum(t_by.get(k, 0.0) for k in keys)
        if ms <= 0:
            return
        if bound == "tensor":
            ach, peak, unit = work / ms / 1e9, pk["bf16"], "TFLOP/s"
        else:
            ach, peak, unit = work / ms / 1e6, pk["hbm"], "GB/s"
        fams.append({"family": name, "entry_points": keys, "ms_per_step": ms, "share_of_step": ms / step_ms, "bound": bound,
                     "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "algorithmic": note})
    n_fwd = 1
    if train:
        # conv_tc is used by the forward(s) and by dgrad; split by call order is not needed for the family roofline
        n_fwd = sum(1 for n, _, _ in records if n == "gw_final_step") / n_steps
        add("conv fwd + dgrad (tcgen05 implicit GEMM)", ["gw_conv_tc"], "tensor", conv_fl * (n_fwd + 1.0),
            "2*Cin*Cout*3*L*B per conv, forward(s) + dgrad")
        add("wgrad (tcgen05 MN-major GEMM)", ["gw_wgrad_tc"], "tensor", conv_fl, "2*Cin*Cout*3*L*B per conv")
        add("GroupNorm/SiLU/cond/FiLM backward", ["gw_gn_bwd", "gw_gn_bwd2"], "hbm", gn_bwd_bytes,
            "bf16: 2 passes over (raw, dout[, dpool/2]) + d_raw write per block")
    else:
        # inference: one kernel per block does the conv AND the GroupNorm/SiLU/cond/FiLM/pool epilogue (conv_gn.cuh); its
        # roofline counts the conv FLOPs only, against the time of the whole fused kernel
        add("fused conv block (tcgen05 implicit GEMM + GroupNorm/SiLU/cond/FiLM/pool epilogue)", ["gw_conv_gn", "gw_conv_gn2", "gw_conv_gn3"], "tensor",
            conv_fl, "2*Cin*Cout*3*L*B per conv (the fused elementwise work is not counted)")
        add("conv fwd (tcgen05 implicit GEMM)", ["gw_conv_tc"], "tensor", conv_fl, "2*Cin*Cout*3*L*B per conv")
    gn_keys = [k for k in ("gw_gn_apply", "gw_gn_apply_stream") if k in t_by]
    n_gn = sum(1 for n, _, _ in records if n in gn_keys) / max(1, n_steps)
    gn_share = min(1.0, n_gn / (len(tab) * n_fwd)) if n_fwd else 0.0        # inference: only the first block is unfused
    gn_work = gn_fwd_bytes * n_fwd if gn_share >= 0.999 else (2.5 * act[0] * 2) * n_gn
    add("GroupNorm/SiLU/cond/FiLM/pool forward", gn_keys, "hbm", gn_work,
        "bf16: raw read + out write (+ pooled write) per unfused block")
    other = step_ms - sum(f["ms_per_step"] for f in fams)
    return fams, step_ms, other, t_by


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_chain_rate(args, steps_sample: int, B_cpu: int, workload: str):
    """Oracle port on the host cores: `B_cpu` waveforms x the first `steps_sample` steps of the chain, extrapolated."""
    import oracle
    from weights import make_state_dict
    n_steps, eta, start_t, _ = WORKLOADS[workload]
    cin = args.cin if args.workload != "train" else 3
    cc = 1 if cin == 3 else 5
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state_dict(cin, cc, seed=0)
    cfg = oracle.ModelCfg(in_ch=cin, cond_in_ch=cc, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    _, y = make_inputs(B_cpu, args.length, cin, seed=1234)
    sched = oracle.build_t_schedule(1000, n_steps, start_t)
    k = min(steps_sample, len(sched))
    t0 = time.perf_counter()
    calls = [0]

    def fwd(xi, ti):
        calls[0] += 1
        if calls[0] > k:
            raise StopIteration
        return oracle.unet_forward(sd, cfg, xi, ti)
    try:
        oracle.ddim_sample(sd, cfg, ab, y, T=1000, steps=n_steps, eta=eta, start_t=start_t, forward_fn=fwd)
    except StopIteration:
        pass
    per_step = (time.perf_counter() - t0) / k
    return B_cpu / (per_step * len(sched)), per_step, k, len(sched)


class CpuTrainer:
    """Reference training step on the host cores through the oracle port (autograd + clip + AdamW + EMA, fp32)."""

    def __init__(self, args, B_cpu: int):
        import oracle
        from weights import make_state_dict, gaussian
        self.oracle = oracle
        cin = 7
        torch.set_num_threads(os.cpu_count() or 1)
        self.sd = make_state_dict(cin, 5, seed=0)
        self.cfg = oracle.ModelCfg(in_ch=cin, cond_in_ch=5, use_selfcond=True)
        self.ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
        self.clean, self.cond = make_inputs(B_cpu, args.length, cin, seed=1234)
        self.mask = torch.ones(B_cpu, 1, args.length)
        self.eps = gaussian((B_cpu, 1, args.length), seed=3)
        g = torch.Generator().manual_seed(5)
        self.t = torch.randint(500, 1000, (B_cpu,), generator=g)
        self.drop = (torch.rand(B_cpu, 1, 1, generator=g) < 0.2).float()
        self.m = {k: torch.zeros_like(v) for k, v in self.sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.sd.items()}
        self.ema = {k: v.clone() for k, v in self.sd.items()}
        self.n, self.B = 0, B_cpu

    def step(self, selfcond: bool):
        o = self.oracle
        loss, grads, _ = o.train_step(self.sd, self.cfg, self.ab, clean_norm=self.clean, cond_stack=self.cond, mask=self.mask,
                                      t=self.t, eps=self.eps, drop=self.drop, selfcond=selfcond)
        grads, _ = o.clip_grad_norm(grads, 1.0)
        self.n += 1
        for k in self.sd:
            self.sd[k], self.m[k], self.v[k] = o.adamw_step(self.sd[k], grads[k], self.m[k], self.v[k], self.n, 2e-4)
            self.ema[k] = o.ema_step(self.ema[k], self.sd[k], 0.999)
        return float(loss)


def cpu_train_rate(args, n_steps: int, B_cpu: int, warm: int = 1):
    tr = CpuTrainer(args, B_cpu)
    coin = random.Random(0)
    for _ in range(warm):
        tr.step(False)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        tr.step(coin.random() < 0.5)
    dt = time.perf_counter() - t0
    return B_cpu * n_steps / dt, dt / n_steps


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    if args.workload == "train":
        B_cpu = 8
        tr = CpuTrainer(args, B_cpu)
        coin = random.Random(0)
        for _ in range(max(1, min(args.warmup, 2))):
            tr.step(False)
        k = max(1, min(args.steps, 40))
        t1 = time.perf_counter()
        for _ in range(k):
            tr.step(coin.random() < 0.5)
        dt = time.perf_counter() - t1
        v = B_cpu * k / dt
        ms = 1e3 * dt / k
        metric, unit = "train samples/sec", "samples/s"
        sample = f"{k} optimisation steps of batch {B_cpu} x {args.length} (in_ch=7, fp32, self-cond coin p=0.5), all host cores"
        wl = f"train: {TRAIN_DESC}, L={args.length}, in_ch=7"
    else:
        B_cpu, k = 8, 10
        vals = []
        for _ in range(max(1, min(args.steps, 3))):
            v_, per_step, kk, n = cpu_chain_rate(args, k, B_cpu, args.workload)
            vals.append(v_)
        v = statistics.median(vals)
        ms = 1e3 * (time.perf_counter() - t0) / max(1, len(vals))
        metric, unit = "waveforms/sec (full reverse chain)", "waveforms/s"
        sample = f"{B_cpu} waveforms x first {k} of {n} chain steps per bench step, extrapolated x{n / k:g}"
        wl = f"{args.workload}: {WORKLOADS[args.workload][3]}, L={args.length}, in_ch={args.cin}"
    line = {"metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": wl, "batch_per_gpu": B_cpu,
                       "note": "CPU oracle port of the reference (torch CPU ops); /root/reference is Python and cannot travel"},
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm: sampling
def bench_sampling(args, workload, B, steps, warmup, world, rank, dev, barrier, pk, with_cpu=True):
    from weights import make_state_dict
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf
    import torch.distributed as dist
    n_steps, eta, start_t, desc = WORKLOADS[workload]
    L = args.length
    cin = args.cin if args.workload != "train" else 3
    cc = 1 if cin == 3 else 5
    model = UNet1D(in_ch=cin, cond_in_ch=cc, use_selfcond=True, compute_dtype=args.dtype)
    model.load_state_dict(make_state_dict(cin, cc, seed=0))
    model = model.to(dev).eval()
    diff = CustomDiffusion(T=1000, device=dev)
    sample0 = rank * B                                # global sample index of this rank's shard (no collective)
    _, y = make_inputs(B, L, cin, seed=1234 + rank)
    y_host = y.pin_memory()
    out_host = torch.empty(B, 1, L).pin_memory()
    eng = model.engine(args.dtype)
    plan = inf.make_sampler_plan(model, diff, B, L, T=1000, steps=n_steps, eta=eta, start_t=start_t, seed=77, sample0=sample0,
                                 compute_dtype=args.dtype)
    y_dev = y_host.to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    x_T = torch.randn(B, 1, L, device=dev, generator=g)
    spg = args.steps_per_graph if args.steps_per_graph > 0 else plan.N

    def chain_resident():
        plan.load_inputs(x_T, y_dev, torch.zeros_like(y_dev), None)
        return plan.run(use_graph=True, steps_per_graph=spg)

    for _ in range(warmup):
        chain_resident()
    barrier()
    clocks = ClockSampler(dev.index)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        chain_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches_per_chain = None                         # counted from the profiled steps below

    def chain_e2e():
        cond = y_host.to(dev, non_blocking=True)
        out = inf.ddim_sample(model, diff, cond, 1000, n_steps, eta, dev, L, False, start_t, "noise", 0.14, 0.0, 1.0, 1.0,
                              "eps", cin, cc, True, 1.0, "const", 0.5, 0.3, 0.0, seed=77, sample0=sample0,
                              compute_dtype=args.dtype)
        out_host.copy_(out, non_blocking=True)
    for _ in range(max(1, warmup // 2)):
        chain_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(steps):
        chain_e2e()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    total_wf = world * B * steps
    value, e2e = total_wf / (ms / 1e3), total_wf / (ms_e2e / 1e3)
    res = None
    if rank == 0:
        # per-family rooflines from 4 eager reverse steps
        tl = TimedLib(eng.lib)
        eng.lib = tl
        plan.load_inputs(x_T, y_dev, torch.zeros_like(y_dev), None)
        for _ in range(2):
            plan.enqueue_step()
        tl.on = True
        for _ in range(4):
            plan.enqueue_step()
        torch.cuda.synchronize()
        tl.on = False
        eng.lib = tl._lib
        launches_per_chain = plan.N * (len(tl.records) // 4) + 1
        fams, step_ms_eager, other, _ = family_rooflines(tl.records, model.spec, plan.Bn, L, pk, 4, train=False)
        step_ms = ms / steps / plan.N
        flops_wf = model.spec.conv_flops(L) * plan.N
        chain_tflops = value * flops_wf / world / 1e12
        conv = fams[0]
        res = {"metric": "waveforms/sec (full reverse chain)", "value": value, "unit": "waveforms/s", "ms_per_chain": ms / steps,
               "workload": f"{workload}: {desc}, L={L}, in_ch={cin}", "batch_per_gpu": B, "chain_steps": plan.N,
               "cuda_graph_steps": spg,
               "e2e": {"value": e2e, "unit": "waveforms/s", "h2d_bytes_per_step": int(y_host.numel() * 4),
                       "d2h_bytes_per_step": int(out_host.numel() * 4)},
               "gpu_launches": launches_per_chain * steps, "clocks": clk,
               "roofline": {"bound": "tensor",
                            "kernel": "conv_gn_kernel (6 launches per reverse step: tcgen05+TMA implicit GEMM with the "
                                      "GroupNorm/SiLU/cond/FiLM/pool epilogue fused in)"
                            if "gw_conv_gn3" in conv["entry_points"] else "conv_tc2_kernel (tcgen05+TMA implicit GEMM)",
                            "achieved": conv["achieved"], "peak": pk["bf16"], "unit": "TFLOP/s", "frac": conv["frac"],
                            "traffic": measured_traffic("reverse_step", "conv_gn_kernel" if "gw_conv_gn3" in conv["entry_points"]
                                                        else "conv_tc2_kernel", plan.Bn, L),
                            "traffic_note": "bytes per launch, ncu capture profiles/r01j_traffic.json (B=256, L=4096, in_ch=3)",
                            "peak_source": pk["src"] + " burst (kernels timed one by one)",
                            "frac_sustained": conv["achieved"] / pk["bf16_sustained"],
                            "step_frac_sustained": chain_tflops / pk["bf16_sustained"],
                            "share_of_step": conv["share_of_step"]},
               "kernels": fams,
               "chain": {"tflops_per_gpu": chain_tflops, "frac_of_sustained_bf16_peak": chain_tflops / pk["bf16_sustained"],
                         "ms_per_reverse_step": step_ms, "flops_per_waveform": flops_wf,
                         "graph_capture_ms": getattr(plan, "graph_capture_ms", None),
                         "graph_capture_note": "stream capture + instantiate of the whole chain as one CUDA graph, host wall clock, "
                                               "once per (batch, length, schedule)"}}
        if with_cpu and world == 1:
            cpu_v, per_step, k, n = cpu_chain_rate(args, 6, 8, workload)
            res["cpu_baseline"] = {"value": cpu_v, "unit": "waveforms/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"8 waveforms x first {k} of {n} chain steps, extrapolated x{n / k:g}"}
    del plan
    inf._PLANS.clear()
    return res


# ------------------------------------------------------------------------------------------------ GPU arm: training
def bench_train(args, world, rank, dev, barrier, pk):
    import torch.distributed as dist
    from weights import make_state_dict
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep
    B, L, cin, cc = args.batch, args.length, 7, 5
    model = UNet1D(in_ch=cin, cond_in_ch=cc, use_selfcond=True, compute_dtype=args.dtype)
    model.load_state_dict(make_state_dict(cin, cc, seed=0))
    model = model.to(dev)
    diff = CustomDiffusion(T=1000, device=dev)
    st = FusedTrainStep(model, diff, B, L, lr=2e-4, weight_decay=1e-4, clip_grad=1.0, ema_decay=0.999, loss="huber",
                        huber_beta=0.5, clamp_inputs=10.0, p_uncond=0.2, dropout_y_only=True, t_min=500, warmup_steps=1000,
                        total_steps=100000, compute_dtype=args.dtype, seed=42, sample0=rank * B)
    clean, cond = make_inputs(B, L, cin, seed=1234 + rank)
    mask = torch.ones(B, 1, L)
    hb = [t.pin_memory() for t in (clean, cond, mask)]
    st.load_batch(hb[0].to(dev), hb[1].to(dev), hb[2].to(dev))
    coin = random.Random(0)                           # train.py:401 coin from a host RNG: identical on every rank
    st.step(selfcond=False)                           # capture both step flavours before anything is timed
    st.step(selfcond=True)
    for _ in range(max(args.warmup, 3)):
        st.step(selfcond=coin.random() < args.p_selfcond)
    seq = [coin.random() < args.p_selfcond for _ in range(args.steps)]
    barrier()
    clocks = ClockSampler(dev.index)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for sc in seq:
        st.step(selfcond=sc)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    loss_resident = float(st.loss)

    # ---- e2e: pinned host batch in, loss out, every step
    loss_host = torch.zeros(args.steps + 4, 1).pin_memory()
    st.prefetch(*hb)
    for i in range(2):
        st.step(selfcond=False, prefetched=True)
        st.prefetch(*hb)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i, sc in enumerate(seq):
        st.step(selfcond=sc, prefetched=True)
        st.prefetch(*hb)                              # next batch's H2D overlaps this step
        loss_host[i].copy_(st.loss, non_blocking=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    total = world * B * args.steps
    value, e2e = total / (ms / 1e3), total / (ms_e2e / 1e3)
    # ---- SURVEY 8(d) config 2 read literally: GLOBAL batch 256 split over the ranks (strong scaling), and the collective alone
    strong, allreduce_us = None, None
    if world > 1:
        bucket = st.bucket[: st.layout.total + 1]
        for _ in range(5):
            dist.all_reduce(bucket)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(50):
            dist.all_reduce(bucket)
        a1.record()
        torch.cuda.synchronize()
        ta = torch.tensor([a0.elapsed_time(a1) / 50 * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        allreduce_us = float(ta[0])
        Bs = max(1, args.batch // world)
        st2 = FusedTrainStep(model, diff, Bs, L, lr=2e-4, weight_decay=1e-4, clip_grad=1.0, ema_decay=0.999, loss="huber",
                             huber_beta=0.5, clamp_inputs=10.0, p_uncond=0.2, dropout_y_only=True, t_min=500, warmup_steps=1000,
                             total_steps=100000, compute_dtype=args.dtype, seed=42, sample0=rank * Bs, share=st)
        st2.load_batch(hb[0][:Bs].to(dev), hb[1][:Bs].to(dev), hb[2][:Bs].to(dev))
        st2.step(selfcond=False)
        st2.step(selfcond=True)
        for _ in range(max(args.warmup, 3)):
            st2.step(selfcond=coin.random() < args.p_selfcond)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for sc in seq:
            st2.step(selfcond=sc)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_s = float(ts[0])
        strong = {"global_batch": Bs * world, "batch_per_gpu": Bs, "ms_per_step": ms_s / args.steps,
                  "value": world * Bs * args.steps / (ms_s / 1e3), "unit": "samples/s", "scaling": "strong",
                  "note": "same step, same graph structure; per-GPU work shrinks with N, so launch latency (one graph of ~120 "
                          "kernels) and the all-reduce weigh more"}
        del st2
    # ---- per-family rooflines: eager steps with every C-ABI call timed (every rank runs them: the step has a collective)
    tl = TimedLib(st.lib)
    st.lib = st.eng.lib = st.bwd.lib = tl
    n_prof = 4
    l0 = st.eng.launches
    st.step(selfcond=False, use_graph=False)
    tl.on = True
    prof_seq = [False, True, False, True][:n_prof]
    for sc in prof_seq:
        st.step(selfcond=sc, use_graph=False)
    torch.cuda.synchronize()
    tl.on = False
    st.lib = st.eng.lib = st.bwd.lib = tl._lib
    launches_per_step = (st.eng.launches - l0) / (n_prof + 1)
    if rank != 0:
        return None
    fams, step_ms_eager, other, t_by = family_rooflines(tl.records, model.spec, B, L, pk, n_prof, train=True)
    frac_sc = sum(seq) / len(seq)
    flops_sample = 3.0 * model.spec.conv_flops(L) + frac_sc * model.spec.conv_flops(L)
    step_ms = ms / args.steps
    tfl = value / world * flops_sample / 1e12
    conv = fams[0]
    # the CPU baseline is timed at N = 1 only (with N ranks the other processes contend for the host cores; the reference arm,
    # `--impl reference`, is a separate process and is the usable CPU number at every N)
    cpu_baseline = None
    if world == 1:
        cpu_v, cpu_step = cpu_train_rate(args, 4, 8)
        cpu_baseline = {"value": cpu_v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": "4 optimisation steps of batch 8 x 4096 (in_ch=7, fp32), all host cores"}
    line = {"metric": "train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"train: {TRAIN_DESC}, L={L}, in_ch={cin}", "batch_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"dp{world} (batch shards; one NCCL all-reduce of the 4.27 MB fp32 gradient bucket per step)",
                       "selfcond_steps": f"{sum(seq)}/{len(seq)} (host coin p={args.p_selfcond}, seed 0)",
                       "cuda_graph": "1 graph per step (pack .. backward, NCCL all-reduce of the bucket, clip+AdamW+EMA, re-pack)",
                       "strong_scaling": "see `strong` (global batch 256 split over the ranks); the headline keeps 256 per GPU",
                       "l2": "per-step activations + gradients ~4 GB >> 126 MB L2 (inputs larger than L2)",
                       "weights": "random-init (numpy PCG64 seed 0), final.* ~ N(0,0.05^2); fp32 master, bf16 GEMM operands",
                       "loss_after": loss_resident},
            "e2e": {"value": e2e, "unit": "samples/s",
                    "h2d_bytes_per_step": int(sum(t_.numel() for t_ in hb) * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "conv_tc2_kernel family (forward convs + dgrad, tcgen05+TMA implicit GEMM)",
                         "achieved": conv["achieved"], "peak": pk["bf16"], "unit": "TFLOP/s", "frac": conv["frac"],
                         "traffic": measured_traffic("train_step", "conv_tc2_kernel", B, L),
                         "traffic_note": "bytes per launch, ncu capture profiles/r01j_traffic.json (B=256, L=4096, in_ch=7)",
                         "peak_source": pk["src"] + " burst (kernels timed one by one, eager)",
                         "frac_sustained": conv["achieved"] / pk["bf16_sustained"],
                         "step_frac_sustained": tfl / pk["bf16_sustained"],
                         "share_of_step": conv["share_of_step"]},
            "strong": strong, "allreduce_us": allreduce_us,
            "kernels": fams,
            "step": {"tflops_per_gpu": tfl, "frac_of_sustained_bf16_peak": tfl / pk["bf16_sustained"],
                     "flops_per_sample": flops_sample, "eager_sum_ms": step_ms_eager, "other_kernels_ms": other},
            "cpu_baseline": cpu_baseline}
    return line


# ------------------------------------------------------------------------------------------------ GPU arm: whitening / scoring
def bench_aux(args, dev, pk):
    """`--workload whiten`: batched train-like whitening of (y, clean) + sigma (inference.py:125-153 / dataloader.py:110-200);
    `--workload score`: the batched scoring kernel (inference.py:11-27, 247-314).  HBM-bound by intent: algorithmic bytes =
    every input read once + every output written once."""
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import scoring, whitening
    B, L = args.batch, args.length
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    y = torch.randn(B, L, device=dev, generator=g) * 3e-3 + 1e-3
    x = torch.randn(B, L, device=dev, generator=g) * 1e-3
    y_host, x_host = y.cpu().pin_memory(), x.cpu().pin_memory()
    if args.workload == "whiten":
        def step_resident():
            y_w, x_w, P = whitening.whiten_train_like(y, x)
            return whitening.sigma(y_w, "std")
        alg_bytes = B * (2 * L * 4 + 2 * L * 4 + (L // 2 + 1) * 8 + L * 4 + 8)      # y, x in; y_w, x_w, P out; sigma pass
        def step_e2e():
            yd, xd = y_host.to(dev, non_blocking=True), x_host.to(dev, non_blocking=True)
            y_w, x_w, P = whitening.whiten_train_like(yd, xd)
            return whitening.sigma(y_w, "std").cpu()
        metric, unit, kernel = "whitened segments/sec", "segments/s", "cuFFT D2Z/Z2D (fp64) + periodogram / spectral-scale / finish / sigma kernels"
        h2d, d2h = 2 * B * L * 4, B * 8
    else:
        sig = torch.full((B,), 1e-3, device=dev)
        def step_resident():
            return scoring.score_batch(y, x, 4096.0, sigma=sig, secs=0.8, max_shift=82)["corr_last"]
        alg_bytes = B * (2 * L * 4 + 12 * 8)
        def step_e2e():
            yd, xd = y_host.to(dev, non_blocking=True), x_host.to(dev, non_blocking=True)
            return scoring.score_batch(yd, xd, 4096.0, sigma=sig, secs=0.8, max_shift=82)["corr_last"].cpu()
        metric, unit, kernel = "scored reconstructions/sec", "reconstructions/s", "score_batch_kernel (one CTA per sample, fp64 accumulation, +-82 lag search)"
        h2d, d2h = 2 * B * L * 4, B * 8
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step_e2e()
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    per = ms / args.steps
    ach = alg_bytes / (per * 1e-3) / 1e9
    return {"metric": metric, "value": B * args.steps / (ms / 1e3), "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.workload == "whiten" else "f32/f64",
            "data": "synthetic", "config": {"workload": f"{args.workload}: batch {B} x {L}", "l2": "inputs + workspace >> 126 MB L2" if B * L * 40 > 126e6 else "fits L2 partly"},
            "e2e": {"value": B * args.steps / (ms_e2e / 1e3), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": None, "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                         "traffic": None, "algorithmic_bytes_per_step": alg_bytes, "peak_source": pk["src"],
                         "note": "whole step (several kernels + cuFFT) against the bytes that must move once; the fp64 FFT workspace "
                                 "passes are not algorithmic bytes" if args.workload == "whiten" else
                                 f"one kernel; its time is the fp64 lag search (2 x 165 lags x L FLOP per sample = "
                                 f"{2 * 165 * L * B / (per * 1e-3) / 1e12:.1f} TFLOP/s fp64, register-tiled out of shared memory), "
                                 "not the 33 KB per sample it reads"}}


def run_ours(args):
    import torch.distributed as dist
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _cabi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load()                                      # fail loudly if the CUDA library is missing
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload in ("whiten", "score"):
        if rank == 0:
            emit(bench_aux(args, dev, pk))
    elif args.workload == "train":
        line = bench_train(args, world, rank, dev, barrier, pk)
        samp = None
        if not args.no_sampling:
            samp = bench_sampling(args, "ddpm1000", args.batch, 2, 3, world, rank, dev, barrier, pk, with_cpu=True)
        if rank == 0:
            line["sampling"] = samp
            emit(line)
    else:
        B = args.batch
        res = bench_sampling(args, args.workload, B, args.steps, max(args.warmup, 3), world, rank, dev, barrier, pk)
        if rank == 0:
            line = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": res["ms_per_chain"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                    "config": {"workload": res["workload"], "batch_per_gpu": B, "global_batch": world * B,
                               "chain_steps": res["chain_steps"], "parallelism": f"dp{world} (batch shards, no collective)",
                               "cuda_graph_steps": res["cuda_graph_steps"],
                               "l2": "activations per reverse step >> 126 MB L2 (inputs larger than L2)",
                               "weights": "random-init (numpy PCG64 seed 0), final.* ~ N(0,0.05^2)"},
                    "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "roofline": res["roofline"],
                    "kernels": res["kernels"], "chain": res["chain"], "cpu_baseline": res.get("cpu_baseline")}
            emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line goes to the process's ORIGINAL stdout; everything else a library prints there (e.g. NCCL's
    "NCCL version ..." banner) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                     # stray prints of native libraries -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train"] + list(WORKLOADS) + ["whiten", "score"])
    ap.add_argument("--batch", type=int, default=256, help="samples / waveforms per GPU")
    ap.add_argument("--length", type=int, default=4096)
    ap.add_argument("--cin", type=int, default=3, choices=[3, 7], help="input channels of the sampling workloads")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--p-selfcond", type=float, default=0.5)
    ap.add_argument("--no-sampling", action="store_true", help="train workload: skip the secondary sampling measurement")
    ap.add_argument("--steps-per-graph", type=int, default=0, help="reverse steps per CUDA graph (0 = whole chain)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
