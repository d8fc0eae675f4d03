#!/usr/bin/env python
"""BASELINE config 5: SNR sweep over synthetic injections, DDIM reconstruction sharded over the ranks (no collective in
the data path), throughput + overlap-vs-oracle on a subset.

    python tools/snr_sweep.py --n 65536 --chunk 1024 --steps 50            # 1 GPU
    torchrun --nproc-per-node 8 tools/snr_sweep.py --n 65536 ...             # 8 GPUs, 8192 injections each

Each rank reconstructs its contiguous shard (parallel.shard_range) in chunks through `inference.ddim_sample`; Philox /
synthetic-data seeds are keyed on the global injection index, so results do not depend on the world size.  Rank 0 also runs
the CPU oracle (the reference restated) on the first `--check` injections and reports the overlap <a,b>/(|a||b|) and the
reference's tail-window Pearson correlation (inference.py:15-18) between the two reconstructions.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from weights import make_state_dict, synthetic_chirps  # noqa: E402


def overlap(a, b):
    a, b = a.double().reshape(a.shape[0], -1), b.double().reshape(b.shape[0], -1)
    return ((a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1) + 1e-30)).float()


def corr(a, b):
    a = a.double().reshape(a.shape[0], -1)
    b = b.double().reshape(b.shape[0], -1)
    a = a - a.mean(1, keepdim=True)
    b = b - b.mean(1, keepdim=True)
    return ((a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1) + 1e-12)).float()


def injections(start, count, L, seed):
    """Injection i (global index) has SNR ~ U[5, 30] and its own chirp / noise realisation."""
    outs = [synthetic_chirps(1, L, snr=5.0, snr_hi=30.0, seed=seed + start + i) for i in range(count)]
    return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}


def run(a):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import scoring
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.parallel import shard_range
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd = make_state_dict(3, 1, seed=0)
    model = UNet1D(in_ch=3, cond_in_ch=1, use_selfcond=True, compute_dtype=a.dtype)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    diff = CustomDiffusion(T=1000, device=dev)
    start, count = shard_range(a.n, rank, world)
    recon, clean, snr, ov_dev, corr_dev = [], [], [], [], []
    torch.cuda.synchronize()
    t_data = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gpu_ms = 0.0
    # one-time setup, reported on its own: sampler plan (FiLM table, coefficient table) + capture of the 50-step chain as one graph
    t0 = time.perf_counter()
    n0 = min(a.chunk, count)
    plan = inf.make_sampler_plan(model, diff, n0, a.length, T=1000, steps=a.steps, eta=a.eta, start_t=a.start_t, seed=a.seed,
                                 sample0=start, compute_dtype=a.dtype)
    plan.load_inputs(torch.zeros(n0, 1, a.length, device=dev), torch.zeros(n0, 1, a.length, device=dev),
                     torch.zeros(n0, 1, a.length, device=dev), None)
    plan.capture(plan.N)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    for c0 in range(0, count, a.chunk):
        n = min(a.chunk, count - c0)
        t0 = time.perf_counter()
        d = injections(start + c0, n, a.length, a.seed)
        t_data += time.perf_counter() - t0
        # inputs of the chunk (measurement, injected clean waveform, sigma) staged through pinned memory before the timed region
        y = d["y_norm"].pin_memory().to(dev, non_blocking=True)
        clean_dev = d["clean_norm"].pin_memory().to(dev, non_blocking=True)
        sigma_dev = d["sigma"].to(dev)
        e0.record()
        x0 = inf.ddim_sample(model, diff, y, 1000, a.steps, a.eta, dev, a.length, False, a.start_t, "noise", 0.14, 0.0, 1.0, 1.0,
                             "eps", 3, 1, True, 1.0, "const", 0.5, 0.3, 0.0, seed=a.seed, sample0=start + c0,
                             compute_dtype=a.dtype)
        # scored on the device (gw_score_batch): overlap with the injected clean waveform, tail-window correlation
        sc = scoring.score_batch(x0, clean_dev, 4096.0, sigma=sigma_dev, secs=0.8, max_shift=1)
        e1.record()
        torch.cuda.synchronize()
        gpu_ms += e0.elapsed_time(e1)
        ov_dev.append(sc["overlap"].cpu())
        corr_dev.append(sc["corr_last"].cpu())
        recon.append(x0.cpu())
        clean.append(d["clean_norm"])
        snr.append(d["snr"])
    recon, clean, snr = torch.cat(recon), torch.cat(clean), torch.cat(snr)
    ov_dev, corr_dev = torch.cat(ov_dev).float(), torch.cat(corr_dev).float()
    t = torch.tensor([gpu_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = None
    if rank == 0:
        ov_clean = ov_dev
        assert float((overlap(recon, clean) - ov_dev).abs().max()) < 1e-5        # device scores == host recomputation
        res = {"n": a.n, "world": world, "steps": a.steps, "length": a.length, "dtype": a.dtype,
               "waveforms_per_s": a.n / (float(t[0]) / 1e3), "gpu_ms_max_rank": float(t[0]), "setup_ms_rank0": setup_ms, "host_data_gen_s_rank0": t_data,
               "tail_corr_vs_clean_mean": float(corr_dev.mean()),
               "overlap_vs_clean_by_snr": {f"{lo}-{lo + 5}": float(ov_clean[(snr >= lo) & (snr < lo + 5)].mean())
                                           for lo in range(5, 30, 5) if ((snr >= lo) & (snr < lo + 5)).any()}}
        if getattr(a, "save_first", None):
            # the first injections of rank 0's shard: re-done on one GPU with the CPU oracle (`--check K --compare-first file`),
            # since the result of an injection does not depend on the world size or on the chunking
            k = min(a.save_first_k, recon.shape[0])
            # x_T of these injections (step 0 of each one's Philox stream), so that tools/sweep_oracle_check.py can redo them on a CPU
            xT = inf.philox_normal(k, a.length, a.seed, 0, 0, dev).cpu()
            torch.save({"recon": recon[:k].clone(), "xT": xT, "n": a.n, "world": world, "steps": a.steps, "seed": a.seed,
                        "eta": a.eta, "start_t": a.start_t, "length": a.length, "dtype": a.dtype}, a.save_first)
        if getattr(a, "compare_first", None):
            ref8 = torch.load(a.compare_first)
            k = min(ref8["recon"].shape[0], recon.shape[0])
            res["same_as_saved_run"] = {"k": k, "saved_world": ref8["world"], "saved_n": ref8["n"],
                                        "bit_equal": bool(torch.equal(ref8["recon"][:k], recon[:k])),
                                        "max_abs_diff": float((ref8["recon"][:k] - recon[:k]).abs().max())}
        if a.check > 0:
            import oracle
            k = min(a.check, recon.shape[0])
            d = injections(0, k, a.length, a.seed)
            cfg = oracle.ModelCfg(in_ch=3, cond_in_ch=1, use_selfcond=True)
            ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
            # the chain's only draw at eta=0 is x_T = step 0 of each injection's Philox stream
            xT = inf.philox_normal(k, a.length, a.seed, 0, 0, dev).cpu()
            torch.set_num_threads(os.cpu_count() or 1)
            ref = oracle.ddim_sample(sd, cfg, ab, d["y_norm"], T=1000, steps=a.steps, eta=a.eta, start_t=a.start_t, noise=[xT])
            tail = slice(int(0.6 * a.length), a.length)
            res["check"] = {"k": k, "overlap_vs_oracle_min": float(overlap(recon[:k], ref).min()),
                            "overlap_vs_oracle_mean": float(overlap(recon[:k], ref).mean()),
                            "tail_corr_vs_oracle_min": float(corr(recon[:k, :, tail], ref[:, :, tail]).min()),
                            "rel_l2_max": float(((recon[:k] - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max())}
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--count", dest="n", type=int, default=65536)
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--length", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--eta", type=float, default=0.0)
    ap.add_argument("--start-t", type=int, default=None)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--check", type=int, default=0, help="injections re-done with the CPU oracle on rank 0")
    ap.add_argument("--save-first", default=None, help="file: keep rank 0's first --save-first-k reconstructions")
    ap.add_argument("--save-first-k", type=int, default=256)
    ap.add_argument("--compare-first", default=None, help="file written by --save-first of another run (e.g. the 8-GPU one)")
    run(ap.parse_args())


if __name__ == "__main__":
    main()
