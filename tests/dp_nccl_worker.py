"""Worker of tests/test_gpu_dp_nccl.py (launched with torch.distributed.run, one process per GPU, NCCL).

Each rank trains on its contiguous shard of the global batch with on-device Philox draws keyed on the GLOBAL sample index;
the flat gradient bucket is all-reduced over NCCL inside FusedTrainStep.step.  Rank 0 compares the reduced gradient and the
parameters after two optimisation steps with the single-GPU run on the full batch (written by the parent test)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build(dtype, B, L, sample0, world_arg=None, **kw):
    from weights import make_state_dict
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep
    m = UNet1D(in_ch=7, cond_in_ch=5, use_selfcond=True, compute_dtype=dtype)
    m.load_state_dict(make_state_dict(7, 5, seed=2), strict=True)
    dev = torch.device("cuda", torch.cuda.current_device())
    st = FusedTrainStep(m.to(dev), CustomDiffusion(T=1000, device=dev), B, L, lr=2e-4, p_uncond=0.2, seed=11, sample0=sample0,
                        compute_dtype=dtype, **kw)
    return m, st


def batch(Bg, L):
    from weights import gaussian, synthetic_chirps
    d = synthetic_chirps(Bg, L, snr=12.0, seed=31)
    cond = torch.cat([d["y_norm"], gaussian((Bg, 4, 1), seed=6).expand(Bg, 4, L).contiguous() * 0.3], dim=1)
    mask = torch.ones(Bg, 1, L)
    mask[1, :, :37] = 0.0
    return d["clean_norm"], cond, mask


def run_steps(st, clean, cond, mask, use_graph):
    st.load_batch(clean.cuda(), cond.cuda(), mask.cuda())
    out = []
    for i in range(2):
        st.step(selfcond=(i == 1), use_graph=use_graph)
        torch.cuda.synchronize()
        out.append({"g": st.flat_g.clone().cpu(), "p": st.flat_p.clone().cpu(), "loss": float(st.loss), "t": st.t.clone().cpu(),
                    "norm": float(st.info[0])})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--Bg", type=int, default=8)
    ap.add_argument("--L", type=int, default=1024)
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"],
                    help="gloo + --same-device: both ranks share cuda:0 (host-staged all-reduce): the DP logic on a one-GPU box")
    ap.add_argument("--same-device", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if a.same_device:
        local = 0
    torch.cuda.set_device(local)
    if a.backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.parallel import current_shard
    sh = current_shard(a.Bg)
    clean, cond, mask = batch(a.Bg, a.L)
    sl = slice(sh.start, sh.start + sh.count)
    res = {}
    for use_graph in (False, True):
        _, st = build(a.dtype, sh.count, a.L, sh.start)
        assert st.world == world
        res[use_graph] = run_steps(st, clean[sl], cond[sl], mask[sl], use_graph)
    if rank == 0:
        ref = torch.load(a.ref)
        rep = {"world": world, "dtype": a.dtype}
        for use_graph in (False, True):
            tag = "graph" if use_graph else "eager"
            for i in range(2):
                r, o = ref[i], res[use_graph][i]
                g_dp = o["g"].double() / world
                rep[f"{tag}.step{i}.grad_rel_l2"] = float((g_dp - r["g"].double()).norm() / r["g"].double().norm())
                rep[f"{tag}.step{i}.t_equal"] = bool(torch.equal(o["t"], r["t"][sl]))
                rep[f"{tag}.step{i}.norm_rel"] = abs(o["norm"] - r["norm"]) / r["norm"]
            p0 = ref["p0"].double()
            du, dr = res[use_graph][1]["p"].double() - p0, ref[1]["p"].double() - p0
            rep[f"{tag}.update_rel_l2"] = float((du - dr).norm() / dr.norm())
            rep[f"{tag}.param_frac_within_1e-6"] = float(((du - dr).abs() <= 1e-6).double().mean())
        torch.save(rep, a.out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
