"""World-size-2 data-parallel logic on CPU (gloo): shard assignment, flat-bucket all-reduce, 1/world averaging.

Each rank differentiates its shard with the oracle, packs the gradients in the product's ParamLayout order, sums the bucket
with the product's `allreduce_flat_` and scales by `grad_scale()`; the result must equal the full-batch oracle gradient
(the loss is a mean of per-sample means, train.py:419-421, so equal shards average exactly)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import oracle
    from weights import gaussian, make_state_dict, synthetic_chirps
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.engine import ModelSpec, ParamLayout
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.parallel import (allreduce_flat_, current_shard,
                                                                                          grad_scale)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    B, L, in_ch, cc = 4, 128, 3, 1
    sd = make_state_dict(in_ch, cc, seed=2)
    cfg = oracle.ModelCfg(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    ab = oracle.alpha_bar_from_betas(oracle.cosine_beta_schedule(1000))
    d = synthetic_chirps(B, L, snr=12.0, seed=31)
    t = torch.tensor([500, 731, 999, 612])
    eps = gaussian((B, 1, L), seed=41)
    mask = torch.ones(B, 1, L)
    sh = current_shard(B)
    assert (sh.rank, sh.world, sh.count) == (rank, world, B // world)
    sl = slice(sh.start, sh.start + sh.count)
    loss, grads, _ = oracle.train_step(sd, cfg, ab, clean_norm=d["clean_norm"][sl], cond_stack=d["y_norm"][sl], mask=mask[sl],
                                       t=t[sl], eps=eps[sl])
    spec = ModelSpec(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True)
    lo = ParamLayout(spec, {k: tuple(v.shape) for k, v in sd.items()})
    flat = torch.zeros(lo.total)
    for k, v in lo.views(flat).items():
        v.copy_(grads[k])
    allreduce_flat_(flat)
    flat *= grad_scale()
    lt = loss.clone().reshape(1)
    dist.all_reduce(lt)
    if rank == 0:
        loss_f, grads_f, _ = oracle.train_step(sd, cfg, ab, clean_norm=d["clean_norm"], cond_stack=d["y_norm"], mask=mask, t=t,
                                               eps=eps)
        worst = 0.0
        for k, v in lo.views(flat).items():
            worst = max(worst, float((v - grads_f[k]).abs().max() / (grads_f[k].abs().max() + 1e-12)))
        torch.save({"worst": worst, "loss_err": abs(float(lt) / world - float(loss_f))}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_gradient_bucket_matches_full_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["worst"] <= 2e-5, res
    assert res["loss_err"] <= 1e-6, res


def test_shard_range_covers_everything():
    from diffusion_models_for_gravitational_waveform_reconstruction_b200.parallel import shard_range
    for n, w in [(8192, 8), (10, 4), (3, 8), (256, 1)]:
        got = []
        for r in range(w):
            s, c = shard_range(n, r, w)
            got += list(range(s, s + c))
        assert got == list(range(n))
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)
