"""The real NCCL data-parallel path on >= 2 GPUs: FusedTrainStep on batch shards (one process per GPU, NCCL all-reduce of
the flat gradient bucket inside the step, eager and CUDA-graph flavours) against the single-GPU run on the full batch.

Skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp_nccl.py -m gpu`
(log committed under profiles/).  Reference semantics: train.py:419-421 (mean of per-sample means) and :445 (the clip uses
the norm of the full-batch gradient)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, dtype, gtol, backend):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import dp_nccl_worker as W
    Bg, L = 8, 1024
    torch.cuda.set_device(0)
    _, st = W.build(dtype, Bg, L, 0)
    assert st.world == 1
    p0 = st.flat_p.clone().cpu()
    clean, cond, mask = W.batch(Bg, L)
    ref = W.run_steps(st, clean, cond, mask, use_graph=False)
    payload = {0: ref[0], 1: ref[1], "p0": p0}
    ref_path, out_path = str(tmp_path / "ref.pt"), str(tmp_path / "out.pt")
    torch.save(payload, ref_path)
    del st
    torch.cuda.empty_cache()
    port = 29600 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_nccl_worker.py"), "--ref", ref_path, "--out", out_path,
           "--dtype", dtype, "--Bg", str(Bg), "--L", str(L), "--backend", backend]
    if backend == "gloo":
        cmd.append("--same-device")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-4000:]
    rep = torch.load(out_path)
    print(rep)
    for tag in ("eager", "graph"):
        for i in range(2):
            assert rep[f"{tag}.step{i}.t_equal"], (tag, i)                  # Philox draws keyed on the global sample index
            assert rep[f"{tag}.step{i}.grad_rel_l2"] <= gtol, (tag, i, rep)
            assert rep[f"{tag}.step{i}.norm_rel"] <= 1e-5, (tag, i, rep)
        # Adam's first updates are ~ lr * sign(g): elements whose gradient is at rounding level may flip, the rest agree
        assert rep[f"{tag}.update_rel_l2"] <= 2e-2, (tag, rep)
        assert rep[f"{tag}.param_frac_within_1e-6"] >= 0.995, (tag, rep)


@pytest.mark.parametrize("dtype,gtol", [("fp32", 2e-5), ("bf16", 2e-5)])
def test_dp_nccl_matches_single_gpu_full_batch(tmp_path, dtype, gtol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(tmp_path, dtype, gtol, "nccl")


@pytest.mark.parametrize("dtype,gtol", [("fp32", 2e-5), ("bf16", 2e-5)])
def test_dp_two_ranks_one_gpu_gloo(tmp_path, dtype, gtol):
    """The same comparison on a ONE-GPU box: two processes share cuda:0 and all-reduce the CUDA bucket through gloo (staged on
    the host) -- shard assignment, Philox keyed on the global sample index, the loss slot of the bucket, the 1/world scale and
    the global-norm clip of the real FusedTrainStep DP path, in the eager and the graph-replay flavour."""
    _run(tmp_path, dtype, gtol, "gloo")
