#!/bin/bash
# 2 GPUs: the real NCCL data-parallel path against the 1-GPU full batch, then the train bench line (weak + strong + all-reduce)
cd "$(dirname "$0")/.."
TAG=r02_2gpu
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
nvidia-smi -L > gpurun_out/smi_$TAG.txt
timeout 200 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x -k nccl > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest nccl exit $?" >> $S
tail -5 gpurun_out/pytest_$TAG.log | cut -c1-600
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 10 --no-sampling > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench exit $?" >> $S
cat $S
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_train_$TAG.json"))
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","allreduce_us")}, d["strong"], d["e2e"])
except Exception as e: print("no bench line", e)
PY
tail -3 gpurun_out/bench_train_$TAG.err | cut -c1-400
