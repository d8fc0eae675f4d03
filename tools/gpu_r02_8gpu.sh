#!/bin/bash
# 8 GPUs: BASELINE configs 2 (train, weak + strong), 3 (DDIM-50 x 8192 segments) and 5 (SNR sweep over 65536 injections)
cd "$(dirname "$0")/.."
TAG=r02_8gpu
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 110 $RUN --master-port 29721 bench.py --gpus 8 --steps 50 --warmup 10 --no-sampling > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench train exit $?" >> $S
timeout 90 $RUN --master-port 29722 bench.py --gpus 8 --workload ddim50 --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_ddim50_$TAG.json 2> gpurun_out/bench_ddim50_$TAG.err; echo "bench ddim50 exit $?" >> $S
timeout 120 $RUN --master-port 29723 tools/snr_sweep.py --count 65536 --chunk 1024 --steps 50 --save-first gpurun_out/sweep_first256_$TAG.pt > gpurun_out/sweep_$TAG.json 2> gpurun_out/sweep_$TAG.err; echo "sweep exit $?" >> $S
cat $S
python - <<PY
import json
for f,ks in (("gpurun_out/bench_train_$TAG.json",("value","ms_per_step","allreduce_us","strong","e2e")),("gpurun_out/bench_ddim50_$TAG.json",("value","ms_per_step","e2e","chain"))):
    try:
        d=json.load(open(f)); print(f, {k:d.get(k) for k in ks})
    except Exception as e: print(f, "ERR", e)
try: print(open("gpurun_out/sweep_$TAG.json").read()[:1500])
except Exception as e: print(e)
PY
tail -2 gpurun_out/sweep_$TAG.err | cut -c1-300
