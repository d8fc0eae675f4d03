#!/bin/bash
# deferred small reductions (parameter-gradient kernels, wgrad fold/scatter, time-MLP backward on a second stream): parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_bench_configs.py tests/test_gpu_bwd_kernels.py tests/test_gpu_dp_nccl.py -q -x 2>&1 | tail -4 > gpurun_out/r02j_pytest.txt
cat gpurun_out/r02j_pytest.txt
: > gpurun_out/r02j_defer_ab.txt
for B in 256 32; do
for F in 0 1; do
  GWB200_DEFER=$F python tools/train_profile.py --B $B --steps 6 2>&1 | grep "graph-replayed" | sed "s/^/B=$B defer=$F: /" | tee -a gpurun_out/r02j_defer_ab.txt
done
done
for F in 0 1; do
  GWB200_DEFER=$F python bench.py --steps 100 --warmup 10 --no-sampling 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench defer=$F', d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])" | tee -a gpurun_out/r02j_defer_ab.txt
done
