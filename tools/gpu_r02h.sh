#!/bin/bash
# programmatic dependent launch across the training-step kernels: parity (GPU suite with the attribute on) and A/B timing
mkdir -p gpurun_out
GWB200_OPTIONS="pdl=1" timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_bwd_kernels.py tests/test_gpu_forward.py tests/test_gpu_chain.py tests/test_gpu_bench_configs.py -q -x 2>&1 | tail -5 > gpurun_out/r02h_pytest_pdl.txt
cat gpurun_out/r02h_pytest_pdl.txt
for B in 32 256; do
  for P in 0 1; do
    python tools/train_profile.py --B $B --opt pdl=$P --steps 6 2>&1 | grep "graph-replayed" | sed "s/^/B=$B pdl=$P: /" | tee -a gpurun_out/r02h_pdl_ab.txt
  done
done
for P in 0 1; do
  GWB200_OPTIONS="pdl=$P" python bench.py --steps 60 --warmup 10 --no-sampling 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench pdl=$P', d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])" | tee -a gpurun_out/r02h_pdl_ab.txt
done
