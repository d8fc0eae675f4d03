#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_chain.py -m gpu -x -q > gpurun_out/pytest_j14.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_j14.log
for o in pdl=1 pdl=0 pdl=1; do
  GWB200_OPTIONS=$o timeout 300 python bench.py --workload ddpm1000 --batch 256 --steps 2 --warmup 3 2>gpurun_out/b_$o.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$o', round(d['value'],1), d['chain']['frac_of_sustained_bf16_peak'])"
done
GWB200_OPTIONS=pdl=1 timeout 300 python bench.py --workload ddim50 --batch 1024 --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ddim50', round(d['value'],1), d['chain']['frac_of_sustained_bf16_peak'])"
