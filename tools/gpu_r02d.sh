#!/bin/bash
# Round-2 run d: analytic-statistics first block.
cd "$(dirname "$0")/.."
TAG=${1:-r02d}
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_bench_configs.py tests/test_gpu_chain.py -m gpu -q --maxfail=12 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
timeout 200 python tools/step_profile.py --B 256 --steps 8 > gpurun_out/step_$TAG.txt 2>&1; echo "step profile exit $?" >> $S
timeout 200 python tools/step_profile.py --B 256 --steps 8 --cin 7 > gpurun_out/step_c7_$TAG.txt 2>&1; echo "step profile c7 exit $?" >> $S
cat $S
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_$TAG.log | tail -20
tail -12 gpurun_out/step_$TAG.txt
tail -12 gpurun_out/step_c7_$TAG.txt
