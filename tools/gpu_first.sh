#!/bin/bash
# First-contact GPU run: per-layer diagnostics for each kernel family, then the gpu test-suite.  Each stage has its own
# timeout so one hang cannot eat the budget.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
export CUDA_LAUNCH_BLOCKING=1
for mode in fp32 bf16_simt bf16_tc; do
  timeout 180 python tools/gpu_diag.py --mode $mode --L 1024 --B 2 --cin 3 --json gpurun_out/diag.jsonl > gpurun_out/diag_$mode.log 2>&1
  echo "diag $mode exit $?" >> gpurun_out/stages.txt
done
timeout 180 python tools/gpu_diag.py --mode bf16_tc --L 4096 --B 3 --cin 7 --json gpurun_out/diag.jsonl > gpurun_out/diag_tc_c7.log 2>&1
echo "diag tc c7 exit $?" >> gpurun_out/stages.txt
unset CUDA_LAUNCH_BLOCKING
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/stages.txt
tail -5 gpurun_out/pytest.log
cat gpurun_out/stages.txt
