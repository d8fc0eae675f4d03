#!/bin/bash
# Round-2 run c: full GPU parity suite after the optimiser / graph / plan-cache changes.
cd "$(dirname "$0")/.."
TAG=${1:-r02c}
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
cat $S
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_$TAG.log | tail -20
