#!/bin/bash
# Iteration run: conv variant sweep, parity tests, short bench.  Usage: bash tools/gpu_iter.sh tag [variant]
cd "$(dirname "$0")/.."
TAG=${1:-it}
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --B 64 --L 4096 --variants 0,2,6,14 --json gpurun_out/conv_small_$TAG.json > gpurun_out/conv_small_$TAG.log 2>&1
echo "conv_bench small exit $?" >> gpurun_out/stages_$TAG.txt
timeout 300 python tools/conv_bench.py --B 256 --L 4096 --variants 0,1,2,6,14 --json gpurun_out/conv_$TAG.json > gpurun_out/conv_$TAG.log 2>&1
echo "conv_bench exit $?" >> gpurun_out/stages_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/stages_$TAG.txt
timeout 600 python bench.py --workload ddim50 --batch 256 --steps 3 --warmup 3 > gpurun_out/bench_ddim50_$TAG.json 2> gpurun_out/bench_ddim50_$TAG.err
echo "bench exit $?" >> gpurun_out/stages_$TAG.txt
cat gpurun_out/stages_$TAG.txt; cat gpurun_out/conv_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
