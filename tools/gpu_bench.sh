#!/bin/bash
# Bench + ncu evidence on one B200.  Usage: bash tools/gpu_bench.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ddpm1000_$TAG.json 2> gpurun_out/bench_ddpm1000_$TAG.err
echo "bench ddpm1000 exit $?" >> gpurun_out/stages_$TAG.txt
timeout 600 python bench.py --workload ddim50 --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_ddim50_$TAG.json 2> gpurun_out/bench_ddim50_$TAG.err
echo "bench ddim50 exit $?" >> gpurun_out/stages_$TAG.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
echo "bench reference exit $?" >> gpurun_out/stages_$TAG.txt
# ncu: launch list of a short bench command (plain run first, as the recipe requires), then one full capture
CMD="python bench.py --workload ddim50 --batch 256 --steps 1 --warmup 3 --steps-per-graph 1"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 340 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?" >> gpurun_out/stages_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel|gn_apply_kernel|conv_in_kernel|final_step_kernel' -s 1700 -c 17 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?" >> gpurun_out/stages_$TAG.txt
cat gpurun_out/stages_$TAG.txt
cat gpurun_out/bench_ddpm1000_$TAG.json | cut -c1-1500
