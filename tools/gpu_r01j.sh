#!/bin/bash
# Round-1 evidence run (session j: head dots in the last decoder, specialised GroupNorm kernels): parity tests, bench lines, per-launch profile, ncu launch
# list with DRAM traffic for one reverse step.  Usage: bash tools/gpu_r01j.sh
cd "$(dirname "$0")/.."
TAG=r01j
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
timeout 900 python bench.py > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench train exit $?" >> $S
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_train_reference_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench reference exit $?" >> $S
timeout 600 python bench.py --workload ddim50 --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_ddim50_$TAG.json 2> gpurun_out/bench_ddim50_$TAG.err; echo "bench ddim50 exit $?" >> $S
timeout 900 python bench.py --workload ddpm1000 --batch 64 --length 16384 --steps 2 --warmup 3 > gpurun_out/bench_L16384_$TAG.json 2> gpurun_out/bench_L16384_$TAG.err; echo "bench L16384 exit $?" >> $S
timeout 300 python tools/step_profile.py --B 256 --steps 8 > gpurun_out/step_$TAG.txt 2>&1; echo "step profile exit $?" >> $S
timeout 300 python tools/train_profile.py > gpurun_out/train_step_$TAG.txt 2>&1; echo "train profile exit $?" >> $S
CMD="python tools/step_profile.py --B 256 --steps 2"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/traffic_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu traffic list exit $?" >> $S
cat $S
tail -3 gpurun_out/pytest_$TAG.log
tail -1 gpurun_out/step_$TAG.txt
cut -c1-400 gpurun_out/bench_train_$TAG.json
