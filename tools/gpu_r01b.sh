#!/bin/bash
# Round-1 evidence run (conv_tc v2): parity tests, per-launch step profile at several batch sizes (L2-residency
# experiment), bench lines, ncu launch list + one full capture.  Usage: bash tools/gpu_r01b.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r01b}
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
for B in 32 64 128 256; do
  timeout 300 python tools/step_profile.py --B $B --steps 6 --json gpurun_out/step_${TAG}_B$B.json > gpurun_out/step_${TAG}_B$B.log 2>&1; echo "step B=$B exit $?" >> $S
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ddpm1000_$TAG.json 2> gpurun_out/bench_ddpm1000_$TAG.err; echo "bench ddpm1000 exit $?" >> $S
timeout 600 python bench.py --workload ddim50 --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_ddim50_$TAG.json 2> gpurun_out/bench_ddim50_$TAG.err; echo "bench ddim50 exit $?" >> $S
CMD="python tools/step_profile.py --B 256 --steps 2"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2_kernel|gn_apply_kernel|conv_in_kernel|final_step_kernel' -s 45 -c 15 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?" >> $S
cat $S
for B in 32 64 128 256; do tail -1 gpurun_out/step_${TAG}_B$B.log; done
cut -c1-600 gpurun_out/bench_ddpm1000_$TAG.json
