#!/bin/bash
# Round-2 evidence run (1 GPU): parity suite, bench lines of BASELINE configs 1, 2, 4 + whiten / score workloads, per-launch
# profiles, ncu launch list with DRAM traffic, one ncu --set full capture of the dominant fused kernel.
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
timeout 600 python bench.py > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench train exit $?" >> $S
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_train_reference_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench reference exit $?" >> $S
timeout 300 python bench.py --workload ddpm1000 --batch 8 --steps 2 --warmup 3 > gpurun_out/bench_ddpm1000_B8_$TAG.json 2> gpurun_out/bench_ddpm1000_B8_$TAG.err; echo "bench config1 (B=8) exit $?" >> $S
timeout 300 python bench.py --workload ddim50 --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_ddim50_$TAG.json 2> gpurun_out/bench_ddim50_$TAG.err; echo "bench ddim50 exit $?" >> $S
timeout 600 python bench.py --workload ddpm1000 --batch 64 --length 16384 --steps 2 --warmup 3 > gpurun_out/bench_L16384_$TAG.json 2> gpurun_out/bench_L16384_$TAG.err; echo "bench L16384 exit $?" >> $S
timeout 200 python bench.py --workload whiten --batch 1024 --steps 20 --warmup 3 > gpurun_out/bench_whiten_$TAG.json 2> gpurun_out/bench_whiten_$TAG.err; echo "bench whiten exit $?" >> $S
timeout 200 python bench.py --workload score --batch 8192 --steps 20 --warmup 3 > gpurun_out/bench_score_$TAG.json 2> gpurun_out/bench_score_$TAG.err; echo "bench score exit $?" >> $S
timeout 200 python tools/step_profile.py --B 256 --steps 8 > gpurun_out/step_$TAG.txt 2>&1; echo "step profile exit $?" >> $S
timeout 200 python tools/train_profile.py > gpurun_out/train_step_$TAG.txt 2>&1; echo "train profile exit $?" >> $S
CMD="python tools/step_profile.py --B 256 --steps 2"
timeout 200 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/traffic_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu traffic list exit $?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:conv_gn_kernel --launch-skip 9 --launch-count 3 -o gpurun_out/prof_cgn_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?" >> $S
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $S
BCMD="python bench.py --steps 2 --warmup 3 --no-sampling"
timeout 300 $BCMD > gpurun_out/plain_bench_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_train_$TAG.csv $BCMD > gpurun_out/ncu_bench_list_$TAG.log 2>&1
echo "ncu bench launch list exit $?" >> $S
cat $S
grep -E "^(FAILED|ERROR)|passed|failed|non-default arch|ddpm1000 |L16384 ddpm" gpurun_out/pytest_$TAG.log | tail -12
for f in bench_train bench_ddpm1000_B8 bench_ddim50 bench_L16384 bench_whiten bench_score; do echo "== $f"; cut -c1-260 gpurun_out/${f}_$TAG.json; done
