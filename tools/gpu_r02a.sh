#!/bin/bash
# Round-2 first run: full GPU parity suite (incl. the new benchmarked-config tests), fused-kernel phase timelines, bench line.
cd "$(dirname "$0")/.."
TAG=r02a
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc >> gpurun_out/smi_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
(cd tools && for l in 1 2 3 4 5 6; do timeout 100 python cgn_timeline.py --layer $l; done) > gpurun_out/timeline_$TAG.txt 2>&1; echo "timeline exit $?" >> $S
timeout 300 python tools/step_profile.py --B 256 --steps 8 > gpurun_out/step_$TAG.txt 2>&1; echo "step profile exit $?" >> $S
timeout 900 python bench.py > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench train exit $?" >> $S
cat $S
tail -15 gpurun_out/pytest_$TAG.log
tail -12 gpurun_out/step_$TAG.txt
cut -c1-300 gpurun_out/bench_train_$TAG.json
