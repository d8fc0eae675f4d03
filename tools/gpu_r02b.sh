#!/bin/bash
# Round-2 run b: CTA-pair (cta_group::2) fused kernels + device-side optimiser schedule / single-graph step.
cd "$(dirname "$0")/.."
TAG=r02b
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
timeout 300 python -m pytest tests/test_gpu_forward.py -m gpu -q -x -k "cta_pair or fused" > gpurun_out/pytest_pair_$TAG.log 2>&1; echo "pytest pair exit $?" >> $S
tail -5 gpurun_out/pytest_pair_$TAG.log
timeout 200 python tools/step_profile.py --B 256 --steps 8 > gpurun_out/step_$TAG.txt 2>&1; echo "step profile exit $?" >> $S
GWB200_OPTIONS="pair2=0" timeout 200 python tools/step_profile.py --B 256 --steps 8 > gpurun_out/step_nopair_$TAG.txt 2>&1; echo "step profile nopair exit $?" >> $S
(cd tools && for l in 1 3 4 5 6; do timeout 100 python cgn_timeline.py --layer $l --brief; done) > gpurun_out/timeline_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
timeout 600 python bench.py --steps 50 > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench train exit $?" >> $S
cat $S
tail -12 gpurun_out/step_$TAG.txt
tail -3 gpurun_out/step_nopair_$TAG.txt
cat gpurun_out/timeline_$TAG.txt
tail -15 gpurun_out/pytest_$TAG.log
cut -c1-300 gpurun_out/bench_train_$TAG.json
