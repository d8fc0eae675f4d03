#!/bin/bash
# Round-2 run f: layer chaining (gw_conv_gn3): parity + chain throughput with and without chaining
cd "$(dirname "$0")/.."
TAG=${1:-r02f}
mkdir -p gpurun_out
S=gpurun_out/stages_$TAG.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_bench_configs.py -m gpu -q --maxfail=6 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $S
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_$TAG.log | tail -8
timeout 300 python bench.py --workload ddpm1000 --batch 256 --steps 2 --warmup 3 > gpurun_out/bench_ddpm_$TAG.json 2> gpurun_out/bench_ddpm_$TAG.err; echo "bench chain exit $?" >> $S
GWB200_CHAIN=0 timeout 300 python bench.py --workload ddpm1000 --batch 256 --steps 2 --warmup 3 > gpurun_out/bench_ddpm_nochain_$TAG.json 2> gpurun_out/bench_ddpm_nochain_$TAG.err; echo "bench nochain exit $?" >> $S
cat $S
python - <<PY
import json
for f in ("gpurun_out/bench_ddpm_$TAG.json","gpurun_out/bench_ddpm_nochain_$TAG.json"):
    try:
        d=json.load(open(f)); print(f, d["value"], d["chain"]["ms_per_reverse_step"], d["chain"]["frac_of_sustained_bf16_peak"], d["e2e"]["value"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_ddpm_$TAG.err
