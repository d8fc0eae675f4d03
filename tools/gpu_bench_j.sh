#!/bin/bash
# bench lines only (session j).  Usage: bash tools/gpu_bench_j.sh
cd "$(dirname "$0")/.."
TAG=r01j
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench train exit $?"
timeout 600 python bench.py --workload ddim50 --batch 1024 --steps 3 --warmup 3 > gpurun_out/bench_ddim50_$TAG.json 2> gpurun_out/bench_ddim50_$TAG.err; echo "bench ddim50 exit $?"
timeout 900 python bench.py --workload ddpm1000 --batch 64 --length 16384 --steps 2 --warmup 3 > gpurun_out/bench_L16384_$TAG.json 2> gpurun_out/bench_L16384_$TAG.err; echo "bench L16384 exit $?"
timeout 900 python bench.py --workload ddpm1000 --batch 1024 --steps 2 --warmup 3 > gpurun_out/bench_ddpm1000_B1024_$TAG.json 2> gpurun_out/bench_ddpm1000_B1024_$TAG.err; echo "bench ddpm1000 B1024 exit $?"
for f in train ddim50 L16384 ddpm1000_B1024; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${f}_$TAG.json").read().strip().splitlines()[-1])
    s=d.get("sampling") or {}
    print("${f}", d["metric"], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "ms/step", round(d.get("ms_per_step",0),3), "roof", round(d["roofline"]["frac"],3), (d.get("chain") or s.get("chain") or {}).get("frac_of_sustained_bf16_peak"), s.get("value"))
except Exception as e: print("${f}", "ERR", e)
PY
done
