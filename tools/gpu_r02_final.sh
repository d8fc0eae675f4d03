#!/bin/bash
# final-state check (1 GPU): full parity suite, smoke, default bench line, reference arm
cd "$(dirname "$0")/.."
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "reference exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_train_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks")}); print(d["sampling"]["value"], d["sampling"]["chain"]["frac_of_sustained_bf16_peak"])
r=json.loads(open("gpurun_out/bench_reference_$TAG.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["kind"])
PY
