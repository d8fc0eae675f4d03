#!/bin/bash
# full GPU suite + the B=8 DDPM-1000 chain with and without layer chaining
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02g_pytest.txt
cat gpurun_out/r02g_pytest.txt
python bench.py --workload ddpm1000 --batch 8 --steps 3 --warmup 3 > gpurun_out/r02g_b8_plain.json 2> gpurun_out/r02g_b8_plain.err
GWB200_CHAIN=1 python bench.py --workload ddpm1000 --batch 8 --steps 3 --warmup 3 > gpurun_out/r02g_b8_chain.json 2> gpurun_out/r02g_b8_chain.err
python - <<'PY'
import json
for f in ["plain", "chain"]:
    try:
        d = json.loads(open(f"gpurun_out/r02g_b8_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
