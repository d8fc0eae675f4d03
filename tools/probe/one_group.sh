# single-group epilogue flavour (runtime-selected): parity subset + B=8 / B=256 chains with the flavour allowed / forbidden
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_chain.py tests/test_gpu_bench_configs.py -q -x 2>&1 | tail -3
for O in "one_group=1" "one_group=0"; do
  GWB200_OPTIONS="$O" python bench.py --workload ddpm1000 --batch 8 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=8 [$O]', round(d['value'],2), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
GWB200_OPTIONS="one_group=1" python bench.py --workload ddpm1000 --batch 16 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=16 [one_group=1]', round(d['value'],2), round(d['ms_per_step'],2))"
GWB200_OPTIONS="one_group=0" python bench.py --workload ddpm1000 --batch 16 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=16 [one_group=0]', round(d['value'],2), round(d['ms_per_step'],2))"
