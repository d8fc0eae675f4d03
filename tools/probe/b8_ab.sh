for C in "0:" "1:" "0:pair2=1" "0:"; do
  DF=${C%%:*}; O=${C#*:}
  GWB200_DIRECT_FIRST=$DF GWB200_OPTIONS="$O" python bench.py --workload ddpm1000 --batch 8 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('direct_first=$DF opts=[$O]', round(d['value'],2), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
