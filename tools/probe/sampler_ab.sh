for C in "0:" "0:pair2=1" "1:pair2=1" "1:" "0:" "0:pair2=1"; do
  CH=${C%%:*}; O=${C#*:}
  GWB200_CHAIN=$CH GWB200_OPTIONS="$O" python bench.py --workload ddpm1000 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chain=$CH opts=[$O]', round(d['value'],2), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
