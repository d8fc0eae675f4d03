for C in "0" "1" "0" "1"; do
  GWB200_FUSE_GN_TRAIN=$C python bench.py --steps 100 --warmup 10 --no-sampling 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse_gn_train=$C', round(d['value'],1), round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
