#!/bin/bash
# ncu launch list (durations + instruction counts) of one reverse step, first-block kernels in focus
cd "$(dirname "$0")/.."
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_forward.py -m gpu -q -k "direct_first" > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
for CIN in 3 7; do
CMD="python tools/step_profile.py --B 256 --steps 2 --cin $CIN"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/ncu_list_c${CIN}_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/ncu_list_c${CIN}_$TAG.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; ik=H.index("Kernel Name"); im=H.index("Metric Name"); iv=H.index("Metric Value"); iid=H.index("ID")
d={}
for r in rows[hdr+1:]:
    if len(r)<=iv: continue
    d.setdefault((int(r[iid]),r[ik][:60]),{})[r[im]]=r[iv]
for (i,k),m in sorted(d.items())[-24:]:
    print(i,k,m.get("gpu__time_duration.sum"),m.get("smsp__inst_executed.sum"),m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),m.get("dram__bytes_read.sum"),m.get("dram__bytes_write.sum"))
PY
done
