#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for F in 1 0 1 0 1 0; do
SPMF_SPLIT8_TWO_PASS=$F timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('two_pass=$F ms/step %.4f e2e ms %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
SPMF_GRAPHS=0 $CMD > gpurun_out/plain.log 2>&1 &&
SPMF_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_e2e/" -k regex:'hot_split' --csv \
    --log-file gpurun_out/launches_split8.csv $CMD > gpurun_out/ncu_list.log 2>&1
grep -c hot_split gpurun_out/launches_split8.csv; grep hot_split gpurun_out/launches_split8.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -4
