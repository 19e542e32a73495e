#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for HD in 0.03 0.02 0.01; do
timeout 200 python bench.py --workload c4 --K 128 --hot-density $HD --steps 24 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('K=128 hd=$HD ms/step %.4f value %.3e tile %.3f H=%s rows %.3f cols %.3f' % (d['ms_per_step'], d['value'], r.get('kernel_ms') or 0, r.get('hot_cols'), r['kernels']['csr_rows']['ms'], r['kernels']['csc_cols']['ms']))"
done
CMD="python bench.py --workload c4 --K 128 --steps 3 --warmup 3 --no-cpu-baseline"
SPMF_GRAPHS=0 $CMD > gpurun_out/plain.log 2>&1 &&
SPMF_GRAPHS=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_timed/" --csv \
    --log-file gpurun_out/launches_k128.csv $CMD > gpurun_out/ncu_list.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_k128.csv 3 gpurun_out/launch_summary_k128.txt | cut -c1-60,97-200 | tail -32
