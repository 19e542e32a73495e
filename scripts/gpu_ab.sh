#!/bin/bash
# A/B of the hybrid (tensor-core hot block) step against the gather-only step + launch list.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-ab0}
for hd in 0.03 0; do
  timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --hot-density $hd > gpurun_out/ab_${TAG}_$hd.json 2> gpurun_out/ab_${TAG}_$hd.err
  echo "hot-density $hd rc=$?"; tail -c 1800 gpurun_out/ab_${TAG}_$hd.json; tail -3 gpurun_out/ab_${TAG}_$hd.err
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_timed/" --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
