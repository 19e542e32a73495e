#!/bin/bash
# ncu --set full of every kernel of ONE timed step (survey: issue utilisation, lane efficiency, stalls)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-a0}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --nvtx --nvtx-include "spmf_timed/" -c ${NCAP:-40} -f -o gpurun_out/all_$TAG $CMD > gpurun_out/ncu_all.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_all.log | cut -c1-200
