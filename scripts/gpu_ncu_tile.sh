#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-t3}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --hot-density 0.03"
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_timed/" \
    -k regex:"${KREGEX:-hot_tile_kernel}" -c ${NCAP:-1} -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
