#!/bin/bash
# ncu --set full of the per-batch preparation kernels (hot split, CSC build) inside the e2e region
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-p0}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_e2e/" \
    -k regex:"${KREGEX:-hot_split_kernel|csc_block_hist|csc_block_scatter}" -c ${NCAP:-3} -f -o gpurun_out/prep_$TAG $CMD > gpurun_out/ncu_prep.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_prep.log
