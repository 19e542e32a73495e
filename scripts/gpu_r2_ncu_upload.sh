#!/bin/bash
# `ncu --set full` of the upload-side kernels of the end-to-end leg (one launch each)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export SPMF_GRAPHS=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_e2e/" \
    -k regex:'hot_split_kernel|csc_block_hist|csc_block_scatter|csc_block_scan|exscan_int' -c 5 -f -o gpurun_out/prof_upload $CMD > gpurun_out/ncu_upload.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_upload.log; ls -la gpurun_out/prof_upload.ncu-rep
