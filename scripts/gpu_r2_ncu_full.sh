#!/bin/bash
# `ncu --set full` of the kernels of one C4 step (one launch each).  Usage: gpurun --timeout 900 -- bash scripts/gpu_r2_ncu_full.sh <tag>
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r2}
export SPMF_GRAPHS=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_timed/" \
    -k regex:'hot_tile_kernel|csr_rows_|csc_cols_kernel|umma_gemm3|backward_dk|gamma_kernel|draw_operands|backward_feat|fill_normal|adam_kernel|rows_finish' -c 12 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_$TAG.ncu-rep
