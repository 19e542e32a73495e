#!/bin/bash
# Round-end evidence on one B200: bench (ours + reference arm), launch list, ncu --set full of the hot kernels.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "ref rc=$?"; tail -c 400 gpurun_out/bench_ref_$TAG.json
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --hot-density 0 > gpurun_out/bench_gather_$TAG.json 2> gpurun_out/bench_gather_$TAG.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_timed/" --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_timed/" \
    -k regex:'hot_tile_kernel|csr_rows_|csc_cols_kernel|umma_gemm3_kernel|backward_dk_kernel|gamma_kernel' -c 7 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
