#!/bin/bash
# Parity tests + bench + ncu evidence on one B200.
# Usage: gpurun --timeout 1500 -- bash scripts/gpu_bench.sh <tag> [full]
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_timed/" --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
if [ "$2" == "full" ]; then
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_timed/" \
    -k regex:'csr_rows_kernel|csc_cols_kernel|backward_dk_kernel|backward_feat_kernel|gamma_grad_kernel' -c 10 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
