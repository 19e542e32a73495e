#!/bin/bash
# Bench + ncu evidence on one B200.  Usage: gpurun --timeout 1500 -- bash scripts/gpu_bench.sh [variant]
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
V=${1:-0}
set -o pipefail
timeout 600 python bench.py --steps 32 --warmup 4 --variant $V > gpurun_out/bench_v$V.json 2> gpurun_out/bench_v$V.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_v$V.json; tail -5 gpurun_out/bench_v$V.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --variant $V"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v$V.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'csr_rows_kernel|csc_cols_kernel' -s 8 -c 4 -f -o gpurun_out/prof_v$V $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
