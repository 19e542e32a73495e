#!/bin/bash
# e2e-focused run: bench + launch list of the streamed (host-resident) loop
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-e2e}
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_e2e/" --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
