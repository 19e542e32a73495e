#!/bin/bash
# Other workloads and the K sweep (BASELINE.json configs[1], [2], [4]) -- parity-test shapes at scale,
# not the headline bench line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for args in "--workload c4 --K 2" "--workload c4 --K 8" "--workload c4 --K 128" "--workload c2" "--workload c3" "--workload c1"; do
  tag=$(echo $args | tr -d ' -')
  timeout 600 python bench.py $args --steps 16 --warmup 4 --no-cpu-baseline > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  echo "$args rc=$?"; tail -2 gpurun_out/sweep_$tag.err
done
