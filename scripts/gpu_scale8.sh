#!/bin/bash
# 8-GPU weak-scaling point + data-parallel parity at world 8.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py 2>&1 | grep -E "dp_check|DP_CHECK|rror" | tee gpurun_out/dp_check_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "gpus=$N rc=$?"; tail -c 400 gpurun_out/scale_$N.json; tail -3 gpurun_out/scale_$N.err
