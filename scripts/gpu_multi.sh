#!/bin/bash
# Multi-GPU checks.  Usage: gpurun --gpus N --timeout 1200 -- bash scripts/gpu_multi.sh N
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py 2>&1 | tail -8 | tee gpurun_out/dp_check_$N.log
for G in 1 $N; do
  if [ $G == 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
  fi
  echo "gpus=$G rc=$?"; tail -c 1200 gpurun_out/scale_$G.json; tail -3 gpurun_out/scale_$G.err
done
SPMF_STREAMS=prio timeout 600 python bench.py --gpus 1 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/prio_1.json 2> gpurun_out/prio_1.err; tail -c 600 gpurun_out/prio_1.json
