#!/bin/bash
# N-GPU bench line of the final configuration (peer-memory exchange kernel + early Adam of the data-free tensors)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final_$N.json 2> gpurun_out/bench_final_$N.err
echo "rc=$?"; tail -1 gpurun_out/bench_final_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=%d ms/step %.4f value %.4e e2e ms %.4f e2e %.4e %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['config']['parallelism']))" || tail -5 gpurun_out/bench_final_$N.err
