#!/bin/bash
# N GPUs: data-parallel parity through the peer-memory exchange kernel and through NCCL, then both bench lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tests/dp_check.py > gpurun_out/dp_check_p2p_$N.log 2>&1; echo "dp_check p2p rc=$?"; grep "dp_check\|DP_CHECK\|Error\|error" gpurun_out/dp_check_p2p_$N.log | tail -8
SPMF_P2P=0 timeout 300 $RUN --master-port 29512 tests/dp_check.py > gpurun_out/dp_check_nccl_$N.log 2>&1; echo "dp_check nccl rc=$?"; grep "dp_check\|DP_CHECK" gpurun_out/dp_check_nccl_$N.log | tail -6
for P in 1 0; do
SPMF_P2P=$P timeout 300 $RUN --master-port 2952$P bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_p2p${P}_$N.json 2> gpurun_out/bench_p2p${P}_$N.err
echo "bench P2P=$P rc=$?"
tail -1 gpurun_out/bench_p2p${P}_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('P2P=$P N=%d ms/step %.4f value %.4e e2e ms %.4f  %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['parallelism']))" || tail -5 gpurun_out/bench_p2p${P}_$N.err
done
