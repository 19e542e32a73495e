#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_graphs.py tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "traject or graph or fit or adam or two_gpus" 2>&1 | tail -5
timeout 300 $RUN --master-port 29511 tests/dp_check.py > gpurun_out/dp_check_early_$N.log 2>&1; echo "dp_check rc=$?"; grep "dp_check\|DP_CHECK\|Error" gpurun_out/dp_check_early_$N.log | tail -6
for E in 1 0; do
SPMF_ADAM_EARLY=$E timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('early=$E N=1 ms/step %.4f e2e ms %.4f launches %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches']))"
SPMF_ADAM_EARLY=$E timeout 300 $RUN --master-port 2952$E bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('early=$E N=%d ms/step %.4f e2e ms %.4f  %s' % (d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['parallelism']))"
done
