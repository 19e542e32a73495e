#!/bin/bash
# Multi-GPU checks (round 2).  Usage: gpurun --gpus N --timeout 900 -- bash scripts/gpu_r2_multi.sh N
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py > gpurun_out/dp_check_r2_$N.log 2>&1
grep -n "dp_check\|DP_CHECK\|Error\|error\|assert" gpurun_out/dp_check_r2_$N.log | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/scale_r2_$N.json 2> gpurun_out/scale_r2_$N.err
echo "gpus=$N rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_r2_$N.json").read().strip().splitlines()[-1])
    print("N=%d ms/step %.4f value %.4e e2e ms %.4f e2e %.4e" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/scale_r2_$N.err
