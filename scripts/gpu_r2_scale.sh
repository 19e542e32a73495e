#!/bin/bash
# bench at N GPUs (round 2).  Usage: gpurun --gpus N --timeout 600 -- bash scripts/gpu_r2_scale.sh N
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/scale_r2_$N.json 2> gpurun_out/scale_r2_$N.err
echo "gpus=$N rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_r2_$N.json").read().strip().splitlines()[-1])
    print("N=%d ms/step %.4f value %.4e e2e ms %.4f e2e %.4e graphs %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("graph_replays")))
except Exception as e:
    print("parse failed", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/scale_r2_$N.err | tail -5
