#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err
  python - <<PY
import json
for l in open("gpurun_out/diag_$name.json"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]; print("$name ms/step %.4f"%j["ms_per_step"], "tile %.4f"%r.get("kernel_ms",0), {k:round(v["ms"],3) for k,v in r["kernels"].items()})
PY
}
run default X=1
run nosplit SPMF_SPLIT_BACKWARD=0
run seq SPMF_STREAMS=seq
env X=1 timeout 600 python bench.py --gpus 1 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/diag_one.json 2>/dev/null
python - <<PY
import json
for l in open("gpurun_out/diag_one.json"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]; print("one-gpu-on-2gpu-box ms/step %.4f"%j["ms_per_step"], "tile %.4f"%r.get("kernel_ms",0), {k:round(v["ms"],3) for k,v in r["kernels"].items()})
PY
