#!/bin/bash
# does host-side CPU contention inflate the event-bracketed kernel times?
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nproc
run() {
  name=$1
  timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/hj_$name.json 2>/dev/null
  python - <<PY
import json
for l in open("gpurun_out/hj_$name.json"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]; print("$name ms/step %.4f e2e %.4f"%(j["ms_per_step"], j["e2e"]["ms_per_step"]), "tile %.4f"%r.get("kernel_ms",0), {k:round(v["ms"],3) for k,v in r["kernels"].items()})
PY
}
run quiet
pids=""
for i in $(seq 1 $(nproc)); do ( while :; do :; done ) & pids="$pids $!"; done
run burn
kill $pids
