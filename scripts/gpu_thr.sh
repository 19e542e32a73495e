#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for hd in ${HDS:-0.08 0.10 0.15}; do
  timeout 600 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --hot-density $hd > gpurun_out/thr_$hd.json 2> gpurun_out/thr_$hd.err
  python - <<PY
import json
for l in open("gpurun_out/thr_$hd.json"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]
        print("hd $hd H", r["hot_cols"], "ms/step %.4f"%j["ms_per_step"], "e2e ms %.4f"%j["e2e"]["ms_per_step"], "tile ms %.4f"%r.get("kernel_ms",0), {k:round(v["ms"],3) for k,v in r["kernels"].items()})
PY
done
