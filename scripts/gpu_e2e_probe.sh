#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-e0}
python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/e2e_$TAG.json 2> gpurun_out/e2e_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/e2e_$TAG.json"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]
        print("ms/step", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "prep ms", r.get("hot_prepare_ms_per_batch"))
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_e2e/" --csv \
    --log-file gpurun_out/launches_e2e_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
