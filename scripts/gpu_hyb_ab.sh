#!/bin/bash
# hybrid tests then A/B bench + launch list
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-h1}
for hd in ${HDS:-0.03 0.015 0}; do
  SPMF_HOT_MODE=${HOTMODE:-2} timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --hot-density $hd > gpurun_out/ab_${TAG}_$hd.json 2> gpurun_out/ab_${TAG}_$hd.err
  echo "hot-density $hd rc=$?"; python - <<PY
import json
for l in open("gpurun_out/ab_${TAG}_$hd.json"):
    if l.startswith("{"):
        j=json.loads(l); r=j["roofline"]
        print("ms/step", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "kern", {k:round(v["ms"],4) for k,v in r["kernels"].items()}, "gemm", r.get("umma_gemm_gradA"), "H", r.get("hot_cols"), "loss", j["config"]["final_loss"])
PY
  tail -3 gpurun_out/ab_${TAG}_$hd.err
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --hot-density ${NCU_HD:-0.03}"
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_timed/" --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
