#!/bin/bash
# Round 2: whole GPU suite + bench on one B200.  Usage: gpurun --timeout 1500 -- bash scripts/gpu_r2_full.sh <tag>
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1100 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    print("tile ms", r.get("kernel_ms"), "frac", r.get("frac"), "rows ms", r["kernels"]["csr_rows"]["ms"], "cols ms", r["kernels"]["csc_cols"]["ms"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_$TAG.err
