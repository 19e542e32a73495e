#!/bin/bash
# guard tests + bench ablations + per-kernel launch list.  Usage: gpurun --timeout 1200 -- bash scripts/gpu_r2_ablate.sh <tag>
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 300 python -m pytest tests/test_gpu_links.py -q --timeout 120 -k guard 2>&1 | tail -3
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"],
          "tile %.3f rows %.3f cols %.3f" % (r.get("kernel_ms") or 0, r["kernels"]["csr_rows"]["ms"], r["kernels"]["csc_cols"]["ms"]))
except Exception as e:
    print("bench parse failed", sys.argv[1], e)
PY
}
timeout 200 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}.err; summ gpurun_out/bench_${TAG}_default.json
SPMF_EXACT_GUARD=0 timeout 200 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_${TAG}_noguard.json 2>> gpurun_out/bench_${TAG}.err; summ gpurun_out/bench_${TAG}_noguard.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_timed/" --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
python scripts/launch_summary.py gpurun_out/launches_$TAG.csv 3 gpurun_out/launch_summary_$TAG.txt | cut -c1-60,97-200 | tail -40
