#!/bin/bash
# Round-2 closing evidence on one B200: whole GPU suite, the default bench line, the reference arm, the
# workload / latent-dim sweep, and the ncu launch list of the timed region (after the same command ran clean).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r2_final}
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    print("tile ms", r.get("kernel_ms"), "frac", r.get("frac"), "rows ms", r["kernels"]["csr_rows"]["ms"], "cols ms", r["kernels"]["csc_cols"]["ms"], "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "reference arm rc=$?"; tail -c 400 gpurun_out/bench_ref_$TAG.json; echo
bash scripts/gpu_r2_sweep.sh $TAG 2>&1 | tail -12
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
SPMF_GRAPHS=0 $CMD > gpurun_out/plain.log 2>&1 &&
SPMF_GRAPHS=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_timed/" --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_$TAG.csv 3 gpurun_out/launch_summary_$TAG.txt | cut -c1-60,97-200 | tail -34
