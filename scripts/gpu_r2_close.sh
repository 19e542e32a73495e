#!/bin/bash
# last check of the closing tree: whole GPU suite, smoke(), default bench line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_close.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_close.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_close.json 2> gpurun_out/bench_close.err
echo "bench rc=$?"; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_close.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "traffic", d["roofline"]["traffic"], "frac", d["roofline"]["frac"])
PY
