#!/bin/bash
# Workload / latent-dim sweep + peak micro-benchmarks on one B200 -> gpurun_out/sweep_<tag>.jsonl
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r2}
OUT=gpurun_out/sweep_$TAG.jsonl
: > $OUT
./scripts/_build/peaks_probe | tee gpurun_out/peaks_$TAG.json
run() { timeout 240 python bench.py --steps 32 --warmup 4 --no-cpu-baseline "$@" 2>> gpurun_out/sweep_$TAG.err | tail -1 >> $OUT; }
run --workload c4
run --workload c1
run --workload c2
run --workload c3
run --workload c4 --K 2
run --workload c4 --K 8
run --workload c4 --K 64
run --workload c4 --K 128
python - <<PY
import json
for l in open("$OUT"):
    try:
        d = json.loads(l)
        print("%-70s ms/step %.4f  value %.3e  e2e %.3e" % (d["config"]["workload"][:70], d["ms_per_step"], d["value"], d["e2e"]["value"]))
    except Exception as e:
        print("bad line", e, l[:80])
PY
tail -3 gpurun_out/sweep_$TAG.err
