#!/bin/bash
# quick knob check on the final binary: tile-kernel waves and hot-column threshold
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 24 --warmup 4 --no-cpu-baseline $EXTRA > gpurun_out/knob_$tag.json 2> gpurun_out/knob_$tag.err
  python - <<PY
import json
for l in open("gpurun_out/knob_$tag.json"):
    if l.startswith("{"):
        j=json.loads(l); print("$tag", "ms/step %.4f" % j["ms_per_step"], "e2e %.4f" % j["e2e"]["ms_per_step"])
PY
}
run w4 SPMF_TILE_WAVES=4
run w8 SPMF_TILE_WAVES=8
EXTRA="--hot-density 0.04" run hd4 SPMF_TILE_WAVES=6
