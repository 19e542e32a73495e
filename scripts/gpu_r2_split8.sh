#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_parity.py -x -q -m gpu -k "u8 or streamed or hybrid_fit" --timeout 200 2>&1 | tail -4
for F in 1 0 0; do
SPMF_SPLIT8_TWO_PASS=$F timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('two_pass=$F ms/step %.4f e2e ms %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
