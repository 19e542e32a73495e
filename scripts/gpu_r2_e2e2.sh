#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_parity.py -x -q -m gpu -k "u8 or streamed or hybrid_fit or dense_ingest or split" --timeout 200 2>&1 | tail -4
for W in c4 c4 c3 c2 c1; do
timeout 200 python bench.py --workload $W --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$W ms/step %.4f e2e ms %.4f h2d %.1f MB fmt %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']/1e6, d['e2e'].get('host_format')))"
done
