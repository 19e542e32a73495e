#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_parity.py -x -q -m gpu -k "u8 or streamed or hybrid_fit or dense_ingest" --timeout 200 2>&1 | tail -8
for F in 0 1; do
SPMF_FUSED_UNPACK8=$F timeout 200 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fused=$F ms/step %.4f e2e ms %.4f h2d %.1f MB' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']/1e6))"
done
BENCH_DEBUG=1 timeout 200 python bench.py --workload c3 --steps 12 --warmup 4 --no-cpu-baseline 2>&1 | grep "e2e iter" | tail -8
