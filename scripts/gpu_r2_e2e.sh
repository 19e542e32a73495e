#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -k "streamed or hybrid_fit" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_hybrid.py -q --timeout 200 -k "fit_reduces" 2>&1 | tail -3
for F in u16 u8; do
BENCH_HOST_FORMAT=$F timeout 200 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$F ms/step %.4f e2e ms %.4f h2d %.1f MB  tile %.3f rows %.3f cols %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']/1e6, r.get('kernel_ms') or 0, r['kernels']['csr_rows']['ms'], r['kernels']['csc_cols']['ms']))"
done
