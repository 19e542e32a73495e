#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for V in "" skip_prep; do
SPMF_DIAG_UPLOAD=$V timeout 200 python bench.py --workload c4 --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('diag=$V ms/step %.4f e2e ms %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
