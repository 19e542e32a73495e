#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python bench.py --workload ${WL:-c3} --steps 40 --warmup 24 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.4f graph_replays %s launches %s' % (d['ms_per_step'], d.get('graph_replays'), d['gpu_launches']))"; }
run SPMF_GRAPHS=0
run SPMF_GRAPHS=1



WL=c1 run SPMF_GRAPHS=0
WL=c1 run SPMF_GRAPHS=1
WL=c2 run SPMF_GRAPHS=0
WL=c2 run SPMF_GRAPHS=1
WL=c4 run SPMF_GRAPHS=0
WL=c4 run SPMF_GRAPHS=1
