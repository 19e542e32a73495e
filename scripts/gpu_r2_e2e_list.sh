#!/bin/bash
# dense-ingest tests, then the launch list of the END-TO-END leg (NVTX range spmf_e2e): the upload-side kernels
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_links.py -x -q -m gpu -k "dense_ingest or dense_ingested" --timeout 200 2>&1 | tail -5
CMD="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline"
SPMF_GRAPHS=0 $CMD > gpurun_out/plain.log 2>&1 &&
SPMF_GRAPHS=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spmf_e2e/" --csv \
    --log-file gpurun_out/launches_e2e.csv $CMD > gpurun_out/ncu_list.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_e2e.csv 3 gpurun_out/launch_summary_e2e.txt | cut -c1-60,97-200 | tail -45
