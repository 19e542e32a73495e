#!/bin/bash
# First GPU check: parity tests + smoke.  Usage: gpurun -- bash scripts/gpu_check.sh
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
