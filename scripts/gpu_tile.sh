#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-t0}
timeout 600 python -m pytest tests/test_gpu_hybrid.py -q -x 2>&1 | tail -40 | tee gpurun_out/pytest_tile_$TAG.log
