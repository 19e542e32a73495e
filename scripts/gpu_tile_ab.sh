#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-t2}
timeout 600 python -m pytest tests/test_gpu_hybrid.py -q -x -k "hybrid_step or fit" 2>&1 | tail -15 | tee gpurun_out/pytest_tile_$TAG.log
HDS="${HDS:-0.03}" HOTMODE=2 NCU_HD=0.03 bash scripts/gpu_hyb_ab.sh $TAG
