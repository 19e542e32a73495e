#!/bin/bash
# tcgen05 GEMM + hybrid-step tests on one B200.  Usage: gpurun --timeout 900 -- bash scripts/gpu_hybrid.sh <tag>
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-h0}
timeout 300 python -m pytest tests/test_gpu_hybrid.py -q -x -k "umma or hot_split" 2>&1 | tail -30 | tee gpurun_out/pytest_hyb_a_$TAG.log
timeout 600 python -m pytest tests/test_gpu_hybrid.py -q -k "not umma and not hot_split" 2>&1 | tail -40 | tee gpurun_out/pytest_hyb_b_$TAG.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -15 | tee gpurun_out/pytest_par_$TAG.log
