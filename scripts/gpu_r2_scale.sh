#!/bin/bash
# Round 2, first GPU check: bench-scale / fitted / trajectory parity + the existing GPU suite.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
free -g | head -2; nproc
timeout 1500 python -m pytest tests/test_gpu_scale.py -q -s 2>&1 | tail -80 | tee gpurun_out/pytest_scale.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py 2>&1 | tail -15 | tee gpurun_out/pytest_rest.log
