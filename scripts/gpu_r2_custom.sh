#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_links.py -x -q -m gpu --timeout 200 -k "custom" 2>&1 | tail -25
