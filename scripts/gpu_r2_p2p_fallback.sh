#!/bin/bash
# one rank fails to map its peers: every rank must fall back to the all-reduce tail together
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
SPMF_P2P_TEST_FAIL_RANK=1 DP_CHECK_EXPECT=nccl-allreduce timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dp_check.py > gpurun_out/dp_check_fallback.log 2>&1
echo "rc=$?"; grep "dp_check\|DP_CHECK\|Error" gpurun_out/dp_check_fallback.log | tail -6
