#!/bin/bash
# ncu --set full of the side-stream kernels (Gamma draws + implicit gradients, data-independent backward)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-g0}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_timed/" \
    -k regex:"${KREGEX:-gamma_kernel|backward_dk_kernel|backward_feat_kernel|draw_operands}" -c ${NCAP:-4} -f -o gpurun_out/side_$TAG $CMD > gpurun_out/ncu_side.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_side.log | cut -c1-300
