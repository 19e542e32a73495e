cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "spmf_timed/" \
    -k regex:'hot_tile_kernel|csr_rows_|csc_cols_kernel|umma_gemm3_kernel|backward_dk_kernel|gamma_kernel' -c 6 -f -o gpurun_out/prof_r1g $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
