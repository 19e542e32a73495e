#!/bin/bash
# dense ingest: tests, then C2/C3 bench lines (e2e through HostDense) next to the u8-CSR transfer
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hybrid.py -x -q -m gpu -k "dense_ingest" > gpurun_out/dense_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/dense_tests.log
tail -15 gpurun_out/dense_tests.log
for w in c3 c2; do
  timeout 600 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/dense_$w.json 2> gpurun_out/dense_$w.err
  BENCH_HOST_FORMAT=u8 timeout 600 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/dense_${w}_u8.json 2> gpurun_out/dense_${w}_u8.err
done
python - <<'PY'
import json
for f in ("dense_c3","dense_c3_u8","dense_c2","dense_c2_u8"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(j["ms_per_step"],4), "e2e ms", round(j["e2e"]["ms_per_step"],4), j["e2e"].get("host_format"), "h2d", j["e2e"]["h2d_bytes_per_step"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
