#!/usr/bin/env python
"""Turn the gpurun_out/ artefacts of scripts/gpu_final.sh into the tracked profiles/ files:
  profiles/<tag>_launches.csv        ncu launch list of the timed region (gpu__time_duration)
  profiles/<tag>_launch_summary.txt  per-kernel totals and share of the step
  profiles/<tag>_ncu_full_summary.txt key metrics of the `ncu --set full` capture
  profiles/r1_traffic.json           DRAM bytes per launch of the hot kernels (bench.py reads it)
Usage: python scripts/make_profiles.py <run tag> <profile tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
run, tag = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
src = os.path.join(ROOT, "gpurun_out", f"launches_{run}.csv")
lines = [l for l in open(src) if l.startswith('"')]
open(os.path.join(out, f"{tag}_launches.csv"), "w").writelines(lines)
rows = list(csv.reader(lines))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
steps = 3
tot = sum(sum(v) for v in agg.values()) / steps / 1000
with open(os.path.join(out, f"{tag}_launch_summary.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, NVTX range spmf_timed ({steps} steps); "
            "times are serialised cold-cache launches: shares, not absolutes\n")
    for n, v in agg.items():
        f.write(f"{n[:96]:96s} n/step={len(v) / steps:5.1f} avg={sum(v) / len(v) / 1000:9.1f} us "
                f"per-step={sum(v) / steps / 1000:8.1f} us share={sum(v) / steps / 1000 / tot:6.1%}\n")
    f.write(f"total per step {tot:.1f} us (serialised)\n")
rep = os.path.join(ROOT, "gpurun_out", f"prof_{run}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
tmp = f"/tmp/{run}_raw.csv"
open(tmp, "w").write(raw)
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), tmp], capture_output=True, text=True).stdout
extra = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
         "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
         "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
         "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
         "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
traffic = {}
names = {"hot_tile_kernel": "hot_tile", "csr_rows_kernel": "csr_rows", "csc_cols_kernel": "csc_cols", "umma_gemm3_kernel": "umma_gemm3"}
with open(os.path.join(out, f"{tag}_ncu_full_summary.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on (one capture per kernel; see scripts/gpu_final.sh)\n")
    f.write(summ)
    f.write("\n---- tensor / shared-memory pipe metrics ----\n")
    for r in rr[2:]:
        kn = r[h.index("Kernel Name")]
        f.write(kn[:100] + "\n")
        for e in extra:
            if e in h:
                f.write(f"  {e:100s} {r[h.index(e)]}\n")
        for key, short in names.items():
            if key in kn and short not in traffic:
                rd = float(r[h.index("dram__bytes_read.sum")]); wr = float(r[h.index("dram__bytes_write.sum")])
                ur, uw = rr[1][h.index("dram__bytes_read.sum")], rr[1][h.index("dram__bytes_write.sum")]
                mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                traffic[short] = rd * mult.get(ur, 1.0) + wr * mult.get(uw, 1.0)
json.dump(traffic, open(os.path.join(out, "r1_traffic.json"), "w"), indent=1)
print(open(os.path.join(out, f"{tag}_launch_summary.txt")).read())
print(traffic)
