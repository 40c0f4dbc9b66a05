#!/bin/bash
# ncu --set full of the ping-pong attention kernel inside a network pass (32 pairs) + the new GEMM error-path test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu -k "depth_multiple or saturate" > gpurun_out/r02_pytest_gemm_errpath.log 2>&1; echo "gemm error-path test exit $?"; tail -2 gpurun_out/r02_pytest_gemm_errpath.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 || { echo "this box crashes the library under ncu: giving up early"; exit 3; }
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_attn_pp -k regex:attn_pp_kernel --launch-skip 3 --launch-count 1 \
  python tools/bench_backbone.py --pairs 32 --chunk 32 --precision 2 --steps 1 > gpurun_out/r02_ncu_attn.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02_attn_pp.ncu-rep --page raw --csv > gpurun_out/r02_attn_pp_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_attn_pp_raw.csv")))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
out = ["| metric | value |", "|---|---|"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60])
    for k in keys:
        if k in hdr:
            print("   ", k, r[hdr.index(k)], units[hdr.index(k)])
            out.append(f"| `{k}` | {r[hdr.index(k)]} {units[hdr.index(k)]} |")
open("gpurun_out/r02_attn_pp_ncu_metrics.md", "w").write("\n".join(out) + "\n")
PY
