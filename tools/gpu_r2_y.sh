#!/bin/bash
# GPU call Y: ncu --set full of the CTA-pair GEMM at precision 2 (fp8 cross terms) on the CLIP QKV shape of a 32-pair pass
mkdir -p gpurun_out
export GEMM_ONE_PRECISION=2 GEMM_ONE_M=36928
timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02y_gemm_tc2_p2 -k regex:gemm_tc2_kernel --launch-skip 2 --launch-count 1 python tools/gemm_one.py > gpurun_out/r02y_ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02y_gemm_tc2_p2.ncu-rep --page raw --csv > gpurun_out/r02y_gemm_tc2_p2_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02y_gemm_tc2_p2_raw.csv")))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__cycles_active.avg", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__cluster_size"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60])
    for k in keys:
        if k in hdr: print("   ", k, r[hdr.index(k)], units[hdr.index(k)])
PY
