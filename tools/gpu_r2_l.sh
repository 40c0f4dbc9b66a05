#!/bin/bash
# GPU call L: ncu --set full of the CTA-pair GEMM on the CLIP QKV shape (what bounds it: L2 hit rate, DRAM, stalls)
mkdir -p gpurun_out
timeout 100 python tools/gemm_one.py > gpurun_out/r02l_plain.log 2>&1; echo "plain exit $?"; tail -3 gpurun_out/r02l_plain.log

ORYON_GEMM_QUAD=0 timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02l_gemm_tc2 -k regex:gemm_tc2_kernel --launch-skip 2 --launch-count 1 python tools/gemm_one.py > gpurun_out/r02l_ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02l_gemm_tc2.ncu-rep --page raw --csv > gpurun_out/r02l_gemm_tc2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02l_gemm_tc2.ncu-rep --page source --csv > gpurun_out/r02l_gemm_tc2_source.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02l_gemm_tc2_raw.csv")))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__cycles_active.avg", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:50])
    for k in keys:
        if k in hdr: print("   ", k, r[hdr.index(k)], units[hdr.index(k)])
PY
