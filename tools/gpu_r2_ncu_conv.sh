#!/bin/bash
# ncu --set full of the 192 x 192 implicit-GEMM convolution (gemm_tc_kernel<32, 3, gather>) inside a network pass
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 || { echo "this box crashes the library under ncu: giving up early"; exit 3; }
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_conv_gather --kernel-name-base demangled -k "regex:gemm_tc_kernel<\(int\)32, \(int\)3, \(bool\)1" --launch-skip 4 --launch-count 1 \
  python tools/bench_backbone.py --pairs 32 --chunk 32 --precision 2 --steps 1 > gpurun_out/r02_ncu_conv.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02_conv_gather.ncu-rep --page raw --csv > gpurun_out/r02_conv_gather_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_conv_gather_raw.csv")))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:70])
    for k in keys:
        if k in hdr: print("   ", k, r[hdr.index(k)], units[hdr.index(k)])
PY
