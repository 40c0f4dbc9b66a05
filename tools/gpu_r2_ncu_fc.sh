#!/bin/bash
# ncu --set full of the CLIP up-projection (fc: bias + QuickGELU + 8-bit cross-term split output) inside a network pass
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import sys; sys.path.insert(0, \".\"); from oryon_b200 import _lib; _lib.load(); print(\"lib loads under ncu\")" > gpurun_out/r02_ncu_probe.log 2>&1 || { echo "this box crashes the library under ncu: giving up early"; exit 3; }
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_gemm_fc_p2 --kernel-name-base demangled -k "regex:gemm_tc2_kernel<\(int\)2" --launch-skip 2 --launch-count 1 \
  python tools/bench_backbone.py --pairs 32 --chunk 32 --precision 2 --steps 1 > gpurun_out/r02_ncu_fc.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02_gemm_fc_p2.ncu-rep --page raw --csv > gpurun_out/r02_gemm_fc_p2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_gemm_fc_p2.ncu-rep --page source --csv > gpurun_out/r02_gemm_fc_p2_source.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_gemm_fc_p2_raw.csv")))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_write.sum", "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for k in keys:
        if k in hdr: print("   ", k, r[hdr.index(k)], units[hdr.index(k)])
src = list(csv.reader(open("gpurun_out/r02_gemm_fc_p2_source.csv", errors="ignore")))
h = src[0]
print(h[:12])
si = [i for i, n in enumerate(h) if "Sampl" in n]
print([h[i] for i in si])
if si:
    c = si[0]
    body = [r for r in src[1:] if len(r) > c and r[c].replace(",", "").isdigit()]
    body.sort(key=lambda r: -int(r[c].replace(",", "")))
    tot = sum(int(r[c].replace(",", "")) for r in body)
    for r in body[:30]:
        print(f"{int(r[c].replace(',', '')) * 100.0 / tot:5.1f}%", " | ".join(r[:3])[:150])
PY
