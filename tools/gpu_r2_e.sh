#!/bin/bash
# GPU call E: attention kernel timeline (ORYON_ATTN_DEBUG clock stamps) + one ncu --set full capture of attn_tc_kernel
mkdir -p gpurun_out
ORYON_ATTN_DEBUG=1 timeout 200 python tools/bench_backbone.py --pairs 16 --steps 1 > gpurun_out/r02e_attn_dbg.json 2> gpurun_out/r02e_attn_dbg.err; echo "attn dbg exit $?"; grep -A 40 "attn_tc dbg" gpurun_out/r02e_attn_dbg.err | head -60
timeout 400 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02e_attn_tc -k regex:attn_tc_kernel --launch-skip 20 --launch-count 1 python tools/bench_backbone.py --pairs 16 --steps 1 > gpurun_out/r02e_ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02e_attn_tc.ncu-rep --page raw --csv > gpurun_out/r02e_attn_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_attn_tc.ncu-rep --page source --csv > gpurun_out/r02e_attn_tc_source.csv 2>/dev/null
ls -la gpurun_out/r02e_*
