#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=driver_version,name --format=csv,noheader; ncu --version | tail -2
N="ncu --metrics gpu__time_duration.sum --clock-control none -c 40"
GEMM_ONE_M=36928 $N python tools/ncu_diag_steps.py > gpurun_out/diag11.log 2>&1; echo "steps M=36928: $?"; grep -v "^ \|^$\|PROF\|WARNING" gpurun_out/diag11.log | tail -12
GEMM_ONE_M=18464 $N python tools/ncu_diag_steps.py > gpurun_out/diag12.log 2>&1; echo "steps M=18464: $?"; grep -v "^ \|^$\|PROF\|WARNING" gpurun_out/diag12.log | tail -12
GEMM_ONE_M=36928 python tools/ncu_diag_steps.py > gpurun_out/diag13.log 2>&1; echo "steps M=36928 without ncu: $?"; tail -4 gpurun_out/diag13.log
