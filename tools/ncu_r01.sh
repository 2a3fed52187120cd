#!/bin/bash
# ncu --set full captures of the kernels that dominate the step (1 GPU, graphs off).  Output: gpurun_out/*.ncu-rep
set -x
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 280 $N -k regex:frontend_clip -c 1 -o gpurun_out/r01_frontend -f python tools/prof_targets.py frontend 8192 > /dev/null 2>&1
timeout 280 $N -k regex:gemm_tcgen05 -s 1 -c 1 -o gpurun_out/r01_gemm_b2a_expand -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 280 $N -k regex:gemm_tcgen05 -s 32 -c 1 -o gpurun_out/r01_gemm_b6b_expand -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 280 $N -k regex:gemm_tcgen05 -s 42 -c 1 -o gpurun_out/r01_gemm_dense1 -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 280 $N -k regex:dwse -s 0 -c 1 -o gpurun_out/r01_dwse_b1a -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 280 $N -k regex:dwse -s 18 -c 1 -o gpurun_out/r01_dwse_b6b -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
