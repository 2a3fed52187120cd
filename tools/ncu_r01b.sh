#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 280 $N -k regex:gemm_tcgen05 -s 1 -c 1 -o gpurun_out/r01b_gemm_b2a_expand -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 280 $N -k regex:dwse -s 1 -c 1 -o gpurun_out/r01b_dwse_b2a -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 280 $N -k regex:dwse -s 15 -c 1 -o gpurun_out/r01b_dwse_b5b -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
