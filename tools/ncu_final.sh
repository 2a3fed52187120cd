#!/bin/bash
# Round-1 profile evidence: launch list of the bench command + full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r01_launches_bench.log 2>&1
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:frontend_clip -c 1 -o gpurun_out/r01f_frontend -f python tools/prof_targets.py frontend 8192 > /dev/null 2>&1
timeout 200 $N -k regex:gemm_tcgen05 -s 1 -c 1 -o gpurun_out/r01f_gemm_b2a_expand -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N -k regex:gemm_tcgen05 -s 70 -c 1 -o gpurun_out/r01f_gemm_dense1 -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N -k regex:dwse -s 1 -c 1 -o gpurun_out/r01f_dwse_b2a -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N -k regex:dwse -s 15 -c 1 -o gpurun_out/r01f_dwse_b5b -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
ls -la gpurun_out/ | tail -12
