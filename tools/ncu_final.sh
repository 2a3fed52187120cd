#!/bin/bash
# Round-1 profile evidence: launch list of the bench command, a full-metric capture of every kernel of one forward pass,
# and source-annotated captures of the dominant kernels.  (Graphs off: kernels inside a graph launch cannot be profiled.)
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r01_launches_bench.log 2>&1
N="ncu --set full --clock-control none"
timeout 600 $N -k regex:"frontend|stem|dwse|gemm|se_scale" -c 86 -o gpurun_out/r01g_forward_all -f python tools/prof_targets.py all 1024 > gpurun_out/r01g_forward_all.log 2>&1
timeout 200 $N --import-source on -k regex:frontend_clip -c 1 -o gpurun_out/r01g_frontend -f python tools/prof_targets.py frontend 8192 > /dev/null 2>&1
timeout 200 $N --import-source on -k regex:gemm_tcgen05 -c 1 -o gpurun_out/r01g_gemm_b2a_expand -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N --import-source on -k regex:dwse -s 1 -c 1 -o gpurun_out/r01g_dwse_b2a -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N --import-source on -k regex:dwse -s 9 -c 1 -o gpurun_out/r01g_dwse_b5b -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
ls -la gpurun_out/ | tail -12
