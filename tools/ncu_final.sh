#!/bin/bash
# Round-1 profile evidence (graphs off: kernels inside a graph launch cannot be profiled):
#   1. launch list of the bench command (durations only)
#   2. a metric table for every kernel of one frontend + embedding forward pass (text CSV)
#   3. full, source-annotated captures of the dominant kernels (.ncu-rep, read offline)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r01_launches_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 500 ncu --metrics $M --clock-control none -k regex:"frontend|stem|dwse|gemm|se_scale" -c 170 --csv --log-file gpurun_out/r01_forward_all_metrics.csv python tools/prof_targets.py all 1024 > gpurun_out/r01_forward_all.log 2>&1
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:frontend_clip -c 1 -o gpurun_out/r01g_frontend -f python tools/prof_targets.py frontend 8192 > /dev/null 2>&1
timeout 200 $N -k regex:gemm_tcgen05 -c 1 -o gpurun_out/r01g_gemm_b2a_expand -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N -k regex:dwse -s 1 -c 1 -o gpurun_out/r01g_dwse_b2a -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N -k regex:dwse -s 9 -c 1 -o gpurun_out/r01g_dwse_b5b -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
du -sh gpurun_out; ls -la gpurun_out/ | tail -14
