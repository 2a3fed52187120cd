#!/bin/bash
# Round-2 profile evidence (graphs off: kernels inside a graph launch cannot be profiled):
#   1. launch list of the bench command (durations only)      2. metric table of every kernel of one forward pass
#   3. --set full captures of the dominant kernels            4. compute-sanitizer memcheck / racecheck logs
#   5. tcgen05 issue microbenchmark                            6. early-chunk sweep of the bench
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-finetune-e2e --no-graph > gpurun_out/r02_launches_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 500 ncu --metrics $M --clock-control none -k regex:"frontend|stem|dwse|gemm|se_scale" -c 170 --csv --log-file gpurun_out/r02_forward_all_metrics.csv python tools/prof_targets.py all 1024 > gpurun_out/r02_forward_all.log 2>&1
KWS_FUSE=2 timeout 300 ncu --metrics $M --clock-control none -k regex:"mbconv" -c 12 --csv --log-file gpurun_out/r02_fused_metrics.csv python tools/prof_targets.py embed 1024 > gpurun_out/r02_fused.log 2>&1
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:gemm_tcgen05 -c 1 -o gpurun_out/r02_gemm_b2a_expand -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 200 $N -k regex:gemm_tcgen05 -s 52 -c 1 -o gpurun_out/r02_gemm_dense_1 -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
KWS_FUSE=1 timeout 200 $N -k regex:mbconv -s 4 -c 1 -o gpurun_out/r02_fused_b5b -f python tools/prof_targets.py embed 1024 > /dev/null 2>&1
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_targets.py 8 > gpurun_out/r02_sanitizer_racecheck.log 2>&1
timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_targets.py 8 > gpurun_out/r02_sanitizer_synccheck.log 2>&1
timeout 60 tools/microbench/mma_chain > gpurun_out/r02_mma_chain.txt 2>&1
for c in 1024 512 256; do timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-finetune-e2e --chunk $c > gpurun_out/r02_bench_chunk$c.json 2>/dev/null; done
tail -3 gpurun_out/r02_sanitizer_*.log; cat gpurun_out/r02_mma_chain.txt | head -8
python - <<PY
import json
for c in (1024, 512, 256):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_chunk{c}.json")); print("chunk", c, "ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    except Exception as e: print(c, e)
PY
du -sh gpurun_out; ls -la gpurun_out/ | tail -25
