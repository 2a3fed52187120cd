#!/bin/bash
# One GPU visit: corrected tcgen05.mma microbenchmark, squeeze-excite kernel A/B (+ memcheck), the GPU test suite, one bench line.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 60 tools/microbench/mma_chain > gpurun_out/mma_chain.txt 2>&1; echo "mma_chain rc $? at $(( $(date +%s) - t0 )) s"
timeout 180 python tools/se_ab.py > gpurun_out/se_ab.txt 2>&1; echo "se_ab rc $? at $(( $(date +%s) - t0 )) s"; tail -4 gpurun_out/se_ab.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc $? at $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/gpu_tests.log
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $? at $(( $(date +%s) - t0 )) s"
timeout 150 compute-sanitizer --tool memcheck python tools/se_ab.py --batch 597 --reps 2 > gpurun_out/se_memcheck.log 2>&1; echo "memcheck rc $? at $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/se_memcheck.log
