#!/bin/bash
# One GPU visit: the GPU test suite, smoke(), one bench line (state after the last code change of the round).
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc $? at $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc $? at $(( $(date +%s) - t0 )) s"; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $? at $(( $(date +%s) - t0 )) s"
