timeout 900 python -m pytest tests/test_embed_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^block\|^stem\|^dense\|^top" | tail -6
timeout 200 python tools/ab.py se240 2>&1 | head -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm" -c 3 -o gpurun_out/r01e_gemm -f python tools/prof_targets.py embed 1024 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
