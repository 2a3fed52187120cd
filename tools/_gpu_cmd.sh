timeout 900 python -m pytest tests/test_embed_gpu.py tests/test_gemm_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^block\|^stem\|^dense\|^top" | tail -6
timeout 200 python tools/ab.py bias_smem_stem 2>&1 | head -3
