timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_embed_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^block\|^stem\|^dense\|^top" | tail -12
timeout 200 python tools/ab.py direct_store
timeout 300 python tools/sweep.py 2>&1 | head -2
