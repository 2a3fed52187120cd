timeout 900 python -m pytest tests/test_embed_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^block\|^stem\|^dense\|^top" | tail -8
timeout 200 python tools/ab.py se_narrow 2>&1 | head -4
