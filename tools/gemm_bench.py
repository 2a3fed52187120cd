"""Micro-benchmark of the tcgen05 GEMM operator on the embedding tower's shapes (CUDA events, median of 20)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multilingual_kws_b200.model import gemm_h16

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
shapes = [("packed4_b2a_expand", 125, 384, 64, 1), ("k64_n96", 500, 96, 64, 1), ("k16_n16", 500, 16, 16, 0), ("b1a_proj", 500, 16, 32, 0), ("b2a_expand", 500, 96, 16, 1), ("b2a_proj", 130, 24, 96, 0), ("b2b_expand", 130, 144, 24, 1),
          ("b3b_expand", 35, 240, 40, 1), ("b4b_expand", 12, 480, 80, 1), ("b5b_expand", 12, 672, 112, 1), ("b5b_proj", 12, 112, 672, 0),
          ("b6b_expand", 4, 1152, 192, 1), ("b6b_proj", 4, 192, 1152, 0), ("b7a_proj", 4, 320, 1152, 0), ("top", 4, 1280, 320, 1),
          ("dense", 1, 2048, 1280, 2), ("dense_1", 1, 2048, 2048, 2), ("dense_2", 1, 1024, 2048, 3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, rows, N, K, act in shapes:
    M = rows * B
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    bias = torch.randn(N, device="cuda") * 0.1
    res = {}
    for tag, kw in (("full", dict(bias=bias, act=act)), ("noact", dict(bias=bias, act=0)), ("f32out", dict(bias=bias, act=act, out_f32=True))):
        for _ in range(3):
            gemm_h16(a, w, **kw)
        ts = []
        for _ in range(15):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); gemm_h16(a, w, **kw); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        res[tag] = float(np.median(ts))
    us = res["full"] * 1e3
    print(json.dumps(dict(name=name, M=M, N=N, K=K, us=round(us, 1), noact_us=round(res["noact"] * 1e3, 1),
                          f32out_us=round(res["f32out"] * 1e3, 1), tflops=round(2.0 * M * N * K / us / 1e6, 1),
                          gbs=round((M * K * 2 + M * N * 2 + N * K * 2) / us / 1e3, 1), tiles=-(-M // 128))))
