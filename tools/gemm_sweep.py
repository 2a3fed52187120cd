"""block_n sweep of the tcgen05 GEMM over every dense contraction of the embedding tower at a given batch
(CUDA events, median of 9, cold = L2 flushed before each launch, warm = operands left in L2 by the previous launch).
Prints, per unique shape, the time of the built-in choice (block_n = 0) and of every explicit candidate."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.model import EmbeddingModel, gemm_h16

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = EmbeddingModel(W.random_init(0))
shapes = {}
for (name, kind, fl, by, n, k, rows) in m.op_info():
    if kind == 1:
        act = 1 if ("expand" in name or "top" in name) else (2 if name.startswith("dense") else 0)
        shapes.setdefault((rows * B, n, k, act), []).append(name)
for blk in W.block_list():                           # the external SE pair of the wide blocks
    if blk["cexp"] > 256:
        se16 = (blk["se"] + 15) // 16 * 16
        shapes.setdefault((B, se16, blk["cexp"], 1), []).append(blk["name"] + "_se_reduce")
        shapes.setdefault((B, blk["cexp"], se16, 4), []).append(blk["name"] + "_se_expand")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, cold):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(9):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return float(np.median(ts))


for (M, N, K, act), names in shapes.items():
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.empty((M, N), dtype=torch.float16, device="cuda")
    row = dict(names=names[0] + (f" (+{len(names) - 1})" if len(names) > 1 else ""), M=M, N=N, K=K, count=len(names))
    cands = [0] + [bn for bn in range(16, 257, 16) if bn <= max(16, (N + 15) // 16 * 16)]
    res = {}
    for bn in cands:
        f = lambda bn=bn: gemm_h16(a, w, bias=bias, act=act, block_n=bn, out=out)
        res[bn] = (round(timed(f, True), 1), round(timed(f, False), 1))
    best_cold = min((v[0], k) for k, v in res.items() if k)
    best_warm = min((v[1], k) for k, v in res.items() if k)
    row.update(default_cold_warm=res[0], best_cold=best_cold, best_warm=best_warm,
               all={k: v for k, v in res.items() if k})
    print(json.dumps(row))
