"""Per-op device time of one embedding forward pass (CUDA events around every launch, plain launches, L2 flushed
before the pass).  Usage: python tools/op_times.py [batch] — prints op name, kind, microseconds (median of 7)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.model import EmbeddingModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = EmbeddingModel(W.random_init(0, randomize_bn=True, residual_gamma_scale=0.3))
feats = torch.rand((B, 49, 40), device="cuda") * 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
runs = []
for i in range(9):
    flush.zero_()
    _, ms = m.forward_timed(feats)
    runs.append(ms)
ms = np.median(np.stack(runs[2:]), axis=0)
info = m.op_info()
tot = 0.0
for (name, kind, fl, by, n, k, rows), t in zip(info, ms):
    print(f"{name:34s} {'stem gemm dwse'.split()[kind]:5s} {t * 1e3:8.1f} us   {by * B / (t * 1e-3) / 1e9:8.1f} GB/s alg")
    tot += t
print(f"total {tot * 1e3:.1f} us over {len(info)} ops")
