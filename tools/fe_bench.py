"""Quick frontend-only timing (CUDA events, L2 flushed between iterations)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multilingual_kws_b200.frontend import MicroFrontend
from multilingual_kws_b200.synthetic import synthetic_pcm

fe = MicroFrontend()
base = synthetic_pcm(256, cfg_id=2)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for B in (32, 1024, 8192, 32768):
    pcm = torch.from_numpy(np.tile(base, (max(1, B // 256), 1))[:B]).cuda()
    out = torch.empty((B, 49, 40), dtype=torch.float32, device="cuda")
    for _ in range(3):
        fe.forward(pcm, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fe.forward(pcm, out=out); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = float(np.median(ts))
    print(json.dumps(dict(B=B, ms=ms, clips_per_s=B / ms * 1e3, algo_GBps=B * 39840 / ms / 1e6)))
