"""Device time of the head kernels (forward+loss+backward, Adam) at the fine-tune batch sizes (CUDA events, median)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multilingual_kws_b200.fewshot import Head

head = Head.keras_init(1024, 18, 3, seed=0)
for B in (64, 512, 4096):
    emb = torch.randn(B, 1024, device="cuda")
    y = torch.randint(0, 3, (B,), device="cuda", dtype=torch.int32)
    for name, fn in (("grad", lambda: head.grad(emb, y)), ("adam", lambda: head.apply_adam(head._flat, 1e-3)),
                     ("forward", lambda: head.forward(emb))):
        for _ in range(5):
            fn()
        ts, hs = [], []
        big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        for _ in range(10):
            big.zero_()                                   # ~160 us of GPU work: the calls below are queued behind it,
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()                                    # so the events bracket device time, not host launch time
            import time
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            hs.append((time.perf_counter() - t0) / 3 * 1e6)
            e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / 3 * 1e3)
        # back-to-back (no memset in between: same shared-memory carve-out from call to call), 50 calls per region
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(50):
            fn()
        e2.record(); torch.cuda.synchronize()
        print(f"B={B:5d} {name:8s} device {np.median(ts):7.1f} us   host {np.median(hs):6.1f} us per call   back-to-back {s2.elapsed_time(e2) / 50 * 1e3:6.1f} us per call")
