"""A/B of the wide layers' squeeze-excite on the GPU: one CUDA-core launch (FC1 + FC2 + gating, the default) against the
round-1 path (two tcgen05 GEMMs + a gating pass, KWS_SE_KERNEL=0): results, forward time (graph replay, L2 flushed) and
the per-op times of the ten affected blocks.

    python tools/se_ab.py [--batch 1024] [--reps 30]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from multilingual_kws_b200 import weights as W                     # noqa: E402
from multilingual_kws_b200.model import EmbeddingModel             # noqa: E402


def build(se_kernel):
    os.environ["KWS_SE_KERNEL"] = str(se_kernel)
    try:
        return EmbeddingModel(W.random_init(3, randomize_bn=True, residual_gamma_scale=0.3))
    finally:
        del os.environ["KWS_SE_KERNEL"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    B = args.batch
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((B, 49, 40), device="cuda", generator=g) * 26.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    models = {"se kernel": build(1), "two GEMMs + gating": build(0)}
    outs, times = {}, {}
    for name, m in models.items():
        out = torch.empty((B, m.output_dim), device="cuda")
        for _ in range(4):
            m.forward_device(x, out=out)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); m.forward_device(x, out=out); e.record()
            torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        outs[name] = out.clone()
        times[name] = float(np.median(ms))
        _, op_ms = m.forward_timed(x)
        _, op_ms = m.forward_timed(x)
        names = [n for n, _ in m.op_names()]
        se_ops = [(n, t) for n, t in zip(names, op_ms) if n.endswith("_se_excite")]
        print(f"{name}: {times[name]:.4f} ms per forward (graph, L2 flushed, median of {args.reps}), launches {m.launches(B)}")
        print("   " + "  ".join(f"{n[:7]}={t * 1e3:.1f}us" for n, t in se_ops))
        print(f"   depthwise + SE ops together: {sum(t for _, t in se_ops) * 1e3:.1f} us (plain launches)")
    a, b = outs["se kernel"].double(), outs["two GEMMs + gating"].double()
    rel = float((a - b).norm() / b.norm())
    print(f"relative difference of the embeddings: {rel:.2e}; finite: {bool(torch.isfinite(a).all())}")
    print("SE_AB", "OK" if rel < 5e-3 and torch.isfinite(a).all() else "FAILED")


if __name__ == "__main__":
    main()
