"""Fused-tail kernel check on the GPU: per-block outputs and final embedding of the fused schedules (modes 1, 2) against
the layer-wise schedule (mode 0) and the torch fp32 oracle, on the two synthetic weight regimes; then device times.

    python tools/fused_check.py [--batch 1024] [--quick]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from multilingual_kws_b200 import weights as W                     # noqa: E402
from multilingual_kws_b200.model import EmbeddingModel             # noqa: E402
from multilingual_kws_b200.synthetic import synthetic_pcm          # noqa: E402
from oracle import effnet_oracle as EO                             # noqa: E402
from oracle.frontend_oracle import FrontendOracle                  # noqa: E402


def rel_err(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--clips", type=int, default=40)
    args = ap.parse_args()
    feats = FrontendOracle().features(synthetic_pcm(args.clips, cfg_id=2), threads=4)
    x = torch.from_numpy(feats).cuda()
    ok = True
    for regime, gamma in (("damped", 0.3), ("undamped", 1.0)):
        w = W.random_init(3, randomize_bn=True, residual_gamma_scale=gamma)
        EO.forward(w, feats, calibrate_bn=True)
        taps = {}
        want = EO.forward(w, feats, taps=taps).numpy()
        taps["top_gap"] = taps["top_activation"].mean(axis=(1, 2))
        model = EmbeddingModel(w)
        names = model.op_names()
        outs = {}
        for mode in (0, 1, 2):
            model.set_fuse(mode)
            try:
                got = model.forward_device(x).cpu().numpy()
                torch.cuda.synchronize()
            except Exception as e:                                  # noqa: BLE001
                print(f"[{regime}] mode {mode}: FAILED {e}")
                ok = False
                continue
            outs[mode] = got
            cos = EO.cosine(got, want).min()
            print(f"[{regime}] fuse mode {mode}: launches {model.launches(args.clips)}  min cosine vs oracle {cos:.6f}  "
                  f"rel err {rel_err(got, want):.5f}" + (f"  vs mode 0 {rel_err(got, outs[0]):.5f}" if 0 in outs and mode else ""))
            if not np.isfinite(got).all():
                print("   non-finite output!")
                ok = False
        if regime == "damped" and not args.quick:
            # block outputs of mode 1 (every fused block is a launch whose last op can be tapped)
            model.set_fuse(1)
            for i, (name, elems) in enumerate(names):
                if not (name.endswith("_out") and name[5] in "4567"):
                    continue
                _, tap = model.forward_device(x, tap_op=i)
                torch.cuda.synchronize()
                ref = taps[name].reshape(feats.shape[0], -1)
                e = rel_err(tap.float().cpu().numpy(), ref)
                print(f"   mode 1 tap {name}: rel err {e:.5f}")
                if not e < 0.02:
                    ok = False
            model.set_fuse(2)
            for i, (name, elems) in enumerate(names):
                if name not in ("block4c_out", "block5c_out", "block6d_out", "block7a_out", "top_gap"):
                    continue
                try:
                    _, tap = model.forward_device(x, tap_op=i)
                    torch.cuda.synchronize()
                except Exception as e:                              # noqa: BLE001
                    print(f"   mode 2 tap {name}: FAILED {e}")
                    continue
                ref = taps[name].reshape(feats.shape[0], -1)
                print(f"   mode 2 tap {name}: rel err {rel_err(tap.float().cpu().numpy(), ref):.5f}")
        # ragged batch sizes through the fused path
        model.set_fuse(2)
        full = model.forward_device(x).clone()
        for b in (1, 3, 33):
            if not torch.equal(model.forward_device(x[:b]), full[:b]):
                d = (model.forward_device(x[:b]) - full[:b]).abs().max().item()
                print(f"   batch {b}: differs from the full batch by {d:.3e}")
        del model
    # timing
    B = args.batch
    reps = -(-B // feats.shape[0])
    xb = torch.from_numpy(np.tile(feats, (reps, 1, 1))[:B]).cuda()
    w = W.random_init(3, randomize_bn=True, residual_gamma_scale=0.3)
    EO.forward(w, feats, calibrate_bn=True)
    model = EmbeddingModel(w)
    out = torch.empty((B, 1024), device="cuda")
    for mode in (0, 1, 2, 3):
        model.set_fuse(mode)
        for _ in range(3):
            model.forward_device(xb, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            model.forward_device(xb, out=out)
        e1.record()
        torch.cuda.synchronize()
        _, ms = model.forward_timed(xb)
        names = model.op_names()
        tail = sum(m for (nm, _), m in zip(names, ms) if nm[:6] in ("block4", "block5", "block6", "block7") or nm == "top_gap")
        print(f"B={B} fuse mode {mode}: {e0.elapsed_time(e1) / n:.4f} ms per forward (graph), launches {model.launches(B)}, "
              f"per-op sum {ms.sum():.4f} ms, tail (block4a..top_gap) {tail:.4f} ms")
        if mode:
            print("   " + "  ".join(f"{nm}={m * 1e3:.1f}us" for (nm, _), m in zip(names, ms)
                                   if m > 0 and (nm[:6] in ("block4", "block5", "block6", "block7") or nm == "top_gap")))
    print("FUSED_CHECK", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(f"({time.time() - t0:.1f} s)")
    sys.exit(rc)
