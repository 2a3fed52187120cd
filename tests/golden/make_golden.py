#!/usr/bin/env python
"""Regenerates the committed fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

What can and cannot be generated: the reference's arithmetic lives in TensorFlow 2.7 (absent, not installable
offline — SURVEY.md §8c), so nothing here comes from running the reference itself.
  * tf_microfrontend_kat.json — hand-written, NOT generated: known-answer vectors of the upstream op's own unit
    tests; they pin the oracle (tests/test_oracle_tf_kat.py).
  * frontend_cfg1.npz, embed_cfg2.npz, head_cfg3.npz — generated below by the pinned oracle on the seeded synthetic
    inputs of SURVEY.md §8d.  The GPU parity tests (tests/test_golden_gpu.py) compare the CUDA path with these files
    without executing anything under oracle/; tests/test_golden_cpu.py checks that the oracle still reproduces them.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from multilingual_kws_b200 import weights as W                      # noqa: E402
from multilingual_kws_b200.synthetic import synthetic_pcm           # noqa: E402
from oracle import effnet_oracle as EO                              # noqa: E402
from oracle import head_oracle as HO                                # noqa: E402
from oracle.frontend_oracle import FrontendOracle                   # noqa: E402

N_FRONTEND, N_EMBED, EMBED_SEED = 32, 8, 3


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def adversarial_pcm() -> np.ndarray:
    """Edge cases next to the §8d mixture: constant extremes, Nyquist square, +-1, single impulse, silence."""
    adv = np.zeros((8, 16000), np.int16)
    adv[0] = -32768
    adv[1] = 32767
    adv[2, ::2], adv[2, 1::2] = -32768, 32767
    adv[3] = 1
    adv[4] = -1
    adv[5, 5000] = -32768
    adv[7] = (np.arange(16000) * 7919 % 65536 - 32768).astype(np.int16)
    return adv


def head_case():
    rng = np.random.default_rng(3)
    p = HO.init_head(3)
    p["b1"] = rng.normal(0, .1, 18).astype(np.float32)
    p["b2"] = rng.normal(0, .1, 3).astype(np.float32)
    y = rng.integers(0, 3, 64)
    emb = (rng.normal(0, 1, (3, 1024))[y] * 0.3 + rng.normal(0, 1.0, (64, 1024))).astype(np.float32)
    return p, emb, y


def main():
    orc = FrontendOracle()
    pcm = synthetic_pcm(N_FRONTEND, cfg_id=1)
    adv = adversarial_pcm()
    np.savez_compressed(os.path.join(HERE, "frontend_cfg1.npz"), pcm_sha256=sha(pcm), features_u16=orc.features_u16(pcm),
                        adversarial_sha256=sha(adv), adversarial_features_u16=orc.features_u16(adv))

    feats = orc.features(synthetic_pcm(40, cfg_id=2), threads=4)
    w = W.random_init(EMBED_SEED, randomize_bn=True, residual_gamma_scale=0.3)
    before = {k: v.copy() for k, v in w.items()}
    EO.forward(w, feats, calibrate_bn=True)
    bn = {k.replace("/", "__"): v for k, v in w.items() if not np.array_equal(v, before[k])}
    emb = EO.forward(w, feats[:N_EMBED]).numpy()
    np.savez_compressed(os.path.join(HERE, "embed_cfg2.npz"), features_sha256=sha(feats[:N_EMBED]),
                        features_u16=np.rint(feats[:N_EMBED] * 25.6).astype(np.uint16), embedding=emb, **bn)

    p, e, y = head_case()
    _, _, probs = HO.forward(p, e)
    loss, acc, g = HO.loss_and_grads(p, e, y)
    p10 = {k: v.copy() for k, v in p.items()}
    hist = HO.train(p10, e, y, steps=10, lr=1e-3)                  # ten Keras-Adam steps
    np.savez_compressed(os.path.join(HERE, "head_cfg3.npz"), emb=e, labels=y.astype(np.int32), probs=probs.astype(np.float32),
                        loss=np.float64(loss), acc=np.float64(acc), **{f"p_{k}": v for k, v in p.items()},
                        **{f"g_{k}": np.asarray(v, np.float32) for k, v in g.items()},
                        **{f"p10_{k}": v for k, v in p10.items()}, loss_history=np.array([h[0] for h in hist]))
    for f in sorted(os.listdir(HERE)):
        print(f"{f:32s} {os.path.getsize(os.path.join(HERE, f)):8d} B")


if __name__ == "__main__":
    main()
