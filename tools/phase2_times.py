"""Device time of the pieces of one phase-2 fine-tune step (frozen embedding, trainable-tail forward, backward, Adam)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.fewshot import Head
from multilingual_kws_b200.finetune import TailTrainer
from multilingual_kws_b200.model import EmbeddingModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = EmbeddingModel(W.random_init(0, randomize_bn=True, residual_gamma_scale=0.3))
tr = TailTrainer(m, Head.keras_init(1024, 18, 3, seed=0))
x = torch.from_numpy(np.random.default_rng(0).uniform(0, 26, (B, 49, 40)).astype(np.float32)).cuda()
y = torch.from_numpy((np.arange(B) % 3).astype(np.int32)).cuda()
tr._stats = torch.zeros(3, device="cuda")


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


x7 = tr.tail_input(x)
print(f"B={B}")
print(f"frozen embedding up to block6d : {t(lambda: tr.tail_input(x)):.3f} ms")
print(f"tail forward (activations kept): {t(lambda: tr.forward_tail(x7, keep=True)):.3f} ms")
print(f"backward (incl. head grad)     : {t(lambda: tr.backward(y)):.3f} ms")
print(f"Adam (14 tensors + head)       : {t(lambda: tr.apply_adam(1e-4)):.3f} ms")
print(f"whole step, eager              : {t(lambda: tr.step(x, y, 1e-4, graph=False)):.3f} ms")
print(f"whole step, CUDA graph         : {t(lambda: tr.step(x, y, 1e-4)):.3f} ms")
