"""oracle/tail_autograd_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

torch.autograd restatement of what Keras / TensorFlow differentiate in phase 2 of the reference's transfer_learn
(multilingual_kws/embedding/transfer_learning.py:97-112, "unfreeze the top 20 layers while leaving BatchNorm layers
frozen"): the last 20 layers of the embedding — block7a, top_conv, the dense tower (architecture:
train_multilingual_embedding.py:66-83 / Keras EfficientNetB0 block()) — followed by the few-shot
head Dense(18, tanh) -> Dense(3) with sparse categorical cross-entropy on the logits (:47-59), BatchNorm in inference
mode.  Gradients are taken with respect to the Keras-shaped kernels / biases.  PARITY UNPINNED against TensorFlow (no TF
offline); the forward agrees with oracle/effnet_oracle.py (tests/test_finetune_gpu.py checks that on the CPU).
Only tests/ and bench.py's checker legs may import this.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_EPS = 1e-3
_SELU_L, _SELU_A = 1.0507009873554805, 1.6732632423543772

TAIL_KEYS = ["block7a_expand_conv/kernel", "block7a_dwconv/depthwise_kernel", "block7a_se_reduce/kernel",
             "block7a_se_reduce/bias", "block7a_se_expand/kernel", "block7a_se_expand/bias", "block7a_project_conv/kernel",
             "top_conv/kernel", "dense/kernel", "dense/bias", "dense_1/kernel", "dense_1/bias", "dense_2/kernel", "dense_2/bias"]


def _bn(w, name, dtype):
    g, b = torch.as_tensor(w[name + "/gamma"], dtype=dtype), torch.as_tensor(w[name + "/beta"], dtype=dtype)
    m, v = torch.as_tensor(w[name + "/moving_mean"], dtype=dtype), torch.as_tensor(w[name + "/moving_variance"], dtype=dtype)
    s = g / torch.sqrt(v + _EPS)
    return s, b - m * s


def make_params(w: Dict[str, np.ndarray], head: Dict[str, np.ndarray], dtype=torch.float64) -> Dict[str, torch.Tensor]:
    p = {k: torch.tensor(np.asarray(w[k]), dtype=dtype, requires_grad=True) for k in TAIL_KEYS}
    for k in ("w1", "b1", "w2", "b2"):
        p["head/" + k] = torch.tensor(np.asarray(head[k]), dtype=dtype, requires_grad=True)
    return p


def tail_forward(w: Dict[str, np.ndarray], p: Dict[str, torch.Tensor], x7, dtype=torch.float64) -> torch.Tensor:
    """x7: block6d output [B,2,2,192] (NHWC) -> embedding [B,1024]."""
    x = torch.as_tensor(np.asarray(x7), dtype=dtype).permute(0, 3, 1, 2)                 # NCHW
    sw = lambda t: t * torch.sigmoid(t)                                                # noqa: E731
    s, sh = _bn(w, "block7a_expand_bn", dtype)
    x = sw(F.conv2d(x, p["block7a_expand_conv/kernel"].permute(3, 2, 0, 1)) * s.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    s, sh = _bn(w, "block7a_bn", dtype)
    k = p["block7a_dwconv/depthwise_kernel"].permute(2, 3, 0, 1)                       # [C,1,3,3]
    x = sw(F.conv2d(x, k, padding=1, groups=k.shape[0]) * s.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    se = x.mean(dim=(2, 3), keepdim=True)
    se = sw(F.conv2d(se, p["block7a_se_reduce/kernel"].permute(3, 2, 0, 1)) + p["block7a_se_reduce/bias"].view(1, -1, 1, 1))
    se = torch.sigmoid(F.conv2d(se, p["block7a_se_expand/kernel"].permute(3, 2, 0, 1)) + p["block7a_se_expand/bias"].view(1, -1, 1, 1))
    x = x * se
    s, sh = _bn(w, "block7a_project_bn", dtype)
    x = F.conv2d(x, p["block7a_project_conv/kernel"].permute(3, 2, 0, 1)) * s.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    s, sh = _bn(w, "top_bn", dtype)
    x = sw(F.conv2d(x, p["top_conv/kernel"].permute(3, 2, 0, 1)) * s.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    x = x.mean(dim=(2, 3))
    x = torch.relu(x @ p["dense/kernel"] + p["dense/bias"])
    x = torch.relu(x @ p["dense_1/kernel"] + p["dense_1/bias"])
    z = x @ p["dense_2/kernel"] + p["dense_2/bias"]
    return _SELU_L * torch.where(z > 0, z, _SELU_A * (torch.exp(z) - 1))


def loss_and_grads(w, p, x7, labels, dtype=torch.float64) -> Tuple[float, float, torch.Tensor, Dict[str, np.ndarray]]:
    """Mean cross-entropy (on the logits), accuracy, the embedding and d(mean loss)/d(parameter) for every entry of p."""
    for t in p.values():
        t.grad = None
    emb = tail_forward(w, p, x7, dtype)
    h = torch.tanh(emb @ p["head/w1"] + p["head/b1"])
    z = h @ p["head/w2"] + p["head/b2"]
    y = torch.as_tensor(np.asarray(labels), dtype=torch.long)
    loss = F.cross_entropy(z, y, reduction="mean")
    loss.backward()
    acc = float((z.argmax(1) == y).double().mean())
    return float(loss), acc, emb.detach(), {k: t.grad.detach().numpy().copy() for k, t in p.items()}


class KerasAdam:
    """tf.keras.optimizers.Adam: m, v moments, eps outside the sqrt, lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)."""

    def __init__(self, params: Dict[str, torch.Tensor], lr, b1=0.9, b2=0.999, eps=1e-7):
        self.p, self.lr, self.b1, self.b2, self.eps, self.t = params, lr, b1, b2, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, grads: Dict[str, np.ndarray]):
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for k, t in self.p.items():
                g = torch.as_tensor(grads[k], dtype=t.dtype)
                self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                t.sub_(lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps))
