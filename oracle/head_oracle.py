"""oracle/head_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy (float64) restatement of the few-shot head and its optimiser as the reference builds them at
multilingual_kws/embedding/transfer_learning.py:47-59,86-93:
    Sequential[frozen embedding, Dense(18, tanh), Dense(3, softmax)]
    compile(Adam(lr), SparseCategoricalCrossentropy(from_logits=False), ["accuracy"]); fit(...)
The layer/optimiser arithmetic is Keras 2.7 (third-party, absent from /root/reference), restated from
its published definitions:
  * Dense: y = act(x @ kernel + bias); glorot_uniform kernel, zero bias.
  * loss: with a softmax-activated output layer Keras 2.7's backend.sparse_categorical_crossentropy
    takes the cached `_keras_logits` of the softmax and evaluates the cross-entropy ON THE LOGITS
    (tf.nn.sparse_softmax_cross_entropy_with_logits): loss_i = logsumexp(z_i) - z_i[y_i]; batch mean.
    (The clip-to-[1e-7, 1-1e-7] path in SURVEY.md App. B.4 is only taken for outputs that are not a
    Keras softmax; both give the gradient softmax(z) - onehot except where the clip saturates.)
  * Adam (non-amsgrad, epsilon 1e-7): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); theta -= lr_t * m / (sqrt(v) + eps).
PARITY UNPINNED against TF itself (cannot run here); pinned by a torch.autograd gradient check and a
torch.optim-free closed-form Adam check in tests/test_oracle_head.py.
"""
from __future__ import annotations

import numpy as np


def glorot_uniform(rng, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def init_head(seed: int = 0, in_dim: int = 1024, hidden: int = 18, classes: int = 3):
    rng = np.random.default_rng(seed)
    return dict(w1=glorot_uniform(rng, in_dim, hidden), b1=np.zeros(hidden, np.float32),
                w2=glorot_uniform(rng, hidden, classes), b2=np.zeros(classes, np.float32))


def forward(p, emb):
    emb = np.asarray(emb, np.float64)
    h = np.tanh(emb @ p["w1"].astype(np.float64) + p["b1"])
    z = h @ p["w2"].astype(np.float64) + p["b2"]
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return h, z, e / e.sum(axis=1, keepdims=True)


def loss_and_grads(p, emb, labels):
    """Mean sparse CE (logits path) and d(mean loss)/d(params); also batch accuracy."""
    emb = np.asarray(emb, np.float64)
    labels = np.asarray(labels)
    B = emb.shape[0]
    h, z, prob = forward(p, emb)
    lse = np.log(np.exp(z).sum(axis=1))
    loss = float((lse - z[np.arange(B), labels]).mean())
    acc = float((prob.argmax(axis=1) == labels).mean())
    dz2 = prob.copy()
    dz2[np.arange(B), labels] -= 1.0
    dz2 /= B
    g = dict(w2=h.T @ dz2, b2=dz2.sum(0))
    dh = dz2 @ p["w2"].astype(np.float64).T
    dz1 = dh * (1.0 - h * h)
    g["w1"] = emb.T @ dz1
    g["b1"] = dz1.sum(0)
    return loss, acc, g


class Adam:
    def __init__(self, params, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0
        self.m = {k: np.zeros_like(v, np.float64) for k, v in params.items()}
        self.v = {k: np.zeros_like(v, np.float64) for k, v in params.items()}

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k in params:
            g = np.asarray(grads[k], np.float64)
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            params[k] = (params[k].astype(np.float64) - lr_t * self.m[k] / (np.sqrt(self.v[k]) + self.eps)).astype(
                np.float32)


def train(p, emb, labels, steps, lr=1e-3):
    """Full-batch Adam steps; returns the per-step (loss, acc) history (pre-update values)."""
    opt = Adam(p, lr)
    hist = []
    for _ in range(steps):
        loss, acc, g = loss_and_grads(p, emb, labels)
        hist.append((loss, acc))
        opt.step(p, g)
    return hist
