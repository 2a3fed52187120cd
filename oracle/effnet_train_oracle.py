"""oracle/effnet_train_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Trains the torch restatement of the embedding network (same architecture as oracle/effnet_oracle.py: Keras
EfficientNetB0 at 49x40x1 + GAP + Dense 2048/2048/1024, reference train_multilingual_embedding.py:66-83) for a few hundred
Adam steps on a synthetic 4-way task, BatchNorm in training mode (batch statistics, moving averages with Keras' momentum
0.99 — the mode of reference train_multilingual_embedding.py:99-133).  The result is a *trained* set of weights: the
third weight regime of the embedding parity tests (random init with damped residual branches, undamped random init,
trained).  The released checkpoint is not available offline, so this is the closest stand-in for "trained weights".
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

_STAGES = [(3, 1, 32, 16, 1, 1), (3, 2, 16, 24, 6, 2), (5, 2, 24, 40, 6, 2), (3, 3, 40, 80, 6, 2),
           (5, 3, 80, 112, 6, 1), (5, 4, 112, 192, 6, 2), (3, 1, 192, 320, 6, 1)]
_EPS, _MOM = 1e-3, 0.99
_SELU_L, _SELU_A = 1.0507009873554805, 1.6732632423543772


def _pad(x, k, stride):
    h, w = x.shape[2], x.shape[3]
    c = k // 2
    if stride == 2:
        return F.pad(x, (c - (1 - w % 2), c, c - (1 - h % 2), c))
    return F.pad(x, (c, c, c, c))


class Net:
    def __init__(self, w: Dict[str, np.ndarray], device, n_classes: int = 4, seed: int = 0):
        self.dev = device
        self.p, self.buf = {}, {}
        for k, v in w.items():
            t = torch.tensor(np.asarray(v, np.float32), device=device)
            if k.endswith(("moving_mean", "moving_variance")) or k.startswith("normalization"):
                self.buf[k] = t
            else:
                self.p[k] = t.requires_grad_(True)
        g = torch.Generator().manual_seed(seed)
        self.p["cls/kernel"] = (torch.randn(w["dense_2/kernel"].shape[1], n_classes, generator=g) * 0.03).to(device).requires_grad_(True)
        self.p["cls/bias"] = torch.zeros(n_classes, device=device, requires_grad=True)

    def bn(self, x, name, training):
        g, b = self.p[name + "/gamma"], self.p[name + "/beta"]
        if training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            with torch.no_grad():
                self.buf[name + "/moving_mean"].mul_(_MOM).add_(mean, alpha=1 - _MOM)
                self.buf[name + "/moving_variance"].mul_(_MOM).add_(var, alpha=1 - _MOM)
        else:
            mean, var = self.buf[name + "/moving_mean"], self.buf[name + "/moving_variance"]
        return (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + _EPS) * g[None, :, None, None] + b[None, :, None, None]

    def conv(self, x, name, stride=1, dw=False):
        k = self.p[name]
        if dw:
            k = k.permute(2, 3, 0, 1)
            return F.conv2d(x, k, stride=stride, groups=k.shape[0])
        return F.conv2d(x, k.permute(3, 2, 0, 1), stride=stride)

    def forward(self, feats, training: bool):
        sw = lambda t: t * torch.sigmoid(t)                                    # noqa: E731
        x = torch.as_tensor(feats, dtype=torch.float32, device=self.dev)[:, None] * (1.0 / 255.0)
        x = sw(self.bn(self.conv(_pad(x, 3, 2), "stem_conv/kernel", stride=2), "stem_bn", training))
        for si, (k, reps, fin, fout, e, s) in enumerate(_STAGES):
            for r in range(reps):
                n = f"block{si + 1}{chr(ord('a') + r)}"
                stride = s if r == 0 else 1
                cin = fin if r == 0 else fout
                inp = x
                if e != 1:
                    x = sw(self.bn(self.conv(x, f"{n}_expand_conv/kernel"), f"{n}_expand_bn", training))
                x = sw(self.bn(self.conv(_pad(x, k, stride), f"{n}_dwconv/depthwise_kernel", stride=stride, dw=True), f"{n}_bn", training))
                se = x.mean(dim=(2, 3), keepdim=True)
                se = sw(self.conv(se, f"{n}_se_reduce/kernel") + self.p[f"{n}_se_reduce/bias"].view(1, -1, 1, 1))
                se = torch.sigmoid(self.conv(se, f"{n}_se_expand/kernel") + self.p[f"{n}_se_expand/bias"].view(1, -1, 1, 1))
                x = self.bn(self.conv(x * se, f"{n}_project_conv/kernel"), f"{n}_project_bn", training)
                if stride == 1 and cin == fout:
                    x = x + inp
        x = sw(self.bn(self.conv(x, "top_conv/kernel"), "top_bn", training)).mean(dim=(2, 3))
        x = torch.relu(x @ self.p["dense/kernel"] + self.p["dense/bias"])
        x = torch.relu(x @ self.p["dense_1/kernel"] + self.p["dense_1/bias"])
        z = x @ self.p["dense_2/kernel"] + self.p["dense_2/bias"]
        emb = _SELU_L * torch.where(z > 0, z, _SELU_A * (torch.exp(z) - 1))
        return emb, emb @ self.p["cls/kernel"] + self.p["cls/bias"]

    def weights(self) -> Dict[str, np.ndarray]:
        out = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in self.p.items() if not k.startswith("cls/")}
        out.update({k: v.detach().cpu().numpy().astype(np.float32) for k, v in self.buf.items()})
        return out


def train(w: Dict[str, np.ndarray], feats: np.ndarray, labels: np.ndarray, steps: int = 200, batch: int = 32, lr: float = 1e-3,
          seed: int = 0, device=None, log=None) -> Dict[str, np.ndarray]:
    """Adam on softmax cross-entropy of a linear classifier over the embedding; returns the trained Keras-named weights
    (moving statistics included)."""
    device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
    net = Net(w, device, int(labels.max()) + 1, seed)
    opt = torch.optim.Adam(list(net.p.values()), lr=lr, eps=1e-7)
    rng = np.random.default_rng(seed)
    y_all = torch.as_tensor(labels, dtype=torch.long, device=device)
    for step in range(steps):
        idx = rng.choice(feats.shape[0], batch, replace=False)
        _, logits = net.forward(feats[idx], training=True)
        loss = F.cross_entropy(logits, y_all[idx])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if log is not None and (step % 25 == 0 or step + 1 == steps):
            log(f"train-oracle step {step}: loss {float(loss):.4f} acc {float((logits.argmax(1) == y_all[idx]).float().mean()):.3f}")
    return net.weights()
