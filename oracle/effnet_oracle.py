"""oracle/effnet_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain PyTorch fp32/fp64 restatement of the embedding network the reference builds with
``tf.keras.applications.EfficientNetB0(include_top=False, weights=None, input_shape=(49,40,1))`` +
GAP + Dense(2048,relu) + Dense(2048,relu) + Dense(1024,selu)
(multilingual_kws/train_multilingual_embedding.py:66-83; cut at "dense_2",
multilingual_kws/embedding/transfer_learning.py:38-43, distance_filtering.py:21-27).

The layer arithmetic lives in Keras 2.7 (`keras/applications/efficientnet.py`), a third-party
dependency absent from /root/reference; this restates its published architecture:
Rescaling(1/255) -> Normalization(un-adapted = identity) -> ZeroPadding2D(correct_pad)+Conv3x3 s2 VALID
-> BN(eps 1e-3) -> swish -> 16 MBConv (expand 1x1/BN/swish, depthwise kxk [stride 2: correct_pad+VALID,
else SAME]/BN/swish, SE with biases, project 1x1/BN, residual) -> Conv1x1 1280/BN/swish.
PARITY UNPINNED AGAINST KERAS: the pretrained checkpoint is a release asset that is not available offline
and TF cannot run here.  Pins that do exist: Keras' parameter counts / output shapes, and
tests/test_oracle_effnet_torchvision.py — torchvision's own EfficientNet-B0 implementation, loaded with this
oracle's weights and given Keras' BN eps / stride-2 padding convention, reproduces every stage output and the
embedding to 1e-9 (float64).  The oracle is deliberately independent of the product's layer table (it
re-derives shapes on the fly).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

_STAGES = [(3, 1, 32, 16, 1, 1), (3, 2, 16, 24, 6, 2), (5, 2, 24, 40, 6, 2), (3, 3, 40, 80, 6, 2),
           (5, 3, 80, 112, 6, 1), (5, 4, 112, 192, 6, 2), (3, 1, 192, 320, 6, 1)]
_EPS = 1e-3
_SELU_L, _SELU_A = 1.0507009873554805, 1.6732632423543772


def _t(w, name, dtype):
    return torch.as_tensor(np.asarray(w[name]), dtype=dtype)


def _bn(x, w, name, dtype):
    g, b = _t(w, name + "/gamma", dtype), _t(w, name + "/beta", dtype)
    m, v = _t(w, name + "/moving_mean", dtype), _t(w, name + "/moving_variance", dtype)
    return (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + _EPS) * g[None, :, None, None] + \
        b[None, :, None, None]


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(x, w, name, dtype, stride=1, groups=1, depthwise=False):
    k = _t(w, name, dtype)                      # HWIO (depthwise: HWC1)
    if depthwise:
        k = k.permute(2, 3, 0, 1)               # [C,1,kh,kw]
        groups = k.shape[0]
    else:
        k = k.permute(3, 2, 0, 1)               # [O,I,kh,kw]
    return F.conv2d(x, k, stride=stride, groups=groups)


def _pad_for(x, ksize, stride):
    h, w = x.shape[2], x.shape[3]
    c = ksize // 2
    if stride == 2:   # keras correct_pad + VALID
        return F.pad(x, (c - (1 - w % 2), c, c - (1 - h % 2), c))
    return F.pad(x, (c, c, c, c))               # SAME, stride 1, odd kernel


def forward(w: Dict[str, np.ndarray], feats, dtype=torch.float32, taps: Optional[dict] = None, n_dense: int = 3,
            calibrate_bn: bool = False) -> torch.Tensor:
    """feats [B,49,40] (or [B,49,40,1]) log-mel features -> embedding [B, units of dense_{n_dense-1}].

    taps: optional dict filled with intermediate activations (NHWC numpy) keyed by Keras layer name.
    calibrate_bn: test helper — overwrite every BN's moving statistics in `w` with the batch statistics
    of this forward pass (gives a 'trained-like' network whose activations have unit scale)."""
    x = torch.as_tensor(np.asarray(feats), dtype=dtype)
    if x.dim() == 4:
        x = x[..., 0]
    x = x[:, None]                                           # NCHW, C=1
    x = x * (1.0 / 255.0)                                    # Rescaling
    mean, var = _t(w, "normalization/mean", dtype), _t(w, "normalization/variance", dtype)
    x = (x - mean.view(1, -1, 1, 1)) / torch.clamp(torch.sqrt(var.view(1, -1, 1, 1)), min=1e-7)

    def tap(name, t):
        if taps is not None:
            taps[name] = t.permute(0, 2, 3, 1).contiguous().numpy()

    def bn(t, name):
        if calibrate_bn:
            w[name + "/moving_mean"] = t.mean(dim=(0, 2, 3)).to(torch.float32).numpy()
            w[name + "/moving_variance"] = t.var(dim=(0, 2, 3), unbiased=False).to(torch.float32).numpy() + 1e-4
        return _bn(t, w, name, dtype)

    x = _swish(bn(_conv(_pad_for(x, 3, 2), w, "stem_conv/kernel", dtype, stride=2), "stem_bn"))
    tap("stem_activation", x)
    for si, (k, reps, fin, fout, e, s) in enumerate(_STAGES):
        for r in range(reps):
            n = f"block{si + 1}{chr(ord('a') + r)}"
            stride = s if r == 0 else 1
            cin = fin if r == 0 else fout
            inp = x
            if e != 1:
                x = _swish(bn(_conv(x, w, f"{n}_expand_conv/kernel", dtype), f"{n}_expand_bn"))
                tap(f"{n}_expand_activation", x)
            x = _swish(bn(_conv(_pad_for(x, k, stride), w, f"{n}_dwconv/depthwise_kernel", dtype, stride=stride,
                                depthwise=True), f"{n}_bn"))
            tap(f"{n}_activation", x)
            se = x.mean(dim=(2, 3), keepdim=True)
            se = _swish(_conv(se, w, f"{n}_se_reduce/kernel", dtype) + _t(w, f"{n}_se_reduce/bias", dtype).view(1, -1, 1, 1))
            se = torch.sigmoid(_conv(se, w, f"{n}_se_expand/kernel", dtype) + _t(w, f"{n}_se_expand/bias", dtype).view(1, -1, 1, 1))
            x = x * se
            tap(f"{n}_se_excite", x)
            x = bn(_conv(x, w, f"{n}_project_conv/kernel", dtype), f"{n}_project_bn")
            if stride == 1 and cin == fout:
                x = x + inp                                   # dropout inactive at inference
            tap(f"{n}_out", x)
    x = _swish(bn(_conv(x, w, "top_conv/kernel", dtype), "top_bn"))
    tap("top_activation", x)
    x = x.mean(dim=(2, 3))                                    # GlobalAveragePooling2D
    acts = [torch.relu, torch.relu, lambda t: _SELU_L * torch.where(t > 0, t, _SELU_A * (torch.exp(t) - 1))]
    for i in range(n_dense):
        nm = "dense" if i == 0 else f"dense_{i}"
        x = acts[i](x @ _t(w, nm + "/kernel", dtype) + _t(w, nm + "/bias", dtype))
        if taps is not None:
            taps[nm] = x.numpy()
    return x


def cosine(a, b) -> np.ndarray:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1) + 1e-30)
