"""Few-shot head (Dense18-tanh -> Dense3-softmax) and the 3-way keyword model on top of the embedding.

Mirrors the `Sequential[xfer, Dense(18, tanh), Dense(3, softmax)]` model the reference builds and trains at
multilingual_kws/embedding/transfer_learning.py:47-59,86-93.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from . import weights as W
from .model import EmbeddingModel


def glorot_uniform(rng, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


class Head:
    """Device-resident head parameters + Adam state (C ABI kws_head_*)."""

    def __init__(self, w1, b1, w2, b2, beta1=0.9, beta2=0.999, eps=1e-7):
        w1, b1, w2, b2 = (np.ascontiguousarray(a, np.float32) for a in (w1, b1, w2, b2))
        self.in_dim, self.hidden = w1.shape
        self.classes = w2.shape[1]
        self._h = ctypes.c_void_p()
        L = _lib.lib()
        _lib.check(L.kws_head_create(ctypes.byref(self._h), self.in_dim, self.hidden, self.classes, w1.ctypes.data,
                                     b1.ctypes.data, w2.ctypes.data, b2.ctypes.data, beta1, beta2, eps), "kws_head_create")
        self.n_params = int(L.kws_head_num_params(self._h))
        self.flat_size = int(L.kws_head_flat_size(self._h))
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._flat = torch.zeros(self.flat_size, dtype=torch.float32, device=self.device)

    @classmethod
    def keras_init(cls, in_dim=1024, hidden=18, classes=3, seed=None) -> "Head":
        """Keras Dense defaults: glorot_uniform kernels, zero biases."""
        rng = np.random.default_rng(seed)
        return cls(glorot_uniform(rng, in_dim, hidden), np.zeros(hidden, np.float32),
                   glorot_uniform(rng, hidden, classes), np.zeros(classes, np.float32))

    @classmethod
    def from_params(cls, p: Dict[str, np.ndarray]) -> "Head":
        return cls(p["w1"], p["b1"], p["w2"], p["b2"])

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().kws_head_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def step_count(self) -> int:
        return int(_lib.lib().kws_head_step_count(self._h))

    def forward(self, emb: torch.Tensor) -> torch.Tensor:
        emb = emb.to(self.device, torch.float32).contiguous()
        out = torch.empty((emb.shape[0], self.classes), dtype=torch.float32, device=self.device)
        if emb.shape[0]:
            _lib.check(_lib.lib().kws_head_forward(self._h, emb.data_ptr(), emb.shape[0], out.data_ptr(),
                                                   _lib.current_stream_ptr()), "kws_head_forward")
        return out

    def grad(self, emb: torch.Tensor, labels: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Flat [dW1|db1|dW2|db2|loss_sum|n_correct|n] for the local batch (sums -> all-reduce(sum)-able)."""
        emb = emb.to(self.device, torch.float32).contiguous()
        if not (isinstance(labels, torch.Tensor) and labels.is_cuda):
            # Keras' sparse categorical cross-entropy rejects labels outside [0, classes); device-resident labels are
            # not read back here (that would synchronise every step): the caller vouches for them
            lab = torch.as_tensor(np.asarray(labels))
            if lab.numel() and (int(lab.min()) < 0 or int(lab.max()) >= self.classes):
                raise ValueError(f"labels must lie in [0, {self.classes}): got [{int(lab.min())}, {int(lab.max())}]")
            labels = lab
        labels = labels.to(self.device, torch.int32).contiguous()
        flat = out if out is not None else self._flat
        _lib.check(_lib.lib().kws_head_grad(self._h, emb.data_ptr(), labels.data_ptr(), emb.shape[0], flat.data_ptr(),
                                            _lib.current_stream_ptr()), "kws_head_grad")
        return flat

    def apply_adam(self, flat: torch.Tensor, lr: float) -> None:
        _lib.check(_lib.lib().kws_head_apply_adam(self._h, flat.data_ptr(), float(lr), _lib.current_stream_ptr()),
                   "kws_head_apply_adam")

    def reset_optimizer(self) -> None:
        _lib.check(_lib.lib().kws_head_reset_optimizer(self._h))

    def get_params(self) -> np.ndarray:
        out = np.zeros(self.n_params, np.float32)
        _lib.check(_lib.lib().kws_head_get_params(self._h, out.ctypes.data))
        return out

    def params_dict(self) -> Dict[str, np.ndarray]:
        f = self.get_params()
        a, b, c = self.in_dim * self.hidden, self.hidden, self.hidden * self.classes
        return dict(w1=f[:a].reshape(self.in_dim, self.hidden), b1=f[a:a + b],
                    w2=f[a + b:a + b + c].reshape(self.hidden, self.classes), b2=f[a + b + c:])


class FewShotModel:
    """Embedding + head: the object `transfer_learn` returns (`.predict`, `.save`, like the Keras model)."""

    def __init__(self, embedding: EmbeddingModel, head: Head):
        self.embedding, self.head = embedding, head
        self.name = "sequential"

    def forward_device(self, feats: torch.Tensor) -> torch.Tensor:
        return self.head.forward(self.embedding.forward_device(feats))

    def predict(self, specs, batch_size: int = 4096, verbose: int = 0) -> np.ndarray:
        x = torch.as_tensor(np.asarray(specs, dtype=np.float32))
        if x.dim() == 4 and x.shape[-1] == 1:
            x = x[..., 0]
        outs = []
        for i in range(0, x.shape[0], batch_size):
            xb = x[i:i + batch_size].pin_memory().to(self.embedding.device, non_blocking=True)
            outs.append(self.forward_device(xb).cpu())
        if not outs:
            return np.zeros((0, self.head.classes), np.float32)
        return torch.cat(outs).numpy()

    def save(self, path: os.PathLike) -> None:
        """``model.save(path)`` (reference run.py:300): ``path/variables/variables.{index,data-*}`` in the object-graph layout
        Keras gives Sequential[embedding, Dense, Dense] (savedmodel.save_keras_model; no saved_model.pb — that is the traced
        TensorFlow graph) + ``path/weights.npz`` for this package's loader."""
        from .savedmodel import save_keras_model
        os.makedirs(str(path), exist_ok=True)
        w = dict(self.embedding.weights)
        hp = self.head.params_dict()
        save_keras_model(path, w, head=hp)
        for k, v in hp.items():
            w["fewshot_head/" + k] = v
        W.save_npz(os.path.join(str(path), "weights.npz"), w)

    @classmethod
    def load(cls, path: os.PathLike, **kw) -> "FewShotModel":
        """Loads ``path/weights.npz`` or, without it, a few-shot Keras SavedModel directory (``variables/variables.index``:
        the trailing Dense(<=32) -> Dense(<=8) pair becomes the head, savedmodel.split_fewshot_variables)."""
        p = str(path)
        if os.path.isdir(p) and not os.path.isfile(os.path.join(p, "weights.npz")) and \
                os.path.isfile(os.path.join(p, "variables", "variables.index")):
            from .savedmodel import load_keras_variables, split_fewshot_variables
            w, hp = split_fewshot_variables({k: np.asarray(v) for k, v in load_keras_variables(p).items()})
            if hp is None:
                raise ValueError(f"{p}: no few-shot head (Dense -> Dense classifier) in the SavedModel")
            return cls(EmbeddingModel({k: np.asarray(v, np.float32) for k, v in w.items()}, **kw), Head.from_params(hp))
        if os.path.isdir(p):
            p = os.path.join(p, "weights.npz")
        w = W.load_npz(p)
        hp = {k.split("/", 1)[1]: w.pop(k) for k in list(w) if k.startswith("fewshot_head/")}
        return cls(EmbeddingModel(w, **kw), Head.from_params(hp))
