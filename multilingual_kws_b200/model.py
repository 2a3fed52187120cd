"""Host-side model objects over the C ABI: the embedding extractor and (later) the few-shot model.

``EmbeddingModel`` mirrors what the reference obtains from
``distance_filtering.embedding_model()`` (multilingual_kws/embedding/distance_filtering.py:12-27):
an object with ``.predict(specs[N,49,40(,1)]) -> [N,1024]`` and a ``.trainable`` attribute.
"""
from __future__ import annotations

import ctypes
import os
import struct
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from . import weights as W


def pack_weights(w: Dict[str, np.ndarray]) -> bytes:
    """Serialise Keras-named fp32 tensors into the 'KWSW0001' container kws_embed_create parses."""
    parts = [b"KWSW0001", struct.pack("<I", len(w))]
    for name, arr in w.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        nb = name.encode("utf-8")
        parts.append(struct.pack("<I", len(nb)))
        parts.append(nb + b"\0" * ((4 - len(nb) % 4) % 4))
        parts.append(struct.pack("<I", a.ndim))
        parts.append(struct.pack(f"<{a.ndim}I", *a.shape) if a.ndim else b"")
        parts.append(a.tobytes())
    return b"".join(parts)


def cut_at(w: Dict[str, np.ndarray], output_layer: str) -> Dict[str, np.ndarray]:
    """Drops every Dense layer after `output_layer` ("dense", "dense_1", ...) and the few-shot head entries."""
    names = sorted({k.split("/")[0] for k in w if k.startswith("dense")}, key=lambda n: int(n.split("_")[1]) if "_" in n else 0)
    if output_layer not in names:
        raise ValueError(f"No such layer: {output_layer}. Existing dense layers are: {names}")
    keep = set(names[:names.index(output_layer) + 1])
    return {k: v for k, v in w.items()
            if not k.startswith("fewshot_head/") and (not k.startswith("dense") or k.split("/")[0] in keep)}


class EmbeddingModel:
    """EfficientNet-B0 embedding tower resident on the current CUDA device (inference mode, BN folded)."""

    def __init__(self, weights: Dict[str, np.ndarray], chunk: Optional[int] = None, dtype: str = "fp16"):
        """dtype: 16-bit activation / tensor-core operand type, "fp16" (default, 11-bit significand) or "bf16"."""
        if dtype not in ("fp16", "bf16"):
            raise ValueError("dtype must be 'fp16' or 'bf16'")
        self.dtype = dtype
        self.weights = weights
        self.trainable = False          # the reference sets embedding.trainable = False (distance_filtering.py:26)
        self.name = "TransferLearnedModel"
        blob = pack_weights(weights)
        self._h = ctypes.c_void_p()
        L = _lib.lib()
        _lib.check(L.kws_embed_create(ctypes.byref(self._h), blob, len(blob), 0 if dtype == "fp16" else 1), "kws_embed_create")
        h, w_, d, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        fl = ctypes.c_double()
        _lib.check(L.kws_embed_info(self._h, ctypes.byref(h), ctypes.byref(w_), ctypes.byref(d), ctypes.byref(n),
                                    ctypes.byref(fl)))
        self.input_hw, self.output_dim, self.n_ops, self.flops_per_clip = (h.value, w_.value), d.value, n.value, fl.value
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ws = None
        if chunk is not None:
            self.set_chunk(chunk)

    @classmethod
    def random(cls, seed: int = 0, **kw) -> "EmbeddingModel":
        return cls(W.random_init(seed), **kw)

    @classmethod
    def load(cls, path: os.PathLike, output_layer: str = "dense_2", **kw) -> "EmbeddingModel":
        """Loads a saved weight container and cuts the dense tower at `output_layer` (the reference cuts its classifier
        at "dense_2", transfer_learning.py:38-43).  `path` is a directory holding weights.npz, the .npz itself, or a
        Keras SavedModel directory (variables/variables.index + data shard, read without TensorFlow by savedmodel.py —
        see the status note there)."""
        p = str(path)
        if os.path.isdir(p) and not os.path.isfile(os.path.join(p, "weights.npz")) and \
                os.path.isfile(os.path.join(p, "variables", "variables.index")):
            from .savedmodel import load_keras_variables, split_fewshot_variables
            # Dense layers are resolved by ORDER (Keras suffixes them per session: dense_7, dense_8 ...); a few-shot
            # model's trailing head is dropped here (FewShotModel.load keeps it)
            w, _ = split_fewshot_variables({k: np.asarray(v) for k, v in load_keras_variables(p).items()})
            w = {k: np.asarray(v, np.float32) for k, v in w.items()}
            return cls(cut_at(w, output_layer), **kw)
        if os.path.isdir(p):
            p = os.path.join(p, "weights.npz")
        if not os.path.isfile(p):
            raise FileNotFoundError(f"no weight container at {p} (expected a directory holding weights.npz or a Keras "
                                    "SavedModel's variables/)")
        return cls(cut_at(W.load_npz(p), output_layer), **kw)

    def save(self, path: os.PathLike) -> None:
        """weights.npz + the variables checkpoint in Keras' object-graph layout (savedmodel.save_keras_model)."""
        from .savedmodel import save_keras_model
        os.makedirs(str(path), exist_ok=True)
        save_keras_model(path, self.weights)
        W.save_npz(os.path.join(str(path), "weights.npz"), self.weights)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().kws_embed_destroy(h)
            except Exception:
                pass
            self._h = None

    def set_chunk(self, chunk: int) -> None:
        _lib.check(_lib.lib().kws_embed_set_chunk(self._h, int(chunk)))
        self._ws = None

    def set_chunk_late(self, chunk: int) -> None:
        _lib.check(_lib.lib().kws_embed_set_chunk_late(self._h, int(chunk)))
        self._ws = None

    def launches(self, batch: int) -> int:
        return int(_lib.lib().kws_embed_launches(self._h, int(batch)))

    def set_graph(self, enable: bool) -> None:
        _lib.check(_lib.lib().kws_embed_set_graph(self._h, int(bool(enable))))

    def set_fuse(self, mode: int) -> None:
        """Tail schedule: 0 layer by layer, 1 one fused launch per MBConv block, 2 runs of blocks per launch (default)."""
        _lib.check(_lib.lib().kws_embed_set_fuse(self._h, int(mode)))

    def op_names(self):
        L = _lib.lib()
        out = []
        for i in range(self.n_ops):
            buf = ctypes.create_string_buffer(128)
            n = ctypes.c_int64()
            _lib.check(L.kws_embed_op_name(self._h, i, buf, 128, ctypes.byref(n)))
            out.append((buf.value.decode(), int(n.value)))
        return out

    def op_info(self):
        """[(name, kind, flops_per_clip, bytes_per_clip, N, K, rows_per_clip)] per op; kind 0 stem, 1 gemm, 2 dw+SE."""
        L = _lib.lib()
        out = []
        for i, (name, _) in enumerate(self.op_names()):
            kind, n, k, r = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            fl, by = ctypes.c_double(), ctypes.c_double()
            _lib.check(L.kws_embed_op_info(self._h, i, ctypes.byref(kind), ctypes.byref(fl), ctypes.byref(by),
                                           ctypes.byref(n), ctypes.byref(k), ctypes.byref(r)))
            out.append((name, kind.value, fl.value, by.value, n.value, k.value, r.value))
        return out

    def forward_timed(self, feats: torch.Tensor):
        """Profiling pass: returns (embeddings, per-op device milliseconds as a numpy array)."""
        feats = feats.to(device=self.device, dtype=torch.float32).contiguous()
        B = feats.shape[0]
        out = torch.empty((B, self.output_dim), dtype=torch.float32, device=self.device)
        ws = self._workspace(B)
        ms = np.zeros(self.n_ops, np.float32)
        _lib.check(_lib.lib().kws_embed_forward_timed(self._h, feats.data_ptr(), B, out.data_ptr(), ws.data_ptr(),
                                                      ws.numel(), ms.ctypes.data, _lib.current_stream_ptr()),
                   "kws_embed_forward_timed")
        return out, ms

    def _workspace(self, batch: int) -> torch.Tensor:
        need = int(_lib.lib().kws_embed_workspace_bytes(self._h, batch))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def workspace_bytes(self, batch: int) -> int:
        return int(_lib.lib().kws_embed_workspace_bytes(self._h, int(batch)))

    def forward_device(self, feats: torch.Tensor, out: Optional[torch.Tensor] = None, tap_op: int = -1,
                       workspace: Optional[torch.Tensor] = None, sm_budget: Optional[tuple] = None):
        """feats: CUDA float32 [B,49,40] (contiguous) -> CUDA float32 [B, output_dim].  With tap_op >= 0 also
        returns that op's output (bf16 [B, elems], fp32 for the last op).  `workspace` (uint8 CUDA tensor of at least
        workspace_bytes(B)) replaces the model's own scratch: forwards that run concurrently on different streams
        must not share one.  `sm_budget` = (head SMs, tail SMs) selects the throughput schedule
        (kws_embed_forward_budget) for callers that overlap several passes; results are identical."""
        if feats.dim() == 4 and feats.shape[-1] == 1:
            feats = feats[..., 0]
        if feats.dim() != 3 or tuple(feats.shape[1:]) != self.input_hw:
            raise ValueError(f"expected [N,{self.input_hw[0]},{self.input_hw[1]}(,1)] features, got {tuple(feats.shape)}")
        feats = feats.to(device=self.device, dtype=torch.float32).contiguous()
        B = feats.shape[0]
        if out is None:
            out = torch.empty((B, self.output_dim), dtype=torch.float32, device=self.device)
        ws = self._workspace(B) if workspace is None else workspace
        if ws.numel() < self.workspace_bytes(B):
            raise ValueError("workspace too small")
        tap = None
        tap_ptr = None
        if tap_op >= 0:
            name, elems = self.op_names()[tap_op]
            last = tap_op == self.n_ops - 1
            tap = torch.empty((B, elems), dtype=torch.float32 if last else (torch.float16 if self.dtype == "fp16" else torch.bfloat16),
                              device=self.device)
            tap_ptr = tap.data_ptr()
        if B and sm_budget is not None and tap_op < 0:
            _lib.check(_lib.lib().kws_embed_forward_budget(self._h, feats.data_ptr(), B, out.data_ptr(), ws.data_ptr(),
                                                           ws.numel(), int(sm_budget[0]), int(sm_budget[1]),
                                                           _lib.current_stream_ptr()), "kws_embed_forward_budget")
        elif B:
            _lib.check(_lib.lib().kws_embed_forward_tap(self._h, feats.data_ptr(), B, out.data_ptr(), ws.data_ptr(),
                                                        ws.numel(), int(tap_op), tap_ptr, _lib.current_stream_ptr()),
                       "kws_embed_forward")
        return (out, tap) if tap_op >= 0 else out

    def forward_until(self, feats: torch.Tensor, tap_op: int) -> torch.Tensor:
        """Runs only the ops up to `tap_op` and returns that op's output (16-bit [B, elems]): the frozen part of the
        network in front of a trainable tail (finetune.TailTrainer)."""
        if feats.dim() == 4 and feats.shape[-1] == 1:
            feats = feats[..., 0]
        feats = feats.to(device=self.device, dtype=torch.float32).contiguous()
        B = feats.shape[0]
        _, elems = self.op_names()[tap_op]
        tap = torch.empty((B, elems), dtype=torch.float16 if self.dtype == "fp16" else torch.bfloat16, device=self.device)
        if B:
            ws = self._workspace(B)
            _lib.check(_lib.lib().kws_embed_forward_until(self._h, feats.data_ptr(), B, ws.data_ptr(), ws.numel(), int(tap_op),
                                                          tap.data_ptr(), _lib.current_stream_ptr()), "kws_embed_forward_until")
        return tap

    def predict(self, specs, batch_size: int = 4096, verbose: int = 0) -> np.ndarray:
        """Keras-style predict: host array [N,49,40] / [N,49,40,1] -> np.float32 [N, output_dim]."""
        x = torch.as_tensor(np.asarray(specs, dtype=np.float32))
        if x.dim() == 4 and x.shape[-1] == 1:
            x = x[..., 0]
        outs = []
        for i in range(0, x.shape[0], batch_size):
            xb = x[i:i + batch_size].pin_memory().to(self.device, non_blocking=True)
            outs.append(self.forward_device(xb).cpu())
        if not outs:
            return np.zeros((0, self.output_dim), np.float32)
        return torch.cat(outs).numpy()

    __call__ = forward_device


def gemm_h16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = 0,
             residual: Optional[torch.Tensor] = None, out_f32: bool = False, gap4: bool = False,
             block_n: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The tcgen05 pointwise/dense operator: act(a @ w.T + bias) (+ residual).  a [M,K], w [N,K] CUDA fp16 or bf16."""
    assert a.dtype in (torch.float16, torch.bfloat16) and w.dtype == a.dtype and a.is_cuda and w.is_cuda
    a, w = a.contiguous(), w.contiguous()
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M // 4 if gap4 else M, N), dtype=torch.float32 if out_f32 else a.dtype, device=a.device)
    _lib.check(_lib.lib().kws_gemm_h16(a.data_ptr(), w.data_ptr(), M, N, K, bias.data_ptr() if bias is not None else None,
                                       int(act), residual.data_ptr() if residual is not None else None, out.data_ptr(),
                                       int(out_f32), int(gap4), int(block_n), 0 if a.dtype == torch.float16 else 1,
                                       _lib.current_stream_ptr()), "kws_gemm_h16")
    return out
