"""Fine-tuning through the top of the embedding: phase 2 of the reference's ``transfer_learn``.

Reference multilingual_kws/embedding/transfer_learning.py:97-112 recompiles the few-shot model with
``Adam(embedding_lr)`` after "unfreez[ing] the top 20 layers while leaving BatchNorm layers frozen" and fits again.
The last 20 layers of the embedding (cut at ``dense_2``) are exactly block7a (13 layers: expand conv / BN / swish,
depthwise conv / BN / swish, SE squeeze / reshape / reduce / expand / excite, project conv / BN), top_conv / top_bn /
top_activation, the global average pool and the three Dense layers — north_star's "head + last block".  (As written the
reference's loop walks ``xfer.layers[-20:]`` of the 3-layer Sequential and therefore flips the whole nested embedding,
BatchNorm included, to trainable; this module implements the stated intent, see DESIGN.md.)

``TailTrainer`` holds fp32 master copies + Adam moments of those layers' kernels / biases on the device, runs their
forward with activations kept for the backward, the backward (every contraction — forward, data gradient, weight
gradient — on the tcgen05 GEMM ``kws_gemm_h16``; the kernels in between are ``kws_train_*``), ONE all-reduce of the flat
fp32 gradient buffer (tail + head, ~10.1 M floats) under torch.distributed, and Keras-Adam.  BatchNorm layers stay frozen:
their inference-mode scale is folded into the 16-bit forward weights and applied to the gradients in the Adam kernel.
Everything below block7a runs through the frozen ``EmbeddingModel`` (its block6d output is the trainable tail's input).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .fewshot import Head
from .model import EmbeddingModel, gemm_h16
from .weights import BN_EPS

ACT_NONE, ACT_SWISH, ACT_RELU, ACT_SELU, ACT_SIGMOID = 0, 1, 2, 3, 4


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _st():
    return _lib.current_stream_ptr()


def _bn_fold(w: Dict[str, np.ndarray], name: str) -> Tuple[np.ndarray, np.ndarray]:
    g, b = w[name + "/gamma"].astype(np.float32), w[name + "/beta"].astype(np.float32)
    mu, var = w[name + "/moving_mean"].astype(np.float32), w[name + "/moving_variance"].astype(np.float32)
    s = g / np.sqrt(var + np.float32(BN_EPS))
    return s.astype(np.float32), (b - mu * s).astype(np.float32)


class _Param:
    """One trainable tensor: fp32 master [rows, cols] (rows = outputs: the GEMM's [N][K] layout), Adam moments, the
    frozen per-row scale, the 16-bit forward copy and (for data gradients) its transpose."""

    def __init__(self, name: str, master: np.ndarray, row_scale: Optional[np.ndarray], dev, need_t: bool, keras_key: str,
                 keras_shape, out32: bool = False):
        self.name, self.keras_key, self.keras_shape = name, keras_key, tuple(keras_shape)
        self.master = torch.from_numpy(np.ascontiguousarray(master, np.float32)).to(dev)
        self.rows, self.cols = self.master.shape
        self.m = torch.zeros_like(self.master)
        self.v = torch.zeros_like(self.master)
        self.scale = None if row_scale is None else torch.from_numpy(np.ascontiguousarray(row_scale, np.float32)).to(dev)
        folded = self.master if self.scale is None else self.master * self.scale.view(-1, 1 if self.scale.numel() == self.rows else self.cols)
        self.w16 = folded.to(torch.float16).contiguous()
        self.w32 = folded.contiguous().clone() if out32 else None
        self.w16_t = torch.empty((self.cols, self.rows), dtype=torch.float16, device=dev) if need_t else None
        self.grad: Optional[torch.Tensor] = None      # view into the flat gradient buffer
        self.refresh_transpose()

    def refresh_transpose(self):
        if self.w16_t is not None:
            _lib.check(_lib.lib().kws_train_transpose_h16(self.w16.data_ptr(), self.rows, self.cols, self.w16_t.data_ptr(),
                                                          self.rows, _st()))


class TailTrainer:
    """Trainable top of the embedding (block7a + top conv + dense tower) and the few-shot head."""

    def __init__(self, embedding: EmbeddingModel, head: Head, loss_scale: float = 256.0, block: str = "block7a"):
        if embedding.dtype != "fp16":
            raise ValueError("fine-tuning through the embedding needs the fp16 model (gradients are loss-scaled halves)")
        self.embedding, self.head = embedding, head
        self.loss_scale = float(loss_scale)
        self.dev = embedding.device
        w = embedding.weights
        names = [n for n, _ in embedding.op_names()]
        prev = {"block7a": "block6d_out"}[block]
        self.tap_op = names.index(prev)
        self.block = block
        ke = w[f"{block}_expand_conv/kernel"]
        self.cin, self.cexp = int(ke.shape[2]), int(ke.shape[3])
        kd = w[f"{block}_dwconv/depthwise_kernel"]
        self.K = int(kd.shape[0])
        self.se = int(w[f"{block}_se_reduce/kernel"].shape[3])
        if self.se % 8:
            raise ValueError("squeeze width must be a multiple of 8 for the tcgen05 GEMM")
        self.cout = int(w[f"{block}_project_conv/kernel"].shape[3])
        self.ctop = int(w["top_conv/kernel"].shape[3])
        self.H = self.W = 2                    # block7a's map at 49x40 input
        self.P = self.H * self.W
        dense = [n for n in ("dense", "dense_1", "dense_2", "dense_3") if n + "/kernel" in w]
        self.dense_names = dense
        if dense != ["dense", "dense_1", "dense_2"]:
            raise ValueError("fine-tuning expects the embedding cut at dense_2 (relu, relu, selu tower)")
        sc_e, self.sh_e = _bn_fold(w, f"{block}_expand_bn")
        sc_d, self.sh_d = _bn_fold(w, f"{block}_bn")
        sc_p, self.sh_p = _bn_fold(w, f"{block}_project_bn")
        sc_t, self.sh_t = _bn_fold(w, "top_bn")
        dev = self.dev
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)     # noqa: E731
        self.sh_e, self.sh_d, self.sh_p, self.sh_t = T(self.sh_e), T(self.sh_d), T(self.sh_p), T(self.sh_t)
        C = self.cexp
        P_ = _Param
        self.p_exp = P_("exp", ke.reshape(self.cin, C).T, sc_e, dev, False, f"{block}_expand_conv/kernel", ke.shape)
        # depthwise: master [K*K][C] (tap-major), the BN scale is per column -> expanded to one factor per element
        self.p_dw = P_("dw", kd.reshape(self.K * self.K, C).reshape(-1, 1), np.tile(sc_d, self.K * self.K), dev, False,
                       f"{block}_dwconv/depthwise_kernel", kd.shape, out32=True)
        self.p_se1 = P_("se1", w[f"{block}_se_reduce/kernel"].reshape(C, self.se).T, None, dev, True,
                        f"{block}_se_reduce/kernel", w[f"{block}_se_reduce/kernel"].shape)
        self.p_se1b = P_("se1b", w[f"{block}_se_reduce/bias"].reshape(1, -1), None, dev, False, f"{block}_se_reduce/bias", (self.se,))
        self.p_se2 = P_("se2", w[f"{block}_se_expand/kernel"].reshape(self.se, C).T, None, dev, True,
                        f"{block}_se_expand/kernel", w[f"{block}_se_expand/kernel"].shape)
        self.p_se2b = P_("se2b", w[f"{block}_se_expand/bias"].reshape(1, -1), None, dev, False, f"{block}_se_expand/bias", (C,))
        kp = w[f"{block}_project_conv/kernel"]
        self.p_proj = P_("proj", kp.reshape(C, self.cout).T, sc_p, dev, True, f"{block}_project_conv/kernel", kp.shape)
        kt = w["top_conv/kernel"]
        self.p_top = P_("top", kt.reshape(self.cout, self.ctop).T, sc_t, dev, True, "top_conv/kernel", kt.shape)
        self.p_dense: List[Tuple[_Param, _Param]] = []
        for n in dense:
            k = w[n + "/kernel"]
            self.p_dense.append((P_(n, k.T, None, dev, True, n + "/kernel", k.shape),
                                 P_(n + "_b", w[n + "/bias"].reshape(1, -1), None, dev, False, n + "/bias", w[n + "/bias"].shape)))
        self.out_dim = int(w[dense[-1] + "/kernel"].shape[1])
        self.params: List[_Param] = [self.p_exp, self.p_dw, self.p_se1, self.p_se1b, self.p_se2, self.p_se2b, self.p_proj,
                                     self.p_top] + [q for pair in self.p_dense for q in pair]
        self.n_tail = sum(p.master.numel() for p in self.params)
        # flat fp32 gradient buffer: [tail parameters ... | head flat (dW1 db1 dW2 db2 loss_sum correct count)]
        self.flat = torch.zeros(self.n_tail + head.flat_size, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.master.numel()].view(p.rows, p.cols)
            off += p.master.numel()
        self.head_flat = self.flat[self.n_tail:]
        self.count = self.head_flat[head.n_params + 2:head.n_params + 3]
        self.t = 0
        self.d_step = torch.zeros(1, dtype=torch.int64, device=dev)       # Adam iteration count and step size on the device:
        self.d_lr_t = torch.zeros(1, dtype=torch.float32, device=dev)     # a captured step replays without the host
        self._scratch = None
        self.saved = None
        self._graph = None

    # ------------------------------------------------------------------ helpers
    def _tr(self, x: torch.Tensor) -> torch.Tensor:
        """[R, C] fp16 -> [C, R] (the weight-gradient GEMMs contract over the batch, which must be the K-major axis)."""
        r, c = x.shape
        r_pad = (r + 7) & ~7                    # the GEMM wants K % 8 == 0: zero columns add nothing to the sums
        out = (torch.empty if r_pad == r else torch.zeros)((c, r_pad), dtype=torch.float16, device=self.dev)
        _lib.check(_lib.lib().kws_train_transpose_h16(x.data_ptr(), r, c, out.data_ptr(), r_pad, _st()))
        return out

    def _wgrad(self, dz: torch.Tensor, x: torch.Tensor, p: _Param) -> None:
        """p.grad[out, in] = sum_rows dz[row, out] x[row, in]   (fp32, written straight into the flat buffer)."""
        gemm_h16(self._tr(dz), self._tr(x), out_f32=True, out=p.grad)

    def _colsum(self, x: torch.Tensor, p: _Param) -> None:
        _lib.check(_lib.lib().kws_train_colsum(x.data_ptr(), x.shape[0], x.shape[1], p.grad.data_ptr(), _st()))

    def _act_bwd(self, kind: int, dy: torch.Tensor, ref: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        dz = torch.empty(ref.shape, dtype=torch.float16, device=self.dev)
        _lib.check(_lib.lib().kws_train_act_bwd(kind, dy.data_ptr(), ref.data_ptr(), ref.numel(), float(scale), dz.data_ptr(), _st()))
        return dz

    def _swish(self, z: torch.Tensor) -> torch.Tensor:
        a = torch.empty_like(z)
        _lib.check(_lib.lib().kws_train_swish_fwd(z.data_ptr(), z.numel(), a.data_ptr(), _st()))
        return a

    # ------------------------------------------------------------------ forward
    def tail_input(self, feats: torch.Tensor) -> torch.Tensor:
        """Frozen part of the network: features -> block6d output [B * P, cin] fp16."""
        return self.embedding.forward_until(feats, self.tap_op).view(-1, self.cin)

    def forward_tail(self, x7: torch.Tensor, keep: bool) -> torch.Tensor:
        L = _lib.lib()
        R = x7.shape[0]
        B, P, C = R // self.P, self.P, self.cexp
        h16 = lambda *s: torch.empty(s, dtype=torch.float16, device=self.dev)     # noqa: E731
        z_e = gemm_h16(x7, self.p_exp.w16, bias=self.sh_e)
        E = self._swish(z_e)
        z_d, D, pooled = h16(R, C), h16(R, C), h16(B, C)
        _lib.check(L.kws_train_dw_fwd(E.data_ptr(), self.p_dw.w32.data_ptr(), self.sh_d.data_ptr(), B, C, self.H, self.W, self.K,
                                      1, self.K // 2, self.K // 2, z_d.data_ptr(), D.data_ptr(), pooled.data_ptr(), _st()))
        s_pre = gemm_h16(pooled, self.p_se1.w16, bias=self.p_se1b.master.view(-1))
        s = self._swish(s_pre)
        g = gemm_h16(s, self.p_se2.w16, bias=self.p_se2b.master.view(-1), act=ACT_SIGMOID)
        Dg = h16(R, C)
        _lib.check(L.kws_train_gate_fwd(D.data_ptr(), g.data_ptr(), B, P, C, Dg.data_ptr(), _st()))
        Pout = gemm_h16(Dg, self.p_proj.w16, bias=self.sh_p)
        Tpre = gemm_h16(Pout, self.p_top.w16, bias=self.sh_t)
        h = h16(B, self.ctop)
        _lib.check(L.kws_train_gap_swish_fwd(Tpre.data_ptr(), B, P, self.ctop, h.data_ptr(), _st()))
        hs = [h]
        for i, (pw, pb) in enumerate(self.p_dense):
            last = i + 1 == len(self.p_dense)
            h = gemm_h16(h, pw.w16, bias=pb.master.view(-1), act=ACT_SELU if last else ACT_RELU, out_f32=last)
            hs.append(h)
        if keep:
            self.saved = dict(x7=x7, z_e=z_e, E=E, z_d=z_d, D=D, pooled=pooled, s_pre=s_pre, s=s, g=g, Dg=Dg, Pout=Pout,
                              Tpre=Tpre, hs=hs, B=B)
        return hs[-1]

    def embed(self, feats: torch.Tensor) -> torch.Tensor:
        """Embedding [B, out_dim] fp32 with the CURRENT (fine-tuned) tail weights."""
        return self.forward_tail(self.tail_input(feats), keep=False)

    def predict_probs(self, feats: torch.Tensor) -> torch.Tensor:
        return self.head.forward(self.embed(feats))

    # ------------------------------------------------------------------ backward
    def backward(self, labels: torch.Tensor) -> None:
        """Fills self.flat with the (loss-scaled, summed over the local batch) gradients of tail + head."""
        L = _lib.lib()
        sv = self.saved
        B, P, C = sv["B"], self.P, self.cexp
        hs = sv["hs"]
        emb = hs[-1]
        self.head.grad(emb, labels, out=self.head_flat)
        demb = torch.empty((B, self.out_dim), dtype=torch.float32, device=self.dev)
        _lib.check(L.kws_head_input_grad(self.head._h, B, demb.data_ptr(), _st()))
        # dense tower, top to bottom
        dz = self._act_bwd(1, demb, emb, self.loss_scale)
        for i in range(len(self.p_dense) - 1, -1, -1):
            pw, pb = self.p_dense[i]
            self._wgrad(dz, hs[i], pw)
            self._colsum(dz, pb)
            dh = gemm_h16(dz, pw.w16_t)
            if i > 0:
                dz = self._act_bwd(0, dh, hs[i])
        # global average pool + top conv
        dT = torch.empty_like(sv["Tpre"])
        _lib.check(L.kws_train_gap_swish_bwd(dh.data_ptr(), sv["Tpre"].data_ptr(), B, P, self.ctop, dT.data_ptr(), _st()))
        self._wgrad(dT, sv["Pout"], self.p_top)
        dP = gemm_h16(dT, self.p_top.w16_t)
        # project conv
        self._wgrad(dP, sv["Dg"], self.p_proj)
        dDg = gemm_h16(dP, self.p_proj.w16_t)
        # squeeze-excite
        dD = torch.empty_like(dDg)
        dg_pre = torch.empty((B, C), dtype=torch.float16, device=self.dev)
        _lib.check(L.kws_train_gate_bwd(dDg.data_ptr(), sv["D"].data_ptr(), sv["g"].data_ptr(), B, P, C, dD.data_ptr(),
                                        dg_pre.data_ptr(), _st()))
        self._wgrad(dg_pre, sv["s"], self.p_se2)
        self._colsum(dg_pre, self.p_se2b)
        ds = gemm_h16(dg_pre, self.p_se2.w16_t)
        ds_pre = self._act_bwd(2, ds, sv["s_pre"])
        self._wgrad(ds_pre, sv["pooled"], self.p_se1)
        self._colsum(ds_pre, self.p_se1b)
        dp = gemm_h16(ds_pre, self.p_se1.w16_t)
        # depthwise conv
        need = int(L.kws_train_dw_bwd_scratch_floats(B, C, self.K))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.float32, device=self.dev)
        dE = torch.empty_like(sv["E"])
        _lib.check(L.kws_train_dw_bwd(dD.data_ptr(), dp.data_ptr(), sv["z_d"].data_ptr(), sv["E"].data_ptr(),
                                      self.p_dw.w32.data_ptr(), B, C, self.H, self.W, self.K, 1, self.K // 2, self.K // 2,
                                      dE.data_ptr(), self.p_dw.grad.data_ptr(), self._scratch.data_ptr(), _st()))
        # expand conv (its input gradient is not needed: block6d stays frozen)
        dz_e = self._act_bwd(2, dE, sv["z_e"])
        self._wgrad(dz_e, sv["x7"], self.p_exp)

    # ------------------------------------------------------------------ optimiser
    def apply_adam(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7, count: bool = True) -> None:
        """One Keras-Adam update of tail + head (one optimiser, one iteration count — the head's betas / eps are its own and
        equal these defaults).  The iteration count and the bias-corrected step size live on the device; count=False leaves
        the host-side mirrors alone (used while the launches are being captured into a graph)."""
        L = _lib.lib()
        if count:
            self.t += 1
            _lib.check(L.kws_head_advance_step_count(self.head._h, 1))
        _lib.check(L.kws_train_lr_step(self.d_step.data_ptr(), float(lr), beta1, beta2, self.d_lr_t.data_ptr(), _st()))
        for p in self.params:
            _lib.check(L.kws_train_adam(p.master.data_ptr(), p.m.data_ptr(), p.v.data_ptr(), p.grad.data_ptr(), p.master.numel(),
                                        p.cols, _ptr(p.scale), self.count.data_ptr(), self.loss_scale, float(lr), 0,
                                        self.d_lr_t.data_ptr(), beta1, beta2, eps, p.w16.data_ptr(), _ptr(p.w32), _st()),
                       "kws_train_adam")
            p.refresh_transpose()
        _lib.check(L.kws_head_apply_adam_dev(self.head._h, self.head_flat.data_ptr(), self.d_lr_t.data_ptr(), _st()))

    def _step_body(self, feats: torch.Tensor, labels: torch.Tensor, lr: float, count: bool = True) -> None:
        import torch.distributed as dist
        self.forward_tail(self.tail_input(feats), keep=True)
        self.backward(labels)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)     # ~10.1 M floats (40 MB): tail + head + loss/acc scalars
        self._stats.copy_(self.head_flat[self.head.n_params:self.head.n_params + 3])
        self.apply_adam(lr, count=count)

    _stats = None

    def step(self, feats: torch.Tensor, labels: torch.Tensor, lr: float, graph: bool = True) -> Tuple[float, float]:
        """One optimisation step on the LOCAL shard `feats`/`labels`; the gradients (sums) are all-reduced across the ranks
        of torch.distributed.  Returns (mean loss, accuracy) of the global batch.

        graph=True: the whole step (frozen embedding, tail forward, backward, all-reduce, Adam: ~150 launches issued from
        Python, ~1 ms of host work against ~0.5 ms of device work) is captured into one CUDA graph per (batch size, lr) on
        first use and replayed afterwards."""
        if self._stats is None:
            self._stats = torch.zeros(3, dtype=torch.float32, device=self.dev)
        feats = feats.to(self.dev, torch.float32)
        labels = labels.to(self.dev, torch.int32)
        if not graph:
            self._step_body(feats.contiguous(), labels.contiguous(), lr)
        else:
            key = (int(feats.shape[0]), float(lr))
            g = self._graph if self._graph is not None and self._graph[0] == key else None
            if g is None:
                # eager step first (allocations, scratch buffers, NCCL warm-up happen outside the capture), then capture
                self._step_body(feats.contiguous(), labels.contiguous(), lr)
                sf, sl = feats.clone().contiguous(), labels.clone().contiguous()
                cg = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(cg):
                    self._step_body(sf, sl, lr, count=False)          # a capture executes nothing
                self._graph = (key, cg, sf, sl)
            else:
                _, cg, sf, sl = g
                sf.copy_(feats)
                sl.copy_(labels)
                cg.replay()
                self.t += 1
                _lib.check(_lib.lib().kws_head_advance_step_count(self.head._h, 1))
        stats = self._stats.tolist()
        cnt = max(stats[2], 1.0)
        return stats[0] / cnt, stats[1] / cnt

    # ------------------------------------------------------------------ export
    def export_weights(self) -> Dict[str, np.ndarray]:
        """Keras-named weights with the fine-tuned tensors written back (BatchNorm parameters unchanged)."""
        out = dict(self.embedding.weights)
        for p in self.params:
            m = p.master.cpu().numpy()
            if p.name == "dw":
                out[p.keras_key] = m.reshape(p.keras_shape).astype(np.float32)
            elif m.shape[0] == 1 and len(p.keras_shape) == 1:
                out[p.keras_key] = m.reshape(p.keras_shape).astype(np.float32)
            else:
                out[p.keras_key] = np.ascontiguousarray(m.T).reshape(p.keras_shape).astype(np.float32)
        return out

    def gradients(self) -> Dict[str, np.ndarray]:
        """Current flat gradients as Keras-shaped arrays of d(mean loss)/d(kernel) (for parity tests): un-scaled, divided
        by the sample count, and referred to the UN-folded kernel (x BN scale)."""
        cnt = max(float(self.count.item()), 1.0)
        out = {}
        for p in self.params:
            g = p.grad / (self.loss_scale * cnt)
            if p.scale is not None:
                g = g * p.scale.view(-1, 1 if p.scale.numel() == p.rows else p.cols)
            g = g.cpu().numpy()
            if p.name == "dw" or (g.shape[0] == 1 and len(p.keras_shape) == 1):
                out[p.keras_key] = g.reshape(p.keras_shape)
            else:
                out[p.keras_key] = np.ascontiguousarray(g.T).reshape(p.keras_shape)
        return out
