"""Weight container for the embedding model, Keras-named, plus deterministic random init.

Architecture source of truth: reference multilingual_kws/train_multilingual_embedding.py:66-83
(`tf.keras.applications.EfficientNetB0(include_top=False, weights=None, input_shape=(49,40,1))`
-> GAP -> Dense(2048,relu) -> Dense(2048,relu) -> Dense(1024,selu); consumers cut at "dense_2",
transfer_learning.py:38-43).  Layer names and tensor layouts are Keras' own (HWIO conv kernels,
[in,out] dense kernels) so a SavedModel importer can fill the same dict.  The released checkpoint is
not available offline; `random_init` follows Keras' initialisers.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

BN_EPS = 1e-3

# (kernel, repeats, filters_in, filters_out, expand_ratio, stride) — Keras DEFAULT_BLOCKS_ARGS at B0
STAGES = [(3, 1, 32, 16, 1, 1), (3, 2, 16, 24, 6, 2), (5, 2, 24, 40, 6, 2), (3, 3, 40, 80, 6, 2),
          (5, 3, 80, 112, 6, 1), (5, 4, 112, 192, 6, 2), (3, 1, 192, 320, 6, 1)]
STEM_FILTERS, TOP_FILTERS = 32, 1280
DENSE_UNITS = (2048, 2048, 1024)            # multilingual embedding; monolingual uses (1024, 1024, 192)
INPUT_HW = (49, 40)


def correct_pad(h: int, w: int, k: int) -> Tuple[Tuple[int, int], Tuple[int, int]]:
    """keras.applications.imagenet_utils.correct_pad for a stride-2 conv."""
    c = k // 2
    return (c - (1 - h % 2), c), (c - (1 - w % 2), c)


def block_list() -> List[dict]:
    """The 16 MBConv blocks with resolved shapes at 49x40 input."""
    h, w = 25, 20            # after the stem
    out = []
    for si, (k, reps, fin, fout, e, s) in enumerate(STAGES):
        for r in range(reps):
            stride = s if r == 0 else 1
            cin = fin if r == 0 else fout
            cexp = cin * e
            if stride == 2:
                (pt, pb), (pl, pr) = correct_pad(h, w, k)
                ho, wo = (h + pt + pb - k) // 2 + 1, (w + pl + pr - k) // 2 + 1
            else:
                pt = pl = k // 2
                ho, wo = h, w
            out.append(dict(name=f"block{si + 1}{chr(ord('a') + r)}", k=k, stride=stride, cin=cin, cout=fout, cexp=cexp,
                            expand=(e != 1), se=max(1, int(cin * 0.25)), h=h, w=w, ho=ho, wo=wo, pad_top=pt, pad_left=pl,
                            residual=(stride == 1 and cin == fout)))
            h, w = ho, wo
    return out


def param_shapes(dense_units=DENSE_UNITS) -> Dict[str, Tuple[int, ...]]:
    sh: Dict[str, Tuple[int, ...]] = {}

    def bn(name, c):
        for p in ("gamma", "beta", "moving_mean", "moving_variance"):
            sh[f"{name}/{p}"] = (c,)

    sh["normalization/mean"] = (1,)
    sh["normalization/variance"] = (1,)
    sh["normalization/count"] = ()
    sh["stem_conv/kernel"] = (3, 3, 1, STEM_FILTERS)
    bn("stem_bn", STEM_FILTERS)
    for b in block_list():
        n = b["name"]
        if b["expand"]:
            sh[f"{n}_expand_conv/kernel"] = (1, 1, b["cin"], b["cexp"])
            bn(f"{n}_expand_bn", b["cexp"])
        sh[f"{n}_dwconv/depthwise_kernel"] = (b["k"], b["k"], b["cexp"], 1)
        bn(f"{n}_bn", b["cexp"])
        sh[f"{n}_se_reduce/kernel"] = (1, 1, b["cexp"], b["se"])
        sh[f"{n}_se_reduce/bias"] = (b["se"],)
        sh[f"{n}_se_expand/kernel"] = (1, 1, b["se"], b["cexp"])
        sh[f"{n}_se_expand/bias"] = (b["cexp"],)
        sh[f"{n}_project_conv/kernel"] = (1, 1, b["cexp"], b["cout"])
        bn(f"{n}_project_bn", b["cout"])
    sh["top_conv/kernel"] = (1, 1, STAGES[-1][3], TOP_FILTERS)
    bn("top_bn", TOP_FILTERS)
    fan = TOP_FILTERS
    for i, u in enumerate(dense_units):
        nm = "dense" if i == 0 else f"dense_{i}"
        sh[f"{nm}/kernel"] = (fan, u)
        sh[f"{nm}/bias"] = (u,)
        fan = u
    return sh


def _trunc_normal(rng, shape, std):
    """Keras truncated_normal: resample outside 2 sigma; VarianceScaling divides std by .87962566."""
    x = rng.normal(0.0, 1.0, size=shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.normal(0.0, 1.0, size=int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


def random_init(seed: int = 0, dense_units=DENSE_UNITS, randomize_bn: bool = False,
                residual_gamma_scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Keras initialisers: conv kernels VarianceScaling(2.0, fan_out, truncated_normal); Dense
    glorot_uniform with zero bias (dense_2: lecun_normal); BN gamma 1, beta 0, mean 0, var 1.
    randomize_bn=True draws non-trivial BN parameters/statistics (for parity tests of the folding).
    residual_gamma_scale < 1 damps the project-BN of blocks that have a skip connection (the usual
    'small residual branch' regime of trained residual nets; a BN network at pure random init amplifies
    perturbations exponentially with depth and is not representative of a trained one)."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    last_dense = "dense" if len(dense_units) == 1 else f"dense_{len(dense_units) - 1}"
    for name, shape in param_shapes(dense_units).items():
        layer, p = name.rsplit("/", 1)
        if layer == "normalization":
            w[name] = np.zeros(shape, np.float32) if p != "variance" else np.ones(shape, np.float32)
        elif p in ("kernel", "depthwise_kernel") and len(shape) == 4:
            kh, kw, cin, cout = shape
            fan_out = kh * kw * cout      # Keras _compute_fans: receptive field x shape[-1] (1 for depthwise)
            w[name] = _trunc_normal(rng, shape, math.sqrt(2.0 / fan_out) / .87962566103423978)
        elif p == "kernel":
            fin, fout = shape
            if layer == last_dense:
                w[name] = _trunc_normal(rng, shape, math.sqrt(1.0 / fin) / .87962566103423978)
            else:
                lim = math.sqrt(6.0 / (fin + fout))
                w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif p == "bias":
            w[name] = np.zeros(shape, np.float32)
        elif p == "gamma":
            w[name] = (rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)).astype(np.float32)
        elif p == "beta":
            w[name] = (rng.normal(0, 0.1, shape) if randomize_bn else np.zeros(shape)).astype(np.float32)
        elif p == "moving_mean":
            w[name] = (rng.normal(0, 0.1, shape) if randomize_bn else np.zeros(shape)).astype(np.float32)
        elif p == "moving_variance":
            w[name] = (rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)).astype(np.float32)
        else:
            raise AssertionError(name)
    if residual_gamma_scale != 1.0:
        for b in block_list():
            if b["residual"]:
                w[b["name"] + "_project_bn/gamma"] = w[b["name"] + "_project_bn/gamma"] * np.float32(residual_gamma_scale)
                w[b["name"] + "_project_bn/beta"] = w[b["name"] + "_project_bn/beta"] * np.float32(residual_gamma_scale)
    if randomize_bn:   # SE biases are trainable too
        for name in w:
            if name.endswith("_se_reduce/bias") or name.endswith("_se_expand/bias"):
                w[name] = rng.normal(0, 0.2, w[name].shape).astype(np.float32)
            if name.endswith("/bias") and name.startswith("dense"):
                w[name] = rng.normal(0, 0.05, w[name].shape).astype(np.float32)
    return w


def count_params(w: Dict[str, np.ndarray]) -> Dict[str, int]:
    conv = sum(v.size for k, v in w.items() if not k.startswith("dense") and not k.startswith("normalization"))
    dense = sum(v.size for k, v in w.items() if k.startswith("dense"))
    return dict(conv_stack=conv, dense_tower=dense, normalization=3, total=conv + dense + 3)


def save_npz(path: str, w: Dict[str, np.ndarray]) -> None:
    np.savez(path, **{k.replace("/", "__"): v for k, v in w.items()})


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k.replace("__", "/"): z[k] for k in z.files}
