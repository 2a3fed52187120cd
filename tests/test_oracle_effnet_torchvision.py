"""Pin of the network oracle against an independent third-party implementation of EfficientNet-B0.

Keras (where the reference's network lives, `train_multilingual_embedding.py:66-83`) cannot run here, so
oracle/effnet_oracle.py was pinned only structurally (parameter counts, output shapes).  torchvision ships its own
implementation of the same published architecture (`torchvision.models.efficientnet_b0`: its MBConv block code, its
squeeze-excitation module, its stage table) and has exactly Keras' trainable-parameter count for the convolutional stack
(4 007 548 at three input channels).  This test loads the ORACLE'S Keras-named weights into torchvision's modules, gives
them Keras' conventions where the two libraries differ by convention rather than by architecture (BatchNorm eps 1e-3,
one input channel, and Keras' `correct_pad` + VALID for the stride-2 convolutions instead of symmetric padding — applied
as a pre-hook in front of torchvision's own convolution), runs torchvision's forward, and requires the result to equal
the oracle's (both in float64, relative difference < 1e-9): block order, expansion / depthwise / squeeze-excite / projection wiring, SE widths,
activation placement, residual rule and BatchNorm arithmetic are then confirmed by code that shares nothing with the
oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.synthetic import synthetic_pcm
from oracle import effnet_oracle as EO
from oracle.frontend_oracle import FrontendOracle

tv = pytest.importorskip("torchvision")


def _bn_load(bn, w, name):
    bn.eps = 1e-3
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(w[name + "/gamma"]))
        bn.bias.copy_(torch.from_numpy(w[name + "/beta"]))
        bn.running_mean.copy_(torch.from_numpy(w[name + "/moving_mean"]))
        bn.running_var.copy_(torch.from_numpy(w[name + "/moving_variance"]))


def _conv_load(conv, kernel, depthwise=False):
    k = torch.from_numpy(np.asarray(kernel))
    k = k.permute(2, 3, 0, 1) if depthwise else k.permute(3, 2, 0, 1)        # HWC1 -> [C,1,kh,kw]; HWIO -> [O,I,kh,kw]
    assert tuple(conv.weight.shape) == tuple(k.shape), (conv.weight.shape, k.shape)
    with torch.no_grad():
        conv.weight.copy_(k)


def _keras_stride2_padding(conv):
    """Keras: ZeroPadding2D(correct_pad(inputs, k)) + Conv2D(strides=2, padding='valid') (efficientnet.py / imagenet_utils.py):
    pad ((k//2 - (1 - H % 2), k//2), (k//2 - (1 - W % 2), k//2)).  torchvision pads k//2 on every side."""
    k = conv.kernel_size[0]
    conv.padding = (0, 0)

    def hook(_, args):
        x, = args
        h, w_ = x.shape[2], x.shape[3]
        return (F.pad(x, (k // 2 - (1 - w_ % 2), k // 2, k // 2 - (1 - h % 2), k // 2)),)

    conv.register_forward_pre_hook(hook)


def torchvision_tower(w):
    net = tv.models.efficientnet_b0(weights=None).features
    stem = torch.nn.Conv2d(1, 32, 3, stride=2, padding=1, bias=False)             # the reference's input has one channel
    net[0][0] = stem
    _conv_load(stem, w["stem_conv/kernel"])
    _bn_load(net[0][1], w, "stem_bn")
    _keras_stride2_padding(stem)
    names = []
    for s in range(1, 8):
        for i, mb in enumerate(net[s]):
            n = f"block{s}{chr(ord('a') + i)}"
            names.append(n)
            layers = list(mb.block)
            if len(layers) == 4:                                                    # expansion present (ratio 6)
                exp = layers.pop(0)
                _conv_load(exp[0], w[f"{n}_expand_conv/kernel"])
                _bn_load(exp[1], w, f"{n}_expand_bn")
            dw, se, proj = layers
            _conv_load(dw[0], w[f"{n}_dwconv/depthwise_kernel"], depthwise=True)
            _bn_load(dw[1], w, f"{n}_bn")
            if dw[0].stride[0] == 2:
                _keras_stride2_padding(dw[0])
            _conv_load(se.fc1, w[f"{n}_se_reduce/kernel"])
            _conv_load(se.fc2, w[f"{n}_se_expand/kernel"])
            with torch.no_grad():
                se.fc1.bias.copy_(torch.from_numpy(w[f"{n}_se_reduce/bias"]))
                se.fc2.bias.copy_(torch.from_numpy(w[f"{n}_se_expand/bias"]))
            _conv_load(proj[0], w[f"{n}_project_conv/kernel"])
            _bn_load(proj[1], w, f"{n}_project_bn")
    _conv_load(net[8][0], w["top_conv/kernel"])
    _bn_load(net[8][1], w, "top_bn")
    return net.eval(), names


@pytest.fixture(scope="module")
def case():
    feats = FrontendOracle().features(synthetic_pcm(12, cfg_id=2), threads=4)
    w = W.random_init(5, randomize_bn=True, residual_gamma_scale=1.0)
    EO.forward(w, feats, calibrate_bn=True)                        # activations of trained-like scale in every layer
    return feats, w


def test_torchvision_has_the_same_blocks(case):
    _, w = case
    net, names = torchvision_tower(w)                              # every copy asserts equal kernel shapes
    assert names == sorted({k.split("_")[0] for k in w if k.startswith("block")})
    # every weight of the conv stack was consumed: torchvision has no parameter the Keras layout lacks and vice versa
    n_tv = sum(p.numel() for p in net.parameters())
    n_keras = sum(v.size for k, v in w.items()
                  if not k.startswith(("dense", "normalization")) and not k.endswith(("moving_mean", "moving_variance")))
    assert n_tv == n_keras == 4007548 - (864 - 288)                # Keras' trainable count, one input channel instead of three


def test_oracle_equals_torchvision_forward(case):
    feats, w = case
    net, _ = torchvision_tower(w)
    net = net.double()                                             # both sides in float64: agreement to ~1e-12, not "close"
    taps = {}
    emb = EO.forward(w, feats, dtype=torch.float64, taps=taps)
    x = torch.from_numpy(feats).double()[:, None] * (1.0 / 255.0)  # Rescaling; Normalization is un-adapted (mean 0, var 1)
    assert np.allclose(w["normalization/mean"], 0) and np.allclose(w["normalization/variance"], 1)
    with torch.no_grad():
        y = x
        outs = []
        for stage in net:
            y = stage(y)
            outs.append(y)
    # per stage (torchvision's Sequential boundaries = the last block of each Keras stage) and the top activation
    last = ["stem_activation", "block1a_out", "block2b_out", "block3b_out", "block4c_out", "block5c_out", "block6d_out",
            "block7a_out", "top_activation"]
    for name, got in zip(last, outs):
        want = torch.from_numpy(taps[name]).permute(0, 3, 1, 2)
        assert got.shape == want.shape, name
        err = float((got - want).norm() / want.norm())
        assert err < 1e-9, (name, err)
    # the tower on top of torchvision's features: GAP -> Dense/ReLU x2 -> Dense/SELU, against the oracle's embedding
    z = outs[-1].mean(dim=(2, 3))
    for i, act in enumerate((torch.relu, torch.relu, torch.selu)):
        nm = "dense" if i == 0 else f"dense_{i}"
        z = act(z @ torch.from_numpy(w[nm + "/kernel"]).double() + torch.from_numpy(w[nm + "/bias"]).double())
    assert float((z - emb).norm() / emb.norm()) < 1e-9


def test_symmetric_padding_would_not_match(case):
    """The one place the two libraries differ by convention: without Keras' correct_pad the even-width maps shift by one
    pixel and the outputs part — the hook above is load-bearing, not decoration."""
    feats, w = case
    net, _ = torchvision_tower(w)
    for m in net.modules():
        if isinstance(m, torch.nn.Conv2d) and m.stride[0] == 2:
            m._forward_pre_hooks.clear()
            m.padding = (m.kernel_size[0] // 2,) * 2
    taps = {}
    EO.forward(w, feats, taps=taps)
    with torch.no_grad():
        y = net(torch.from_numpy(feats)[:, None] * (1.0 / 255.0))
    want = torch.from_numpy(taps["top_activation"]).permute(0, 3, 1, 2)
    assert y.shape == want.shape
    assert float((y - want).norm() / want.norm()) > 0.05
