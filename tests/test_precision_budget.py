"""Where the 16-bit embedding path loses cosine, measured on the CPU with the fp32 oracle and simulated fp16 roundings.

On the trained-like network every single source of rounding costs < 3e-5 of cosine and all of them together 7e-5.  On the
UNDAMPED random-init network (no damping of the residual branches: a chaotic map, perturbations grow ~150x from the first
block to the embedding) rounding only the WEIGHTS to fp16 — activations and arithmetic in fp32 — already gives ~0.9989:
the 0.999 tolerance is unreachable there for any path whose tensor-core operands are 16 bit (kind::f16, and kind::tf32
has the same 10-bit significand), which is what north_star mandates for the pointwise contractions.  That regime is
therefore reported with its measured bound, and parity is asserted on the trained-like and trained regimes."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import effnet_oracle as EO
from oracle.frontend_oracle import FrontendOracle
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.synthetic import synthetic_pcm


def _h(t):
    return t.to(torch.float16).to(torch.float32)


def forward_rounded(w, feats, rw=False, ract=False):
    """fp32 forward with BN folded; optional fp16 rounding of the 1x1 / dense weights (rw) and of every stored activation (ract)."""
    t = lambda n: torch.as_tensor(w[n], dtype=torch.float32)                                   # noqa: E731
    sw = lambda v: v * torch.sigmoid(v)                                                        # noqa: E731
    r = (lambda v: _h(v)) if ract else (lambda v: v)

    def conv(x, kname, bnname, stride=1, dw=False):
        g, b, m, v = t(bnname + "/gamma"), t(bnname + "/beta"), t(bnname + "/moving_mean"), t(bnname + "/moving_variance")
        s = g / torch.sqrt(v + 1e-3)
        k = t(kname)
        k = (k.permute(2, 3, 0, 1) if dw else k.permute(3, 2, 0, 1)) * s[:, None, None, None]
        if rw and not dw:
            k = _h(k)
        return F.conv2d(x, k, stride=stride, groups=k.shape[0] if dw else 1) + (b - m * s)[None, :, None, None]

    x = torch.as_tensor(feats, dtype=torch.float32)[:, None] * (1 / 255.0)
    x = r(sw(conv(EO._pad_for(x, 3, 2), "stem_conv/kernel", "stem_bn", stride=2)))
    for si, (k, reps, fin, fout, e, s) in enumerate(EO._STAGES):
        for rep in range(reps):
            n = f"block{si + 1}{chr(ord('a') + rep)}"
            stride, cin, inp = (s if rep == 0 else 1), (fin if rep == 0 else fout), x
            if e != 1:
                x = r(sw(conv(x, f"{n}_expand_conv/kernel", f"{n}_expand_bn")))
            x = sw(conv(EO._pad_for(x, k, stride), f"{n}_dwconv/depthwise_kernel", f"{n}_bn", stride=stride, dw=True))
            se = x.mean(dim=(2, 3), keepdim=True)
            se = sw(F.conv2d(se, t(f"{n}_se_reduce/kernel").permute(3, 2, 0, 1)) + t(f"{n}_se_reduce/bias").view(1, -1, 1, 1))
            se = torch.sigmoid(F.conv2d(se, t(f"{n}_se_expand/kernel").permute(3, 2, 0, 1)) + t(f"{n}_se_expand/bias").view(1, -1, 1, 1))
            x = r(x * se)
            x = conv(x, f"{n}_project_conv/kernel", f"{n}_project_bn")
            if stride == 1 and cin == fout:
                x = x + inp
            x = r(x)
    x = r(sw(conv(x, "top_conv/kernel", "top_bn")).mean(dim=(2, 3)))
    acts = [torch.relu, torch.relu, lambda v: EO._SELU_L * torch.where(v > 0, v, EO._SELU_A * (torch.exp(v) - 1))]
    for i in range(3):
        nm = "dense" if i == 0 else f"dense_{i}"
        kk = t(nm + "/kernel")
        x = acts[i](x @ (_h(kk) if rw else kk) + t(nm + "/bias"))
        if i < 2:
            x = r(x)
    return x.numpy()


def test_rounding_budget_of_the_two_synthetic_regimes():
    feats = FrontendOracle().features(synthetic_pcm(24, cfg_id=2), threads=4)
    res = {}
    for name, gamma in (("damped", 0.3), ("undamped", 1.0)):
        w = W.random_init(3, randomize_bn=True, residual_gamma_scale=gamma)
        EO.forward(w, feats, calibrate_bn=True)
        want = EO.forward(w, feats, dtype=torch.float64).numpy()
        assert EO.cosine(forward_rounded(w, feats), want).min() > 1 - 1e-8       # the restatement itself is exact
        res[name] = (EO.cosine(forward_rounded(w, feats, rw=True), want).min(),
                     EO.cosine(forward_rounded(w, feats, rw=True, ract=True), want).min())
    print(res)
    assert res["damped"][0] > 0.99995 and res["damped"][1] > 0.9998        # weights only; weights + activations
    assert res["undamped"][0] < 0.9995                                      # weights alone already cost > 5e-4 here
    assert res["undamped"][1] > 0.99                                        # and everything together stays bounded
