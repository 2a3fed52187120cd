"""Pins for the numpy head/Adam oracle: gradients vs torch.autograd, Adam vs its closed form."""
import numpy as np
import torch

from oracle import head_oracle as HO


def test_gradients_match_autograd():
    rng = np.random.default_rng(0)
    p = HO.init_head(1)
    p["b1"] = rng.normal(0, .1, 18).astype(np.float32)
    p["b2"] = rng.normal(0, .1, 3).astype(np.float32)
    emb = rng.normal(0, 1, (37, 1024)).astype(np.float32)
    y = rng.integers(0, 3, 37)
    loss, acc, g = HO.loss_and_grads(p, emb, y)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    h = torch.tanh(torch.tensor(emb, dtype=torch.float64) @ tp["w1"] + tp["b1"])
    z = h @ tp["w2"] + tp["b2"]
    tl = torch.nn.functional.cross_entropy(z, torch.tensor(y))
    tl.backward()
    assert abs(loss - tl.item()) < 1e-12
    for k in p:
        assert np.allclose(g[k], tp[k].grad.numpy(), rtol=1e-9, atol=1e-12), k
    assert 0.0 <= acc <= 1.0


def test_adam_first_steps_closed_form():
    p = dict(x=np.array([1.0, -2.0], np.float32))
    opt = HO.Adam(p, lr=1e-3)
    g = dict(x=np.array([0.5, -0.25]))
    opt.step(p, g)
    # t=1: m_hat = g, v_hat = g^2 -> step = lr * g/(|g| + eps*sqrt(1-b2)/... ) ~= lr*sign(g)
    assert np.allclose(p["x"], [1.0 - 1e-3, -2.0 + 1e-3], atol=2e-9)
    x0 = p["x"].copy()
    opt.step(p, g)
    assert np.allclose(p["x"], x0 - 1e-3 * np.sign(g["x"]), atol=1e-8)     # constant gradient -> unit steps


def test_training_reduces_loss():
    rng = np.random.default_rng(2)
    centers = rng.normal(0, 1, (3, 1024))
    y = rng.integers(0, 3, 96)
    emb = (centers[y] + rng.normal(0, 2.0, (96, 1024))).astype(np.float32)
    p = HO.init_head(3)
    hist = HO.train(p, emb, y, 30, lr=1e-3)
    assert hist[-1][0] < hist[0][0] * 0.5 and hist[-1][1] > 0.9
