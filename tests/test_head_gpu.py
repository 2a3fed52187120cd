"""GPU parity of the fused head kernels (forward, loss+backward, Keras-Adam) against the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as HO

pytestmark = pytest.mark.gpu


def make(seed=0, B=64):
    rng = np.random.default_rng(seed)
    p = HO.init_head(seed)
    p["b1"] = rng.normal(0, .1, 18).astype(np.float32)
    p["b2"] = rng.normal(0, .1, 3).astype(np.float32)
    centers = rng.normal(0, 1, (3, 1024))
    y = rng.integers(0, 3, B)
    emb = (centers[y] * 0.3 + rng.normal(0, 1.0, (B, 1024))).astype(np.float32)
    return p, emb, y


@pytest.mark.parametrize("B", [1, 31, 64, 512, 1000])
def test_forward_and_grads(kws_lib, B):
    from multilingual_kws_b200.fewshot import Head
    p, emb, y = make(1, B)
    head = Head.from_params(p)
    d_emb, d_y = torch.from_numpy(emb).cuda(), torch.from_numpy(y.astype(np.int32)).cuda()
    probs = head.forward(d_emb).cpu().numpy()
    _, _, want_p = HO.forward(p, emb)
    assert np.abs(probs - want_p).max() < 2e-5
    flat = head.grad(d_emb, d_y).cpu().numpy().astype(np.float64)
    loss, acc, g = HO.loss_and_grads(p, emb, y)
    n = head.n_params
    assert flat[n + 2] == B
    assert abs(flat[n] / B - loss) < 1e-4 * max(1.0, abs(loss))
    assert abs(flat[n + 1] / B - acc) < 1e-9
    want = np.concatenate([g["w1"].ravel(), g["b1"], g["w2"].ravel(), g["b2"]]) * B      # sums, not means
    assert np.abs(flat[:n] - want).max() < 2e-4 * max(1.0, np.abs(want).max())


def test_grad_is_deterministic(kws_lib):
    from multilingual_kws_b200.fewshot import Head
    p, emb, y = make(2, 700)
    head = Head.from_params(p)
    d_emb, d_y = torch.from_numpy(emb).cuda(), torch.from_numpy(y.astype(np.int32)).cuda()
    a = head.grad(d_emb, d_y).clone()
    for _ in range(3):
        assert torch.equal(head.grad(d_emb, d_y), a)


def test_100_adam_steps_track_oracle(kws_lib):
    """BASELINE config 3 shape: batch 512, 100 steps, lr 1e-3 — loss curve and final accuracy."""
    from multilingual_kws_b200.fewshot import Head
    p, emb, y = make(3, 512)
    head = Head.from_params(p)
    d_emb, d_y = torch.from_numpy(emb).cuda(), torch.from_numpy(y.astype(np.int32)).cuda()
    hist = []
    for _ in range(100):
        flat = head.grad(d_emb, d_y)
        head.apply_adam(flat, 1e-3)
        f = flat.cpu().numpy()
        hist.append((f[head.n_params] / 512, f[head.n_params + 1] / 512))
    want = HO.train(p, emb, y, 100, lr=1e-3)           # mutates p
    got_l, want_l = np.array([h[0] for h in hist]), np.array([h[0] for h in want])
    assert np.abs(got_l - want_l).max() < 2e-3
    assert abs(hist[-1][1] - want[-1][1]) <= 0.005                   # +-0.5 pp (north_star)
    flatp = head.get_params()
    wantp = np.concatenate([p["w1"].ravel(), p["b1"], p["w2"].ravel(), p["b2"]])
    assert np.abs(flatp - wantp).max() < 5e-4
    assert head.step_count == 100
