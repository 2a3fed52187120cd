"""Pin of the two Adam restatements used as oracles (oracle/head_oracle.py Adam, oracle/tail_autograd_oracle.py KerasAdam)
against torch.optim.Adam.

tf.keras.optimizers.Adam (reference transfer_learning.py:56,102; TF 2.7 `ResourceApplyAdam`) updates
    theta -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)                       ("epsilon hat" form)
torch.optim.Adam updates
    theta -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
which is the same expression with eps replaced by eps * sqrt(1 - b2^t).  So a torch optimiser whose eps is set to
eps / sqrt(1 - b2^t) before step t must follow the Keras trajectory exactly — moments, bias corrections and step size are
then confirmed by an independent implementation, and the placement of eps is the one documented difference."""
import numpy as np
import torch

from oracle import head_oracle as HO
from oracle import tail_autograd_oracle as TO


def _problem(seed=0):
    rng = np.random.default_rng(seed)
    p0 = {"a": rng.normal(size=(7, 5)), "b": rng.normal(size=(5,)) * 1e-3}
    grads = [{k: rng.normal(size=v.shape) * (10.0 ** rng.uniform(-9, 0)) for k, v in p0.items()} for _ in range(40)]
    return p0, grads                                     # gradient magnitudes from far below eps to O(1)


def _torch_reference(p0, grads, lr, eps, b1=0.9, b2=0.999):
    params = [torch.nn.Parameter(torch.tensor(v, dtype=torch.float64)) for v in p0.values()]
    opt = torch.optim.Adam(params, lr=lr, betas=(b1, b2), eps=eps)
    for t, g in enumerate(grads, 1):
        opt.param_groups[0]["eps"] = eps / np.sqrt(1.0 - b2 ** t)
        for q, k in zip(params, p0):
            q.grad = torch.tensor(g[k], dtype=torch.float64)
        opt.step()
    return {k: q.detach().numpy() for q, k in zip(params, p0)}


def test_tail_oracle_adam_equals_torch_with_rescaled_eps():
    p0, grads = _problem()
    want = _torch_reference(p0, grads, 1e-3, 1e-7)
    p = {k: torch.tensor(v, dtype=torch.float64) for k, v in p0.items()}
    opt = TO.KerasAdam(p, 1e-3)
    for g in grads:
        opt.step(g)
    for k in p0:
        assert np.abs(p[k].numpy() - want[k]).max() < 1e-12, k


def test_head_oracle_adam_equals_torch_with_rescaled_eps():
    p0, grads = _problem(1)
    want = _torch_reference(p0, grads, 1e-3, 1e-7)
    p = {k: v.astype(np.float32) for k, v in p0.items()}         # the head oracle stores fp32 parameters
    opt = HO.Adam(p, 1e-3)
    for g in grads:
        opt.step(p, g)
    for k in p0:
        assert np.abs(p[k] - want[k]).max() < 5e-6, k            # 40 roundings of the parameters to fp32


def test_eps_placement_matters_at_small_gradients():
    """With torch's own (fixed) eps the trajectories differ where |g| ~ eps: the rescaling above is not a no-op."""
    p0, grads = _problem(2)
    grads = [{k: g * 1e-7 for k, g in step.items()} for step in grads]
    params = [torch.nn.Parameter(torch.tensor(v, dtype=torch.float64)) for v in p0.values()]
    opt = torch.optim.Adam(params, lr=1e-3, eps=1e-7)
    p = {k: torch.tensor(v, dtype=torch.float64) for k, v in p0.items()}
    keras = TO.KerasAdam(p, 1e-3)
    for g in grads:
        for q, k in zip(params, p0):
            q.grad = torch.tensor(g[k], dtype=torch.float64)
        opt.step()
        keras.step(g)
    diff = max(float(np.abs(p[k].numpy() - q.detach().numpy()).max()) for q, k in zip(params, p0))
    assert diff > 1e-4
