"""Backward through the top of the embedding (block7a + top conv + dense tower) + head: the CUDA path against
torch.autograd on the fp64 restatement (oracle/tail_autograd_oracle.py), then Adam steps against Keras-Adam on it.
Tolerances: the CUDA path keeps activations and activation gradients in fp16 (fp32 accumulation)."""
import numpy as np
import pytest
import torch

from oracle import effnet_oracle as EO
from oracle import head_oracle as HO
from oracle import tail_autograd_oracle as TO
from oracle.frontend_oracle import FrontendOracle
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.synthetic import synthetic_pcm

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


@pytest.fixture(scope="module")
def setup(kws_lib):
    from multilingual_kws_b200.fewshot import Head
    from multilingual_kws_b200.finetune import TailTrainer
    from multilingual_kws_b200.model import EmbeddingModel
    feats = FrontendOracle().features(synthetic_pcm(32, cfg_id=3), threads=4)
    w = W.random_init(7, randomize_bn=True, residual_gamma_scale=0.3)
    EO.forward(w, feats, calibrate_bn=True)
    hp = HO.init_head(1)
    model = EmbeddingModel(w)
    trainer = TailTrainer(model, Head.from_params(hp))
    labels = (np.arange(32) % 3).astype(np.int32)
    return trainer, w, hp, feats, labels


def test_tail_forward_matches_inference_path_and_oracle(setup):
    trainer, w, hp, feats, labels = setup
    x = torch.from_numpy(feats).cuda()
    emb = trainer.embed(x).cpu().numpy()
    want = EO.forward(w, feats).numpy()
    assert EO.cosine(emb, want).min() >= 0.999
    assert EO.cosine(emb, trainer.embedding.forward_device(x).cpu().numpy()).min() >= 0.9999


def test_gradients_match_autograd(setup):
    trainer, w, hp, feats, labels = setup
    x = torch.from_numpy(feats).cuda()
    x7 = trainer.tail_input(x)
    trainer.forward_tail(x7, keep=True)
    trainer.backward(torch.from_numpy(labels).cuda())
    torch.cuda.synchronize()
    got = trainer.gradients()
    p = TO.make_params(w, hp)
    x7_host = x7.float().cpu().numpy().reshape(32, 2, 2, -1)          # same (fp16-rounded) tail input on both sides
    loss, acc, emb, want = TO.loss_and_grads(w, p, x7_host, labels)
    n_p = trainer.head.n_params
    stats = trainer.head_flat[n_p:n_p + 3].cpu().numpy()
    assert abs(stats[0] / stats[2] - loss) < 2e-3 and stats[2] == 32
    report = []
    for k in TO.TAIL_KEYS:
        e = rel_err(got[k], want[k])
        report.append(f"{k}: rel err {e:.4f}  |g| {np.linalg.norm(want[k]):.3e}")
        assert got[k].shape == want[k].shape, k
    print("\n".join(report))
    # the two sides differ by the forward pass's 16-bit roundings (embedding rel. error ~0.7 % -> softmax residuals
    # p - y differ by ~1-2 %), which every gradient inherits; a wrong formula shows up as an O(1) error
    for k in TO.TAIL_KEYS:
        a, b = got[k].ravel().astype(np.float64), want[k].ravel()
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
        assert rel_err(got[k], want[k]) < 0.08 and cos > 0.997, (k, cos, "\n".join(report))
    # head gradients (fp32 path)
    hf = trainer.head_flat[:n_p].cpu().numpy() / 32
    assert rel_err(hf[:1024 * 18], want["head/w1"]) < 2e-3


def test_adam_steps_track_keras_adam(setup):
    from multilingual_kws_b200.fewshot import Head
    from multilingual_kws_b200.finetune import TailTrainer
    from multilingual_kws_b200.model import EmbeddingModel
    _, w, hp, feats, labels = setup
    trainer = TailTrainer(EmbeddingModel(w), Head.from_params(hp))
    x = torch.from_numpy(feats).cuda()
    y = torch.from_numpy(labels).cuda()
    x7_host = trainer.tail_input(x).float().cpu().numpy().reshape(32, 2, 2, -1)
    p = TO.make_params(w, hp)
    opt = TO.KerasAdam(p, 1e-4)
    losses_g, losses_o = [], []
    for _ in range(12):
        loss_g, _ = trainer.step(x, y, 1e-4)
        loss_o, _, _, grads = TO.loss_and_grads(w, p, x7_host, labels)
        opt.step(grads)
        losses_g.append(loss_g)
        losses_o.append(loss_o)
    print("gpu   ", np.round(losses_g, 4))
    print("oracle", np.round(losses_o, 4))
    assert losses_o[-1] < losses_o[0] * 0.9                      # the steps do train
    assert np.abs(np.array(losses_g) - np.array(losses_o)).max() < 0.02
    assert trainer.t == 12 and trainer.head.step_count == 12 and int(trainer.d_step.item()) == 12   # 1 eager + 11 replays
    # graph replay == eager launches: a second trainer stepping without the graph lands on the same parameters
    eager = TailTrainer(EmbeddingModel(w), Head.from_params(hp))
    for _ in range(12):
        eager.step(x, y, 1e-4, graph=False)
    for pg, pe in zip(trainer.params, eager.params):
        assert torch.equal(pg.master, pe.master), pg.name
    new = trainer.export_weights()
    for k in ("dense_2/kernel", "top_conv/kernel", "block7a_expand_conv/kernel", "block7a_dwconv/depthwise_kernel"):
        moved = rel_err(p[k].detach().numpy(), w[k])             # how far training moved the tensor
        assert rel_err(new[k], p[k].detach().numpy()) < 0.25 * moved + 1e-4, k
    # the exported weights load into the inference model and reproduce the trainer's own forward
    emb_inf = EmbeddingModel(new).forward_device(x).cpu().numpy()
    assert EO.cosine(emb_inf, trainer.embed(x).cpu().numpy()).min() >= 0.9999
