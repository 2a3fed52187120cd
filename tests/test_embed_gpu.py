"""GPU parity of the embedding tower against the torch fp32 oracle: per-op taps and final cosine >= 0.999.

Synthetic weights (the released checkpoint is not available offline): Keras initialisers + randomised BN
parameters, residual branches damped (project-BN gamma x0.3, the regime of trained residual nets) and BN moving
statistics calibrated on the batch so activations have unit scale.  A BN network at *pure* random init is
chaotic (it amplifies any perturbation ~100x by the last layer, see DESIGN.md "Precision"); that case is
measured too, with the bound it can meet.
"""
import numpy as np
import pytest
import torch

from oracle import effnet_oracle as EO
from oracle.frontend_oracle import FrontendOracle
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.synthetic import synthetic_pcm

pytestmark = pytest.mark.gpu


def make_net(seed, feats, residual_gamma_scale=0.3, **kw):
    w = W.random_init(seed, randomize_bn=True, residual_gamma_scale=residual_gamma_scale, **kw)
    EO.forward(w, feats, calibrate_bn=True)            # trained-like activation scales
    return w


@pytest.fixture(scope="module")
def feats():
    return FrontendOracle().features(synthetic_pcm(40, cfg_id=2), threads=4)


@pytest.fixture(scope="module")
def setup(kws_lib, feats):
    from multilingual_kws_b200.model import EmbeddingModel
    w = make_net(3, feats)
    taps = {}
    want = EO.forward(w, feats, taps=taps).numpy()
    taps["top_gap"] = taps["top_activation"].mean(axis=(1, 2))
    return EmbeddingModel(w), w, want, taps


def rel_err(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def test_every_op_matches_oracle_tap(setup, feats):
    model, _, want, taps = setup
    x = torch.from_numpy(feats).cuda()
    report = []
    for i, (name, elems) in enumerate(model.op_names()):
        _, tap = model.forward_device(x, tap_op=i)
        torch.cuda.synchronize()
        ref = taps[name].reshape(feats.shape[0], -1)
        assert ref.shape[1] == elems, (name, ref.shape, elems)
        e = rel_err(tap.float().cpu().numpy(), ref)
        report.append((name, e))
        assert e < 0.02, f"op {i} {name}: relative error {e:.4f}\n" + "\n".join(f"{n}: {v:.5f}" for n, v in report)
    print("\n".join(f"{n}: {v:.5f}" for n, v in report))


def test_embedding_cosine(setup, feats):
    model, _, want, _ = setup
    got = model.predict(feats)
    assert got.shape == want.shape == (feats.shape[0], 1024) and got.dtype == np.float32
    cos = EO.cosine(got, want)
    pair = EO.cosine(want[0], want[1])
    print(f"min cosine {cos.min():.6f}  rel err {rel_err(got, want):.5f}  (cosine between two different clips {pair:.3f})")
    assert pair < 0.9                                           # the embedding is not collapsed: parity is meaningful
    assert cos.min() >= 0.999, cos.min()                        # the tolerance north_star states
    assert rel_err(got, want) < 0.03


def test_chunking_and_batch_sizes_agree(setup, feats):
    model = setup[0]
    x = torch.from_numpy(feats).cuda()
    full = model.forward_device(x).cpu()
    for chunk in (7, 16):
        model.set_chunk(chunk)
        assert torch.equal(model.forward_device(x).cpu(), full)
    model.set_chunk(256)
    model.set_chunk_late(9)
    assert torch.equal(model.forward_device(x).cpu(), full)
    model.set_chunk_late(2048)
    model.set_graph(False)
    assert torch.equal(model.forward_device(x).cpu(), full)
    model.set_graph(True)
    for b in (1, 3, 33):
        assert torch.equal(model.forward_device(x[:b]).cpu(), full[:b])
    assert model.forward_device(x[:0]).shape == (0, 1024)
    assert model.predict(feats[:5, :, :, None]).shape == (5, 1024)


def test_bf16_mode_and_chaotic_net_bounds(kws_lib, setup, feats):
    """Documented bounds outside the headline configuration: bf16 storage (8-bit significand) on the same net, and
    fp16 on an undamped random-init BN net (chaotic regime)."""
    from multilingual_kws_b200.model import EmbeddingModel
    _, w, want, _ = setup
    cos_bf16 = EO.cosine(EmbeddingModel(w, dtype="bf16").predict(feats), want).min()
    wc = make_net(3, feats, residual_gamma_scale=1.0)
    want_c = EO.forward(wc, feats).numpy()
    cos_chaotic = EO.cosine(EmbeddingModel(wc).predict(feats), want_c).min()
    print(f"bf16 on the trained-like net: min cosine {cos_bf16:.5f}; fp16 on the chaotic net: {cos_chaotic:.5f}")
    assert cos_bf16 >= 0.99
    # Undamped random initialisation is chaotic: tests/test_precision_budget.py shows on the CPU that rounding ONLY the
    # weights to fp16 (all arithmetic and activations fp32) already lands at 0.9989 there — no 16-bit-operand tensor-core
    # path can meet 0.999 on it; the regimes that matter are the trained-like one above and the trained one below.
    assert cos_chaotic >= 0.995


def test_fused_tail_schedules_agree(setup, feats):
    """kws_embed_set_fuse: the tail (blocks 4a..7a + top conv) as one tcgen05 launch per MBConv block (1) or per run of
    blocks (2) against the layer-wise schedule (0) and the oracle: block outputs, embedding, ragged batches."""
    model, _, want, taps = setup
    x = torch.from_numpy(feats).cuda()
    base = model.forward_device(x).cpu().numpy()
    names = model.op_names()
    try:
        for mode in (1, 2):
            model.set_fuse(mode)
            assert model.launches(feats.shape[0]) < 40
            got = model.forward_device(x).cpu().numpy()
            assert np.isfinite(got).all()
            assert EO.cosine(got, want).min() >= 0.999 and rel_err(got, base) < 0.01, mode
            for b in (1, 3, 33):
                assert torch.equal(model.forward_device(x[:b]).cpu(), torch.from_numpy(got[:b])), (mode, b)
        model.set_fuse(1)                      # every fused block is a launch whose output can be tapped
        for i, (name, elems) in enumerate(names):
            if name.endswith("_out") and name[5] in "4567":
                _, tap = model.forward_device(x, tap_op=i)
                assert rel_err(tap.float().cpu().numpy(), taps[name].reshape(feats.shape[0], -1)) < 0.02, name
        model.set_fuse(2)
        i = [n for n, _ in names].index("top_gap")
        _, tap = model.forward_device(x, tap_op=i)
        assert rel_err(tap.float().cpu().numpy(), taps["top_gap"].reshape(feats.shape[0], -1)) < 0.02
    finally:
        model.set_fuse(0)


def test_trained_weight_regime(kws_lib, feats):
    """Third weight regime: the torch restatement TRAINED for 150 Adam steps (BatchNorm in training mode) on a synthetic
    4-way task from plain Keras initialisation (no damping, no calibration) — oracle/effnet_train_oracle.py.  The
    tolerance north_star states (cosine >= 0.999) has to hold on trained weights."""
    from multilingual_kws_b200.model import EmbeddingModel
    from oracle import effnet_train_oracle as TR
    train_feats = FrontendOracle().features(synthetic_pcm(256, cfg_id=7), threads=4)
    labels = np.arange(256) % 4                                 # the four clip kinds of the synthetic generator
    w = TR.train(W.random_init(21), train_feats, labels, steps=150, batch=32, lr=1e-3, seed=1, log=print)
    # 150 steps at Keras' BN momentum 0.99 leave the moving statistics 22 % of the way at their initial values (the
    # inference-mode network then collapses to a constant): re-estimate them over the training set, the usual last step
    EO.forward(w, train_feats, calibrate_bn=True)
    want = EO.forward(w, feats).numpy()
    got = EmbeddingModel(w).predict(feats)
    cos = EO.cosine(got, want)
    # a trained embedding carries a large component common to all clips (cosine between clips of different classes
    # 0.95 ... 0.995 from run to run: cuDNN training is not deterministic), so parity is also checked on what
    # distinguishes the clips: the embeddings minus their mean
    mean = want.mean(axis=0)
    spread = np.linalg.norm(want - mean) / np.linalg.norm(want)
    cen = EO.cosine(got - mean, want - mean)
    print(f"trained regime: min cosine {cos.min():.6f} rel err {rel_err(got, want):.5f}; spread of the embeddings around "
          f"their mean {spread:.3f}, min cosine of the centred embeddings {cen.min():.5f}")
    assert spread > 0.02                                        # not collapsed to a constant
    assert cos.min() >= 0.999 and cen.min() >= 0.995


def test_monolingual_head_sizes(kws_lib, feats):
    """train_monolingual_embedding.py:93-98 uses 1024/1024/192 dense units."""
    from multilingual_kws_b200.model import EmbeddingModel
    w = make_net(5, feats[:16], dense_units=(1024, 1024, 192))
    got = EmbeddingModel(w).predict(feats[:16])
    want = EO.forward(w, feats[:16]).numpy()
    assert got.shape == (16, 192) and EO.cosine(got, want).min() >= 0.999


def test_throughput_schedule_same_results(setup, feats):
    """kws_embed_forward_budget (tail of the network sized for a subset of the SMs): identical embeddings."""
    model = setup[0]
    x = torch.from_numpy(np.tile(feats, (8, 1, 1))).cuda()
    want = model.forward_device(x).clone()
    for budget in ((132, 80), (0, 32), (64, 0), (148, 148)):
        assert torch.equal(model.forward_device(x, sm_budget=budget), want), budget
