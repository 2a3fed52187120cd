"""GPU parity of the embedding tower against the torch fp32 oracle: per-op taps and final cosine >= 0.999."""
import numpy as np
import pytest
import torch

from oracle import effnet_oracle as EO
from oracle.frontend_oracle import FrontendOracle
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.synthetic import synthetic_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(kws_lib):
    from multilingual_kws_b200.model import EmbeddingModel
    feats = FrontendOracle().features(synthetic_pcm(40, cfg_id=2), threads=4)
    w = W.random_init(3, randomize_bn=True)
    EO.forward(w, feats, calibrate_bn=True)            # trained-like activation scales
    taps = {}
    want = EO.forward(w, feats, taps=taps).numpy()
    taps["top_gap"] = taps["top_activation"].mean(axis=(1, 2))
    return EmbeddingModel(w), feats, want, taps


def rel_err(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def test_every_op_matches_oracle_tap(setup):
    model, feats, want, taps = setup
    x = torch.from_numpy(feats).cuda()
    report = []
    for i, (name, elems) in enumerate(model.op_names()):
        _, tap = model.forward_device(x, tap_op=i)
        torch.cuda.synchronize()
        ref = taps[name].reshape(feats.shape[0], -1)
        assert ref.shape[1] == elems, (name, ref.shape, elems)
        e = rel_err(tap.float().cpu().numpy(), ref)
        report.append((name, e))
        assert e < 0.05, f"op {i} {name}: relative error {e:.4f}\n" + "\n".join(f"{n}: {v:.5f}" for n, v in report)
    print("\n".join(f"{n}: {v:.5f}" for n, v in report))


def test_embedding_cosine(setup):
    model, feats, want, _ = setup
    got = model.predict(feats)
    assert got.shape == want.shape == (feats.shape[0], 1024) and got.dtype == np.float32
    cos = EO.cosine(got, want)
    assert cos.min() >= 0.999, cos.min()                       # the tolerance north_star states
    assert rel_err(got, want) < 0.03


def test_chunking_and_batch_sizes_agree(setup):
    model, feats, _, _ = setup
    x = torch.from_numpy(feats).cuda()
    full = model.forward_device(x).cpu()
    for chunk in (7, 16):
        model.set_chunk(chunk)
        assert torch.equal(model.forward_device(x).cpu(), full)
    model.set_chunk(256)
    for b in (1, 3, 33):
        assert torch.equal(model.forward_device(x[:b]).cpu(), full[:b])
    assert model.forward_device(x[:0]).shape == (0, 1024)
    assert model.predict(feats[:5, :, :, None]).shape == (5, 1024)


def test_keras_default_init_weights(kws_lib):
    """Un-calibrated Keras initialisation (BN identity): activations shrink layer by layer; still cosine-parity."""
    from multilingual_kws_b200.model import EmbeddingModel
    feats = FrontendOracle().features(synthetic_pcm(8, cfg_id=7))
    w = W.random_init(0)
    got = EmbeddingModel(w).predict(feats)
    want = EO.forward(w, feats).numpy()
    assert EO.cosine(got, want).min() >= 0.999


def test_monolingual_head_sizes(kws_lib):
    """train_monolingual_embedding.py:93-98 uses 1024/1024/192 dense units."""
    from multilingual_kws_b200.model import EmbeddingModel
    feats = FrontendOracle().features(synthetic_pcm(4, cfg_id=8))
    w = W.random_init(5, dense_units=(1024, 1024, 192), randomize_bn=True)
    got = EmbeddingModel(w).predict(feats)
    want = EO.forward(w, feats).numpy()
    assert got.shape == (4, 192) and EO.cosine(got, want).min() >= 0.999
