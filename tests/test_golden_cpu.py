"""The committed fixtures under tests/golden/ are reproducible: the seeded input generators have not drifted and
the oracle still produces the stored outputs (tests/golden/make_golden.py is the generating script)."""
import hashlib
import os
import sys

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import make_golden as MG  # noqa: E402

from oracle import head_oracle as HO  # noqa: E402
from oracle.frontend_oracle import FrontendOracle  # noqa: E402
from multilingual_kws_b200.synthetic import synthetic_pcm  # noqa: E402


def test_frontend_fixture_reproducible():
    g = np.load(os.path.join(GOLDEN, "frontend_cfg1.npz"))
    pcm, adv = synthetic_pcm(MG.N_FRONTEND, cfg_id=1), MG.adversarial_pcm()
    assert MG.sha(pcm) == str(g["pcm_sha256"]) and MG.sha(adv) == str(g["adversarial_sha256"])
    orc = FrontendOracle()
    assert np.array_equal(orc.features_u16(pcm), g["features_u16"])
    assert np.array_equal(orc.features_u16(adv), g["adversarial_features_u16"])
    assert g["features_u16"].shape == (MG.N_FRONTEND, 49, 40)
    assert not g["adversarial_features_u16"][6].any()              # silence → all-zero features


def test_embed_fixture_inputs_reproducible():
    g = np.load(os.path.join(GOLDEN, "embed_cfg2.npz"))
    feats = FrontendOracle().features(synthetic_pcm(MG.N_EMBED, cfg_id=2))
    assert np.array_equal(np.rint(feats * 25.6).astype(np.uint16), g["features_u16"])
    assert g["embedding"].shape == (MG.N_EMBED, 1024) and np.isfinite(g["embedding"]).all()


def test_head_fixture_reproducible():
    g = np.load(os.path.join(GOLDEN, "head_cfg3.npz"))
    p, e, y = MG.head_case()
    assert np.array_equal(e, g["emb"]) and np.array_equal(y, g["labels"])
    loss, acc, grads = HO.loss_and_grads(p, e, y)
    assert abs(loss - float(g["loss"])) < 1e-12 and acc == float(g["acc"])
    hist = HO.train(p, e, y, steps=10, lr=1e-3)
    assert np.allclose([h[0] for h in hist], g["loss_history"], rtol=0, atol=1e-9)
    for k in ("w1", "b1", "w2", "b2"):
        assert np.allclose(p[k], g[f"p10_{k}"], rtol=0, atol=1e-7)
