"""Structure pins for the embedding network: parameter counts, shapes, MACs."""
import numpy as np

from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.model import pack_weights


def test_param_counts_match_keras():
    c = W.count_params(W.random_init(0))
    assert c["conv_stack"] == 4_048_988 and c["dense_tower"] == 8_918_016 and c["total"] == 12_967_007


def test_block_shapes():
    b = {x["name"]: x for x in W.block_list()}
    assert len(b) == 16
    assert (b["block2a"]["ho"], b["block2a"]["wo"], b["block2a"]["cexp"], b["block2a"]["se"]) == (13, 10, 96, 4)
    assert (b["block3a"]["ho"], b["block3a"]["wo"], b["block3a"]["pad_top"], b["block3a"]["pad_left"]) == (7, 5, 2, 1)
    assert (b["block6a"]["ho"], b["block6a"]["wo"], b["block6a"]["pad_top"], b["block6a"]["pad_left"]) == (2, 2, 1, 2)
    assert (b["block7a"]["cexp"], b["block7a"]["cout"], b["block7a"]["se"]) == (1152, 320, 48)
    assert [x["residual"] for x in W.block_list()].count(True) == 9
    macs = sum(x["h"] * x["w"] * x["cin"] * x["cexp"] * x["expand"] + x["ho"] * x["wo"] * x["cexp"] * x["cout"]
               for x in W.block_list()) + 4 * 320 * 1280
    assert macs == 21_398_528                                   # pointwise MACs / clip


def test_container_roundtrip(tmp_path):
    w = W.random_init(1, randomize_bn=True)
    W.save_npz(str(tmp_path / "w.npz"), w)
    w2 = W.load_npz(str(tmp_path / "w.npz"))
    assert set(w) == set(w2) and all(np.array_equal(w[k], w2[k]) for k in w)
    blob = pack_weights(w)
    assert blob[:8] == b"KWSW0001" and len(blob) > 4 * 12_967_007
