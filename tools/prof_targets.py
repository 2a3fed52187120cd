"""Small driver for ncu captures: a few frontend launches and embedding forwards (graphs off so every kernel is a
plain launch).  Usage: ncu ... python tools/prof_targets.py [frontend|embed|all] [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.frontend import MicroFrontend
from multilingual_kws_b200.model import EmbeddingModel
from multilingual_kws_b200.synthetic import synthetic_pcm

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pcm = torch.from_numpy(np.tile(synthetic_pcm(256, cfg_id=2), (-(-B // 256), 1))[:B]).cuda()
fe = MicroFrontend()
feats = fe.forward(pcm)
if what in ("frontend", "all"):
    for _ in range(3):
        fe.forward(pcm, out=feats)
if what in ("embed", "all"):
    m = EmbeddingModel(W.random_init(0, randomize_bn=True, residual_gamma_scale=0.3))
    m.set_graph(False)
    for _ in range(2):
        m.forward_device(feats)
torch.cuda.synchronize()
print("done")
