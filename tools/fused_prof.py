"""One forward pass at fuse mode 1 without CUDA graphs (for ncu: every fused block is its own launch)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multilingual_kws_b200 import weights as W                     # noqa: E402
from multilingual_kws_b200.model import EmbeddingModel             # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = W.random_init(3, randomize_bn=True, residual_gamma_scale=0.3)
model = EmbeddingModel(w)
model.set_fuse(mode)
model.set_graph(False)
x = torch.from_numpy(np.random.default_rng(0).uniform(0, 26, (B, 49, 40)).astype(np.float32)).cuda()
for _ in range(2):
    out = model.forward_device(x)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
