"""A/B of build-time/env knobs: embedding forward at B=1024 (chunk 1024 / late 4096), median of 20 graph replays with
an L2 flush in between.  Usage: KWS_NO_PDL=1 python tools/ab.py [label]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.frontend import MicroFrontend
from multilingual_kws_b200.model import EmbeddingModel
from multilingual_kws_b200.synthetic import synthetic_pcm

label = sys.argv[1] if len(sys.argv) > 1 else "default"
fe = MicroFrontend()
w = W.random_init(0, randomize_bn=True, residual_gamma_scale=0.3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (1024,):
    pcm = torch.from_numpy(np.tile(synthetic_pcm(256, cfg_id=2), (-(-B // 256), 1))[:B]).cuda()
    feats = fe.forward(pcm)
    m = EmbeddingModel(w, chunk=1024)
    m.set_chunk_late(4096)
    out = torch.empty((B, 1024), device="cuda")
    for graph in (True, False):
        m.set_graph(graph)
        for _ in range(3):
            m.forward_device(feats, out=out)
        ts = []
        for _ in range(20):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); m.forward_device(feats, out=out); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        print(json.dumps(dict(label=label, graph=graph, B=B, ms_median=round(float(np.median(ts)), 4), ms_min=round(float(np.min(ts)), 4),
                              launches=m.launches(B), checksum=float(out.double().sum().item()))))
