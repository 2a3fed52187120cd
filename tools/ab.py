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
    m.set_graph(False)
    _, ms = m.forward_timed(feats)
    _, ms = m.forward_timed(feats)
    short = lambda n: n.replace("block", "b").replace("_activation", "").replace("_se_excite", "_dw")
    print("  per-op us (chunk 1024 / late 4096, plain launches with events in between):",
          [(short(i[0]), int(round(float(t) * 1000))) for t, i in zip(ms, m.op_info())])
    # consecutive batches on alternating streams (separate buffers + workspaces): region time / K
    m.set_graph(True)
    for ns in (1, 2, 3):
        strs = [torch.cuda.Stream() for _ in range(ns)]
        bufs = [(feats.clone(), torch.empty_like(out), torch.empty(m.workspace_bytes(B), dtype=torch.uint8, device="cuda")) for _ in range(ns)]
        def run(k):
            f, o, w_ = bufs[k % ns]
            with torch.cuda.stream(strs[k % ns]):
                fe.forward(pcm, out=f)
                m.forward_device(f, out=o, workspace=w_)
        for k in range(2 * ns): run(k)
        torch.cuda.synchronize()
        K = 24
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for st in strs: st.wait_stream(torch.cuda.current_stream())
        for k in range(K):
            flush.zero_()
            strs[k % ns].wait_stream(torch.cuda.current_stream())
            run(k)
        for st in strs: torch.cuda.current_stream().wait_stream(st)
        e.record(); torch.cuda.synchronize()
        print(json.dumps(dict(streams=ns, ms_per_batch_frontend_plus_embed=round(s.elapsed_time(e) / K, 4), same_result=bool(torch.equal(bufs[0][1], out)))))
