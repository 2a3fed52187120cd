"""Quick sweep of schedule knobs (chunk sizes, batch) with CUDA-event timing; prints per-kernel-class ms."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.frontend import MicroFrontend
from multilingual_kws_b200.model import EmbeddingModel
from multilingual_kws_b200.synthetic import synthetic_pcm

fe = MicroFrontend()
w = W.random_init(0, randomize_bn=True, residual_gamma_scale=0.3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
kinds = {0: "stem", 1: "gemm", 2: "dwse"}
for B in (1024,):
    pcm = torch.from_numpy(np.tile(synthetic_pcm(256, cfg_id=2), (-(-B // 256), 1))[:B]).cuda()
    feats = fe.forward(pcm)
    for chunk, late in ((256, 2048), (512, 1024), (512, 4096), (1024, 512), (1024, 1024), (1024, 4096)):
        m = EmbeddingModel(w, chunk=chunk)
        m.set_chunk_late(late)
        out = torch.empty((B, 1024), device="cuda")
        for _ in range(3):
            m.forward_device(feats, out=out)
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); m.forward_device(feats, out=out); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        _, ms = m.forward_timed(feats)
        _, ms = m.forward_timed(feats)
        agg = {}
        for (name, kind, *_), t in zip(m.op_info(), ms):
            agg[kinds[kind]] = agg.get(kinds[kind], 0) + float(t)
        print(json.dumps(dict(B=B, chunk=chunk, late=late, ms=round(float(np.median(ts)), 3), utt_s=round(B / np.median(ts) * 1e3),
                              launches=m.launches(B), per_kind={k: round(v, 3) for k, v in agg.items()})))
        if B == 1024 and chunk == 256 and late == 2048:
            print("  all ops (us):", [(n.replace("block", "b").replace("_activation", "").replace("_se_excite", "_dw"), int(round(float(t) * 1000))) for t, n in zip(ms, [i[0] for i in m.op_info()])])
