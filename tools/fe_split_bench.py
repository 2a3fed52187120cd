"""frontend_clip_kernel (one CTA per clip) vs the frame-reuse pair (per-frame magnitudes + per-window tail) run over
B independent clips laid end to end (hop = clip length): same bits, different parallelisation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multilingual_kws_b200.frontend import FEATURE_SCALE, MicroFrontend
from multilingual_kws_b200.synthetic import synthetic_pcm

fe = MicroFrontend()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (256, 1024, 4096, 8192):
    pcm = torch.from_numpy(np.tile(synthetic_pcm(256, cfg_id=2), (-(-B // 256), 1))[:B]).cuda()
    flat = torch.cat([pcm.reshape(-1), torch.zeros(8, dtype=torch.int16, device="cuda")])   # last offset must be < n - clip
    out = torch.empty((B, 49, 40), dtype=torch.float32, device="cuda")

    def a():
        fe.forward(pcm, out=out)

    def b():
        st = fe.stream_prepare(flat)
        return st.windows(16000, 16000, 0, B, FEATURE_SCALE)
    ref = fe.forward(pcm).clone()
    got = b()
    same = torch.equal(ref, got)
    res = []
    for fn in (a, b):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        res.append(float(np.median(ts)))
    print(f"B={B}: clip kernel {res[0] * 1e3:.1f} us, frame-mags + window-tail {res[1] * 1e3:.1f} us, bit-equal {same}")
