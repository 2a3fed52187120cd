"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck): frontend (clip + streaming kernels),
embedding forward in the layer-wise and the fused schedules, head step, one phase-2 fine-tune step, augmentation kernel,
streaming post-processor.  Plain launches (graphs off).  Usage: compute-sanitizer --tool memcheck python tools/sanitize_targets.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.fewshot import FewShotModel, Head
from multilingual_kws_b200.finetune import TailTrainer
from multilingual_kws_b200.frontend import FEATURE_SCALE, MicroFrontend
from multilingual_kws_b200.model import EmbeddingModel
from multilingual_kws_b200.synthetic import synthetic_pcm, synthetic_stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
pcm = torch.from_numpy(synthetic_pcm(B, cfg_id=2)).cuda()
fe = MicroFrontend()
feats = fe.forward(pcm)
audio = torch.from_numpy(synthetic_stream(16000 * 3, cfg_id=5)).cuda()
st = fe.stream_prepare(audio)
win = st.windows(16000, 1600, 0, 8, FEATURE_SCALE)
m = EmbeddingModel(W.random_init(0, randomize_bn=True, residual_gamma_scale=0.3))
m.set_graph(False)
outs = []
for mode in (0, 1, 2):
    m.set_fuse(mode)
    outs.append(m.forward_device(feats).clone())
m.set_fuse(0)
head = Head.keras_init(1024, 18, 3, seed=0)
labels = torch.from_numpy((np.arange(B) % 3).astype(np.int32)).cuda()
flat = head.grad(outs[0], labels)
head.apply_adam(flat, 1e-3)
probs = FewShotModel(m, head).forward_device(win)
tr = TailTrainer(m, head)
loss, acc = tr.step(feats, labels, 1e-4)
from multilingual_kws_b200.embedding.single_target_recognize_commands import detect_stream_device
pr = torch.softmax(torch.randn(400, 3, device="cuda"), 1)
det = detect_stream_device(pr, (np.arange(400) * 20).tolist(), ["_silence_", "_unknown_", "kw"], 100, [0.3, 0.5], 500, 4)
torch.cuda.synchronize()
print("sanitize targets done: modes agree", float((outs[0] - outs[1]).abs().max()), float((outs[0] - outs[2]).abs().max()), "loss", loss)
