"""torchrun check (N >= 2 GPUs): window-sharded streaming inference == single-process result, on every rank.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_stream_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.embedding import batch_streaming_analysis as sa
from multilingual_kws_b200.embedding import input_data
from multilingual_kws_b200.fewshot import FewShotModel, Head
from multilingual_kws_b200.model import EmbeddingModel
from multilingual_kws_b200.synthetic import synthetic_stream

s = input_data.standard_microspeech_model_settings(3)
model = FewShotModel(EmbeddingModel(W.random_init(1, randomize_bn=True, residual_gamma_scale=0.3)), Head.keras_init(1024, 18, 3, seed=1))
ok = True
for T, stride_ms in ((16000 * 7 + 333, 20), (16000 * 3, 100), (16000 + 320, 20), (16000, 20)):
    audio = synthetic_stream(T, cfg_id=5).astype(np.float32) / 32768.0
    sharded = sa.stream_inferences(model, s, audio, 16000, 1000, stride_ms)
    # single-process result: temporarily hide the process group from the helper
    real = sa._dist
    sa._dist = lambda: None
    full = sa.stream_inferences(model, s, audio, 16000, 1000, stride_ms)
    sa._dist = real
    same = sharded.shape == full.shape and np.array_equal(sharded, full)
    ok &= same
    if rank == 0:
        print(f"T={T} stride={stride_ms}ms windows={full.shape[0]} sharded==single: {same}")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_STREAM_CHECK", "OK" if int(flag.item()) == 1 else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
