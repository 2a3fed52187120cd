"""world_size-2 gloo test (CPU) of the data-parallel fine-tune plumbing: contiguous batch shards, flat
[grads | loss_sum | n_correct | n] buffers summed with ONE all-reduce, identical Adam on every rank.
The per-shard gradient producer here is the numpy oracle (the CUDA kernel needs a GPU); what is under test is the
sharding arithmetic and the collective, i.e. that the N-rank result equals the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import head_oracle as HO


def shard(n, rank, world):
    return n * rank // world, n * (rank + 1) // world          # same split as transfer_learning.train_step


def flat_from_oracle(p, emb, y):
    n = emb.shape[0]
    if n == 0:
        return np.zeros(18507 + 3)
    loss, acc, g = HO.loss_and_grads(p, emb, y)
    return np.concatenate([g["w1"].ravel() * n, g["b1"] * n, g["w2"].ravel() * n, g["b2"] * n, [loss * n, acc * n, n]])


def worker(rank, world, port, emb, y, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = HO.init_head(0)
    opt = HO.Adam(p, 1e-3)
    hist = []
    for step in range(5):
        lo, hi = shard(emb.shape[0], rank, world)
        flat = torch.from_numpy(flat_from_oracle(p, emb[lo:hi], y[lo:hi]))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)            # the ONE collective per step
        f = flat.numpy()
        n = f[-1]
        g = dict(w1=f[:18432].reshape(1024, 18) / n, b1=f[18432:18450] / n, w2=f[18450:18504].reshape(18, 3) / n,
                 b2=f[18504:18507] / n)
        opt.step(p, g)
        hist.append((f[-3] / n, f[-2] / n))
    out_q.put((rank, hist, p["w1"].copy(), p["b2"].copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 33])
def test_two_ranks_equal_single_process(n):
    rng = np.random.default_rng(0)
    emb = rng.normal(0, 1, (n, 1024)).astype(np.float32)
    y = rng.integers(0, 3, n)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, emb, y, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = HO.init_head(0)
    want_hist = HO.train(p, emb, y, 5, lr=1e-3)
    for rank, hist, w1, b2 in res:
        assert np.allclose(hist, want_hist, rtol=1e-9, atol=1e-12)
        assert np.allclose(w1, p["w1"], atol=1e-7) and np.allclose(b2, p["b2"], atol=1e-7)
    assert np.array_equal(res[0][2], res[1][2])                # ranks stay bit-identical


def test_shards_partition_the_batch():
    for n in (0, 1, 7, 512, 8192):
        for world in (1, 2, 4, 8):
            edges = [shard(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
