"""Host-buffer entry point of the hot path: pinned int16 PCM in, pinned fp32 embeddings out.

Work is cut into jobs of at most `sub_batch` clips.  Every job is three stream-ordered pieces:

    H2D  (upload stream)    pinned PCM        -> device slot s
    run  (compute stream)   frontend kernel + embedding graph on slot s
    D2H  (download stream)  device embeddings -> pinned result rows

with `depth` device slots cycled round-robin and `streams` compute streams taken in turn.  Consecutive jobs therefore
overlap on the device as well: the late layers of the network work on small feature maps and are latency-bound
(they leave most SMs idle), the early layers are throughput-bound, and the two phases of neighbouring jobs fill
each other's gaps.  Every slot has its own embedding workspace.  The two copy directions have their own streams (and their own copy
engines on the device), so the upload of job k+1 and the download of job k-1 overlap the kernels of job k, across
sub-batches of one call and across consecutive calls: `run_host()` only enqueues.  The embedding's CUDA graphs are
keyed by (device buffers, batch), so cycling a fixed set of slots replays them instead of re-capturing.

Slot reuse is ordered by events only (no host synchronisation inside `run_host`):
    upload(k)  waits  ran(k - depth)        the PCM slot has been consumed by the frontend
    run(k)     waits  uploaded(k), downloaded(k - depth), and whatever preceded run_host on the caller's stream
    download(k) waits ran(k)
(depth is a multiple of streams, so job k and job k - depth also share a compute stream.)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import hostmem
from .frontend import FEATURE_SCALE, MicroFrontend
from .model import EmbeddingModel


class EmbedPipeline:
    # (head SMs, tail SMs) of the throughput schedule when several compute streams overlap (0 = all): the tail of the
    # network (after block3b) is launch/latency-bound and sized for 72 of the 148 SMs so that it leaves room for the
    # throughput-bound head of the neighbouring job (measured: profiles/README.md item 11)
    SM_BUDGET = (132, 72)

    def __init__(self, frontend: MicroFrontend, model: EmbeddingModel, n_samples: int = 16000, sub_batch: int = 512,
                 depth: int = 4, streams: int = 2, sm_budget: Optional[tuple] = None):
        if depth < 2 or streams < 1 or depth % streams:
            raise ValueError("EmbedPipeline needs at least two device slots, depth a multiple of streams")
        self.fe, self.model, self.n, self.sub, self.depth = frontend, model, int(n_samples), int(sub_batch), int(depth)
        dev = model.device
        self._compute = [torch.cuda.Stream(device=dev) for _ in range(int(streams))]
        self._ws = [torch.empty(model.workspace_bytes(self.sub), dtype=torch.uint8, device=dev) for _ in range(depth)]
        frames = frontend.num_frames(self.n)
        self._pcm = [torch.empty((self.sub, self.n), dtype=torch.int16, device=dev) for _ in range(depth)]
        self._feat = [torch.empty((self.sub, frames, frontend.num_channels), dtype=torch.float32, device=dev)
                      for _ in range(depth)]
        self._emb = [torch.empty((self.sub, model.output_dim), dtype=torch.float32, device=dev) for _ in range(depth)]
        self._up = torch.cuda.Stream(device=dev)
        self._down = torch.cuda.Stream(device=dev)
        self._uploaded = [torch.cuda.Event() for _ in range(depth)]
        self._ran = [torch.cuda.Event() for _ in range(depth)]
        self._downloaded = [torch.cuda.Event() for _ in range(depth)]
        self._jobs = 0                      # jobs enqueued since construction (slot = job % depth)
        self.sm_budget = sm_budget if sm_budget is not None else (self.SM_BUDGET if int(streams) > 1 else None)
        # The library captures a forward pass into a CUDA graph the second time it meets the same buffers (one-off
        # buffers are not worth a capture).  The slots are fixed, so meet each of them twice now: the first full job on
        # every slot is then already a graph replay instead of a capture inside the caller's latency.
        for s in range(depth):
            with torch.cuda.stream(self._compute[s % len(self._compute)]):
                self._feat[s].zero_()
                for _ in range(2):
                    self.model.forward_device(self._feat[s], out=self._emb[s], workspace=self._ws[s], sm_budget=self.sm_budget)
        for c in self._compute:
            c.synchronize()

    def alloc_input(self, batch: int) -> torch.Tensor:
        """int16 [batch, n_samples] upload buffer in write-combined pinned memory (fill it with PCM, do not read it
        back on the CPU): uploads from it run at the PCIe line rate, about twice as fast as from `pin_memory()`."""
        return hostmem.upload_buffer((int(batch), self.n), torch.int16)

    def alloc_output(self, batch: int) -> torch.Tensor:
        """fp32 [batch, out_dim] pinned result buffer."""
        return hostmem.download_buffer((int(batch), self.model.output_dim), torch.float32)

    def run_host(self, pcm_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pcm_host: page-locked int16 [B, n_samples] (`alloc_input()` or any pinned tensor); returns pinned fp32
        [B, out_dim].

        Asynchronous: the call enqueues copies and kernels and returns.  The rows are valid after `join()` followed by
        a synchronisation of the current stream (or simply `synchronize()`); neither buffer may be modified before
        that.  Consecutive calls overlap each other's copies and kernels."""
        if pcm_host.dtype != torch.int16 or pcm_host.dim() != 2 or pcm_host.shape[1] != self.n:
            raise ValueError(f"pcm_host must be int16 [B, {self.n}]")
        B = pcm_host.shape[0]
        if out_host is None:
            out_host = self.alloc_output(B)
        caller = torch.cuda.current_stream()
        for b0 in range(0, B, self.sub):
            nb = min(self.sub, B - b0)
            k, s = self._jobs, self._jobs % self.depth
            compute = self._compute[k % len(self._compute)]
            compute.wait_stream(caller)
            with torch.cuda.stream(self._up):
                if k >= self.depth:
                    self._up.wait_event(self._ran[s])
                self._pcm[s][:nb].copy_(pcm_host[b0:b0 + nb], non_blocking=True)
                self._uploaded[s].record(self._up)
            compute.wait_event(self._uploaded[s])
            if k >= self.depth:
                compute.wait_event(self._downloaded[s])
            with torch.cuda.stream(compute):
                self.fe.forward(self._pcm[s][:nb], out_scale=FEATURE_SCALE, out=self._feat[s][:nb])
                self.model.forward_device(self._feat[s][:nb], out=self._emb[s][:nb], workspace=self._ws[s],
                                          sm_budget=self.sm_budget)
                self._ran[s].record(compute)
            with torch.cuda.stream(self._down):
                self._down.wait_event(self._ran[s])
                out_host[b0:b0 + nb].copy_(self._emb[s][:nb], non_blocking=True)
                self._downloaded[s].record(self._down)
            self._jobs += 1
        return out_host

    def join(self) -> None:
        """Order the current stream after every copy enqueued so far (an event recorded on it afterwards covers the
        last download)."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self._down)
        cur.wait_stream(self._up)
        for c in self._compute:
            cur.wait_stream(c)

    def synchronize(self) -> None:
        """Block the host until every enqueued job has delivered its rows."""
        for c in self._compute:
            c.synchronize()
        self._down.synchronize()
        self._up.synchronize()
