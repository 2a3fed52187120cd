"""Host-buffer entry point of the hot path: pinned int16 PCM in, pinned fp32 embeddings out.

The batch is cut into sub-batches; the H2D copy of sub-batch i+1 and the D2H copy of sub-batch i-1 run on a copy
stream while sub-batch i is in the frontend + embedding kernels on the compute stream (double-buffered device
buffers, so the embedding's CUDA graphs are replayed, not re-captured)."""
from __future__ import annotations

from typing import Optional

import torch

from .frontend import FEATURE_SCALE, MicroFrontend
from .model import EmbeddingModel


class EmbedPipeline:
    def __init__(self, frontend: MicroFrontend, model: EmbeddingModel, n_samples: int = 16000, sub_batch: int = 512):
        self.fe, self.model, self.n, self.sub = frontend, model, int(n_samples), int(sub_batch)
        dev = model.device
        frames = frontend.num_frames(self.n)
        self._pcm = [torch.empty((self.sub, self.n), dtype=torch.int16, device=dev) for _ in range(2)]
        self._feat = [torch.empty((self.sub, frames, frontend.num_channels), dtype=torch.float32, device=dev) for _ in range(2)]
        self._emb = [torch.empty((self.sub, model.output_dim), dtype=torch.float32, device=dev) for _ in range(2)]
        self._copy = torch.cuda.Stream(device=dev)
        self._h2d = [torch.cuda.Event() for _ in range(2)]
        self._done = [torch.cuda.Event() for _ in range(2)]
        self._d2h = [torch.cuda.Event() for _ in range(2)]

    def run_host(self, pcm_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pcm_host: pinned int16 [B, n_samples]; returns pinned fp32 [B, out_dim] (valid after the call's final sync
        by the caller, e.g. torch.cuda.synchronize() or an event on the current stream)."""
        B = pcm_host.shape[0]
        if out_host is None:
            out_host = torch.empty((B, self.model.output_dim), dtype=torch.float32).pin_memory()
        compute = torch.cuda.current_stream()
        self._copy.wait_stream(compute)
        k = 0
        for b0 in range(0, B, self.sub):
            nb = min(self.sub, B - b0)
            i = k & 1
            with torch.cuda.stream(self._copy):
                if k >= 2:
                    self._copy.wait_event(self._done[i])          # buffer i free again (its compute finished)
                self._pcm[i][:nb].copy_(pcm_host[b0:b0 + nb], non_blocking=True)
                self._h2d[i].record(self._copy)
            compute.wait_event(self._h2d[i])
            if k >= 2:
                compute.wait_event(self._d2h[i])                  # previous result in buffer i has left the device
            self.fe.forward(self._pcm[i][:nb], out_scale=FEATURE_SCALE, out=self._feat[i][:nb])
            self.model.forward_device(self._feat[i][:nb], out=self._emb[i][:nb])
            self._done[i].record(compute)
            with torch.cuda.stream(self._copy):
                self._copy.wait_event(self._done[i])
                out_host[b0:b0 + nb].copy_(self._emb[i][:nb], non_blocking=True)
                self._d2h[i].record(self._copy)
            k += 1
        compute.wait_stream(self._copy)
        return out_host
