"""Host wrapper of the fused sm_100a micro-frontend kernel (C ABI kws_frontend_*).

Mirrors TF's ``frontend_op.audio_microfrontend`` as the reference calls it
(multilingual_kws/embedding/input_data.py:25-33) but takes whole batches that stay on the GPU.
PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib

# Attributes of the TF op the reference leaves at their defaults.
OP_DEFAULTS = dict(lower_band_limit=125.0, upper_band_limit=7500.0, smoothing_bits=10, even_smoothing=0.025,
                   odd_smoothing=0.06, min_signal_remaining=0.05, enable_pcan=True, pcan_strength=0.95,
                   pcan_offset=80.0, gain_bits=21, enable_log=True, scale_shift=6)

FEATURE_SCALE = 10.0 / 256.0   # input_data.py:34


class MicroFrontend:
    """One immutable frontend configuration living on the current CUDA device."""

    def __init__(self, sample_rate: int = 16000, window_size_ms: int = 30, window_step_ms: int = 20,
                 num_channels: int = 40, **attrs):
        cfg = dict(OP_DEFAULTS)
        unknown = set(attrs) - set(cfg)
        if unknown:
            raise TypeError(f"unknown audio_microfrontend attributes: {sorted(unknown)}")
        cfg.update(attrs)
        self.sample_rate = int(sample_rate)
        self.window_size_ms = int(window_size_ms)   # the op's attrs are ints; TF's wrapper applies int()
        self.window_step_ms = int(window_step_ms)
        self.num_channels = int(num_channels)
        self.window_size = self.window_size_ms * self.sample_rate // 1000
        self.window_step = self.window_step_ms * self.sample_rate // 1000
        self._h = ctypes.c_void_p()
        L = _lib.lib()
        _lib.check(L.kws_frontend_create(
            ctypes.byref(self._h), self.sample_rate, self.window_size_ms, self.window_step_ms, self.num_channels,
            cfg["lower_band_limit"], cfg["upper_band_limit"], int(cfg["smoothing_bits"]), cfg["even_smoothing"],
            cfg["odd_smoothing"], cfg["min_signal_remaining"], int(bool(cfg["enable_pcan"])), cfg["pcan_strength"],
            cfg["pcan_offset"], int(cfg["gain_bits"]), int(bool(cfg["enable_log"])), int(cfg["scale_shift"])),
            "kws_frontend_create")
        self.device = torch.device("cuda", torch.cuda.current_device())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().kws_frontend_destroy(h)
            except Exception:
                pass
            self._h = None

    def num_frames(self, n_samples: int) -> int:
        return int(_lib.lib().kws_frontend_num_frames(self._h, int(n_samples)))

    def tables_bytes(self) -> bytes:
        need = ctypes.c_size_t()
        L = _lib.lib()
        _lib.check(L.kws_frontend_tables(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _lib.check(L.kws_frontend_tables(self._h, buf, need.value, None))
        return buf.raw

    def forward(self, pcm: torch.Tensor, out_scale: float = FEATURE_SCALE, out: Optional[torch.Tensor] = None,
                raw_u16: bool = False) -> torch.Tensor:
        """pcm: CUDA int16 [B, n] (or [n]) -> float32 [B, frames, C] = uint16 * out_scale.

        With raw_u16=True returns the op's raw uint16 values (as a torch.uint16 tensor)."""
        if pcm.dtype != torch.int16 or not pcm.is_cuda:
            raise TypeError("pcm must be a CUDA int16 tensor")
        squeeze = pcm.dim() == 1
        if squeeze:
            pcm = pcm[None]
        if pcm.dim() != 2:
            raise ValueError("audio is not a vector")   # the op's InvalidArgument text
        pcm = pcm.contiguous()
        B, n = pcm.shape
        frames = self.num_frames(n)
        if raw_u16:
            res = torch.empty((B, frames, self.num_channels), dtype=torch.uint16, device=pcm.device)
            f32_ptr, u16_ptr = None, res.data_ptr()
        else:
            res = out if out is not None else torch.empty((B, frames, self.num_channels), dtype=torch.float32,
                                                          device=pcm.device)
            if res.shape != (B, frames, self.num_channels) or res.dtype != torch.float32 or not res.is_contiguous():
                raise ValueError("out has the wrong shape/dtype/layout")
            f32_ptr, u16_ptr = res.data_ptr(), None
        if B and frames:
            _lib.check(_lib.lib().kws_frontend_forward(self._h, pcm.data_ptr(), B, n, float(out_scale), f32_ptr, u16_ptr,
                                                       _lib.current_stream_ptr()), "kws_frontend_forward")
        return res[0] if squeeze else res

    __call__ = forward

    # ---- streaming (batch_streaming_analysis.py:66-115) ----
    def stream_num_windows(self, total_samples: int, clip_samples: int, hop_samples: int) -> int:
        return int(_lib.lib().kws_frontend_stream_num_windows(self._h, int(total_samples), int(clip_samples),
                                                               int(hop_samples)))

    def stream_prepare(self, pcm: torch.Tensor) -> "StreamState":
        """Computes the per-frame magnitudes of a long signal once (frame-reuse path)."""
        if pcm.dtype != torch.int16 or not pcm.is_cuda or pcm.dim() != 1:
            raise TypeError("pcm must be a 1-D CUDA int16 tensor")
        pcm = pcm.contiguous()
        nbytes = int(_lib.lib().kws_frontend_stream_scratch_bytes(self._h, pcm.numel()))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=pcm.device)
        return StreamState(self, pcm, scratch)


class StreamState:
    def __init__(self, fe: MicroFrontend, pcm: torch.Tensor, scratch: torch.Tensor):
        self.fe, self.pcm, self.scratch, self.ready = fe, pcm, scratch, False

    def windows(self, clip_samples: int, hop_samples: int, first: int, count: int,
                out_scale: float = FEATURE_SCALE) -> torch.Tensor:
        fe = self.fe
        frames = fe.num_frames(clip_samples)
        out = torch.empty((count, frames, fe.num_channels), dtype=torch.float32, device=self.pcm.device)
        _lib.check(_lib.lib().kws_frontend_stream(
            fe._h, self.pcm.data_ptr(), self.pcm.numel(), int(clip_samples), int(hop_samples), int(first), int(count),
            float(out_scale), out.data_ptr() if count else None, self.scratch.data_ptr(), int(self.ready),
            _lib.current_stream_ptr()), "kws_frontend_stream")
        self.ready = True
        return out


def float_audio_to_int16(audio: torch.Tensor) -> torch.Tensor:
    """tf.cast(tf.multiply(audio, 32768), tf.int16) (input_data.py:23): truncate toward zero; +1.0
    wraps to -32768 like the x86 TF kernel does."""
    return (audio.to(torch.float32) * 32768.0).to(torch.int32).to(torch.int16)


def float_audio_to_int16_np(audio: np.ndarray) -> np.ndarray:
    return (np.asarray(audio, dtype=np.float32) * np.float32(32768.0)).astype(np.int32).astype(np.int16)
