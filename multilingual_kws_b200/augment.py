"""Device-resident waveform augmentation (``kws_augment_pcm`` / ``kws_spec_mask``, csrc/augment.cu).

The host draws the random DECISIONS of reference ``AudioDataset.augment`` / ``spec_augment``
(multilingual_kws/embedding/input_data.py:275-364) and packs them into one 32-byte plan item per clip; the samples
live in int16 banks on the GPU and never travel: a fine-tune batch costs a few KB of plan upload instead of
32 KB of float audio per clip.  No CPU fallback: the numpy augmentation in ``embedding/input_data.py`` is the host
mirror of the reference's API, this module is the device path the training pipeline uses.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib

AUG_ITEM = np.dtype([("mode", "<i4"), ("fg_index", "<i4"), ("shift", "<i4"), ("bg_index", "<i4"), ("bg_offset", "<i4"),
                     ("volume", "<f4"), ("reserved", "<i4", (2,))])
assert AUG_ITEM.itemsize == 32
MODE_CLIP, MODE_SILENCE, MODE_MIX = 0, 1, 2


def plan_item(mode=MODE_CLIP, fg_index=-1, shift=0, bg_index=-1, bg_offset=0, volume=0.0) -> np.void:
    it = np.zeros((), AUG_ITEM)
    it["mode"], it["fg_index"], it["shift"] = mode, fg_index, shift
    it["bg_index"], it["bg_offset"], it["volume"] = bg_index, bg_offset, np.float32(volume)
    return it


def exact_int16(wave: np.ndarray) -> np.ndarray:
    """float32 samples that came from 16-bit PCM (decode_wav: s / 32768) back to int16, refusing anything else."""
    w = np.asarray(wave, np.float32)
    s = np.rint(w * np.float32(32768.0))
    if not (np.array_equal(s.astype(np.float32) / np.float32(32768.0), w) and s.min(initial=0) >= -32768 and s.max(initial=0) <= 32767):
        raise ValueError("clip bank holds 16-bit PCM; this waveform is not exactly representable as int16 / 32768")
    return s.astype(np.int16)


class ClipBank:
    """Growable int16 [capacity, stride] tensor on the device; ``add`` returns the row index."""

    def __init__(self, n_samples: int, capacity: int = 64, device: Optional[torch.device] = None):
        self.n_samples = int(n_samples)
        self.stride = (self.n_samples + 7) // 8 * 8
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.data = torch.zeros((max(1, capacity), self.stride), dtype=torch.int16, device=self.device)
        self.count = 0

    def add(self, wave: np.ndarray) -> int:
        s = exact_int16(wave)
        if s.shape != (self.n_samples,):
            raise ValueError(f"clip bank rows hold {self.n_samples} samples, got {s.shape}")
        if self.count == self.data.shape[0]:
            grown = torch.zeros((2 * self.data.shape[0], self.stride), dtype=torch.int16, device=self.device)
            grown[:self.count] = self.data[:self.count]
            self.data = grown
        self.data[self.count, :self.n_samples] = torch.from_numpy(s).to(self.device)
        self.count += 1
        return self.count - 1


class DeviceAugmenter:
    def __init__(self, n_samples: int, background_data: Optional[np.ndarray] = None, device: Optional[torch.device] = None):
        """background_data: float32 [n_bg, max_len], zero padded (AudioDataset.background_data)."""
        self.n_samples = int(n_samples)
        if self.n_samples % 8:
            raise ValueError("desired_samples must be a multiple of 8")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.clips = ClipBank(self.n_samples, device=self.device)
        if background_data is not None and len(background_data):
            bg = exact_int16(background_data)
            stride = (bg.shape[1] + 7) // 8 * 8
            self.bg = torch.zeros((bg.shape[0], stride), dtype=torch.int16, device=self.device)
            self.bg[:, :bg.shape[1]] = torch.from_numpy(bg).to(self.device)
        else:
            self.bg = torch.zeros((0, 8), dtype=torch.int16, device=self.device)

    def run(self, plan: np.ndarray, return_audio: bool = False):
        """plan: AUG_ITEM array [B] -> int16 PCM [B, n_samples] on the device (and the float32 waveform on request)."""
        plan = np.ascontiguousarray(plan, dtype=AUG_ITEM)
        B = plan.shape[0]
        uses_fg = plan["mode"] != MODE_SILENCE
        uses_bg = plan["mode"] != MODE_CLIP
        if ((plan["mode"] < 0) | (plan["mode"] > 2)).any():
            raise ValueError("augmentation plan: unknown mode")
        if (uses_fg & ((plan["fg_index"] < 0) | (plan["fg_index"] >= self.clips.count))).any():
            raise ValueError("augmentation plan: foreground index outside the clip bank")
        if (uses_bg & ((plan["bg_index"] < 0) | (plan["bg_index"] >= self.bg.shape[0]) | (plan["bg_offset"] < 0) |
                       (plan["bg_offset"] + self.n_samples > self.bg.shape[1]))).any():
            raise ValueError("augmentation plan: background window outside the background bank")
        d_plan = torch.from_numpy(plan.view(np.uint8).reshape(B, 32)).to(self.device)
        pcm = torch.empty((B, self.n_samples), dtype=torch.int16, device=self.device)
        audio = torch.empty((B, self.n_samples), dtype=torch.float32, device=self.device) if return_audio else None
        _lib.check(_lib.lib().kws_augment_pcm(
            self.clips.data.data_ptr(), self.clips.count, self.clips.stride, self.bg.data_ptr(), self.bg.shape[0],
            self.bg.shape[1], d_plan.data_ptr(), B, self.n_samples, pcm.data_ptr(),
            audio.data_ptr() if return_audio else None, _lib.current_stream_ptr()), "kws_augment_pcm")
        return (pcm, audio) if return_audio else pcm


def spec_mask_(feats: torch.Tensor, bands: np.ndarray) -> torch.Tensor:
    """In-place spec_augment masks: feats float32 [B, T, F] on the device, bands int32 [B, 8]."""
    bands = np.ascontiguousarray(bands, dtype=np.int32)
    B, T, F = feats.shape
    if bands.shape != (B, 8):
        raise ValueError(f"spec_mask_: need bands of shape ({B}, 8), got {bands.shape}")
    if not (feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous()):
        raise ValueError("spec_mask_: feats must be a contiguous float32 CUDA tensor")
    d_b = torch.from_numpy(bands).to(feats.device)
    _lib.check(_lib.lib().kws_spec_mask(feats.data_ptr(), B, T, F, d_b.data_ptr(), _lib.current_stream_ptr()), "kws_spec_mask")
    return feats
