"""Drop-in mirror of reference ``multilingual_kws/embedding/input_data.py`` on the B200 kernels.

Same public names, argument meaning and error behaviour; the arithmetic runs in libkws_b200.so:
  to_micro_spectrogram / file2spec      -> fused fixed-point frontend kernel   (ref :19-47)
  prepare_model_settings / standard_... -> identical dict                       (ref :63-138)
  add_background, SpecAugParams, AudioDataset (augment, spec_augment, init_*)   (ref :141-556)
Differences a caller can see: tensors are numpy / torch instead of tf.Tensor; datasets are small Python
iterables (``.shuffle().repeat().batch()``) whose ``batch`` evaluates the frontend for the whole batch
in one kernel launch; random draws come from numpy's Generator (the reference itself mixes the global
TF RNG with its own generator, SURVEY.md §5.9f, so streams were never reproducible across versions).
"""
from __future__ import annotations

import glob
import math
import os
import struct
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..augment import AUG_ITEM, MODE_CLIP, MODE_MIX, MODE_SILENCE, DeviceAugmenter, plan_item, spec_mask_
from ..frontend import FEATURE_SCALE, MicroFrontend, float_audio_to_int16_np

SILENCE_LABEL = "_silence_"
SILENCE_INDEX = 0
UNKNOWN_WORD_LABEL = "_unknown_"
UNKNOWN_WORD_INDEX = 1

_frontends = {}


def _frontend_for(model_settings) -> MicroFrontend:
    sample_rate = model_settings["sample_rate"]
    window_size_ms = (model_settings["window_size_samples"] * 1000) / sample_rate      # ref :21
    window_step_ms = (model_settings["window_stride_samples"] * 1000) / sample_rate    # ref :22
    key = (torch.cuda.current_device(), sample_rate, int(window_size_ms), int(window_step_ms),
           model_settings["fingerprint_width"])
    fe = _frontends.get(key)
    if fe is None:
        fe = MicroFrontend(sample_rate=sample_rate, window_size_ms=int(window_size_ms),
                           window_step_ms=int(window_step_ms), num_channels=model_settings["fingerprint_width"])
        _frontends[key] = fe
    return fe


def to_micro_spectrogram(model_settings, audio):
    """float audio in [-1, 1], shape [n] (or a batch [B, n]) -> float32 features [frames, channels]
    (or [B, frames, channels]) as a numpy array; torch CUDA input gives torch CUDA output."""
    fe = _frontend_for(model_settings)
    if isinstance(audio, torch.Tensor):
        if audio.dim() not in (1, 2):
            raise ValueError("audio is not a vector")
        pcm = (audio.to(torch.float32) * 32768.0).to(torch.int32).to(torch.int16).cuda()
        return fe.forward(pcm, out_scale=FEATURE_SCALE)
    a = np.asarray(audio)
    if a.ndim not in (1, 2):
        raise ValueError("audio is not a vector")
    pcm = torch.from_numpy(float_audio_to_int16_np(a)).cuda()
    return fe.forward(pcm, out_scale=FEATURE_SCALE).cpu().numpy()


def decode_wav(data: bytes, desired_channels: int = 1, desired_samples: int = -1) -> Tuple[np.ndarray, int]:
    """tf.audio.decode_wav for 16-bit PCM RIFF files: float32 [samples, channels] in [-1, 1), sample rate."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("Header mismatch: Expected RIFF/WAVE")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            pcm = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError("Bad WAV file: missing fmt or data chunk")
    audio_format, channels, rate, _, _, bits = fmt
    if audio_format != 1 or bits != 16:
        raise ValueError("Can only read 16-bit PCM WAV files")
    x = np.frombuffer(pcm[:len(pcm) // (2 * channels) * 2 * channels], dtype="<i2").reshape(-1, channels)
    x = x.astype(np.float32) * np.float32(1.0 / 32768.0)
    if desired_channels > 0:
        x = x[:, :desired_channels] if channels >= desired_channels else np.repeat(x[:, :1], desired_channels, axis=1)
    if desired_samples > 0:
        if x.shape[0] >= desired_samples:
            x = x[:desired_samples]
        else:
            x = np.concatenate([x, np.zeros((desired_samples - x.shape[0], x.shape[1]), np.float32)])
    return x, rate


def encode_wav(path: str, audio: np.ndarray, sample_rate: int = 16000) -> None:
    pcm = np.clip(np.rint(np.asarray(audio, np.float64) * 32768.0), -32768, 32767).astype("<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVEfmt " +
                struct.pack("<IHHIIHH", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16) + b"data" +
                struct.pack("<I", len(pcm)) + pcm)


def _read_wav(filepath, desired_samples: int = -1) -> np.ndarray:
    with open(os.fspath(filepath), "rb") as f:
        audio, _ = decode_wav(f.read(), desired_channels=1, desired_samples=desired_samples)
    return audio[:, 0]


def file2spec(model_settings, filepath):
    """there's a version of this that adds bg noise in AudioDataset"""
    audio = _read_wav(filepath, model_settings["desired_samples"])
    return to_micro_spectrogram(model_settings, audio)


def files2specs(model_settings, filepaths: Sequence) -> np.ndarray:
    """Batched file2spec: one frontend launch for all files -> [N, frames, channels]."""
    if len(filepaths) == 0:
        return np.zeros((0, model_settings["spectrogram_length"], model_settings["fingerprint_width"]), np.float32)
    audio = np.stack([_read_wav(f, model_settings["desired_samples"]) for f in filepaths])
    return to_micro_spectrogram(model_settings, audio)


def _next_power_of_two(x):
    return 1 if x == 0 else 2 ** (int(x) - 1).bit_length()


def prepare_model_settings(label_count, sample_rate, clip_duration_ms, window_size_ms, window_stride_ms,
                           feature_bin_count, preprocess):
    """Calculates common settings needed for all models (reference input_data.py:63-126)."""
    desired_samples = int(sample_rate * clip_duration_ms / 1000)
    window_size_samples = int(sample_rate * window_size_ms / 1000)
    window_stride_samples = int(sample_rate * window_stride_ms / 1000)
    length_minus_window = desired_samples - window_size_samples
    if length_minus_window < 0:
        spectrogram_length = 0
    else:
        spectrogram_length = 1 + int(length_minus_window / window_stride_samples)
    if preprocess == "average":
        fft_bin_count = 1 + (_next_power_of_two(window_size_samples) / 2)
        average_window_width = int(math.floor(fft_bin_count / feature_bin_count))
        fingerprint_width = int(math.ceil(fft_bin_count / average_window_width))
    elif preprocess == "mfcc":
        average_window_width = -1
        fingerprint_width = feature_bin_count
    elif preprocess == "micro":
        average_window_width = -1
        fingerprint_width = feature_bin_count
    else:
        raise ValueError('Unknown preprocess mode "%s" (should be "mfcc",'
                         ' "average", or "micro")' % (preprocess))
    fingerprint_size = fingerprint_width * spectrogram_length
    return {
        "desired_samples": desired_samples,
        "window_size_samples": window_size_samples,
        "window_stride_samples": window_stride_samples,
        "spectrogram_length": spectrogram_length,
        "fingerprint_width": fingerprint_width,
        "fingerprint_size": fingerprint_size,
        "label_count": label_count,
        "sample_rate": sample_rate,
        "preprocess": preprocess,
        "average_window_width": average_window_width,
    }


def standard_microspeech_model_settings(label_count: int):
    return prepare_model_settings(label_count=label_count, sample_rate=16000, clip_duration_ms=1000,
                                  window_size_ms=30, window_stride_ms=20, feature_bin_count=40, preprocess="micro")


def add_background(foreground_audio, background_audio, background_volume):
    """reference input_data.py:141-157: float32 arithmetic throughout; the two mean squares are float32 squares
    accumulated in float64 and rounded once (TF's reduce_mean sums float32 in an unspecified tree order, so its last
    bit is not reproducible anyway) — the same definition the device kernel uses (csrc/augment.cu), which makes the
    host mirror and the GPU path agree bit for bit."""
    fg = np.asarray(foreground_audio, np.float32)
    bg = np.asarray(background_audio, np.float32)
    foreground_rms = np.sqrt(np.float32(np.mean(np.square(fg), dtype=np.float64)))
    background_rms = np.sqrt(np.float32(np.mean(np.square(bg), dtype=np.float64)))
    snr_scaling = np.float32(foreground_rms / background_rms) if background_rms > 0 else np.float32(0.0)
    bg_add = bg * snr_scaling * np.float32(background_volume) + fg
    return np.clip(bg_add, -1.0, 1.0).astype(np.float32)


@dataclass(frozen=True)
class SpecAugParams:
    percentage: float = 80.0
    frequency_n_range: int = 2   # how many augmentations to include, inclusive
    frequency_max_px: int = 2    # how large each mask should be (pixels)
    time_n_range: int = 2
    time_max_px: int = 2


class _Dataset:
    """Minimal stand-in for the tf.data pipeline the reference builds: an iterable of (waveform, label_id,
    spec-aug masks) elements with shuffle/repeat/batch/take; ``batch`` runs the frontend per batch on the GPU."""

    def __init__(self, owner: "AudioDataset", make_elements, n_elements: Optional[int], training: bool):
        self.owner, self._make, self._n, self.training = owner, make_elements, n_elements, training
        self._shuffle = 0
        self._repeat = False
        self._batch = None
        self._take = None
        self.fast = None          # {"rows": bank rows of the files, "labels": their label ids} -> vectorised batches

    def _clone(self):
        d = _Dataset(self.owner, self._make, self._n, self.training)
        d._shuffle, d._repeat, d._batch, d._take = self._shuffle, self._repeat, self._batch, self._take
        d.fast = self.fast
        return d

    def shuffle(self, buffer_size: int):
        d = self._clone(); d._shuffle = int(buffer_size); return d

    def repeat(self):
        d = self._clone(); d._repeat = True; return d

    def batch(self, batch_size: int):
        d = self._clone(); d._batch = int(batch_size); return d

    def take(self, n: int):
        d = self._clone(); d._take = int(n); return d

    def __len__(self):
        if self._n is None or self._repeat:
            raise TypeError("dataset has no finite length")
        return self._n if self._batch is None else -(-self._n // self._batch)

    def _elements(self) -> Iterator:
        rng = self.owner.gen
        while True:
            it = self._make()
            if self._shuffle > 1:
                buf = []
                for e in it:
                    buf.append(e)
                    if len(buf) >= self._shuffle:
                        yield buf.pop(int(rng.integers(0, len(buf))))
                while buf:
                    yield buf.pop(int(rng.integers(0, len(buf))))
            else:
                yield from it
            if not self._repeat:
                return

    def _finish(self, waves: List[np.ndarray], labels: List[int]):
        o = self.owner
        if o.device_augment:      # elements are augmentation plan items: samples are mixed on the GPU (augment.py)
            pcm = o._augmenter().run(np.stack(waves).astype(AUG_ITEM, copy=False))
            specs = _frontend_for(o.model_settings).forward(pcm, out_scale=FEATURE_SCALE)
        else:
            specs = to_micro_spectrogram(o.model_settings, torch.from_numpy(np.stack(waves)))    # [n, T, F] on the GPU
        if self.training:
            specs = o._spec_augment_batch(specs)
        return specs[..., None], torch.as_tensor(np.asarray(labels, np.int64))

    def _fast_batches(self):
        """Vectorised training batches (AudioDataset(device_augment="batched")): the random DECISIONS of `augment` /
        `map_spec_aug` for a whole batch are drawn as numpy arrays (same policy and probabilities, reference
        input_data.py:275-369: time shift, then silence | unknown | background mix | plain, then spec-augment masks), the
        samples never leave the GPU.  The per-element generator costs ~27 us of Python per clip (14 ms per 512-clip batch
        against 0.34 ms of device time); this path costs a few numpy calls per batch.  The stream of random numbers differs
        from the per-element path (which the host/device equivalence tests pin), the distributions do not."""
        o, rng = self.owner, self.owner.gen
        rows, labs = self.fast["rows"], self.fast["labels"]
        nf, B = rows.shape[0], self._batch
        unk_rows = o._unknown_rows()
        sil_id, unk_id = o.label_id(SILENCE_LABEL), o.label_id(UNKNOWN_WORD_LABEL)
        desired = o.model_settings["desired_samples"]
        m = o.max_time_shift_samples
        p = o.spec_aug_params
        fe = _frontend_for(o.model_settings)
        order, pos, count = None, nf, 0
        while self._take is None or count < self._take:
            idx = np.empty(B, np.int64)
            filled = 0
            while filled < B:                      # shuffle buffer >= number of files: a fresh permutation per pass
                if pos >= nf:
                    order, pos = (rng.permutation(nf) if self._shuffle > 1 else np.arange(nf)), 0
                take = min(B - filled, nf - pos)
                idx[filled:filled + take] = order[pos:pos + take]
                pos += take
                filled += take
            u = rng.uniform(0, 1, (4, B))
            sil = u[0] < o.silence_percentage / 100
            unk = ~sil & (unk_rows.shape[0] > 0) & (u[1] < o.unknown_percentage / 100)
            mix = ~sil & ~unk & (u[2] < o.background_frequency)
            shift = rng.integers(-m, m, (2, B)) if m > 0 else np.zeros((2, B), np.int64)
            bg_index = rng.integers(0, o.background_sizes.shape[0], B)
            bg_off = rng.integers(0, o.background_sizes[bg_index] - desired)
            plan = np.zeros(B, AUG_ITEM)
            plan["mode"] = np.where(sil, MODE_SILENCE, np.where(mix, MODE_MIX, MODE_CLIP))
            fg = np.where(unk, unk_rows[rng.integers(0, max(unk_rows.shape[0], 1), B)] if unk_rows.shape[0] else 0, rows[idx])
            plan["fg_index"] = np.where(sil, -1, fg)
            plan["shift"] = np.where(sil, 0, np.where(unk, shift[1], shift[0]))
            uses_bg = sil | mix
            plan["bg_index"] = np.where(uses_bg, bg_index, -1)
            plan["bg_offset"] = np.where(uses_bg, bg_off, 0)
            plan["volume"] = np.where(sil, rng.uniform(0, 1, B), np.where(mix, rng.uniform(0, o.background_volume_range, B), 0.0))
            labels = np.where(sil, sil_id, np.where(unk, unk_id, labs[idx])).astype(np.int64)
            specs = fe.forward(o._augmenter().run(plan), out_scale=FEATURE_SCALE)
            n, t, f = specs.shape
            bands = np.zeros((B, 8), np.int32)
            apply = u[3] < p.percentage / 100
            for col, n_range, max_px, extent in ((0, p.frequency_n_range, p.frequency_max_px, f), (4, p.time_n_range, p.time_max_px, t)):
                if n_range > 2:
                    raise ValueError("batched spec-augment supports up to two bands per axis")
                n_b = rng.integers(0, n_range + 1, B)
                for j in range(n_range):
                    size = rng.integers(1, max_px + 1, B)
                    start = rng.integers(0, extent - size)
                    on = apply & (j < n_b)
                    bands[on, col + 2 * j] = start[on]
                    bands[on, col + 2 * j + 1] = size[on]
            specs = spec_mask_(specs.contiguous(), bands)
            yield specs[..., None], torch.from_numpy(labels)
            count += 1

    def __iter__(self):
        count = 0
        if self.fast is not None and self.training and self._batch is not None and self._repeat:
            yield from self._fast_batches()
            return
        if self._batch is None:
            for wave, label in self._elements():
                if self._take is not None and count >= self._take:
                    return
                s, l = self._finish([wave], [label])
                yield s[0], l[0]
                count += 1
            return
        waves, labels = [], []
        for wave, label in self._elements():
            waves.append(wave); labels.append(label)
            if len(waves) == self._batch:
                if self._take is not None and count >= self._take:
                    return
                yield self._finish(waves, labels)
                count += 1
                waves, labels = [], []
        if waves and (self._take is None or count < self._take):
            yield self._finish(waves, labels)


class AudioDataset:
    def __init__(self, model_settings, commands, background_data_dir, unknown_files, time_shift_ms=100,
                 background_frequency=0.8, background_volume_range=0.1, silence_percentage=10.0,
                 unknown_percentage=10.0, spec_aug_params=SpecAugParams(), seed=None, device_augment=False) -> None:
        """Arguments as the reference (:173-187).  device_augment=True (not in the reference) keeps the decoded clips
        in an int16 bank on the GPU and runs time shift / background mix / int16 cast / spec-augment masks there
        (augment.py); the random draws, and therefore the batches, are identical to the host path for the same seed.
        device_augment="batched" additionally draws the decisions of a whole training batch at once (_fast_batches)."""
        self.device_augment = bool(device_augment)
        self.batched = device_augment == "batched"
        self._unk_rows = None
        self._aug = None
        self._bank_rows = {}
        self.model_settings = model_settings
        self.get_background_data(background_data_dir)
        self.max_time_shift_samples = self.timeshift_samples(time_shift_ms=time_shift_ms)
        self.background_frequency = background_frequency   # freq. between 0-1
        self.background_volume_range = background_volume_range
        # order-sensitive (unknown, then silence) so labels are always [silence, unknown, word1, word2, ...]
        commands = list(commands)
        self.unknown_percentage = unknown_percentage
        self.unknown_files = list(unknown_files)
        if len(self.unknown_files) > 0 and self.unknown_percentage > 0:
            commands = [UNKNOWN_WORD_LABEL] + commands
        self.silence_percentage = silence_percentage        # pct between 0-100
        if self.silence_percentage > 0:
            commands = [SILENCE_LABEL] + commands
        self.commands = np.asarray(commands)
        self.spec_aug_params = spec_aug_params
        self.gen = np.random.default_rng(seed if seed else None)

    def timeshift_samples(self, time_shift_ms=100):
        return int((time_shift_ms * self.model_settings["sample_rate"]) / 1000)

    # ---- augmentations (reference :227-369)
    def random_background_sample(self, background_volume=1.0):
        desired_samples = self.model_settings["desired_samples"]
        background_index = int(self.gen.integers(0, self.background_sizes.shape[0]))
        wav_length = int(self.background_sizes[background_index])
        background_offset = int(self.gen.integers(0, wav_length - desired_samples))
        clipped = self.background_data[background_index, background_offset:background_offset + desired_samples]
        return (clipped * np.float32(background_volume)).reshape(desired_samples)

    def random_timeshift(self, audio):
        desired_samples = self.model_settings["desired_samples"]
        amount = int(self.gen.integers(-self.max_time_shift_samples, self.max_time_shift_samples))
        if amount > 0:
            padded, offset = np.concatenate([np.zeros(amount, np.float32), audio]), 0      # pad beginning
        else:
            padded, offset = np.concatenate([audio, np.zeros(-amount, np.float32)]), -amount
        return padded[offset:offset + desired_samples]

    def get_unknown(self):
        return self.decode_audio(self.unknown_files[int(self.gen.integers(0, len(self.unknown_files)))])

    def augment(self, audio, label):
        audio = self.random_timeshift(audio) if self.max_time_shift_samples > 0 else audio
        if self.gen.uniform(0, 1) < (self.silence_percentage / 100):
            background_volume = self.gen.uniform(0, 1)
            label = SILENCE_LABEL
            audio = self.random_background_sample(background_volume)
        elif len(self.unknown_files) > 0 and self.gen.uniform(0, 1) < (self.unknown_percentage / 100):
            audio = self.get_unknown()
            audio = self.random_timeshift(audio) if self.max_time_shift_samples > 0 else audio
            label = UNKNOWN_WORD_LABEL
        elif self.gen.uniform(0, 1) < self.background_frequency:      # mix in background?
            background_volume = self.gen.uniform(0, self.background_volume_range)
            audio = add_background(audio, self.random_background_sample(), background_volume)
        return audio, label

    def _spec_aug_bands(self, time_max: int, freq_max: int):
        """One draw of reference spec_augment (:306-364): ([(f_start, f_size), ...], [(t_start, t_size), ...])."""
        p = self.spec_aug_params
        fbands, tbands = [], []
        freq_n = int(self.gen.integers(0, p.frequency_n_range + 1))      # both counts are drawn before any band (:318-323)
        time_n = int(self.gen.integers(0, p.time_n_range + 1))
        for _ in range(freq_n):
            size = int(self.gen.integers(1, p.frequency_max_px + 1))
            fbands.append((int(self.gen.integers(0, freq_max - size)), size))
        for _ in range(time_n):
            size = int(self.gen.integers(1, p.time_max_px + 1))
            tbands.append((int(self.gen.integers(0, time_max - size)), size))
        return fbands, tbands

    def _spec_aug_mask(self, time_max: int, freq_max: int) -> np.ndarray:
        """The same draw as a multiplicative 0/1 mask [T, F]."""
        fbands, tbands = self._spec_aug_bands(time_max, freq_max)
        mask = np.ones((time_max, freq_max), np.float32)
        for start, size in fbands:
            mask[:, start:start + size] = 0.0
        for start, size in tbands:
            mask[start:start + size, :] = 0.0
        return mask

    def spec_augment(self, spectrogram):
        s = np.asarray(spectrogram)
        return s * self._spec_aug_mask(s.shape[0], s.shape[1])

    def map_spec_aug(self, spectrogram, label_id):
        if self.gen.uniform(0, 1) < (self.spec_aug_params.percentage / 100):
            spectrogram = self.spec_augment(spectrogram)
        return spectrogram, label_id

    def _spec_augment_batch(self, specs: torch.Tensor) -> torch.Tensor:
        n, t, f = specs.shape
        p = self.spec_aug_params
        if self.device_augment and p.frequency_n_range <= 2 and p.time_n_range <= 2:
            bands = np.zeros((n, 8), np.int32)       # 16 bytes of plan per clip instead of a [T, F] float mask
            for i in range(n):
                if self.gen.uniform(0, 1) < (p.percentage / 100):
                    fb, tb = self._spec_aug_bands(t, f)
                    bands[i, :2 * len(fb)] = np.asarray(fb, np.int32).reshape(-1)
                    bands[i, 4:4 + 2 * len(tb)] = np.asarray(tb, np.int32).reshape(-1)
            return spec_mask_(specs.contiguous(), bands)
        masks = np.ones((n, t, f), np.float32)
        for i in range(n):
            if self.gen.uniform(0, 1) < (self.spec_aug_params.percentage / 100):
                masks[i] = self._spec_aug_mask(t, f)
        return specs * torch.from_numpy(masks).to(specs.device)

    # ---- device path: the same decisions as a plan item (augment.py), same order of random draws as augment()
    def _augmenter(self) -> DeviceAugmenter:
        if self._aug is None:
            self._aug = DeviceAugmenter(self.model_settings["desired_samples"], self.background_data)
        return self._aug

    def _bank_row(self, file_path) -> int:
        """Row of the device clip bank holding decode_audio(file_path); decoded and uploaded on first use."""
        key = os.fspath(file_path)
        row = self._bank_rows.get(key)
        if row is None:
            row = self._bank_rows[key] = self._augmenter().clips.add(self.decode_audio(key))
        return row

    def _unknown_rows(self) -> np.ndarray:
        """Bank rows of every unknown file (decoded and uploaded once)."""
        if self._unk_rows is None:
            self._unk_rows = np.asarray([self._bank_row(f) for f in self.unknown_files], np.int64)
        return self._unk_rows

    def _draw_timeshift(self) -> int:
        return int(self.gen.integers(-self.max_time_shift_samples, self.max_time_shift_samples))

    def _draw_background(self):
        background_index = int(self.gen.integers(0, self.background_sizes.shape[0]))
        wav_length = int(self.background_sizes[background_index])
        return background_index, int(self.gen.integers(0, wav_length - self.model_settings["desired_samples"]))

    def augment_plan(self, fg_row: int, label):
        shift = self._draw_timeshift() if self.max_time_shift_samples > 0 else 0
        if self.gen.uniform(0, 1) < (self.silence_percentage / 100):
            volume = self.gen.uniform(0, 1)
            bi, bo = self._draw_background()
            return plan_item(MODE_SILENCE, bg_index=bi, bg_offset=bo, volume=volume), SILENCE_LABEL
        if len(self.unknown_files) > 0 and self.gen.uniform(0, 1) < (self.unknown_percentage / 100):
            row = self._bank_row(self.unknown_files[int(self.gen.integers(0, len(self.unknown_files)))])
            shift = self._draw_timeshift() if self.max_time_shift_samples > 0 else 0
            return plan_item(MODE_CLIP, fg_index=row, shift=shift), UNKNOWN_WORD_LABEL
        if self.gen.uniform(0, 1) < self.background_frequency:
            volume = self.gen.uniform(0, self.background_volume_range)
            bi, bo = self._draw_background()
            return plan_item(MODE_MIX, fg_index=fg_row, shift=shift, bg_index=bi, bg_offset=bo, volume=volume), label
        return plan_item(MODE_CLIP, fg_index=fg_row, shift=shift), label

    # ---- data access (reference :375-434)
    def get_background_data(self, background_dir):
        data, sizes = [], []
        for wav_path in sorted(glob.glob(os.path.join(os.fspath(background_dir), "*.wav"))):
            a = _read_wav(wav_path)
            sizes.append(a.shape[0]); data.append(a)
        if not sizes:
            raise ValueError(f"no background *.wav files in {background_dir}")
        bgdata = np.zeros((len(sizes), max(sizes)), dtype=np.float32)
        for i, wav in enumerate(data):
            bgdata[i, :wav.shape[0]] = wav
        self.background_data = bgdata
        self.background_sizes = np.asarray(sizes)

    def decode_audio(self, file_path):
        return _read_wav(file_path, self.model_settings["desired_samples"])

    def get_label(self, file_path):
        return os.fspath(file_path).split(os.path.sep)[-2]

    def get_waveform_and_label(self, file_path):
        return self.decode_audio(file_path), self.get_label(file_path)

    def get_single_target_waveforms(self, file_path):
        return self.decode_audio(file_path), str(self.commands[-1])

    def label_id(self, label) -> int:
        return int(np.argmax(self.commands == label))        # argmax(label == commands), 0 if absent

    def get_spectrogram_and_label_id(self, audio, label):
        return to_micro_spectrogram(self.model_settings, audio), self.label_id(label)

    def add_channel(self, spectrogram, label_id):
        return np.asarray(spectrogram)[..., None], label_id

    def file2spec_w_bg(self, filepath):
        return to_micro_spectrogram(self.model_settings, self._add_bg(self.decode_audio(filepath)))

    def _add_bg(self, audio):
        background_volume = self.gen.uniform(0, self.background_volume_range)
        return add_background(audio, self.random_background_sample(), background_volume)

    # ---- dataset builders (reference :447-556)
    def _build(self, files, loader, is_training, extra=None) -> _Dataset:
        files = [os.fspath(f) for f in files]

        def make_device():
            label_of = self.get_label if loader == self.get_waveform_and_label else (lambda _f: str(self.commands[-1]))
            for f in files:
                item, label = plan_item(MODE_CLIP, fg_index=self._bank_row(f)), label_of(f)
                if is_training:
                    item, label = self.augment_plan(int(item["fg_index"]), label)
                yield item, self.label_id(label)
            if extra is not None:
                for item, label in extra():
                    yield item, self.label_id(label)

        def make():
            for f in files:
                audio, label = loader(f)
                if is_training:
                    audio, label = self.augment(audio, label)
                yield np.asarray(audio, np.float32), self.label_id(label)
            if extra is not None:
                for audio, label in extra():
                    yield np.asarray(audio, np.float32), self.label_id(label)

        if self.device_augment:
            make = make_device

        n = len(files) + (extra.n if extra is not None else 0)
        ds = _Dataset(self, make, n, is_training)
        if self.batched and is_training and extra is None and files:
            label_of = self.get_label if loader == self.get_waveform_and_label else (lambda _f: str(self.commands[-1]))
            ds.fast = dict(rows=np.asarray([self._bank_row(f) for f in files], np.int64),
                           labels=np.asarray([self.label_id(label_of(f)) for f in files], np.int64))
        return ds

    def init_single_target(self, AUTOTUNE, files, is_training):
        """assumes a single-target model, reads label from self.commands"""
        return self._build(files, self.get_single_target_waveforms, is_training)

    def init_from_parent_dir(self, AUTOTUNE, files, is_training):
        """uses the parent dir as the label name"""
        return self._build(files, self.get_waveform_and_label, is_training)

    def _random_silence(self):
        if self.device_augment:
            volume = self.gen.uniform(0, 1)
            bi, bo = self._draw_background()
            return plan_item(MODE_SILENCE, bg_index=bi, bg_offset=bo, volume=volume), SILENCE_LABEL
        return self.random_background_sample(self.gen.uniform(0, 1)), SILENCE_LABEL

    def _random_unknown(self):
        if self.device_augment:
            f = self.unknown_files[int(self.gen.integers(0, len(self.unknown_files)))]
            return plan_item(MODE_CLIP, fg_index=self._bank_row(f)), UNKNOWN_WORD_LABEL
        return self.get_unknown(), UNKNOWN_WORD_LABEL

    def eval_with_silence_unknown(self, AUTOTUNE, files, label_from_parent_dir: bool):
        n_silent = int(len(files) * self.silence_percentage / 100)
        n_unknown = int(len(files) * self.unknown_percentage / 100)
        if not label_from_parent_dir:
            assert self.commands.shape[0] == 3, "model does not support both silence and unknown"

        def extra():
            for _ in range(n_silent):
                yield self._random_silence()
            for _ in range(n_unknown):
                yield self._random_unknown()

        extra.n = n_silent + n_unknown
        loader = self.get_waveform_and_label if label_from_parent_dir else self.get_single_target_waveforms
        return self._build(files, loader, False, extra)
