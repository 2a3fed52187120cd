"""Synthetic 1 s / 16 kHz PCM clips for parity tests and bench.py.

A mixture chosen so every branch of the fixed-point frontend fires:
  1/4 Gaussian noise sigma=3000;   1/4 three sines + light noise;
  1/4 silence-then-burst (input_shift = 15 and the log's value<=1 branch);
  1/4 full-scale square / noise (input_shift = 0, the 64-bit sqrt path, uint16 saturation).
"""
from __future__ import annotations

import numpy as np


def synthetic_pcm(batch: int, n_samples: int = 16000, cfg_id: int = 1, sample_rate: int = 16000) -> np.ndarray:
    rng = np.random.default_rng(1234 + cfg_id)
    out = np.zeros((batch, n_samples), dtype=np.int16)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    for b in range(batch):
        kind = b % 4
        if kind == 0:
            x = rng.normal(0.0, 3000.0, n_samples)
        elif kind == 1:
            x = rng.normal(0.0, 200.0, n_samples)
            for _ in range(3):
                f = rng.uniform(100.0, 7000.0)
                a = rng.uniform(500.0, 12000.0)
                x = x + a * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
        elif kind == 2:
            x = np.zeros(n_samples)
            half = n_samples // 2
            x[half:] = rng.normal(0.0, rng.uniform(50.0, 8000.0), n_samples - half)
        else:
            if (b // 4) % 2 == 0:
                period = int(rng.integers(8, 400))
                x = np.where((np.arange(n_samples) // period) % 2 == 0, 32767.0, -32767.0)
            else:
                x = rng.choice(np.array([-32768.0, 32767.0]), n_samples)
        out[b] = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    return out


def synthetic_stream(n_samples: int, cfg_id: int = 5) -> np.ndarray:
    """A long stream made by tiling the clip generator."""
    clips = synthetic_pcm(64, 16000, cfg_id)
    reps = -(-n_samples // clips.size)
    return np.tile(clips.reshape(-1), reps)[:n_samples].copy()
