"""Floating-point statement of what the micro-frontend computes — TEST INFRASTRUCTURE, never imported by the product.

The integer oracle (`microfrontend_ref.c`) restates TF 2.7's `audio_microfrontend` operation by operation.  This file
states the same signal chain from its MATHEMATICAL definition instead — no fixed point, no lookup tables, no shared code
or constants with the oracle beyond the op's attributes — so that the oracle can be pinned at the reference's production
configuration (16 kHz, 30 ms / 20 ms, 40 channels, 512-point FFT: `input_data.py:25-33`), where no upstream
known-answer vector exists.  In "true" units (the integer pipeline carries magnitudes / 8 and compensates later):

    window      w[i] = 0.5 - 0.5 cos(2 pi (i + 0.5) / N)                                  (window_util.c)
    spectrum    X = rfft(frame * w, 512)
    filterbank  S[c] = sqrt(sum_k tri_c(mel(k * sr / 512)) |X[k]|^2), triangles peaked at      (filterbank_util.c)
                mel_lo + (c + 1) * (mel_hi - mel_lo) / (C + 1), mel(f) = 1127 ln(1 + f / 700)
    noise       E <- a S + (1 - a) E (a = 0.025 even / 0.06 odd channels, E starts at 0);        (noise_reduction.c)
                D = max(S - E, 0.05 S)
    PCAN        x = D (E + 80)^-0.95;  y = x^2 / 4 if x < 2 else x - 1                          (pcan_gain_control*.c)
    log         out = 64 ln(512 y) if 512 y > 1 else 0                                           (log_scale.c)

The integer pipeline deviates from this where ITS OWN resolution ends, and the comparison (tests/test_oracle_float_model.py)
is restricted accordingly: (1) the 16-bit block-floating FFT has a noise floor ~30 dB under the frame's peak; (2) the
value entering the log is a multiple of 8, so outputs below 64 ln 256 are coarsely quantised; (3) the PCAN gain is an
int16 with 21 fractional bits, i.e. a SMALL integer for loud stationary input (E > ~56 000: gain < 64).
"""
import numpy as np


def float_frontend(pcm, sample_rate=16000, window=480, step=320, channels=40, lower_hz=125.0, upper_hz=7500.0,
                   even_smoothing=0.025, odd_smoothing=0.06, min_signal_remaining=0.05, pcan_strength=0.95,
                   pcan_offset=80.0, bin_shift=0):
    """pcm int16 [B, n] -> (features [B, frames, C] in the op's uint16 units, magnitudes S, noise estimates E).
    `bin_shift` moves every FFT bin's frequency by that many bins (sensitivity checks only)."""
    pcm = np.asarray(pcm, np.float64)
    B, n = pcm.shape
    frames = (n - window) // step + 1
    nfft = 1
    while nfft < window:
        nfft <<= 1
    i = np.arange(window)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * (i + 0.5) / window)
    idx = np.arange(frames)[:, None] * step + i[None, :]
    power = np.abs(np.fft.rfft(pcm[:, idx] * w, nfft, axis=-1)) ** 2            # [B, frames, nfft/2 + 1]

    def mel(f):
        return 1127.0 * np.log1p(f / 700.0)

    m_lo, m_hi = mel(lower_hz), mel(upper_hz)
    spacing = (m_hi - m_lo) / (channels + 1)
    edges = m_lo + spacing * np.arange(channels + 2)       # edges[c], edges[c + 1] (peak), edges[c + 2] of channel c
    fm = mel(np.maximum(np.arange(nfft // 2 + 1) + bin_shift, 0) * (sample_rate / nfft))
    tri = np.zeros((channels, nfft // 2 + 1))
    for c in range(channels):
        up = (fm - edges[c]) / (edges[c + 1] - edges[c])
        down = (edges[c + 2] - fm) / (edges[c + 2] - edges[c + 1])
        tri[c] = np.clip(np.minimum(up, down), 0.0, None)
    S = np.sqrt(power @ tri.T)
    a = np.where(np.arange(channels) % 2 == 0, even_smoothing, odd_smoothing)
    E = np.zeros((B, channels))
    out, Eh = np.zeros_like(S), np.zeros_like(S)
    for t in range(frames):
        E = a * S[:, t] + (1.0 - a) * E
        D = np.maximum(S[:, t] - E, min_signal_remaining * S[:, t])
        x = D * (E + pcan_offset) ** -pcan_strength
        y = np.where(x < 2.0, x * x / 4.0, x - 1.0)
        v = 512.0 * y
        out[:, t] = np.where(v > 1.0, 64.0 * np.log(np.maximum(v, 1e-300)), 0.0)
        Eh[:, t] = E
    return out, S, Eh


def lut_gain(lut, estimate_true_units):
    """The oracle's int16 PCAN gain for a noise estimate given in true units: pcan_gain_control.c WideDynamicFunction on
    the oracle's own 125-entry table (piecewise quadratic per octave), input = estimate / 8 * 2^smoothing_bits.  Only used
    to show that the loud-regime deviation from the float model IS this integer (tests/test_oracle_float_model.py)."""
    x = int(round(estimate_true_units * 128.0))
    if x <= 2:
        return int(lut[x])
    interval = x.bit_length()
    p = 4 * interval - 6
    frac = ((x << (11 - interval)) if interval < 11 else (x >> (interval - 11))) & 0x3FF
    r = (int(lut[p + 2]) * frac) >> 5
    r += int(lut[p + 1]) << 5
    r *= frac
    r = (r + (1 << 14)) >> 15
    r += int(lut[p])
    return ((r + 32768) % 65536) - 32768
