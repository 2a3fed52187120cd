"""ctypes wrapper over oracle/microfrontend_ref.c — TEST INFRASTRUCTURE ONLY.

The CPU restatement of TF 2.7's ``audio_microfrontend`` op as the reference calls it in
``multilingual_kws/embedding/input_data.py:19-35``.  Pinned on the upstream op's own known-answer tests (see the C
file's header and tests/test_oracle_tf_kat.py).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkws_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "microfrontend_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkws_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.kws_ref_frontend_create.restype = ctypes.c_void_p
        L.kws_ref_frontend_create.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_float,
            ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.kws_ref_frontend_destroy.argtypes = [ctypes.c_void_p]
        L.kws_ref_frontend_num_frames.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.kws_ref_frontend_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p]
        L.kws_ref_frontend_frame_magnitudes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.kws_ref_frontend_frame_finish.argtypes = [ctypes.c_void_p] * 4
        L.kws_ref_frontend_tables.argtypes = [ctypes.c_void_p] * 10
        L.kws_ref_fftr.argtypes = [ctypes.c_void_p] * 3
        L.kws_ref_sqrt64.restype = ctypes.c_uint32
        L.kws_ref_sqrt64.argtypes = [ctypes.c_uint64]
        _lib = L
    return _lib


# Defaults of the TF python wrapper with the reference's overrides.
DEFAULTS = dict(sample_rate=16000, window_ms=30, step_ms=20, num_channels=40, lower_hz=125.0,
                upper_hz=7500.0, smoothing_bits=10, even_smoothing=0.025, odd_smoothing=0.06,
                min_signal_remaining=0.05, enable_pcan=1, pcan_strength=0.95, pcan_offset=80.0,
                gain_bits=21, enable_log=1, scale_shift=6)


class FrontendOracle:
    def __init__(self, **overrides):
        cfg = dict(DEFAULTS)
        cfg.update(overrides)
        self.cfg = cfg
        L = lib()
        self._h = L.kws_ref_frontend_create(
            cfg["sample_rate"], cfg["window_ms"], cfg["step_ms"], cfg["num_channels"], cfg["lower_hz"],
            cfg["upper_hz"], cfg["smoothing_bits"], cfg["even_smoothing"], cfg["odd_smoothing"],
            cfg["min_signal_remaining"], cfg["enable_pcan"], cfg["pcan_strength"], cfg["pcan_offset"],
            cfg["gain_bits"], cfg["enable_log"], cfg["scale_shift"])
        if not self._h:
            raise ValueError("oracle: bad frontend configuration")
        self.num_channels = cfg["num_channels"]
        self.window_size = cfg["window_ms"] * cfg["sample_rate"] // 1000
        self.window_step = cfg["step_ms"] * cfg["sample_rate"] // 1000

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.kws_ref_frontend_destroy(h)
            self._h = None

    def num_frames(self, n_samples: int) -> int:
        return lib().kws_ref_frontend_num_frames(self._h, n_samples)

    def features_u16(self, pcm: np.ndarray, threads: int = 1) -> np.ndarray:
        """pcm int16 [B, n] → uint16 [B, frames, C] (raw op output, before ×10/256)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if pcm.ndim == 1:
            pcm = pcm[None]
        B, n = pcm.shape
        frames = self.num_frames(n)
        out = np.zeros((B, frames, self.num_channels), dtype=np.uint16)
        if B == 0 or frames == 0:
            return out
        L = lib()

        def run(lo, hi):
            L.kws_ref_frontend_batch(self._h, pcm[lo:hi].ctypes.data, hi - lo, n, None, out[lo:hi].ctypes.data)

        if threads <= 1 or B < 2 * threads:
            run(0, B)
        else:
            edges = np.linspace(0, B, threads + 1).astype(int)
            with ThreadPoolExecutor(threads) as ex:  # ctypes drops the GIL during the C call
                list(ex.map(lambda p: run(*p), zip(edges[:-1], edges[1:])))
        return out

    def features(self, pcm: np.ndarray, threads: int = 1) -> np.ndarray:
        """float32 features exactly as input_data.py:34 returns them: uint16 * (10/256)."""
        return self.features_u16(pcm, threads).astype(np.float32) * np.float32(10.0 / 256.0)

    def frame_magnitudes(self, frame: np.ndarray) -> np.ndarray:
        frame = np.ascontiguousarray(frame, dtype=np.int16)
        assert frame.shape == (self.window_size,)
        out = np.zeros(self.num_channels, dtype=np.uint32)
        lib().kws_ref_frontend_frame_magnitudes(self._h, frame.ctypes.data, out.ctypes.data)
        return out

    def tables(self) -> dict:
        fft_size = 1
        while fft_size < self.window_size:
            fft_size <<= 1
        nc = fft_size // 2
        t = dict(window=np.zeros(self.window_size, np.int16), twiddles=np.zeros((nc, 2), np.int16),
                 super_twiddles=np.zeros((nc // 2, 2), np.int16), bin_band=np.zeros(nc + 1, np.int16),
                 bin_weight=np.zeros(nc + 1, np.int16), bin_unweight=np.zeros(nc + 1, np.int16),
                 gain_lut=np.zeros(125, np.int16), log_lut=np.zeros(130, np.uint16),
                 scalars=np.zeros(16, np.int32))
        lib().kws_ref_frontend_tables(self._h, *[t[k].ctypes.data for k in (
            "window", "twiddles", "super_twiddles", "bin_band", "bin_weight", "bin_unweight", "gain_lut",
            "log_lut", "scalars")])
        s = t["scalars"]
        t.update(window_size=int(s[0]), window_step=int(s[1]), fft_size=int(s[2]), start_index=int(s[3]),
                 end_index=int(s[4]), even_smoothing=int(s[5]), odd_smoothing=int(s[6]),
                 min_signal_remaining=int(s[7]), snr_shift=int(s[8]), correction_bits=int(s[9]))
        return t

    def fftr(self, timedata: np.ndarray) -> np.ndarray:
        timedata = np.ascontiguousarray(timedata, dtype=np.int16)
        n = timedata.shape[0]
        out = np.zeros((n // 2 + 1, 2), np.int16)
        lib().kws_ref_fftr(self._h, timedata.ctypes.data, out.ctypes.data)
        return out


def sqrt64(x: int) -> int:
    return int(lib().kws_ref_sqrt64(ctypes.c_uint64(x)))
