"""CPU check of the CUDA frontend's half-warp choreography: the phases of csrc/frontend_core.cuh are
replayed lane by lane on the host (tests/emu) and must equal the oracle bit for bit.  Also checks the
tables libkws_b200.so builds against the oracle's independently built ones."""
import ctypes

import numpy as np
import pytest

from oracle.frontend_oracle import DEFAULTS, FrontendOracle, sqrt64
from multilingual_kws_b200.synthetic import synthetic_pcm

ARGT = [ctypes.c_int] * 4 + [ctypes.c_float] * 2 + [ctypes.c_int] + [ctypes.c_float] * 3 + [ctypes.c_int] + \
    [ctypes.c_float] * 2 + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_void_p]


def emulate(E, pcm, **kw):
    cfg = dict(DEFAULTS)
    cfg.update(kw)
    E.emu_frontend_batch.argtypes = ARGT
    B, n = pcm.shape
    orc = FrontendOracle(**kw)
    fr = orc.num_frames(n)
    out = np.zeros((B, fr, cfg["num_channels"]), np.uint16)
    r = E.emu_frontend_batch(cfg["sample_rate"], cfg["window_ms"], cfg["step_ms"], cfg["num_channels"], cfg["lower_hz"],
                             cfg["upper_hz"], cfg["smoothing_bits"], cfg["even_smoothing"], cfg["odd_smoothing"],
                             cfg["min_signal_remaining"], cfg["enable_pcan"], cfg["pcan_strength"], cfg["pcan_offset"],
                             cfg["gain_bits"], cfg["enable_log"], cfg["scale_shift"], pcm.ctypes.data, B, n,
                             out.ctypes.data, None)
    assert r == fr
    return out, orc.features_u16(pcm)


def test_synthetic_bit_exact(emu_lib):
    a, b = emulate(emu_lib, synthetic_pcm(48))
    assert np.array_equal(a, b)


def test_adversarial_bit_exact(emu_lib):
    rng = np.random.default_rng(7)
    adv = rng.integers(-32768, 32768, (12, 16000)).astype(np.int16)
    adv[0] = -32768
    adv[1] = 32767
    adv[2, ::2] = -32768
    adv[2, 1::2] = 32767
    adv[3] = 1
    adv[4] = -1
    adv[5] = 0
    adv[5, 5000] = -32768
    adv[6] = (rng.integers(0, 2, 16000) * 65535 - 32768).astype(np.int16)
    a, b = emulate(emu_lib, adv)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("kw", [dict(enable_pcan=0), dict(enable_log=0), dict(num_channels=32, window_ms=25, step_ms=10),
                                dict(num_channels=10, lower_hz=20.0, upper_hz=4000.0), dict(scale_shift=4, gain_bits=19)])
def test_other_configurations(emu_lib, kw):
    a, b = emulate(emu_lib, synthetic_pcm(6, cfg_id=3), **kw)
    assert np.array_equal(a, b)


def test_isqrt_matches_reference_bit_loop(emu_lib):
    emu_lib.emu_isqrt64_round.restype = ctypes.c_uint32
    emu_lib.emu_isqrt64_round.argtypes = [ctypes.c_uint64]
    rng = np.random.default_rng(3)
    xs = [0, 1, 2, 3, 65535 ** 2 + 65535, 65535 ** 2 + 65536, 2 ** 32 - 1, 2 ** 32, 2 ** 64 - 1, (2 ** 32 - 1) ** 2,
          (2 ** 32 - 1) ** 2 + 2 ** 32]
    xs += [int(v) for v in rng.integers(0, 2 ** 63, 4000, dtype=np.uint64)]
    xs += [int(v) for v in rng.integers(0, 2 ** 34, 4000, dtype=np.uint64)]
    for x in xs:
        assert emu_lib.emu_isqrt64_round(x) == sqrt64(x), x
