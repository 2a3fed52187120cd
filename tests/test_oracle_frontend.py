"""Pins for the CPU oracle of the micro-frontend (oracle/microfrontend_ref.c).

The reference ships no tests or golden vectors for this path and TF cannot run
here, so the oracle is pinned by closed forms, analytic invariants and numpy consistency checks.
"""
import numpy as np
import pytest

from oracle.frontend_oracle import FrontendOracle, sqrt64
from multilingual_kws_b200.synthetic import synthetic_pcm


@pytest.fixture(scope="module")
def orc():
    return FrontendOracle()


def test_scalars_and_band_layout(orc):
    t = orc.tables()
    assert (t["window_size"], t["window_step"], t["fft_size"]) == (480, 320, 512)
    assert (t["start_index"], t["end_index"]) == (5, 241)
    assert (t["even_smoothing"], t["odd_smoothing"], t["min_signal_remaining"]) == (409, 983, 819)
    assert (t["snr_shift"], t["correction_bits"]) == (6, 3)
    bb = t["bin_band"]
    widths = [int((bb[5:241] == i).sum()) for i in range(41)]
    # SURVEY.md App. A.4 (numpy-float32 emulation done independently during the survey)
    assert widths == [1, 2, 1, 2, 2, 2, 2, 2, 3, 2, 3, 3, 3, 3, 3, 4, 4, 3, 5, 4, 5, 5, 5, 5, 6, 6, 7, 7, 7, 8, 8, 9, 9,
                      9, 11, 10, 12, 12, 13, 13, 15]
    w, u = t["bin_weight"][5:241].astype(int), t["bin_unweight"][5:241].astype(int)
    assert ((w + u) >= 4095).all() and ((w + u) <= 4097).all() and (w >= 0).all() and (w <= 4096).all()


def test_window_closed_form(orc):
    t = orc.tables()
    i = np.arange(480)
    ref = np.floor(4096 * (0.5 - 0.5 * np.cos(2 * np.pi * (i + 0.5) / 480)) + 0.5).astype(int)
    # float32 `arg` in the C source perturbs at most the last unit of a handful of coefficients
    assert np.abs(t["window"].astype(int) - ref).max() <= 1
    assert t["window"].max() == 4096 and t["window"][0] == 0
    assert (t["window"] == t["window"][::-1]).mean() > 0.95


def test_log_lut_known_values(orc):
    lut = orc.tables()["log_lut"].astype(int)
    # leading / peak / trailing entries of TF's kLogLut as recalled in SURVEY.md App. A.7
    assert list(lut[:12]) == [0, 224, 442, 654, 861, 1063, 1259, 1450, 1636, 1817, 1992, 2163]
    assert lut.max() == 5641 and int(lut.argmax()) == 57
    assert list(lut[126:130]) == [282, 142, 0, 0]


def test_twiddles(orc):
    t = orc.tables()
    i = np.arange(256)
    assert np.array_equal(t["twiddles"][:, 0], np.floor(.5 + 32767 * np.cos(-2 * np.pi * i / 256)).astype(np.int16))
    assert np.array_equal(t["twiddles"][:, 1], np.floor(.5 + 32767 * np.sin(-2 * np.pi * i / 256)).astype(np.int16))
    assert tuple(t["twiddles"][0]) == (32767, 0) and tuple(t["twiddles"][64]) == (0, -32767)


def test_pcan_lut_monotone(orc):
    lut = orc.tables()["gain_lut"].astype(int)
    g0 = int(2 ** 21 * (80.0 ** -0.95) + 0.5)
    assert abs(lut[0] - g0) <= 1
    y0 = lut[2::4][:31]          # y0 of every interval: gain falls as the noise estimate grows
    assert (np.diff(y0) <= 0).all() and y0[-1] >= 0


def test_sqrt_matches_math():
    rng = np.random.default_rng(0)
    xs = [0, 1, 2, 3, 4, 15, 16, 17, 65535 ** 2, 65535 ** 2 + 65535, 65535 ** 2 + 65536, 2 ** 32 - 1, 2 ** 32,
          2 ** 32 + 1, 2 ** 48 - 1, 2 ** 64 - 1] + [int(v) for v in rng.integers(0, 2 ** 50, 2000, dtype=np.uint64)]
    import math
    for x in xs:
        r = math.isqrt(x)
        cap = 0xFFFFFFFF if x >> 32 else 0xFFFF
        if x - r * r > r and r != cap:
            r += 1
        assert sqrt64(x) == r, x


def test_fft_consistent_with_numpy(orc):
    rng = np.random.default_rng(1)
    worst = 0.0
    for trial in range(20):
        x = np.zeros(512)
        x[:480] = 6000 * np.sin(2 * np.pi * rng.uniform(2, 250) * np.arange(480) / 512 + rng.uniform(0, 6)) \
            + rng.normal(0, 800, 480)
        x = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
        F = orc.fftr(x)
        got = F[:, 0].astype(float) + 1j * F[:, 1]
        ref = np.fft.rfft(x.astype(float)) / 512
        worst = max(worst, np.abs(got - ref).max())
        assert np.argmax(np.abs(got[1:-1])) == np.argmax(np.abs(ref[1:-1]))
    assert worst < 6.0     # a scaled FFT with bounded rounding error


def test_invariants(orc):
    z = orc.features_u16(np.zeros((2, 16000), np.int16))
    assert z.shape == (2, 49, 40) and z.max() == 0
    f = orc.features(synthetic_pcm(8))
    assert f.shape == (8, 49, 40) and f.dtype == np.float32
    q = f / np.float32(0.0390625)
    assert np.array_equal(q, np.rint(q))            # integer multiples of 10/256
    assert f.max() < 30.0
    # clips are independent and deterministic
    assert np.array_equal(orc.features_u16(synthetic_pcm(8)[3:4])[0], orc.features_u16(synthetic_pcm(8))[3])
    assert np.array_equal(orc.features_u16(synthetic_pcm(8), threads=4), orc.features_u16(synthetic_pcm(8)))


def test_shapes_and_short_inputs(orc):
    assert orc.num_frames(16000) == 49 and orc.num_frames(480) == 1 and orc.num_frames(479) == 0
    assert orc.num_frames(800) == 2 and orc.num_frames(799) == 1
    assert orc.features_u16(np.zeros((3, 100), np.int16)).shape == (3, 0, 40)
    assert orc.features_u16(np.zeros((0, 16000), np.int16)).shape == (0, 49, 40)


def test_sine_lights_expected_channel(orc):
    t = orc.tables()
    n = np.arange(16000)
    for bin_ in (20, 60, 150):
        x = (12000 * np.sin(2 * np.pi * (bin_ * 31.25) * n / 16000)).astype(np.int16)
        mags = orc.frame_magnitudes(x[:480])
        band = int(t["bin_band"][bin_])         # band i feeds channels i-1 (weight) and i (unweight)
        assert int(np.argmax(mags)) in (band - 1, band)


def test_trailing_samples_unused(orc):
    pcm = synthetic_pcm(4)
    a = orc.features_u16(pcm)
    pcm2 = pcm.copy()
    pcm2[:, 15840:] = 0                      # (49-1)*320+480 = 15840: the last 160 samples are never read
    assert np.array_equal(a, orc.features_u16(pcm2))
