"""Pins of the CPU oracle against the UPSTREAM op's own known-answer tests (TF 2.7 microfrontend unit tests).

`tests/golden/tf_microfrontend_kat.json` holds the vectors and their provenance.  The whole chain — Hann window,
block-floating shift, kissfft FIXED_POINT=16 real FFT, mel filterbank, 64-bit sqrt, noise reduction, PCAN gain
control, log scale — has to be right for the 4x2 end-to-end matrix to come out; the module-level vectors localise
a failure.  The configuration differs from the reference's (1 kHz / 25 ms / 2 channels → 32-point FFT instead of
16 kHz / 30 ms / 40 channels → 512) but every line of code is shared, only table sizes change.
"""
import json
import os

import numpy as np
import pytest

from oracle.frontend_oracle import FrontendOracle, lib, sqrt64

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tf_microfrontend_kat.json")))


def audio():
    return np.array(KAT["audio_pattern"] * KAT["audio_repeats"], dtype=np.int16)


def oracle(**kw):
    return FrontendOracle(**KAT["config"], **kw)


def test_op_testSimple_end_to_end():
    got = oracle().features_u16(audio())[0]
    assert got.tolist() == KAT["op_testSimple_uint16"]


def test_op_testSimpleFloatScaled():
    got = oracle().features_u16(audio())[0].astype(np.float32) / np.float32(64.0)
    assert got.tolist() == KAT["op_testSimpleFloatScaled_out_scale_64"]


def test_window_test_coefficients():
    t = oracle().tables()
    assert t["window"].tolist() == KAT["window_test_coefficients"]
    assert t["fft_size"] == KAT["window_test_fft_size"]
    assert t["correction_bits"] == -1                  # MSB(32) - 1 - kFilterbankBits/2


def test_filterbank_sqrt():
    work = KAT["filterbank_test_sqrt_work"]
    assert [sqrt64(w) for w in work[1:]] == KAT["filterbank_test_sqrt_expected"]
    # the first frame of the op-test audio produces exactly these magnitudes
    assert oracle().frame_magnitudes(audio()[:25]).tolist() == KAT["filterbank_test_sqrt_expected"]


@pytest.mark.parametrize("pcan,log,key", [(0, 0, "noise_reduction_test_expected_signal"),
                                          (1, 0, "pcan_test_expected"), (1, 1, "log_scale_test_expected")])
def test_noise_reduction_pcan_log_chain(pcan, log, key):
    o = oracle(enable_pcan=pcan, enable_log=log)
    sig = np.array(KAT["noise_reduction_test_signal"], np.uint32)
    est = np.zeros(2, np.uint32)
    out = np.zeros(2, np.uint16)
    lib().kws_ref_frontend_frame_finish(o._h, sig.ctypes.data, est.ctypes.data, out.ctypes.data)
    assert est.tolist() == KAT["noise_reduction_test_expected_estimate"]
    assert sig.tolist() == (KAT[key] if not log else KAT["pcan_test_expected"])
    if pcan:
        assert out.tolist() == KAT[key]
