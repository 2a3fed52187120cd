"""CPU tests of the host-side mirror of the reference API (no GPU compute): settings dict, WAV codec,
add_background, label ordering, spec-augment masks, the streaming recogniser."""
import collections
import os

import numpy as np
import pytest

from multilingual_kws_b200.embedding import input_data
from multilingual_kws_b200.embedding.single_target_recognize_commands import (RecognizeResult,
                                                                             SingleTargetRecognizeCommands, detect_stream)
from multilingual_kws_b200.frontend import float_audio_to_int16_np


def test_model_settings_keys_and_values():
    s = input_data.standard_microspeech_model_settings(3)
    # reference input_data.py:115-126, values from SURVEY.md §3.1
    assert s == {"desired_samples": 16000, "window_size_samples": 480, "window_stride_samples": 320,
                 "spectrogram_length": 49, "fingerprint_width": 40, "fingerprint_size": 1960, "label_count": 3,
                 "sample_rate": 16000, "preprocess": "micro", "average_window_width": -1}
    a = input_data.prepare_model_settings(12, 16000, 1000, 30, 10, 40, "average")
    assert a["fingerprint_width"] == 43 and a["average_window_width"] == 6 and a["spectrogram_length"] == 98
    assert input_data.prepare_model_settings(2, 16000, 10, 30, 10, 40, "mfcc")["spectrogram_length"] == 0
    with pytest.raises(ValueError, match='Unknown preprocess mode "foo"'):
        input_data.prepare_model_settings(3, 16000, 1000, 30, 20, 40, "foo")


def test_wav_roundtrip_pad_truncate(tmp_path):
    rng = np.random.default_rng(0)
    x = (rng.integers(-32768, 32768, 12345) / 32768.0).astype(np.float32)
    p = str(tmp_path / "a.wav")
    input_data.encode_wav(p, x)
    a, sr = input_data.decode_wav(open(p, "rb").read(), desired_channels=1)
    assert sr == 16000 and a.shape == (12345, 1) and np.array_equal(a[:, 0], x)
    assert np.array_equal(float_audio_to_int16_np(a[:, 0]), np.rint(x * 32768).astype(np.int16))   # exact PCM round trip
    b, _ = input_data.decode_wav(open(p, "rb").read(), desired_channels=1, desired_samples=16000)
    assert b.shape == (16000, 1) and np.array_equal(b[:12345, 0], x) and not b[12345:].any()
    c, _ = input_data.decode_wav(open(p, "rb").read(), desired_channels=1, desired_samples=100)
    assert c.shape == (100, 1)
    with pytest.raises(ValueError):
        input_data.decode_wav(b"not a wav file at all")


def test_float_to_int16_cast_quirk():
    # tf.cast(audio * 32768, int16): truncation toward zero; +1.0 wraps
    x = np.array([0.0, 0.5, -0.5, 0.99999, -1.0, 1.0, 1e-6, -1e-6], np.float32)
    assert float_audio_to_int16_np(x).tolist() == [0, 16384, -16384, 32767, -32768, -32768, 0, 0]


def test_add_background():
    rng = np.random.default_rng(1)
    fg = rng.normal(0, 0.1, 16000).astype(np.float32)
    bg = rng.normal(0, 0.3, 16000).astype(np.float32)
    out = input_data.add_background(fg, bg, 0.1)
    snr = np.sqrt(np.mean(fg ** 2)) / np.sqrt(np.mean(bg ** 2))
    assert np.allclose(out, np.clip(fg + bg * snr * 0.1, -1, 1), atol=1e-6)
    assert np.array_equal(input_data.add_background(fg, np.zeros_like(bg), 0.5), fg)   # bg_rms == 0 -> snr 0
    assert np.abs(input_data.add_background(fg * 20, bg, 1.0)).max() <= 1.0


def _make_dataset(tmp_path, **kw):
    bg = tmp_path / "_background_noise_"
    bg.mkdir()
    rng = np.random.default_rng(2)
    input_data.encode_wav(str(bg / "noise.wav"), rng.normal(0, 0.1, 40000))
    unk = []
    for i in range(3):
        p = tmp_path / f"unk{i}.wav"
        input_data.encode_wav(str(p), rng.normal(0, 0.05, 16000))
        unk.append(str(p))
    return input_data.AudioDataset(input_data.standard_microspeech_model_settings(3), ["tiempo"], str(bg), unk, seed=7, **kw)


def test_label_ordering_and_augment(tmp_path):
    ds = _make_dataset(tmp_path, unknown_percentage=50.0)
    assert ds.commands.tolist() == ["_silence_", "_unknown_", "tiempo"]            # reference :196-206
    assert ds.label_id("tiempo") == 2 and ds.label_id("_unknown_") == 1 and ds.label_id("nope") == 0
    assert ds.max_time_shift_samples == 1600
    audio = np.linspace(-0.5, 0.5, 16000).astype(np.float32)
    labels = collections.Counter()
    for _ in range(300):
        a, l = ds.augment(audio, "tiempo")
        assert a.shape == (16000,) and a.dtype == np.float32 and np.abs(a).max() <= 1.0
        labels[l] += 1
    # 10 % silence, then 50 % of the rest unknown (reference :283-297)
    assert 10 <= labels["_silence_"] <= 60 and 90 <= labels["_unknown_"] <= 190 and labels["tiempo"] >= 80
    no_sil = input_data.AudioDataset(ds.model_settings, ["w"], str(tmp_path / "_background_noise_"), [], silence_percentage=0)
    assert no_sil.commands.tolist() == ["w"]


def test_spec_aug_mask_shapes(tmp_path):
    ds = _make_dataset(tmp_path)
    for _ in range(200):
        m = ds._spec_aug_mask(49, 40)
        assert m.shape == (49, 40) and set(np.unique(m)) <= {0.0, 1.0}
        cols, rows = (m.min(axis=0) == 0).sum(), (m.min(axis=1) == 0).sum()
        assert cols <= 4 or rows == 49 and rows <= 4 or cols == 40           # <= 2 masks x <= 2 px each way
    s = np.ones((49, 40), np.float32)
    assert ds.spec_augment(s).shape == (49, 40)


def naive_recognizer(inferences, times, labels, avg_ms, thr, sup_ms, min_count, target_id):
    """Straight restatement of reference single_target_recognize_commands.py:94-207 for cross-checking."""
    prev, prev_label, prev_time, found = [], "_silence_", -np.inf, []
    for row, t in zip(inferences, times):
        prev.append((t, row))
        while t - avg_ms > prev[0][0]:
            prev.pop(0)
        n = len(prev)
        if n < min_count or t - prev[0][0] < avg_ms / 4:
            continue
        score = sum(float(r[target_id]) / n for _, r in prev)
        label = labels[target_id] if score > thr else "_silence_"
        since = np.inf if (prev_label == "_silence_" or prev_time == -np.inf) else t - prev_time
        new = False
        if score > thr and label != prev_label and since > sup_ms:
            prev_label, prev_time, new = label, t, True
        elif score < thr and label == "_silence_" and since > sup_ms:
            prev_label, prev_time, new = label, t, True
        if new and label != "_silence_":
            found.append((label, int(t), score))
    return found


def test_recognizer_matches_restatement():
    rng = np.random.default_rng(3)
    W = 3000
    p = rng.dirichlet([1, 1, 0.3], W).astype(np.float32)
    for start in range(100, W, 400):                        # bursts of the target
        p[start:start + 40] = [0.02, 0.03, 0.95]
    times = [int(o * 1000 / 16000) for o in range(0, W * 320, 320)]
    labels = ["_silence_", "_unknown_", "kw"]
    for thr in (0.5, 0.9):
        got = detect_stream(p, times, labels, 100, thr, 500, 4, 2)
        want = naive_recognizer(p, times, labels, 100, thr, 500, 4, 2)
        assert [(w, t) for w, t, _ in got] == [(w, t) for w, t, _ in want] and len(got) >= 5
        assert np.allclose([s for *_, s in got], [s for *_, s in want])


def test_recognizer_errors():
    rc = SingleTargetRecognizeCommands(["a", "b", "c"], 100, 0.5, 500, 4, 2)
    el = RecognizeResult()
    with pytest.raises(ValueError, match="should contain 3 elements"):
        rc.process_latest_result(np.zeros(2), 0, el)
    rc.process_latest_result(np.zeros(3), 100, el)
    assert el.found_command == "_silence_" and el.score == 0.0 and not el.is_new_command
    with pytest.raises(ValueError, match="increasing time order"):
        rc.process_latest_result(np.zeros(3), 50, el)


def test_missing_library_message(monkeypatch):
    from multilingual_kws_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libkws_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()


def test_bench_cpu_leg_helpers():
    """bench.py sizes its CPU legs from a probe so `--impl reference` ends within its budget on any host."""
    import time
    import bench
    assert 1 <= bench.host_threads() <= (os.cpu_count() or 1)

    def fake_step(n):
        time.sleep(0.0005 * n)                       # 0.5 ms per clip

    n = bench.bounded_cpu_sample(fake_step, max_clips=1024, budget_s=2.0, n_steps=4, probe=8)
    assert 8 <= n <= 1024 and 0.25 * 1000 <= n <= 4 * 1000 or n == 1024      # ~1000 clips fit 0.5 s per step
    assert bench.bounded_cpu_sample(fake_step, max_clips=16, budget_s=100.0, n_steps=1, probe=8) == 16
    tr, name = bench.load_traffic()
    assert tr is None or ("kernels" in tr and name.startswith("profiles"))


def test_augment_plan_draws_match_host_augment(monkeypatch, tmp_path):
    """Device-path plan items consume the random stream exactly like the host augment(): same seed -> same generator
    state after every element, same labels (no GPU needed: the clip bank is stubbed)."""
    from multilingual_kws_b200.embedding import input_data
    rng = np.random.default_rng(0)
    bgd = tmp_path / "_background_noise_"
    bgd.mkdir()
    input_data.encode_wav(str(bgd / "a.wav"), rng.normal(0, 0.05, 40000))
    files = []
    for i in range(6):
        p = tmp_path / f"c{i}.wav"
        input_data.encode_wav(str(p), rng.normal(0, 0.1, 16000))
        files.append(str(p))
    s = input_data.standard_microspeech_model_settings(3)
    a = input_data.AudioDataset(s, ["hola"], str(bgd), files[3:], unknown_percentage=40.0, silence_percentage=20.0, seed=5)
    b = input_data.AudioDataset(s, ["hola"], str(bgd), files[3:], unknown_percentage=40.0, silence_percentage=20.0, seed=5,
                                device_augment=True)
    monkeypatch.setattr(b, "_bank_row", lambda f: files.index(os.fspath(f)))
    modes = set()
    for k in range(200):
        f = files[k % 3]
        _, la = a.augment(a.decode_audio(f), "hola")
        item, lb = b.augment_plan(files.index(f), "hola")
        assert la == lb
        assert a.gen.bit_generator.state == b.gen.bit_generator.state
        modes.add(int(item["mode"]))
        if lb == input_data.UNKNOWN_WORD_LABEL:
            assert int(item["fg_index"]) >= 3
    assert modes == {0, 1, 2}


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the oracle port on the host cores) prints one well-formed JSON line without a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-sample", "48", "--ref-budget-s", "6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "utterances/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_chunk_row_layout():
    """Rows of the inferences matrix the reference returns for long audio (its inverted chunk branch,
    batch_streaming_analysis.py:72-86): restated here line by line and compared with reference_chunk_rows."""
    from multilingual_kws_b200.embedding.batch_streaming_analysis import reference_chunk_rows

    def reference(n, sr, clip, stride, max_sec):
        audio = np.arange(n)
        max_chunk = max_sec * sr
        chunks = []
        if n < max_chunk:
            chunks.append(audio)
        else:
            for offset in range(0, n, max_chunk):
                if offset + max_chunk > n:
                    chunks.append(audio[offset:offset + max_chunk])
                else:
                    chunks.append(audio[offset:])
        return [(int(c[0]), int(np.ceil((c.shape[0] - clip) / stride))) for c in chunks]

    for n in (16000 * 5, 16000 * 1200 - 1, 16000 * 1200, 16000 * 1800, 16000 * 1800 + 123, 16000 * 3000):
        assert reference_chunk_rows(n, 16000, 16000, 320, 1200) == reference(n, 16000, 16000, 320, 1200), n
    # 30 minutes at the default 20 ms hop (BASELINE config 5): 89 950 windows + the 29 950 of the last 10 minutes again
    assert reference_chunk_rows(16000 * 1800, 16000, 16000, 320, 1200) == [(0, 89950), (19200000, 29950)]
