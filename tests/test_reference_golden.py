"""Host mirrors vs golden vectors produced by EXECUTING THE REFERENCE (tests/golden/make_reference_golden.py):
model settings (a1), streaming post-processor step by step (a13), the streaming driver's detection lists (a12),
accuracy statistics and tpr/fpr (f1)."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from multilingual_kws_b200.embedding import input_data
from multilingual_kws_b200.embedding.accuracy_utils import StreamingAccuracyStats
from multilingual_kws_b200.embedding.single_target_recognize_commands import (RecognizeResult,
                                                                             SingleTargetRecognizeCommands, detect_stream)
from multilingual_kws_b200.embedding.tpr_fpr import get_groundtruth, tpr_fpr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
EXP = json.load(open(os.path.join(GOLDEN, "reference_postproc.json")))
ARR = np.load(os.path.join(GOLDEN, "reference_postproc.npz"))
KEYWORD = EXP["keyword"]
LABELS = [input_data.SILENCE_LABEL, input_data.UNKNOWN_WORD_LABEL, KEYWORD]
CASES = [(c, thr) for c in EXP["cases"] for thr in c["thresholds"]]
IDS = [f"seed{c['seed']}-thr{t}" for c, t in CASES]


def times_of(case):
    stride = int(case["stride_ms"] * 16000 / 1000)
    return [int(o * 1000 / 16000) for o in range(0, case["n_samples"] - 16000, stride)]


def test_model_settings_match_reference():
    ms = EXP["model_settings"]
    for c in ms["cases"]:
        assert input_data.prepare_model_settings(*c["args"]) == c["out"]
    assert input_data.standard_microspeech_model_settings(3) == ms["standard"]
    with pytest.raises(ValueError) as e:
        input_data.prepare_model_settings(3, 16000, 1000, 30, 20, 40, "bogus")
    assert str(e.value) == ms["bogus_error"]


@pytest.mark.parametrize("case,thr", CASES, ids=IDS)
def test_recognizer_step_by_step(case, thr):
    probs, times = ARR[f"probs_{case['seed']}"], times_of(case)
    assert len(times) == case["windows"] and [times[0], times[-1]] == case["times_first_last"]
    want = case["per_threshold"][repr(thr)]
    want_score = ARR[f"scores_{case['seed']}_{thr!r}"]
    rc = SingleTargetRecognizeCommands(LABELS, case["average_window_duration_ms"], thr, case["suppression_ms"],
                                       case["minimum_count"], target_id=2)
    el = RecognizeResult()
    new_steps, kw_steps = [], []
    for i, (row, t) in enumerate(zip(probs, times)):          # float32 rows, as model.predict returns them
        rc.process_latest_result(row, t, el)
        if el.is_new_command:
            new_steps.append(i)
        if el.found_command == KEYWORD:
            kw_steps.append(i)
        assert el.score == want_score[i], (i, el.score, want_score[i])      # bit-exact float64
    assert new_steps == want["is_new_steps"] and kw_steps == want["keyword_steps"]
    found = detect_stream(probs, times, LABELS, case["average_window_duration_ms"], thr, case["suppression_ms"],
                          case["minimum_count"], target_id=2)
    assert [[w, t, s] for w, t, s in found] == want["found_words_w_confidences"]


@pytest.mark.parametrize("case,thr", CASES, ids=IDS)
def test_accuracy_stats_and_rates(case, thr, tmp_path):
    want = case["per_threshold"][repr(thr)]
    gt_file = tmp_path / "gt.txt"
    gt_file.write_text(case["ground_truth_file"])
    tol = case["time_tolerance_ms"]
    stats = StreamingAccuracyStats(target_keyword=KEYWORD)
    stats.read_ground_truth_file(gt_file)
    found, sofar, states = want["found_words"], [], []
    for w in found:
        sofar.append(w)
        stats.calculate_accuracy_stats(sofar, w[1], tol)
        states.append(stats.delta())
    stats.calculate_accuracy_stats(found, -1, tol)
    assert states == want["stats"]["delta_states"]
    for k, v in want["stats"]["counters"].items():
        assert getattr(stats, k) == v, k
    with contextlib.redirect_stdout(io.StringIO()):
        printed = stats.print_accuracy_stats()
    if want["stats"]["info"] is None:
        assert printed is None
    else:
        assert printed[0] == want["stats"]["info"] and printed[1] == want["stats"]["stat"]
    gt_sorted = [list(x) for x in stats._gt_occurrence]
    kw_times = [t for lab, t in gt_sorted if lab == KEYWORD]
    with contextlib.redirect_stdout(io.StringIO()):
        if want["tpr_fpr"] is not None:
            got = tpr_fpr(KEYWORD, thr, found, kw_times, duration_s=case["n_samples"] / 16000, time_tolerance_ms=tol,
                          num_nontarget_words=max(1, len(gt_sorted) - len(kw_times)))
            assert got == want["tpr_fpr"]
        assert get_groundtruth(want["found_words_w_confidences"], [KEYWORD], gt_sorted, time_tolerance_ms=tol) == \
            want["get_groundtruth"]
        assert get_groundtruth([], [], gt_sorted) is None


def test_streaming_driver_with_existing_inferences(tmp_path):
    """calculate_streaming_accuracy(existing_inferences=...) of the mirror vs the reference's own driver."""
    import struct
    from multilingual_kws_b200.embedding import batch_streaming_analysis as bsa
    for case in EXP["cases"]:
        n = case["n_samples"]
        wav = tmp_path / f"s{case['seed']}.wav"
        hdr = b"RIFF" + struct.pack("<I", 36 + 2 * n) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16) + \
            b"data" + struct.pack("<I", 2 * n)
        wav.write_bytes(hdr + bytes(2 * n))
        gt = tmp_path / f"gt{case['seed']}.txt"
        gt.write_text(case["ground_truth_file"])
        flags = bsa.StreamFlags(wav=wav, ground_truth=gt, target_keyword=KEYWORD, detection_thresholds=case["thresholds"],
                                clip_stride_ms=case["stride_ms"], average_window_duration_ms=case["average_window_duration_ms"],
                                suppression_ms=case["suppression_ms"], time_tolerance_ms=case["time_tolerance_ms"],
                                minimum_count=case["minimum_count"])
        with contextlib.redirect_stdout(io.StringIO()):
            results, _ = bsa.calculate_streaming_accuracy(None, input_data.standard_microspeech_model_settings(3), [flags],
                                                          existing_inferences=ARR[f"probs_{case['seed']}"])
        (_, res), = results
        for thr in case["thresholds"]:
            want = case["per_threshold"][repr(thr)]
            assert res[thr][0] == want["found_words"]
            assert res[thr][1] == want["found_words_w_confidences"]
