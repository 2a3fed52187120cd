#!/usr/bin/env python
"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN CODE (run in the build container, where /root/reference
exists: python tests/golden/make_reference_golden.py).  `_reference_stub` turns every tensorflow import into a mock,
so only the reference's pure Python / numpy logic runs — exactly the parts of the hot path that are not TensorFlow:

  a1   input_data.prepare_model_settings / standard_microspeech_model_settings        (input_data.py:63-138)
  a13  SingleTargetRecognizeCommands.process_latest_result, step by step              (single_target_recognize_commands.py:94-207)
  a12  calculate_streaming_accuracy with existing_inferences: window times, thresholds (batch_streaming_analysis.py:50-179)
  f1   StreamingAccuracyStats (accuracy_utils.py:93-204), tpr_fpr / get_groundtruth    (tpr_fpr.py:1-138)

Scores are handed to the reference as float64 holding float32 values: under the numpy the reference pins (1.19, via
TF 2.7) `float32_scalar / int` promotes to float64; numpy 2 keeps float32.  float64 inputs give the pinned behaviour
on both.  Output: reference_postproc.npz (inputs) + reference_postproc.json (expected outputs).
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _reference_stub  # noqa: E402

_reference_stub.install()
import tensorflow as tf  # noqa: E402  (the mock)
import multilingual_kws.embedding.accuracy_utils as AU  # noqa: E402
import multilingual_kws.embedding.batch_streaming_analysis as BSA  # noqa: E402
import multilingual_kws.embedding.input_data as ID  # noqa: E402
import multilingual_kws.embedding.single_target_recognize_commands as RC  # noqa: E402
import multilingual_kws.embedding.tpr_fpr as TPR  # noqa: E402

KEYWORD = "merhaba"
LABELS = [ID.SILENCE_LABEL, ID.UNKNOWN_WORD_LABEL, KEYWORD]

# (seed, windows, stride_ms, avg_window_ms, suppression_ms, minimum_count, thresholds, time_tolerance_ms)
CASES = [
    (0, 1500, 20, 100, 500, 4, [0.3, 0.5, 0.7, 0.9], 750),      # the reference's StreamFlags defaults
    (1, 1200, 20, 1000, 1500, 3, [0.4, 0.6], 750),              # TF's recognize_commands defaults
    (2, 600, 100, 500, 500, 1, [0.5], 1500),                    # coarse hop, minimum_count 1
    (3, 900, 20, 100, 0, 4, [0.2, 0.8], 100),                   # no suppression, tight tolerance
    (4, 400, 20, 50, 500, 4, [0.5], 750),                       # window shorter than minimum_count results: never fires
]


def synthetic_inferences(seed: int, W: int, stride_ms: int):
    """Softmax rows with keyword bursts (ground truth), a few unknown-word bursts and noise; float32."""
    rng = np.random.default_rng(9000 + seed)
    t = np.arange(W) * stride_ms
    logit = rng.normal(0.0, 1.0, (W, 3))
    logit[:, 0] += 1.5
    gt = []
    pos = 700
    while pos < t[-1] - 1500:
        kind = rng.choice(["kw", "kw", "unk", "miss"])
        width = rng.uniform(150, 500)
        bump = np.exp(-0.5 * ((t - pos) / width) ** 2)
        if kind == "kw":
            logit[:, 2] += bump * rng.uniform(3, 9)
            gt.append((KEYWORD, pos + rng.uniform(-200, 200)))
        elif kind == "unk":
            logit[:, 1] += bump * rng.uniform(3, 8)
            gt.append((ID.UNKNOWN_WORD_LABEL, pos + rng.uniform(-200, 200)))
        else:
            gt.append((KEYWORD, float(pos)))                     # spoken but not recognised -> false negative
        if rng.random() < 0.25:                                  # a spurious detection with no ground truth nearby
            logit[:, 2] += np.exp(-0.5 * ((t - (pos + 1800)) / 200.0) ** 2) * 7
        pos += rng.uniform(2500, 6000)
    e = np.exp(logit - logit.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32), gt


def run_recognizer(probs64, times, avg_ms, thr, sup_ms, min_count):
    rc = RC.SingleTargetRecognizeCommands(labels=LABELS, average_window_duration_ms=avg_ms, detection_threshold=thr,
                                          suppression_ms=sup_ms, minimum_count=min_count, target_id=2)
    el = RC.RecognizeResult()
    is_new, is_kw, score = [], [], []
    for row, t in zip(probs64, times):
        rc.process_latest_result(row, t, el)
        is_new.append(bool(el.is_new_command))
        is_kw.append(el.found_command == KEYWORD)
        score.append(float(el.score))
    return is_new, is_kw, score


def run_stats(found_words, gt_file, tol):
    """The bookkeeping loop of calculate_streaming_accuracy (batch_streaming_analysis.py:140-170), keeping the stats."""
    stats = AU.StreamingAccuracyStats(target_keyword=KEYWORD)
    stats.read_ground_truth_file(gt_file)
    states, sofar = [], []
    for w in found_words:
        sofar.append(w)
        stats.calculate_accuracy_stats(sofar, w[1], tol)
        states.append(stats.delta())
    stats.calculate_accuracy_stats(found_words, -1, tol)
    with contextlib.redirect_stdout(io.StringIO()):
        printed = stats.print_accuracy_stats()
    counters = {k: getattr(stats, k) for k in ("_how_many_gt", "_how_many_gt_matched", "_how_many_fp", "_how_many_c", "_how_many_w",
                                               "_how_many_fn", "_how_many_gt_target", "_how_many_gt_unknown_or_silence")}
    counters["_which_matched"] = dict(stats._which_matched)
    counters["_which_wrong"] = dict(stats._which_wrong)
    return dict(delta_states=states, counters=counters, info=None if printed is None else printed[0],
                stat=None if printed is None else printed[1])


def main():
    arrays, expected = {}, {"keyword": KEYWORD, "cases": []}
    # ---- a1: model settings
    settings = []
    for args in [(3, 16000, 1000, 30, 20, 40, "micro"), (12, 16000, 1000, 30.0, 10.0, 40, "mfcc"),
                 (12, 16000, 1000, 30.0, 10.0, 40, "average"), (3, 8000, 1500, 25, 10, 32, "micro"),
                 (3, 16000, 20, 30, 20, 40, "micro")]:
        settings.append(dict(args=list(args), out=ID.prepare_model_settings(*args)))
    try:
        ID.prepare_model_settings(3, 16000, 1000, 30, 20, 40, "bogus")
        err = None
    except ValueError as e:
        err = str(e)
    expected["model_settings"] = dict(cases=settings, bogus_error=err, standard=ID.standard_microspeech_model_settings(3))

    tmp = tempfile.mkdtemp()
    for seed, W, stride_ms, avg_ms, sup_ms, min_count, thresholds, tol in CASES:
        probs, gt = synthetic_inferences(seed, W, stride_ms)
        arrays[f"probs_{seed}"] = probs
        gt_file = os.path.join(tmp, f"gt_{seed}.txt")
        with open(gt_file, "w") as f:
            for lab, tm in gt:
                f.write(f"{lab},{tm:.3f}\n")
            f.write("malformed line without comma\n")
        gt_lines = open(gt_file).read()
        sample_rate = 16000
        stride = int(stride_ms * sample_rate / 1000)
        n_samples = 16000 + stride * W - (stride - 1)            # range(0, n - 16000, stride) has exactly W entries
        times = [int(o * 1000 / sample_rate) for o in range(0, n_samples - 16000, stride)]
        assert len(times) == W
        case = dict(seed=seed, windows=W, stride_ms=stride_ms, average_window_duration_ms=avg_ms, suppression_ms=sup_ms,
                    minimum_count=min_count, thresholds=thresholds, time_tolerance_ms=tol, n_samples=n_samples,
                    ground_truth_file=gt_lines, times_first_last=[times[0], times[-1]], per_threshold={})
        # ---- a12: the reference's own driver, wav decode mocked, inferences supplied
        audio_obj, sr_obj = tf.audio.decode_wav.return_value = (type("A", (), {})(), type("S", (), {})())
        audio_obj.numpy = lambda n=n_samples: np.zeros((n, 1), np.float32)
        sr_obj.numpy = lambda: sample_rate
        flags = BSA.StreamFlags(wav="synthetic.wav", ground_truth=gt_file, target_keyword=KEYWORD,
                                detection_thresholds=thresholds, clip_stride_ms=stride_ms,
                                average_window_duration_ms=avg_ms, suppression_ms=sup_ms, time_tolerance_ms=tol,
                                minimum_count=min_count)
        with contextlib.redirect_stdout(io.StringIO()):
            results, _ = BSA.calculate_streaming_accuracy(None, ID.standard_microspeech_model_settings(3), [flags],
                                                          existing_inferences=probs.astype(np.float64))
        (_, res_thresh), = results
        for thr in thresholds:
            found, found_conf = res_thresh[thr]
            is_new, is_kw, score = run_recognizer(probs.astype(np.float64), times, avg_ms, thr, sup_ms, min_count)
            # a13 step-by-step trace must agree with the driver's detection list
            assert [[KEYWORD, t] for t, n, k in zip(times, is_new, is_kw) if n and k] == found
            gt_sorted = sorted([[lab, round(tm)] for lab, tm in gt], key=lambda it: it[1])
            kw_times = [t for lab, t in gt_sorted if lab == KEYWORD]
            with contextlib.redirect_stdout(io.StringIO()):
                rates = TPR.tpr_fpr(KEYWORD, thr, found, kw_times, duration_s=n_samples / sample_rate, time_tolerance_ms=tol,
                                    num_nontarget_words=max(1, len(gt_sorted) - len(kw_times))) if kw_times else None
                dets = TPR.get_groundtruth(found_conf, [KEYWORD], gt_sorted, time_tolerance_ms=tol)
            arrays[f"scores_{seed}_{thr!r}"] = np.array(score, np.float64)       # recognize_element.score at every step
            case["per_threshold"][repr(thr)] = dict(
                found_words=found, found_words_w_confidences=found_conf,
                is_new_steps=[i for i, n in enumerate(is_new) if n], keyword_steps=[i for i, k in enumerate(is_kw) if k],
                stats=run_stats(found, gt_file, tol), tpr_fpr=rates, get_groundtruth=dets)
        expected["cases"].append(case)
    np.savez_compressed(os.path.join(HERE, "reference_postproc.npz"), **arrays)
    with open(os.path.join(HERE, "reference_postproc.json"), "w") as f:
        json.dump(expected, f)
    for c in expected["cases"]:
        print(c["seed"], {k: len(v["found_words"]) for k, v in c["per_threshold"].items()},
              {k: v["stats"]["counters"]["_how_many_c"] for k, v in c["per_threshold"].items()})
    for fn in ("reference_postproc.npz", "reference_postproc.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "B")


if __name__ == "__main__":
    main()
