"""Streaming accuracy bookkeeping with the semantics of reference
``multilingual_kws/embedding/accuracy_utils.py:24-245`` (``StreamingAccuracyStats``), used by
``calculate_streaming_accuracy`` (batch_streaming_analysis.py:140-170).

Re-implemented, not copied: ground truth is kept as two parallel sorted arrays and every "first ground truth inside
the tolerance window" query is a bisection instead of a scan from the start of the list.  Attribute names, the
counters' meaning, ``delta()``'s strings / ``ValueError`` and ``print_accuracy_stats()``'s return value are the
reference's; parity is checked against vectors produced by executing the reference (tests/golden/reference_postproc.*).
"""
from __future__ import annotations

import bisect
import math
from typing import Dict, List, Sequence

from . import input_data

_NON_TARGET = (input_data.SILENCE_LABEL, input_data.UNKNOWN_WORD_LABEL)


class StreamingAccuracyStats:
    def __init__(self, target_keyword: str):
        self.target_keyword = target_keyword
        self._gt_occurrence: List[list] = []        # [label, time_ms], time-ordered (kept for API compatibility)
        self._gt_times: List[int] = []
        self._how_many_gt = self._how_many_gt_matched = 0
        self._how_many_fp = self._how_many_c = self._how_many_w = self._how_many_fn = 0
        self._previous_c = self._previous_w = self._previous_fp = 0
        self._how_many_gt_target = self._how_many_gt_unknown_or_silence = 0
        self._which_matched: Dict[str, int] = {}
        self._which_wrong: Dict[str, int] = {}

    # ------------------------------------------------------------------ ground truth
    def read_ground_truth_file(self, file_name) -> None:
        """`label,time_ms` per line; lines without exactly two fields are skipped; times rounded half-to-even."""
        with open(file_name, "r") as f:
            for line in f:
                fields = line.strip().split(",")
                if len(fields) == 2:
                    self._gt_occurrence.append([fields[0], round(float(fields[1]))])
        self._gt_occurrence.sort(key=lambda item: item[1])          # stable, like sorted()
        self._gt_times = [t for _, t in self._gt_occurrence]

    def set_ground_truth(self, occurrences: Sequence[Sequence]) -> None:
        """Same as read_ground_truth_file for an in-memory list of (label, time_ms)."""
        self._gt_occurrence = sorted(([lab, round(float(t))] for lab, t in occurrences), key=lambda item: item[1])
        self._gt_times = [t for _, t in self._gt_occurrence]

    # ------------------------------------------------------------------ statistics
    def delta(self) -> str:
        d_fp, d_w, d_c = (self._how_many_fp - self._previous_fp, self._how_many_w - self._previous_w,
                          self._how_many_c - self._previous_c)
        if d_fp == 1:
            state = "(False Positive)"
        elif d_c == 1:
            state = "(Correct)"
        elif d_w == 1:
            state = "(Wrong)"
        else:
            raise ValueError("Unexpected state in statistics")
        self._previous_c, self._previous_w, self._previous_fp = self._how_many_c, self._how_many_w, self._how_many_fp
        return state

    def calculate_accuracy_stats(self, found_words, up_to_time_ms, time_tolerance_ms) -> None:
        horizon = math.inf if up_to_time_ms == -1 else up_to_time_ms + time_tolerance_ms
        times, occ = self._gt_times, self._gt_occurrence
        n_seen = bisect.bisect_right(times, horizon)                 # ground truths with time <= horizon
        self._how_many_gt = n_seen
        self._how_many_gt_unknown_or_silence = sum(1 for lab, _ in occ[:n_seen] if lab in _NON_TARGET)
        self._how_many_gt_target = sum(1 for lab, _ in occ[:n_seen] if lab == self.target_keyword)

        words = [input_data.SILENCE_LABEL, input_data.UNKNOWN_WORD_LABEL, self.target_keyword]
        self._which_matched = {w: 0 for w in words}
        self._which_wrong = {w: 0 for w in words}
        fp = correct = wrong = 0
        used_times = set()
        for found in found_words:
            label, when = found[0], found[1]
            i = bisect.bisect_left(times, when - time_tolerance_ms)  # first ground truth not before the window
            if i < n_seen and times[i] <= when + time_tolerance_ms:
                gt_label, gt_time = occ[i]
                if gt_label == label and gt_time not in used_times:
                    correct += 1
                    self._which_matched[label] += 1
                else:
                    wrong += 1
                    if gt_label in _NON_TARGET and label == self.target_keyword:
                        self._which_wrong[gt_label] += 1
                used_times.add(gt_time)
            else:
                fp += 1
        self._how_many_fp, self._how_many_c, self._how_many_w = fp, correct, wrong
        self._how_many_gt_matched = len(used_times)

        # false negatives: ground truths strictly before the horizon with no detection strictly inside +- tolerance
        found_times = sorted(f[1] for f in found_words)
        missed = 0
        for t in times[:bisect.bisect_left(times, horizon)]:
            j = bisect.bisect_right(found_times, t - time_tolerance_ms)      # first detection with time > t - tol
            if not (j < len(found_times) and found_times[j] < t + time_tolerance_ms):
                missed += 1
        self._how_many_fn = missed

    def print_accuracy_stats(self):
        if self._how_many_gt == 0:
            print("No ground truth yet, {}false positives".format(self._how_many_fp))
            return None
        pct = lambda n: n / self._how_many_gt * 100                          # noqa: E731
        info = ("{:.1f}% matched, {:.1f}% correct, {:.1f}% wrong, {:.1f}% false positive, {:.1f}% false negative, "
                "{:.1f} howmanyfp, {:.1f} howmanyfn").format(pct(self._how_many_gt_matched), pct(self._how_many_c),
                                                             pct(self._how_many_w), pct(self._how_many_fp),
                                                             pct(self._how_many_fn), self._how_many_fp, self._how_many_fn)
        print(info)
        stat = {
            "correct_match_percentage": pct(self._how_many_c),
            "wrong_match_percentage": pct(self._how_many_w),
            "howmanyfp": self._how_many_fp,
            "howmanyfn": self._how_many_fn,
            "wrong": dict(self._which_wrong),
            "matched": dict(self._which_matched),
            "num_groundtruth_target": self._how_many_gt_target,
            "num_groundtruth_unknown_or_silence": self._how_many_gt_unknown_or_silence,
        }
        return info, stat
