"""True/false-positive rates of a detection list against ground truth, with the semantics of reference
``multilingual_kws/embedding/tpr_fpr.py:1-138``.  Re-implemented around one helper (`_any_within`) that keeps the
reference's scan-and-break order semantics; parity against vectors produced by executing the reference
(tests/golden/reference_postproc.*)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence


def _any_within(times: Iterable, centre, tolerance) -> bool:
    """True if the in-order scan of `times` meets a value in [centre - tol, centre + tol] before one above it
    (the reference breaks at the first value beyond the window, so order matters for unsorted input)."""
    for t in times:
        if t > centre + tolerance:
            return False
        if t >= centre - tolerance:
            return True
    return False


def get_groundtruth(found_words: Sequence, targets: Sequence[str], groundtruth: Sequence, time_tolerance_ms=1500):
    """Per-detection tp / fp records and per-ground-truth fn records for targets[0] (the reference returns from inside
    its loop over targets, so only the first target is ever processed; no targets -> None)."""
    for target in targets:
        gt_times = [t for k, t in groundtruth if k == target]
        found = [f for f in found_words if f[0] == target]
        print("gt target occurences", len(gt_times))
        print("num found targets", len(found))
        found_times = [f[1] for f in found]
        records: List[dict] = [dict(keyword=target, time_ms=t, groundtruth="fn") for t in gt_times
                               if not _any_within(found_times, t, time_tolerance_ms)]
        for _, when, confidence in found:
            kind = "tp" if _any_within(gt_times, when, time_tolerance_ms) else "fp"
            records.append(dict(keyword=target, time_ms=when, confidence=confidence, groundtruth=kind))
        return records
    return None


def tpr_fpr(keyword, thresh, found_words, gt_target_times_ms, duration_s, time_tolerance_ms,
            num_nontarget_words: Optional[int] = None) -> dict:
    found_times = [t for w, t in found_words if w == keyword]
    n_gt = len(gt_target_times_ms)
    false_negatives = sum(1 for t in gt_target_times_ms if not _any_within(found_times, t, time_tolerance_ms))
    true_positives = sum(1 for t in found_times if _any_within(gt_target_times_ms, t, time_tolerance_ms))
    if true_positives > n_gt:
        print("WARNING: weird timing issue")
        true_positives = n_gt
    false_positives = len(found_times) - true_positives
    result = dict(keyword=keyword, tpr=true_positives / n_gt, thresh=thresh, true_positives=true_positives,
                  false_positives=false_positives, false_negatives=false_negatives,
                  false_rejections_per_instance=false_negatives / n_gt,
                  false_accepts_per_hour=false_positives / duration_s * 3600, groundtruth_positives=n_gt)
    if num_nontarget_words is not None:
        result["fpr"] = false_positives / num_nontarget_words
    return result
