"""Streaming post-processor with the semantics of reference
``multilingual_kws/embedding/single_target_recognize_commands.py:94-207`` (itself derived from TF's
speech_commands ``recognize_commands.py``): average the softmax rows that fall inside a trailing time window,
fire when the target's mean score clears the threshold, the label changed and the suppression interval passed.

Re-implemented (not copied): the window is kept as a deque with running column sums, and
``detect_stream`` runs the whole recurrence over an inference matrix in one call.
"""
from __future__ import annotations

import collections
from typing import List, Sequence, Tuple

import numpy as np

SILENCE = "_silence_"


class RecognizeResult:
    """Mutable result slot filled by process_latest_result (same three attributes as the reference)."""

    def __init__(self):
        self.found_command = SILENCE
        self.score = 0
        self.is_new_command = False


class SingleTargetRecognizeCommands:
    def __init__(self, labels: Sequence[str], average_window_duration_ms, detection_threshold, suppression_ms,
                 minimum_count, target_id):
        self._labels = list(labels)
        self._target_id = target_id
        self._average_window_duration_ms = average_window_duration_ms
        self._detection_threshold = detection_threshold
        self._suppression_ms = suppression_ms
        self._minimum_count = minimum_count
        self._label_count = len(self._labels)
        self._window = collections.deque()            # (time_ms, scores)
        self._previous_top_label = SILENCE
        self._previous_top_time = -np.inf

    def process_latest_result(self, latest_results, current_time_ms, recognize_element: RecognizeResult) -> None:
        latest_results = np.asarray(latest_results)
        if latest_results.shape[0] != self._label_count:
            raise ValueError("The results for recognition should contain {} elements, but there are {} produced".format(
                self._label_count, latest_results.shape[0]))
        if self._window and current_time_ms < self._window[0][0]:
            raise ValueError("Results must be fed in increasing time order, but receive a timestamp of {}, which was "
                             "earlier than the previous one of {}".format(current_time_ms, self._window[0][0]))
        self._window.append((current_time_ms, latest_results))
        limit = current_time_ms - self._average_window_duration_ms
        while self._window[0][0] < limit:              # strictly older than the window start is dropped
            self._window.popleft()
        count = len(self._window)
        span = current_time_ms - self._window[0][0]
        if count < self._minimum_count or span < self._average_window_duration_ms / 4:
            recognize_element.found_command = self._previous_top_label
            recognize_element.score = 0.0
            recognize_element.is_new_command = False
            return
        # mean of the target column, accumulated in arrival order as score/count (matches the reference's float sums)
        mean = 0.0
        for _, scores in self._window:
            mean += float(scores[self._target_id]) / count      # float64 division, as numpy 1.x promoted it
        label = self._labels[self._target_id] if mean > self._detection_threshold else SILENCE
        if self._previous_top_label == SILENCE or self._previous_top_time == -np.inf:
            since = np.inf
        else:
            since = current_time_ms - self._previous_top_time
        fire_word = mean > self._detection_threshold and label != self._previous_top_label
        fire_silence = mean < self._detection_threshold and label == SILENCE
        if (fire_word or fire_silence) and since > self._suppression_ms:
            self._previous_top_label = label
            self._previous_top_time = current_time_ms
            recognize_element.is_new_command = True
        else:
            recognize_element.is_new_command = False
        recognize_element.found_command = label
        recognize_element.score = mean


def detect_stream(inferences: np.ndarray, times_ms: Sequence[int], labels: Sequence[str], average_window_duration_ms,
                  detection_threshold, suppression_ms, minimum_count, target_id=2) -> List[Tuple[str, int, float]]:
    """Runs the recogniser over a whole [W, n_labels] inference matrix; returns (word, time_ms, score) detections
    (new, non-silence commands), i.e. `all_found_words_w_confidences` of batch_streaming_analysis.py:140-163."""
    rc = SingleTargetRecognizeCommands(labels, average_window_duration_ms, detection_threshold, suppression_ms,
                                       minimum_count, target_id)
    el = RecognizeResult()
    found = []
    for row, t in zip(np.asarray(inferences), times_ms):
        rc.process_latest_result(row, t, el)
        if el.is_new_command and el.found_command != SILENCE:
            found.append((el.found_command, int(t), float(el.score)))
    return found


def detect_stream_device(inferences, times_ms: Sequence[int], labels: Sequence[str], average_window_duration_ms,
                         detection_thresholds: Sequence[float], suppression_ms, minimum_count, target_id=2,
                         return_scores: bool = False):
    """The same recurrence on the GPU for a whole sweep of thresholds in one call (``kws_stream_detect``,
    csrc/stream_detect.cu): `inferences` is the float32 [W, n_labels] softmax matrix (a CUDA tensor as the head left
    it, or anything `torch.as_tensor` accepts, which is uploaded).  Returns {threshold: [(word, time_ms, score), ...]}
    — per threshold exactly what `detect_stream` returns, scores bit-identical (same float64 operation order) — and,
    with return_scores, the per-step `recognize_element.score` array as well.  No CPU fallback."""
    import torch

    from .. import _lib
    labels = list(labels)
    if labels[target_id] == SILENCE:
        raise ValueError("the target label must not be the silence label")
    probs = torch.as_tensor(inferences)
    if probs.dtype != torch.float32:
        probs = probs.float()
    probs = probs.cuda().contiguous()
    if probs.dim() != 2 or probs.shape[1] != len(labels):
        raise ValueError("The results for recognition should contain {} elements, but there are {} produced".format(
            len(labels), probs.shape[-1]))
    W = probs.shape[0]
    t_host = np.asarray(times_ms, dtype=np.int64)
    if t_host.shape != (W,):
        raise ValueError(f"need one timestamp per inference row ({W}), got {t_host.shape}")
    if W > 1 and (np.diff(t_host) < 0).any():
        i = int(np.argmax(np.diff(t_host) < 0)) + 1
        raise ValueError("Results must be fed in increasing time order, but receive a timestamp of {}, which was "
                         "earlier than the previous one of {}".format(int(t_host[i]), int(t_host[i - 1])))
    thr_host = np.asarray(list(detection_thresholds), dtype=np.float64)
    dev = probs.device
    times = torch.from_numpy(t_host).to(dev)
    thr = torch.from_numpy(thr_host).to(dev)
    scores = torch.empty(W, dtype=torch.float64, device=dev)
    valid = torch.empty(W, dtype=torch.uint8, device=dev)
    cap = max(W, 1)
    found_idx = torch.empty((len(thr_host), cap), dtype=torch.int32, device=dev)
    found_cnt = torch.empty(len(thr_host), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().kws_stream_detect(
        probs.data_ptr(), W, probs.shape[1], int(target_id), times.data_ptr(), float(average_window_duration_ms),
        float(suppression_ms), int(minimum_count), thr.data_ptr(), len(thr_host), scores.data_ptr(), valid.data_ptr(),
        found_idx.data_ptr(), found_cnt.data_ptr(), cap, _lib.current_stream_ptr()), "kws_stream_detect")
    counts = found_cnt.cpu().numpy()
    idx_host = found_idx.cpu().numpy()
    scores_host = scores.cpu().numpy()
    word = labels[target_id]
    out = {}
    for k, th in enumerate(detection_thresholds):
        ids = idx_host[k, :counts[k]]
        out[th] = [(word, int(t_host[i]), float(scores_host[i])) for i in ids]
    return (out, scores_host) if return_scores else out
