"""Streaming (sliding-window) inference, mirroring reference
``multilingual_kws/embedding/batch_streaming_analysis.py:27-241``.

The reference slides a 1 s window with a 20 ms stride over the wav and calls the CPU frontend once per window
from a Python loop (:108-115) before one ``model.predict``.  Here the whole signal goes to the GPU once, the
per-frame frontend work is computed once per 20 ms frame and shared by all windows (bit-identical, see
kws_frontend_stream), windows are embedded + classified in large batches, and the post-processing recurrence runs
on the device for all detection thresholds at once (kws_stream_detect) when the softmax rows were produced there.
Ground-truth bookkeeping follows the reference (accuracy_utils.StreamingAccuracyStats), results keep its structure.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import input_data
from .accuracy_utils import StreamingAccuracyStats
from .single_target_recognize_commands import detect_stream, detect_stream_device
from ..frontend import FEATURE_SCALE, float_audio_to_int16_np


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


@dataclass(frozen=True)
class StreamFlags:
    wav: os.PathLike
    ground_truth: os.PathLike
    target_keyword: str
    detection_thresholds: List[float]
    clip_duration_ms: int = 1000
    clip_stride_ms: int = 20  # window_stride_ms in model_settings
    average_window_duration_ms: int = 100
    suppression_ms: int = 500
    time_tolerance_ms: int = 750
    minimum_count: int = 4
    max_chunk_length_sec: int = 1200

    def labels(self) -> List[str]:
        return [input_data.SILENCE_LABEL, input_data.UNKNOWN_WORD_LABEL, self.target_keyword]


@dataclass
class StreamTarget:
    target_lang: str
    target_word: str
    model_path: os.PathLike
    stream_flags: List[StreamFlags]
    destination_result_pkl: Optional[os.PathLike] = None
    destination_result_inferences: Optional[os.PathLike] = None


def stream_inferences(model, model_settings, audio: np.ndarray, sample_rate: int, clip_duration_ms: int,
                      clip_stride_ms: int, window_batch: int = 8192, keep_on_device: bool = False):
    """softmax rows [W, n_labels] for every window offset in range(0, len - clip, stride)
    (batch_streaming_analysis.py:66-117; the chunk branches there add up to the un-chunked result).

    Under torch.distributed (one process per GPU) the window range is sharded contiguously: rank r uploads only the
    samples its windows [W r / n, W (r+1) / n) cover, and the rows are all-gathered, so every rank returns the full
    matrix (windows are independent; a window's features depend on its own samples only)."""
    clip = int(clip_duration_ms * sample_rate / 1000)
    stride = int(clip_stride_ms * sample_rate / 1000)
    fe = input_data._frontend_for(model_settings)
    pcm_host = float_audio_to_int16_np(audio)
    W = fe.stream_num_windows(pcm_host.shape[0], clip, stride)
    n_labels = model.head.classes if hasattr(model, "head") else model.output_dim
    if W <= 0:
        return torch.zeros((0, n_labels), device="cuda") if keep_on_device else np.zeros((0, n_labels), np.float32)
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    lo, hi = W * rank // world, W * (rank + 1) // world
    outs = []
    if hi > lo:
        # window offsets are range(0, n - clip, stride): the last offset must be < n - clip, hence the extra sample
        pcm = torch.from_numpy(np.ascontiguousarray(pcm_host[lo * stride:(hi - 1) * stride + clip + 1])).cuda()
        st = fe.stream_prepare(pcm)
        for w0 in range(0, hi - lo, window_batch):
            feats = st.windows(clip, stride, w0, min(window_batch, hi - lo - w0), FEATURE_SCALE)
            outs.append(model.forward_device(feats).clone())
    probs = torch.cat(outs) if outs else torch.zeros((0, n_labels), device="cuda")
    if dist is not None:
        most = -(-W // world)                                     # shard sizes differ by at most one row
        padded = torch.zeros((most, n_labels), dtype=probs.dtype, device=probs.device)
        padded[:probs.shape[0]] = probs
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded)
        probs = torch.cat([parts[r][:W * (r + 1) // world - W * r // world] for r in range(world)])
    return probs if keep_on_device else probs.cpu().numpy()


def reference_chunk_rows(n_samples: int, sample_rate: int, clip_samples: int, stride_samples: int, max_chunk_length_sec: int):
    """Row layout of the `inferences` matrix the REFERENCE returns for a wav of n_samples
    (batch_streaming_analysis.py:72-123).  Its chunk branch is inverted: for audio of at least max_chunk_length_sec the first
    "chunk" is the whole signal and every later offset k * max_chunk contributes audio[offset : offset + max_chunk]
    again, so the saved matrix is the W window rows followed by the windows of those tails.  Returns [(sample offset,
    number of windows)] per chunk; the first entry is always (0, W).  Only the first W rows are ever read back
    (the post-processing loop indexes by window, :131-140)."""
    max_chunk = int(max_chunk_length_sec * sample_rate)

    def n_windows(length):
        return max(0, -(-(length - clip_samples) // stride_samples))

    if n_samples < max_chunk:
        return [(0, n_windows(n_samples))]
    out = []
    for offset in range(0, n_samples, max_chunk):
        if offset + max_chunk > n_samples:
            out.append((offset, n_windows(min(max_chunk, n_samples - offset))))
        else:
            out.append((offset, n_windows(n_samples - offset)))
    return out


def calculate_streaming_accuracy(model, model_settings, flag_list, existing_inferences=None):
    assert len(set([f.wav for f in flag_list])) == 1, "can only process one wav"
    assert len(set([f.clip_duration_ms for f in flag_list])) == 1, "cannot vary"
    assert len(set([f.clip_stride_ms for f in flag_list])) == 1, "cannot vary"
    f0 = flag_list[0]
    with open(os.fspath(f0.wav), "rb") as fh:
        audio, sample_rate = input_data.decode_wav(fh.read(), desired_channels=1)
    audio = audio[:, 0]
    clip_duration_samples = int(f0.clip_duration_ms * sample_rate / 1000)
    clip_stride_samples = int(f0.clip_stride_ms * sample_rate / 1000)
    audio_data_end = audio.shape[0] - clip_duration_samples
    times = [int(o * 1000 / sample_rate) for o in range(0, audio_data_end, clip_stride_samples)]
    device_probs = None
    if existing_inferences is not None:
        inferences = existing_inferences
    else:
        device_probs = stream_inferences(model, model_settings, audio, sample_rate, f0.clip_duration_ms,
                                         f0.clip_stride_ms, keep_on_device=True)
        inferences = device_probs.cpu().numpy()
        # the matrix the reference RETURNS (and eval_stream_test saves) repeats the windows of the chunk tails for audio
        # of >= max_chunk_length_sec (its inverted chunk branch, see reference_chunk_rows): same rows, same order
        extra = []
        for offset, n_w in reference_chunk_rows(audio.shape[0], sample_rate, clip_duration_samples, clip_stride_samples,
                                                f0.max_chunk_length_sec)[1:]:
            if n_w <= 0:
                continue
            if offset % clip_stride_samples == 0:        # those windows are windows of the whole signal: same rows
                w0 = offset // clip_stride_samples
                extra.append(inferences[w0:w0 + n_w])
            else:
                extra.append(stream_inferences(model, model_settings, audio[offset:offset + int(f0.max_chunk_length_sec * sample_rate)],
                                               sample_rate, f0.clip_duration_ms, f0.clip_stride_ms))
        if extra:
            inferences = np.concatenate([inferences] + extra, axis=0)
    results = []
    for FLAGS in flag_list:
        res_thresh = {}
        if device_probs is not None:      # softmax rows are on the GPU: the whole threshold sweep is one launch
            sweep = detect_stream_device(device_probs[:len(times)], times, FLAGS.labels(), FLAGS.average_window_duration_ms,
                                         FLAGS.detection_thresholds, FLAGS.suppression_ms, FLAGS.minimum_count, target_id=2)
        for threshold in FLAGS.detection_thresholds:
            if device_probs is not None:
                found = sweep[threshold]
            else:                         # caller-supplied host matrix (reference: existing_inferences)
                found = detect_stream(inferences, times, FLAGS.labels(), FLAGS.average_window_duration_ms, threshold,
                                      FLAGS.suppression_ms, FLAGS.minimum_count, target_id=2)
            all_found_words = [[w, t] for w, t, _ in found]
            # accuracy statistics, updated detection by detection as the reference does (:150-170)
            stats = StreamingAccuracyStats(target_keyword=FLAGS.target_keyword)
            stats.read_ground_truth_file(FLAGS.ground_truth)
            for n in range(1, len(all_found_words) + 1):
                stats.calculate_accuracy_stats(all_found_words[:n], all_found_words[n - 1][1], FLAGS.time_tolerance_ms)
                stats.delta()
            print(f"results for {threshold:0.2f}")
            stats.calculate_accuracy_stats(all_found_words, -1, FLAGS.time_tolerance_ms)
            stats.print_accuracy_stats()
            res_thresh[threshold] = (all_found_words, [[w, t, s] for w, t, s in found])
        results.append((FLAGS, res_thresh))
    return results, inferences


def eval_stream_test(st: StreamTarget, live_model=None):
    from ..fewshot import FewShotModel
    model = live_model if live_model is not None else FewShotModel.load(st.model_path)
    model_settings = input_data.standard_microspeech_model_settings(label_count=3)
    if st.destination_result_pkl is not None and os.path.isfile(st.destination_result_pkl):
        print("results already present", st.destination_result_pkl, flush=True)
        return
    loaded = None
    if st.destination_result_inferences is not None and os.path.isfile(st.destination_result_inferences):
        print("inferences already present", flush=True)
        loaded = np.load(st.destination_result_inferences)
    results = {}
    results[st.target_word], inferences = calculate_streaming_accuracy(model, model_settings, st.stream_flags, loaded)
    if st.destination_result_pkl is not None:
        with open(st.destination_result_pkl, "wb") as fh:
            pickle.dump(results, fh)
    if loaded is None and st.destination_result_inferences is not None:
        np.save(st.destination_result_inferences, inferences)
    return results
