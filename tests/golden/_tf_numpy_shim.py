"""A numpy-backed stand-in for the few `tf.*` calls the reference's AUGMENTATION code makes — used ONLY by
tests/golden/make_reference_augment_golden.py (build container, where /root/reference exists) to EXECUTE the reference's
own `add_background`, `AudioDataset.random_timeshift / random_background_sample / get_unknown / augment / spec_augment /
map_spec_aug` (multilingual_kws/embedding/input_data.py:141-157, 227-369) without TensorFlow.

What this can and cannot pin: the reference's program logic — which random draws are made, in which order and with which
bounds, the branch structure, padding / slicing / masking semantics, operand order and float32 typing — runs as written.
TensorFlow's own arithmetic does not: `reduce_mean` here is numpy's float32 mean, TF's summation order is unspecified, so
the mixed-background samples are compared with a one-ulp tolerance, everything else bit for bit.

Random numbers: every `tf.random.uniform` / `Generator.uniform` call is served from a TAPE that records
(kind, low, high, value); the test replays the tape through the repo's mirror and requires the same sequence of
requests."""
import types

import numpy as np

float32, int32, int16, int64 = np.float32, np.int32, np.int16, np.int64
dtypes = types.SimpleNamespace(float32=np.float32, int32=np.int32, int16=np.int16, int64=np.int64)


class Tape:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.rec = []

    def draw(self, lo, hi, dtype):
        if np.issubdtype(np.dtype(dtype), np.integer):
            v = int(self.rng.integers(int(lo), int(hi)))
            self.rec.append(["i", int(lo), int(hi), v])
            return np.int32(v)
        v = np.float32(self.rng.uniform(float(lo), float(hi)))          # TF draws float32 scalars
        self.rec.append(["f", float(lo), float(hi), float(v)])
        return v


TAPE = None          # the active tape (set by the generator script before each case)
FILES = {}           # path -> float32 waveform, for tf.io.read_file / tf.audio.decode_wav


class Generator:     # tf.random.Generator
    def uniform(self, shape, minval=0, maxval=None, dtype=np.float32):
        assert list(shape) == []
        return TAPE.draw(minval, maxval, dtype)


def _uniform(shape, minval=0, maxval=None, dtype=np.float32, seed=None, name=None):
    assert list(shape) == []
    return TAPE.draw(minval, maxval, dtype)


random = types.SimpleNamespace(uniform=_uniform, Generator=Generator)


def constant(value, dtype=None):
    if dtype is None:
        dtype = np.float32 if isinstance(value, float) else np.int32
    return np.asarray(value, dtype)[()]


def convert_to_tensor(value, dtype=None):
    return np.asarray(value, dtype)


def sqrt(x): return np.sqrt(x)
def square(x): return np.square(x)
def reduce_mean(x): return np.mean(x, dtype=np.asarray(x).dtype)
def greater(a, b): return a > b
def cond(pred, true_fn, false_fn): return true_fn() if bool(pred) else false_fn()
def divide(a, b): return np.divide(a, b)
def multiply(a, b): return np.multiply(a, b)
def add(a, b): return np.add(a, b)
def clip_by_value(x, lo, hi): return np.clip(x, lo, hi)
def reshape(x, shape): return np.reshape(x, shape)
def expand_dims(x, axis): return np.expand_dims(x, axis)
def stack(values, axis=0): return np.stack([np.asarray(v) for v in values], axis=axis)
def squeeze(x, axis=None): return np.squeeze(x, axis=axis)
def shape(x): return np.asarray(np.shape(x), np.int32)
def concat(values, axis): return np.concatenate(values, axis=axis)
def ones(shape_, dtype=np.float32): return np.ones([int(s) for s in shape_], dtype)
def zeros(shape_, dtype=np.float32): return np.zeros([int(s) for s in shape_], dtype)
def gather(params, indices): return params[int(indices)]
def function(fn=None, **kw): return fn if fn is not None else (lambda f: f)


def pad(tensor, paddings, mode="CONSTANT", constant_values=0):
    assert mode == "CONSTANT"
    return np.pad(tensor, np.asarray(paddings), constant_values=constant_values)


def slice(input_, begin, size):          # noqa: A001  (tf.slice)
    idx = tuple(np.s_[int(b):int(b) + int(s)] for b, s in zip(begin, size))
    return input_[idx]


def while_loop(cond_fn, body, loop_vars):
    v = tuple(loop_vars)
    while bool(cond_fn(*v)):
        v = tuple(body(*v))
    return v


def _read_file(path):
    return path                           # the "binary" is just the key into FILES


def _decode_wav(contents, desired_channels=-1, desired_samples=-1):
    a = np.asarray(FILES[str(contents)], np.float32)
    if desired_samples is not None and desired_samples > 0:          # tf.audio.decode_wav: zero-pad / truncate
        out = np.zeros(desired_samples, np.float32)
        out[:min(desired_samples, a.shape[0])] = a[:desired_samples]
        a = out
    return a[:, None], np.int32(16000)


io = types.SimpleNamespace(read_file=_read_file)
audio = types.SimpleNamespace(decode_wav=_decode_wav)
