"""The host mirror of the augmentation policy against the REFERENCE'S OWN CODE.

tests/golden/reference_augment.{npz,json} were produced by executing the reference's `add_background`,
`AudioDataset.augment` (with `random_timeshift`, `random_background_sample`, `get_unknown`) and `map_spec_aug` /
`spec_augment` (input_data.py:141-157, 227-369) unmodified on a numpy stand-in for the `tf.*` calls they make
(tests/golden/make_reference_augment_golden.py; TensorFlow itself cannot run here).  Every random draw the reference
made is on a tape; the mirror replays the tape and must (1) ask for exactly the same draws — kind, bounds, order —
(2) return the same label and (3) the same samples: bit for bit on every path except the background mix, where the two
mean squares are float32 sums in an order TensorFlow does not specify (one-ulp tolerance there).  The device kernels
(`kws_augment_pcm`, `kws_spec_mask`) are in turn bit-exact against this mirror (tests/test_augment_gpu.py)."""
import json
import os

import numpy as np
import pytest

from multilingual_kws_b200.embedding import input_data

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N = 400


def corpus(seed=123):
    """Same generator as tests/golden/make_reference_augment_golden.py::inputs."""
    rng = np.random.default_rng(seed)
    bg = np.zeros((3, 1000), np.float32)
    sizes = np.asarray([1000, 700, 850])
    for i, n in enumerate(sizes):
        bg[i, :n] = (rng.normal(0, 0.2, n) * rng.uniform(0.2, 1.5)).astype(np.float32)
    bg[2, :850] = 0.0
    unknown = {f"/u/{i}.wav": rng.normal(0, 0.3, n).astype(np.float32) for i, n in enumerate((400, 250, 520))}
    clips = [np.clip(rng.normal(0, a, N), -1, 1).astype(np.float32) for a in (0.05, 0.3, 0.9, 0.6)]
    clips[3][:] = np.where(np.arange(N) % 2, 1.0, -1.0).astype(np.float32) * 0.999
    specs = [rng.uniform(0, 26, (49, 40)).astype(np.float32) for _ in range(2)]
    return bg, sizes, unknown, clips, specs


class TapeGen:
    """Stands in for the mirror's numpy Generator: serves the recorded values and checks every request against the tape."""

    def __init__(self, tape):
        self.tape, self.pos = tape, 0

    def _next(self, kind, lo, hi):
        assert self.pos < len(self.tape), f"the mirror asks for more draws than the reference made ({len(self.tape)})"
        k, tlo, thi, v = self.tape[self.pos]
        assert (k, tlo, thi) == (kind, lo, hi), f"draw {self.pos}: reference {k}[{tlo}, {thi}) vs mirror {kind}[{lo}, {hi})"
        self.pos += 1
        return v

    def integers(self, lo, hi):
        return self._next("i", int(lo), int(hi))

    def uniform(self, lo, hi):
        return self._next("f", float(lo), float(hi))

    def done(self):
        return self.pos == len(self.tape)


def make_dataset(cfg, bg, sizes, unknown, tape):
    ds = object.__new__(input_data.AudioDataset)
    ds.model_settings = {"desired_samples": N, "sample_rate": 4000}
    ds.max_time_shift_samples = cfg["shift"]
    ds.background_frequency = cfg["background_frequency"]
    ds.background_volume_range = cfg["background_volume_range"]
    ds.silence_percentage = cfg["silence_percentage"]
    ds.unknown_percentage = cfg["unknown_percentage"]
    ds.unknown_files = list(unknown) if cfg["unknown"] else []
    ds.spec_aug_params = input_data.SpecAugParams(**cfg.get("spec", {}))
    ds.background_data, ds.background_sizes = bg, sizes
    ds.gen = TapeGen(tape)

    def decode_audio(path):                                         # decode_wav(desired_samples): zero-pad / truncate
        out = np.zeros(N, np.float32)
        a = unknown[path]
        out[:min(N, a.shape[0])] = a[:N]
        return out

    ds.decode_audio = decode_audio
    return ds


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "reference_augment.json")) as fh:
        meta = json.load(fh)
    return meta, np.load(os.path.join(HERE, "reference_augment.npz")), corpus()


def test_augment_replays_the_reference(golden):
    meta, arrays, (bg, sizes, unknown, clips, _) = golden
    seen, exact, n_mix = set(), 0, 0
    for case in (c for c in meta["cases"] if c["kind"] == "augment"):
        cfg = meta["configs"][case["config"]]
        ds = make_dataset(cfg, bg, sizes, unknown, case["tape"])
        audio, label = ds.augment(clips[case["clip"]].copy(), "word")
        assert ds.gen.done(), f"{case['key']}: the reference made {len(case['tape'])} draws, the mirror {ds.gen.pos}"
        assert label == case["label"], case["key"]
        want = arrays[case["key"]]
        audio = np.asarray(audio)
        assert audio.dtype == np.float32 and audio.shape == want.shape
        # which branch ran: label and the number of draws identify it; only the background mix involves a mean
        mixed = label == "word" and len(case["tape"]) >= (5 if cfg["shift"] else 4) + (1 if ds.unknown_files else 0)
        if mixed:
            n_mix += 1
            assert np.abs(audio - want).max() <= 2.4e-7, (case["key"], np.abs(audio - want).max())     # 2 ulp at 1.0
            exact += int(np.array_equal(audio, want))
        else:
            assert np.array_equal(audio, want), case["key"]
        seen.add((label, mixed))
    assert seen == {("_silence_", False), ("_unknown_", False), ("word", True), ("word", False)}   # every branch
    assert n_mix >= 20 and exact >= n_mix // 2


def test_add_background_matches_the_reference(golden):
    meta, arrays, (bg, _, _, clips, _) = golden
    for case in (c for c in meta["cases"] if c["kind"] == "add_background"):
        got = input_data.add_background(clips[case["clip"]], bg[case["bg"], :N], np.float32(case["volume"]))
        want = arrays[case["key"]]
        assert got.dtype == np.float32
        # one ulp of the float32 rms ratio, carried by the scaled background term (which reaches ~3 in the loud case)
        fg64, bg64 = clips[case["clip"]].astype(np.float64), bg[case["bg"], :N].astype(np.float64)
        snr = np.sqrt((fg64 ** 2).mean() / (bg64 ** 2).mean()) if bg64.any() else 0.0
        tol = 2.0 ** -22 * max(1.0, 1.0 + np.abs(bg64).max() * snr * case["volume"])
        assert np.abs(got - want).max() <= tol, (case, np.abs(got - want).max(), tol)
        if case["bg"] == 2:                                         # silent background: snr_scaling = 0 -> foreground unchanged
            assert np.array_equal(got, clips[case["clip"]])
        if case["volume"] > 1:                                      # loud mix: clip_by_value is active
            assert (np.abs(want) == 1.0).sum() > 10 and np.array_equal(np.abs(got) == 1.0, np.abs(want) == 1.0)


def test_spec_augment_replays_the_reference(golden):
    meta, arrays, (bg, sizes, unknown, _, specs) = golden
    masked = 0
    for case in (c for c in meta["cases"] if c["kind"] == "map_spec_aug"):
        ds = make_dataset(dict(meta["configs"][0], spec=case["spec_cfg"]), bg, sizes, unknown, case["tape"])
        spec = specs[case["spec"]]
        out, lab = ds.map_spec_aug(spec.copy(), 2)
        assert ds.gen.done() and lab == 2, case["key"]
        zero = np.unpackbits(arrays[case["key"]])[:spec.size].reshape(spec.shape).astype(bool)
        out = np.asarray(out, np.float32)
        assert np.array_equal(out == 0, zero), case["key"]
        assert np.array_equal(out[~zero], spec[~zero]), case["key"]
        masked += int(zero.any())
    assert masked >= 30


def test_device_plan_makes_the_reference_draws(golden):
    """`augment_plan` (the decisions handed to kws_augment_pcm as a 32-byte item per clip) consumes the reference's tape
    exactly like `augment` does, and the item carries the tape's values: shift, branch, background file / offset / volume."""
    from multilingual_kws_b200.augment import MODE_CLIP, MODE_MIX, MODE_SILENCE
    meta, _, (bg, sizes, unknown, _, _) = golden
    modes = set()
    for case in (c for c in meta["cases"] if c["kind"] == "augment"):
        cfg = meta["configs"][case["config"]]
        tape = case["tape"]
        ds = make_dataset(cfg, bg, sizes, unknown, tape)
        ds._bank_row = lambda path: 100 + list(unknown).index(path)           # no device clip bank in this test
        item, label = ds.augment_plan(7, "word")
        assert ds.gen.done() and label == case["label"], case["key"]
        vals = [t[3] for t in tape]
        first_shift = vals[0] if cfg["shift"] else 0
        mode = int(item["mode"])
        modes.add(mode)
        if label == "_silence_":
            assert mode == MODE_SILENCE
            assert (float(item["volume"]), int(item["bg_index"]), int(item["bg_offset"])) == \
                (np.float32(vals[-3]), vals[-2], vals[-1])
        elif label == "_unknown_":
            assert mode == MODE_CLIP
            unk_index = vals[-2] if cfg["shift"] else vals[-1]
            assert int(item["fg_index"]) == 100 + unk_index
            assert int(item["shift"]) == (vals[-1] if cfg["shift"] else 0)     # the unknown clip gets its OWN time shift
        elif mode == MODE_MIX:
            assert (int(item["fg_index"]), int(item["shift"])) == (7, first_shift)
            assert (float(item["volume"]), int(item["bg_index"]), int(item["bg_offset"])) == \
                (np.float32(vals[-3]), vals[-2], vals[-1])
        else:
            assert mode == MODE_CLIP and (int(item["fg_index"]), int(item["shift"])) == (7, first_shift)
    assert modes == {MODE_CLIP, MODE_MIX, MODE_SILENCE}
