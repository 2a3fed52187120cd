#!/usr/bin/env python
"""Golden vectors for the augmentation policy, produced by EXECUTING THE REFERENCE'S OWN CODE:
`add_background`, `AudioDataset.random_timeshift`, `random_background_sample`, `get_unknown`, `augment`, `spec_augment`,
`map_spec_aug` (multilingual_kws/embedding/input_data.py:141-157, 227-369) run unmodified on top of a numpy stand-in for
the handful of `tf.*` calls they make (tests/golden/_tf_numpy_shim.py).  Run in the build container, where
/root/reference exists:  python tests/golden/make_reference_augment_golden.py

Every random draw the reference makes is served from a tape and recorded as (kind, low, high, value).  The fixture holds
the tapes and the reference's outputs; tests/test_reference_augment_golden.py replays the tapes through the repo's host
mirror, which must ask for the same draws in the same order and return the same audio / labels / masks.

Clips are 400 samples long (the code under test is generic in `desired_samples`) to keep the fixture small.
Output: reference_augment.npz (expected audio; for the spectrogram masks the packed set of zeroed cells) + reference_augment.json (tapes, labels, configuration)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _reference_stub  # noqa: E402
import _tf_numpy_shim as shim  # noqa: E402

_reference_stub.install(shim)
import multilingual_kws.embedding.input_data as ID  # noqa: E402  (the reference, with numpy underneath)

N = 400                      # desired_samples of the miniature configuration
SHIFT = 40                   # max_time_shift_samples (time_shift_ms = 100 at 16 kHz would be 1600 of 16000)


def inputs(seed=123):
    """Deterministic miniature corpus: regenerated identically by the test."""
    rng = np.random.default_rng(seed)
    bg = np.zeros((3, 1000), np.float32)
    sizes = np.asarray([1000, 700, 850])
    for i, n in enumerate(sizes):
        bg[i, :n] = (rng.normal(0, 0.2, n) * rng.uniform(0.2, 1.5)).astype(np.float32)
    bg[2, :850] = 0.0                                               # an all-zero background: the snr_scaling = 0 branch
    unknown = {f"/u/{i}.wav": rng.normal(0, 0.3, n).astype(np.float32) for i, n in enumerate((400, 250, 520))}
    clips = [np.clip(rng.normal(0, a, N), -1, 1).astype(np.float32) for a in (0.05, 0.3, 0.9, 0.6)]
    clips[3][:] = np.where(np.arange(N) % 2, 1.0, -1.0).astype(np.float32) * 0.999     # loud: the clip_by_value branch
    specs = [rng.uniform(0, 26, (49, 40)).astype(np.float32) for _ in range(2)]
    return bg, sizes, unknown, clips, specs


def make_dataset(cfg, bg, sizes, unknown):
    ds = object.__new__(ID.AudioDataset)                            # the constructor only reads files and seeds TF's RNG
    ds.model_settings = {"desired_samples": N, "sample_rate": 4000}
    ds.max_time_shift_samples = cfg["shift"]
    ds.background_frequency = cfg["background_frequency"]
    ds.background_volume_range = cfg["background_volume_range"]
    ds.silence_percentage = cfg["silence_percentage"]
    ds.unknown_percentage = cfg["unknown_percentage"]
    ds.unknown_files = list(unknown) if cfg["unknown"] else []
    ds.spec_aug_params = ID.SpecAugParams(**cfg.get("spec", {}))
    ds.background_data, ds.background_sizes = bg, sizes
    ds.gen = shim.Generator()
    return ds


CONFIGS = [
    dict(name="reference defaults", shift=SHIFT, background_frequency=0.8, background_volume_range=0.1,
         silence_percentage=10.0, unknown_percentage=10.0, unknown=True),
    dict(name="transfer_learn settings", shift=SHIFT, background_frequency=0.8, background_volume_range=0.1,
         silence_percentage=10.0, unknown_percentage=50.0, unknown=True),
    dict(name="no unknown files, loud background", shift=SHIFT, background_frequency=1.0, background_volume_range=3.0,
         silence_percentage=30.0, unknown_percentage=50.0, unknown=False),
    dict(name="no time shift", shift=0, background_frequency=0.5, background_volume_range=0.1,
         silence_percentage=0.0, unknown_percentage=100.0, unknown=True),
]


def main():
    bg, sizes, unknown, clips, specs = inputs()
    shim.FILES.update(unknown)
    arrays, cases = {}, []
    for ci, cfg in enumerate(CONFIGS):
        ds = make_dataset(cfg, bg, sizes, unknown)
        for k in range(24):
            shim.TAPE = shim.Tape(1000 * ci + k)
            audio, label = ds.augment(clips[k % len(clips)].copy(), "word")
            key = f"aug_{ci}_{k}"
            arrays[key] = np.asarray(audio, np.float32)
            assert arrays[key].shape == (N,) and np.asarray(audio).dtype == np.float32, (key, np.asarray(audio).dtype)
            cases.append(dict(kind="augment", config=ci, clip=k % len(clips), tape=shim.TAPE.rec, label=str(label), key=key))
    # add_background alone, including the silent-background branch
    for k, (f, b, v) in enumerate([(0, 0, 0.1), (1, 1, 0.5), (2, 2, 0.7), (3, 0, 2.5)]):
        out = ID.add_background(clips[f], bg[b, :N], np.float32(v))
        arrays[f"mix_{k}"] = np.asarray(out, np.float32)
        cases.append(dict(kind="add_background", clip=f, bg=b, volume=v, key=f"mix_{k}"))
    # spec_augment / map_spec_aug with the reference's default parameters and a heavier setting
    for si, spec_cfg in enumerate([{}, dict(percentage=100.0, frequency_n_range=4, frequency_max_px=5, time_n_range=3, time_max_px=6)]):
        ds = make_dataset(dict(CONFIGS[0], spec=spec_cfg), bg, sizes, unknown)
        for k in range(30):
            shim.TAPE = shim.Tape(50000 + 100 * si + k)
            out, lab = ds.map_spec_aug(specs[k % 2].copy(), 2)
            key = f"spec_{si}_{k}"
            out = np.asarray(out, np.float32)
            assert np.array_equal(out[out != 0], specs[k % 2][out != 0])          # a 0 / 1 mask was applied, nothing else
            arrays[key] = np.packbits(out == 0)                                   # the zeroed cells are the whole result
            cases.append(dict(kind="map_spec_aug", spec_cfg=spec_cfg, spec=k % 2, tape=shim.TAPE.rec, key=key))
    np.savez_compressed(os.path.join(HERE, "reference_augment.npz"), **arrays)
    with open(os.path.join(HERE, "reference_augment.json"), "w") as fh:
        json.dump(dict(n=N, configs=CONFIGS, cases=cases), fh)
    labels = [c["label"] for c in cases if c["kind"] == "augment"]
    print(f"{len(cases)} cases; labels {dict((l, labels.count(l)) for l in set(labels))}; "
          f"draws per augment call {sorted(set(len(c['tape']) for c in cases if c['kind'] == 'augment'))}")


if __name__ == "__main__":
    main()
