"""Device-side augmentation (kws_augment_pcm / kws_spec_mask, §8 f3) vs the host mirror of the reference's
augment / add_background / random_timeshift / spec_augment (input_data.py:141-157, 227-369): every mode of the plan
against numpy float32 arithmetic (bit-exact PCM), and whole AudioDataset batches device path == host path for the
same seed."""
import numpy as np
import pytest
import torch

from multilingual_kws_b200.augment import (AUG_ITEM, MODE_CLIP, MODE_MIX, MODE_SILENCE, DeviceAugmenter, exact_int16,
                                           plan_item, spec_mask_)
from multilingual_kws_b200.frontend import float_audio_to_int16_np
from multilingual_kws_b200.synthetic import synthetic_pcm

pytestmark = pytest.mark.gpu
N = 16000


def host_shift(audio, amount):
    if amount > 0:
        return np.concatenate([np.zeros(amount, np.float32), audio])[:N]
    return np.concatenate([audio, np.zeros(-amount, np.float32)])[-amount:-amount + N]


def host_item(clips, bg, it):
    from multilingual_kws_b200.embedding.input_data import add_background
    mode = int(it["mode"])
    fg = host_shift(clips[int(it["fg_index"])], int(it["shift"])) if mode != MODE_SILENCE else None
    b = bg[int(it["bg_index"]), int(it["bg_offset"]):int(it["bg_offset"]) + N] if mode != MODE_CLIP else None
    if mode == MODE_CLIP:
        return fg
    if mode == MODE_SILENCE:
        return b * np.float32(it["volume"])
    return add_background(fg, b, it["volume"])


@pytest.fixture(scope="module")
def world(kws_lib):
    rng = np.random.default_rng(21)
    clips = synthetic_pcm(12, cfg_id=7).astype(np.float32) / np.float32(32768.0)
    bg = np.zeros((3, 60003), np.float32)
    for i, n in enumerate((60003, 31000, 16001)):
        bg[i, :n] = (np.clip(rng.normal(0, 0.05 * (i + 1), n), -1, 1) * 32768).astype(np.int16).astype(np.float32) / 32768
    bg[2] = 0.0                                                   # an all-zero recording: rms 0 -> snr 0 branch
    aug = DeviceAugmenter(N, bg)
    for c in clips:
        aug.clips.add(c)
    return dict(clips=clips, bg=bg, aug=aug, rng=rng)


def test_every_mode_bit_exact(world):
    rng, aug = world["rng"], world["aug"]
    items = []
    for shift in (0, 1, -1, 7, -8, 1599, -1600, 15999, -15999, 16000, -16000):
        items.append(plan_item(MODE_CLIP, fg_index=int(rng.integers(0, 12)), shift=shift))
    for k in range(40):
        bi = int(rng.integers(0, 3))
        length = (60003, 31000, 16001)[bi]
        off = int(rng.integers(0, length - N))
        vol = rng.uniform(0, 1)
        mode = MODE_SILENCE if k % 3 == 0 else MODE_MIX
        items.append(plan_item(mode, fg_index=int(rng.integers(0, 12)), shift=int(rng.integers(-1600, 1600)), bg_index=bi,
                               bg_offset=off, volume=vol if mode == MODE_SILENCE else vol * 0.1))
    items.append(plan_item(MODE_MIX, fg_index=3, shift=0, bg_index=0, bg_offset=5, volume=30.0))     # saturates the clip
    items.append(plan_item(MODE_MIX, fg_index=11, shift=100, bg_index=0, bg_offset=44003, volume=0.1))  # last window
    plan = np.stack(items)
    pcm, audio = aug.run(plan, return_audio=True)
    pcm, audio = pcm.cpu().numpy(), audio.cpu().numpy()
    saw_wrap = False
    for i, it in enumerate(plan):
        want = host_item(world["clips"], world["bg"], it)
        assert np.array_equal(audio[i], want), (i, it, np.abs(audio[i] - want).max())
        want_pcm = float_audio_to_int16_np(want)
        assert np.array_equal(pcm[i], want_pcm), (i, it)
        saw_wrap |= bool((want == 1.0).any())
    assert saw_wrap                                              # +1.0 -> -32768, like the reference's cast


def test_plan_validation(world):
    aug = world["aug"]
    with pytest.raises(ValueError):
        aug.run(np.stack([plan_item(MODE_CLIP, fg_index=12)]))
    with pytest.raises(ValueError):
        aug.run(np.stack([plan_item(MODE_SILENCE, bg_index=1, bg_offset=60008 - N + 1)]))
    with pytest.raises(ValueError):
        exact_int16(np.array([0.1], np.float32))
    assert aug.run(np.zeros(0, AUG_ITEM)).shape == (0, N)


def test_spec_mask(kws_lib):
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.uniform(1, 20, (6, 49, 40)).astype(np.float32)).cuda()
    bands = np.zeros((6, 8), np.int32)
    bands[1] = [3, 2, 0, 0, 0, 0, 0, 0]
    bands[2] = [0, 1, 38, 2, 47, 2, 0, 1]
    bands[3] = [0, 0, 0, 0, 10, 1, 11, 2]
    bands[5] = [39, 1, 5, 2, 48, 1, 0, 2]
    want = x.cpu().numpy().copy()
    for b in range(6):
        for s, n in (bands[b, 0:2], bands[b, 2:4]):
            want[b, :, s:s + n] = 0
        for s, n in (bands[b, 4:6], bands[b, 6:8]):
            want[b, s:s + n, :] = 0
    assert np.array_equal(spec_mask_(x, bands).cpu().numpy(), want)


def test_dataset_device_path_equals_host_path(kws_lib, tmp_path):
    """Same seed -> same random decisions -> identical batches, labels and spec-augment masks on both paths."""
    from multilingual_kws_b200.embedding import input_data
    pcm = synthetic_pcm(30, cfg_id=13)
    rng = np.random.default_rng(1)

    def wavs(sub, idx):
        d = tmp_path / sub
        d.mkdir()
        out = []
        for i in idx:
            input_data.encode_wav(str(d / f"c{i}.wav"), pcm[i].astype(np.float64) / 32768.0)
            out.append(str(d / f"c{i}.wav"))
        return out

    train, val, unk = wavs("hola", range(5)), wavs("hola_val", range(5, 11)), wavs("other", range(11, 30))
    bgd = tmp_path / "_background_noise_"
    bgd.mkdir()
    input_data.encode_wav(str(bgd / "a.wav"), rng.normal(0, 0.05, 40000))
    input_data.encode_wav(str(bgd / "b.wav"), rng.normal(0, 0.2, 17000))
    s = input_data.standard_microspeech_model_settings(3)
    out = {}
    for dev in (False, True):
        ds = input_data.AudioDataset(s, ["hola"], str(bgd), unk, unknown_percentage=40.0, silence_percentage=20.0, seed=11,
                                     device_augment=dev)
        tr = ds.init_single_target(-1, train, is_training=True).shuffle(1000).repeat().batch(32)
        it = iter(tr)
        batches = [next(it) for _ in range(4)]
        ev = list(ds.eval_with_silence_unknown(-1, val, label_from_parent_dir=False).batch(64))
        out[dev] = [(x.cpu().numpy(), y.cpu().numpy()) for x, y in batches + ev]
    labels_seen = set()
    for (xh, yh), (xd, yd) in zip(out[False], out[True]):
        assert np.array_equal(yh, yd)
        assert np.array_equal(xh, xd)
        labels_seen |= set(yh.tolist())
    assert labels_seen == {0, 1, 2}                              # silence, unknown and target branches all exercised


def test_batched_training_path_statistics(kws_lib, tmp_path):
    """device_augment="batched" (what transfer_learn uses): whole-batch decision draws.  Same policy as `augment`
    (reference input_data.py:275-304): branch frequencies, label / mode consistency, masks, determinism per seed."""
    from multilingual_kws_b200.embedding import input_data
    pcm = synthetic_pcm(30, cfg_id=13)
    rng = np.random.default_rng(1)

    def wavs(sub, idx):
        d = tmp_path / sub
        d.mkdir()
        out = []
        for i in idx:
            input_data.encode_wav(str(d / f"c{i}.wav"), pcm[i].astype(np.float64) / 32768.0)
            out.append(str(d / f"c{i}.wav"))
        return out

    train, unk = wavs("hola", range(5)), wavs("other", range(11, 30))
    bgd = tmp_path / "_background_noise_"
    bgd.mkdir()
    input_data.encode_wav(str(bgd / "a.wav"), rng.normal(0, 0.05, 40000))
    input_data.encode_wav(str(bgd / "b.wav"), rng.normal(0, 0.2, 17000))
    s = input_data.standard_microspeech_model_settings(3)

    def batches(seed, n):
        ds = input_data.AudioDataset(s, ["hola"], str(bgd), unk, unknown_percentage=40.0, silence_percentage=20.0, seed=seed,
                                     spec_aug_params=input_data.SpecAugParams(percentage=80), device_augment="batched")
        tr = ds.init_single_target(-1, train, is_training=True).shuffle(1000).repeat().batch(512)
        assert tr.fast is not None
        it = iter(tr)
        return [next(it) for _ in range(n)]

    got = batches(5, 8)
    x = torch.cat([b[0] for b in got]).cpu().numpy()
    y = torch.cat([b[1] for b in got]).numpy()
    assert x.shape == (4096, 49, 40, 1) and y.dtype == np.int64 and np.isfinite(x).all()
    frac = np.bincount(y, minlength=3) / y.size                  # [silence, unknown, target]
    assert abs(frac[0] - 0.20) < 0.03 and abs(frac[1] - 0.8 * 0.4) < 0.03 and abs(frac[2] - 0.8 * 0.6) < 0.03, frac
    # spec-augment: ~80 % of the clips get a mask draw; a masked column / row is exactly zero
    zero_cols = (np.abs(x[..., 0]).sum(axis=1) == 0).any(axis=1)
    zero_rows = (np.abs(x[..., 0]).sum(axis=2) == 0).any(axis=1)
    assert 0.3 < zero_cols.mean() < 0.8 and 0.3 < zero_rows.mean() < 0.9
    again = batches(5, 2)
    assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(got[:2], again))     # same seed -> same batches
    other = batches(6, 1)
    assert not torch.equal(other[0][1], got[0][1])
