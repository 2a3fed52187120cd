"""GPU tests of the reference-shaped Python API: file2spec, AudioDataset batches, transfer_learn, evaluators,
save/load, streaming inference (frame reuse == per-window recompute) and the embedding extractor."""
import os

import numpy as np
import pytest
import torch

from oracle import effnet_oracle as EO
from oracle.frontend_oracle import FrontendOracle
from multilingual_kws_b200 import weights as W
from multilingual_kws_b200.synthetic import synthetic_pcm, synthetic_stream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(tmp_path_factory, kws_lib):
    from multilingual_kws_b200.embedding import input_data
    root = tmp_path_factory.mktemp("kws")
    rng = np.random.default_rng(0)
    pcm = synthetic_pcm(40, cfg_id=11)

    def wavs(sub, idx):
        d = root / sub
        d.mkdir(parents=True, exist_ok=True)
        out = []
        for i in idx:
            p = d / f"clip{i}.wav"
            input_data.encode_wav(str(p), pcm[i].astype(np.float64) / 32768.0)
            out.append(str(p))
        return out

    train, val, unk = wavs("tiempo", range(0, 5)), wavs("tiempo_val", range(5, 13)), wavs("other", range(13, 33))
    bg = root / "_background_noise_"
    bg.mkdir()
    input_data.encode_wav(str(bg / "noise.wav"), rng.normal(0, 0.05, 48000))
    feats = FrontendOracle().features(pcm, threads=4)
    w = W.random_init(1, randomize_bn=True, residual_gamma_scale=0.3)
    w["dense_3/kernel"] = rng.normal(0, 0.02, (1024, 761)).astype(np.float32)      # the classifier layer consumers cut away
    w["dense_3/bias"] = np.zeros(761, np.float32)
    EO.forward(w, feats, calibrate_bn=True)
    base = root / "base_model"
    base.mkdir()
    W.save_npz(str(base / "weights.npz"), w)
    return dict(root=root, pcm=pcm, feats=feats, train=train, val=val, unk=unk, bg=str(bg), base=str(base), w=w)


def test_file2spec_matches_oracle(world):
    from multilingual_kws_b200.embedding import input_data
    s = input_data.standard_microspeech_model_settings(3)
    spec = input_data.file2spec(s, world["train"][2])
    assert spec.shape == (49, 40) and spec.dtype == np.float32               # ipynb:1058
    assert np.array_equal(spec, world["feats"][2])
    specs = input_data.files2specs(s, world["val"])
    assert np.array_equal(specs, world["feats"][5:13])
    with pytest.raises(ValueError, match="audio is not a vector"):
        input_data.to_micro_spectrogram(s, np.zeros((2, 2, 16000), np.float32))


def test_embedding_extractor(world):
    from multilingual_kws_b200.embedding.distance_filtering import embedding_model
    emb = embedding_model(world["base"])                                     # cut at dense_2 (dense_3 dropped)
    assert emb.trainable is False and emb.output_dim == 1024
    out = emb.predict(world["feats"][:6, :, :, None])
    assert out.shape == (6, 1024)                                            # ipynb:1059 shape (N, 1024)
    w_cut = {k: v for k, v in world["w"].items() if not k.startswith("dense_3")}
    assert EO.cosine(out, EO.forward(w_cut, world["feats"][:6]).numpy()).min() >= 0.999
    with pytest.raises(ValueError, match="No such layer"):
        embedding_model(world["base"], base_model_output="dense_9")


def test_dataset_batches(world):
    from multilingual_kws_b200.embedding import input_data
    s = input_data.standard_microspeech_model_settings(3)
    ds = input_data.AudioDataset(s, ["tiempo"], world["bg"], world["unk"], unknown_percentage=50.0, seed=3)
    tr = ds.init_single_target(-1, world["train"], is_training=True).shuffle(1000).repeat().batch(16)
    specs, labels = next(iter(tr))
    assert specs.shape == (16, 49, 40, 1) and specs.is_cuda and labels.shape == (16,)
    assert set(labels.tolist()) <= {0, 1, 2}
    va = ds.init_single_target(-1, world["val"], is_training=False).batch(5)
    got = [(s_.shape[0], l.tolist()) for s_, l in va]
    assert [n for n, _ in got] == [5, 3] and all(l == [2] * n for n, l in got)
    first = next(iter(va))[0][..., 0].cpu().numpy()
    assert np.array_equal(first, world["feats"][5:10])                       # eval path is un-augmented & bit-exact
    ev = ds.eval_with_silence_unknown(-1, world["val"], label_from_parent_dir=False).batch(64)
    s_, l = next(iter(ev))
    assert s_.shape[0] == 8 + 0 + 4 and l.tolist()[-4:] == [1, 1, 1, 1]     # 10 % silence -> 0, 50 % unknown -> 4


def test_transfer_learn_end_to_end(world, tmp_path):
    from multilingual_kws_b200.embedding import input_data, transfer_learning
    from multilingual_kws_b200.fewshot import FewShotModel
    s = input_data.standard_microspeech_model_settings(3)
    csv = str(tmp_path / "log.csv")
    name, model, details = transfer_learning.transfer_learn(
        target="tiempo", train_files=world["train"], val_files=world["val"], unknown_files=world["unk"], num_epochs=2,
        num_batches=1, batch_size=8, primary_lr=0.001, backprop_into_embedding=False, embedding_lr=0, model_settings=s,
        base_model_path=world["base"], base_model_output="dense_2", UNKNOWN_PERCENTAGE=50.0, bg_datadir=world["bg"],
        csvlog_dest=csv, verbose=0)
    assert name.startswith("xfer_epochs_2_bs_8_nbs_1_val_acc_") and name.endswith("_target_tiempo")
    assert set(details) == {"num_epochs", "batch_size", "num_batches", "val_accuracy", "target"}
    assert model.head.step_count == 2 * 8                                    # steps_per_epoch = batch_size * num_batches
    lines = open(csv).read().strip().splitlines()
    assert lines[0] == "epoch,accuracy,loss,val_accuracy,val_loss" and len(lines) == 3
    preds = model.predict(world["feats"][:7, :, :, None])
    assert preds.shape == (7, 3) and np.allclose(preds.sum(1), 1, atol=1e-5)
    tgt, all_preds = transfer_learning.evaluate_files_single_target(world["val"], 2, model, s)
    assert tgt.shape == (8,) and np.array_equal(tgt, all_preds[:, 2])
    d = transfer_learning.evaluate_files_multiclass(world["val"], 2, model, s)
    assert len(d["correct"]) + len(d["incorrect"]) == 8
    out = str(tmp_path / "saved")
    model.save(out)
    again = FewShotModel.load(out)
    assert np.array_equal(again.predict(world["feats"][:7]), preds)
    # phase 2 (reference transfer_learning.py:97-112): the top 20 layers of the embedding train too, fresh Adam(embedding_lr)
    before = {k: v.copy() for k, v in model.embedding.weights.items()}
    name2, model2, details2 = transfer_learning.transfer_learn(
        "tiempo", world["train"], world["val"], world["unk"], 1, 1, 4, 1e-3, True, 1e-4, s, world["base"], "dense_2",
        bg_datadir=world["bg"], verbose=0)
    assert name2.startswith("xfer_epochs_1_bs_4_nbs_1_val_acc_") and model2.head.step_count == 4     # optimiser was reset
    w2 = model2.embedding.weights
    changed = [k for k in w2 if not np.array_equal(w2[k], before[k])]
    assert changed and all(k.startswith(("block7a_", "top_conv", "dense")) and "_bn/" not in k for k in changed), changed
    assert {"dense_2/kernel", "top_conv/kernel", "block7a_expand_conv/kernel", "block7a_se_reduce/bias"} <= set(changed)
    p2 = model2.predict(world["feats"][:7, :, :, None])
    assert p2.shape == (7, 3) and np.allclose(p2.sum(1), 1, atol=1e-5)


def test_training_learns_separable_classes(world):
    """Fine-tune on features whose labels are decodable from the embedding: accuracy must rise like the oracle's."""
    from multilingual_kws_b200.embedding.transfer_learning import train_step
    from multilingual_kws_b200.fewshot import FewShotModel, Head
    from multilingual_kws_b200.model import EmbeddingModel
    from oracle import head_oracle as HO
    emb_model = EmbeddingModel({k: v for k, v in world["w"].items() if not k.startswith("dense_3")})
    feats = torch.from_numpy(world["feats"]).cuda()
    labels = torch.from_numpy((np.arange(40) % 4 % 3).astype(np.int64)).cuda()       # clip kind -> class
    p = HO.init_head(5)
    model = FewShotModel(emb_model, Head.from_params(p))
    hist = [train_step(model, feats, labels, 1e-2) for _ in range(60)]
    emb = emb_model.forward_device(feats).cpu().numpy()
    want = HO.train(p, emb, labels.cpu().numpy(), 60, lr=1e-2)
    assert hist[-1][0] < hist[0][0] * 0.6 and hist[-1][1] >= 0.9
    assert abs(hist[-1][1] - want[-1][1]) <= 0.026 and abs(hist[-1][0] - want[-1][0]) < 0.02     # 1 clip of 40 = 2.5 pp


def test_streaming_matches_per_window_recompute(world, tmp_path):
    from multilingual_kws_b200.embedding import batch_streaming_analysis as sa
    from multilingual_kws_b200.embedding import input_data
    from multilingual_kws_b200.fewshot import FewShotModel, Head
    from multilingual_kws_b200.model import EmbeddingModel
    s = input_data.standard_microspeech_model_settings(3)
    T = 16000 * 4 + 777
    x = synthetic_stream(T, cfg_id=5)
    wav = str(tmp_path / "stream.wav")
    input_data.encode_wav(wav, x.astype(np.float64) / 32768.0)
    model = FewShotModel(EmbeddingModel({k: v for k, v in world["w"].items() if not k.startswith("dense_3")}),
                         Head.keras_init(1024, 18, 3, seed=1))
    gt = tmp_path / "gt.txt"
    gt.write_text("tiempo,1200\n_unknown_,2600\n")
    flags = [sa.StreamFlags(wav=wav, ground_truth=str(gt), target_keyword="tiempo", detection_thresholds=[0.3, 0.9],
                            clip_stride_ms=100)]
    results, inferences = sa.calculate_streaming_accuracy(model, s, flags)
    offsets = list(range(0, T - 16000, 1600))
    assert inferences.shape == (len(offsets), 3)
    wins = np.stack([x[o:o + 16000] for o in offsets])
    want = model.predict(FrontendOracle().features(wins, threads=4))          # reference: frontend per window, then predict
    assert np.array_equal(inferences, want)
    (fl, res), = results
    assert set(res) == {0.3, 0.9} and all(len(v) == 2 for v in res.values())
    r2, _ = sa.calculate_streaming_accuracy(model, s, flags, existing_inferences=inferences)
    assert r2[0][1] == res            # device post-processor (first call) == host recogniser (existing_inferences)


def test_host_pipeline_matches_direct_path(world):
    """EmbedPipeline.run_host (pinned buffers, overlapped copies, sub-batches) == frontend + embedding on the device."""
    from multilingual_kws_b200.frontend import MicroFrontend
    from multilingual_kws_b200.model import EmbeddingModel
    from multilingual_kws_b200.pipeline import EmbedPipeline
    fe = MicroFrontend()
    model = EmbeddingModel({k: v for k, v in world["w"].items() if not k.startswith("dense_3")})
    pcm = torch.from_numpy(world["pcm"])
    want = model.forward_device(fe.forward(pcm.cuda())).cpu()
    for sub in (7, 16, 64):
        pipe = EmbedPipeline(fe, model, n_samples=16000, sub_batch=sub)
        for _ in range(2):                                     # second call reuses buffers / graphs
            out = pipe.run_host(pcm.pin_memory())
            torch.cuda.synchronize()
            assert torch.equal(out, want)


def test_host_pipeline_back_to_back_calls(world):
    """run_host only enqueues: several calls in flight (slots recycled across calls, no host sync in between) must
    each deliver their own rows; join() orders the caller's stream after the last download."""
    from multilingual_kws_b200.frontend import MicroFrontend
    from multilingual_kws_b200.model import EmbeddingModel
    from multilingual_kws_b200.pipeline import EmbedPipeline
    fe = MicroFrontend()
    model = EmbeddingModel({k: v for k, v in world["w"].items() if not k.startswith("dense_3")})
    pcm = torch.from_numpy(world["pcm"])
    n = pcm.shape[0]
    inputs = [torch.roll(pcm, shifts=i, dims=0).contiguous() for i in range(7)]
    want = [model.forward_device(fe.forward(x.cuda())).cpu() for x in inputs]
    pinned = [x.pin_memory() for x in inputs]
    for sub, depth, streams in ((n, 2, 1), (n, 4, 2), (max(1, n // 3), 3, 3)):
        pipe = EmbedPipeline(fe, model, n_samples=16000, sub_batch=sub, depth=depth, streams=streams)
        if depth >= 3:                                          # write-combined upload buffers from the C ABI
            pinned = [pipe.alloc_input(n) for _ in inputs]
            for dst, src in zip(pinned, inputs):
                dst.copy_(src)
                assert dst.is_pinned() and dst.shape == src.shape and dst.dtype == torch.int16
        outs = [pipe.run_host(x) for x in pinned]
        pipe.join()
        done = torch.cuda.Event()
        done.record()
        done.synchronize()
        for o, w_ in zip(outs, want):
            assert torch.equal(o, w_)
        outs = [pipe.run_host(x) for x in pinned[:2]]
        pipe.synchronize()
        assert torch.equal(outs[0], want[0]) and torch.equal(outs[1], want[1])
    with pytest.raises(ValueError):
        pipe.run_host(torch.zeros((2, 100), dtype=torch.int16).pin_memory())


def test_grouped_steps_equal_single_steps(world):
    """train_steps_grouped (one embedding forward for G steps; the embedding is frozen) == G x train_step, bit for bit."""
    from multilingual_kws_b200.embedding.transfer_learning import train_step, train_steps_grouped
    from multilingual_kws_b200.fewshot import FewShotModel, Head
    from multilingual_kws_b200.model import EmbeddingModel
    from oracle import head_oracle as HO
    emb_model = EmbeddingModel({k: v for k, v in world["w"].items() if not k.startswith("dense_3")})
    feats = torch.from_numpy(world["feats"]).cuda()
    rng = np.random.default_rng(2)
    batches = []
    for n in (16, 16, 7, 16, 1, 16):
        idx = torch.from_numpy(rng.integers(0, 40, n)).cuda()
        batches.append((feats[idx][..., None], torch.from_numpy(rng.integers(0, 3, n)).cuda()))
    a = FewShotModel(emb_model, Head.from_params(HO.init_head(7)))
    b = FewShotModel(emb_model, Head.from_params(HO.init_head(7)))
    want = [train_step(a, x[..., 0], y, 1e-2) for x, y in batches]
    got = train_steps_grouped(b, batches[:4], 1e-2) + train_steps_grouped(b, batches[4:], 1e-2)
    assert got == want
    assert np.array_equal(a.head.get_params(), b.head.get_params())


def test_streaming_sharded_over_two_gpus():
    """Window-sharded streaming inference under torchrun (2 ranks, NCCL all-gather) == the single-process result."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dist_stream_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "DIST_STREAM_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
