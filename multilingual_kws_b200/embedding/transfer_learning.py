"""Drop-in mirror of reference ``multilingual_kws/embedding/transfer_learning.py`` on the B200 kernels.

``transfer_learn`` keeps the reference signature and return triple (ref :14-123).  Per step it does what
``xfer.fit`` does there with the embedding frozen (:43): frontend (fused kernel) -> embedding forward (tcgen05
GEMMs + depthwise/SE kernels, BN in inference mode) -> head forward/backward (one kernel) -> Adam (one kernel).
Under ``torch.distributed`` the global batch is sharded across ranks and the flat head gradient (+ loss/accuracy
scalars) is summed with ONE NCCL all-reduce per step.
"""
from __future__ import annotations

import csv
import glob
import logging
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import input_data
from ..fewshot import FewShotModel, Head
from ..model import EmbeddingModel

AUTOTUNE = -1   # placeholder for tf.data.experimental.AUTOTUNE in the reference's call sites


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shared_seed() -> Optional[int]:
    """Under torch.distributed every rank must start from the SAME head and draw the SAME global batches (each then
    computes its shard of them): rank 0 draws a seed and broadcasts it.  Single process: None (fresh entropy, like the
    reference's unseeded Keras initialisers / tf.data shuffles)."""
    dist = _dist()
    if dist is None:
        return None
    t = torch.zeros(1, dtype=torch.int64, device="cuda")
    if dist.get_rank() == 0:
        t[0] = int(np.random.SeedSequence().generate_state(1)[0])
    dist.broadcast(t, src=0)
    return int(t.item())


def train_step(model: FewShotModel, specs: torch.Tensor, labels: torch.Tensor, lr: float):
    """One optimisation step on a GLOBAL batch (every rank passes the same batch; each computes its shard).
    Returns (mean loss, accuracy) of the global batch as python floats."""
    dist = _dist()
    if dist is not None:
        r, ws = dist.get_rank(), dist.get_world_size()
        n = specs.shape[0]
        lo, hi = n * r // ws, n * (r + 1) // ws
        specs, labels = specs[lo:hi], labels[lo:hi]
    flat = model.head._flat
    if specs.shape[0] > 0:
        emb = model.embedding.forward_device(specs)
        model.head.grad(emb, labels, out=flat)
    else:
        flat.zero_()
    if dist is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)      # head grads + loss/acc scalars, 74 KB, NCCL over NVLink
    model.head.apply_adam(flat, lr)
    n_p = model.head.n_params
    stats = flat[n_p:n_p + 3].tolist()
    cnt = max(stats[2], 1.0)
    return stats[0] / cnt, stats[1] / cnt


def train_steps_grouped(model: FewShotModel, batches, lr: float):
    """`len(batches)` consecutive optimisation steps with ONE embedding forward.

    The embedding is frozen while the head trains (reference transfer_learning.py:43 `embedding.trainable = False`), so
    its outputs do not depend on the head's updates: the spectrograms of G consecutive steps go through the tower as one
    batch of G x batch_size clips (where the kernels are throughput- instead of launch-bound), then the G head steps
    (forward + loss + backward, all-reduce under torch.distributed, Adam) run in order on their slices.  Per-clip
    embeddings do not depend on how clips are batched, so every step computes exactly what `train_step` computes.
    batches: list of (specs [n,49,40(,1)] CUDA, labels [n] CUDA int); returns [(mean loss, accuracy)] per step, read
    from the device once for the whole group."""
    dist = _dist()
    shards = []
    for specs, labels in batches:
        if dist is not None:
            r, ws = dist.get_rank(), dist.get_world_size()
            n = specs.shape[0]
            lo, hi = n * r // ws, n * (r + 1) // ws
            specs, labels = specs[lo:hi], labels[lo:hi]
        shards.append((specs[..., 0] if specs.dim() == 4 else specs, labels))
    sizes = [s.shape[0] for s, _ in shards]
    emb_all = model.embedding.forward_device(torch.cat([s for s, _ in shards])) if sum(sizes) else None
    flat = model.head._flat
    n_p = model.head.n_params
    stats = torch.empty((len(shards), 3), dtype=torch.float32, device=flat.device)
    off = 0
    for i, ((_, labels), n) in enumerate(zip(shards, sizes)):
        if n > 0:
            model.head.grad(emb_all[off:off + n], labels, out=flat)
        else:
            flat.zero_()
        off += n
        if dist is not None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        stats[i].copy_(flat[n_p:n_p + 3])
        model.head.apply_adam(flat, lr)
    out = []
    for loss_sum, correct, cnt in stats.tolist():
        cnt = max(cnt, 1.0)
        out.append((loss_sum / cnt, correct / cnt))
    return out


def evaluate(model: FewShotModel, ds) -> Dict[str, float]:
    """Loss / accuracy over a (batched) dataset, forward only."""
    loss_sum, correct, n = 0.0, 0, 0
    for specs, labels in ds:
        probs = model.forward_device(specs.cuda())
        labels = labels.cuda()
        p = probs.gather(1, labels.view(-1, 1).long()).clamp_min(1e-30)
        loss_sum += float(-torch.log(p).sum())
        correct += int((probs.argmax(1) == labels).sum())
        n += labels.numel()
    return dict(loss=loss_sum / max(n, 1), accuracy=correct / max(n, 1))


class _TrainerView:
    """What `evaluate` needs of a model, served by a TailTrainer (the fine-tuned tail is not in the frozen model yet)."""

    def __init__(self, trainer):
        self.trainer = trainer

    def forward_device(self, specs: torch.Tensor) -> torch.Tensor:
        if specs.dim() == 4:
            specs = specs[..., 0]
        return self.trainer.predict_probs(specs.contiguous())


def train_step_embedding(trainer, specs: torch.Tensor, labels: torch.Tensor, lr: float):
    """Phase-2 step on a GLOBAL batch: each rank takes its shard, runs forward + backward through the trainable top of
    the embedding and the head, all-reduces the flat gradient buffer (TailTrainer.step) and applies Adam."""
    dist = _dist()
    if specs.dim() == 4:
        specs = specs[..., 0]
    if dist is not None:
        r, ws = dist.get_rank(), dist.get_world_size()
        n = specs.shape[0]
        lo, hi = n * r // ws, n * (r + 1) // ws
        specs, labels = specs[lo:hi], labels[lo:hi]
        if hi == lo:
            raise ValueError("phase-2 fine-tuning needs at least one clip per rank and step")
    return trainer.step(specs.contiguous(), labels, lr)


def fit(model: FewShotModel, train_ds, validation_data, steps_per_epoch: int, epochs: int, lr: float,
        csvlog_dest=None, verbose=1, group_clips: int = 4096, trainer=None) -> Dict[str, List[float]]:
    """Keras-style fit of the head.  Steps are executed in groups that share one embedding forward of about
    `group_clips` clips (`train_steps_grouped`; group_clips = 0: one forward per step) — same updates, same history.
    With `trainer` (a finetune.TailTrainer: phase 2, the top of the embedding trains too) every step is its own
    forward / backward pass."""
    if trainer is not None:
        group_clips = 0
        model = _TrainerView(trainer)
    history = {"loss": [], "accuracy": [], "val_loss": [], "val_accuracy": []}
    it = iter(train_ds)
    writer = None
    fh = None
    if csvlog_dest is not None:
        fh = open(csvlog_dest, "w", newline="")
        writer = csv.writer(fh)
        writer.writerow(["epoch", "accuracy", "loss", "val_accuracy", "val_loss"])   # Keras CSVLogger column order
    for epoch in range(epochs):
        loss_sum, acc_sum = 0.0, 0.0
        done = 0
        while done < steps_per_epoch:
            specs, labels = next(it)
            group = [(specs.cuda(), labels.cuda())]
            want = max(1, group_clips // max(int(specs.shape[0]), 1)) if group_clips > 0 else 1
            while len(group) < min(want, steps_per_epoch - done):
                specs, labels = next(it)
                group.append((specs.cuda(), labels.cuda()))
            results = ([train_step_embedding(trainer, *group[0], lr)] if trainer is not None
                       else train_steps_grouped(model, group, lr))
            for loss, acc in results:
                loss_sum += loss
                acc_sum += acc
            done += len(group)
        val = evaluate(model, validation_data)
        history["loss"].append(loss_sum / steps_per_epoch)
        history["accuracy"].append(acc_sum / steps_per_epoch)
        history["val_loss"].append(val["loss"])
        history["val_accuracy"].append(val["accuracy"])
        if verbose:
            print(f"Epoch {epoch + 1}/{epochs} - loss: {history['loss'][-1]:.4f} - accuracy: {history['accuracy'][-1]:.4f}"
                  f" - val_loss: {val['loss']:.4f} - val_accuracy: {val['accuracy']:.4f}", flush=True)
        if writer:
            writer.writerow([epoch, history["accuracy"][-1], history["loss"][-1], val["accuracy"], val["loss"]])
            fh.flush()
    if fh:
        fh.close()
    return history


def transfer_learn(
    target,
    train_files,
    val_files,
    unknown_files,
    num_epochs,
    num_batches,
    batch_size,
    primary_lr,
    backprop_into_embedding,
    embedding_lr,
    model_settings: Dict,
    base_model_path: os.PathLike,
    base_model_output: str,
    UNKNOWN_PERCENTAGE: float = 50.0,
    bg_datadir: os.PathLike = "/home/mark/tinyspeech_harvard/speech_commands/_background_noise_/",
    csvlog_dest: Optional[os.PathLike] = None,
    verbose=1,
):
    """this only works for single-target models: see audio_dataset and CATEGORIES"""
    embedding = EmbeddingModel.load(base_model_path, output_layer=base_model_output)
    embedding.trainable = False

    CATEGORIES = 3  # silence + unknown + target_keyword
    seed = shared_seed()          # identical head initialisation and batch draws on every rank (None: single process)
    xfer = FewShotModel(embedding, Head.keras_init(embedding.output_dim, 18, CATEGORIES, seed=seed))

    audio_dataset = input_data.AudioDataset(
        model_settings=model_settings,
        commands=[target],
        background_data_dir=bg_datadir,
        unknown_files=unknown_files,
        unknown_percentage=UNKNOWN_PERCENTAGE,
        spec_aug_params=input_data.SpecAugParams(percentage=80),
        seed=seed,
        device_augment="batched",  # clips decoded once into a device bank; shift / mix / masks run on the GPU, the
                                   # decisions of a whole batch are drawn at once (input_data._Dataset._fast_batches)
    )
    init_train_ds = audio_dataset.init_single_target(AUTOTUNE, train_files, is_training=True)
    init_val_ds = audio_dataset.init_single_target(AUTOTUNE, val_files, is_training=False)
    train_ds = init_train_ds.shuffle(buffer_size=1000).repeat().batch(batch_size)
    val_ds = init_val_ds.batch(batch_size)

    # steps_per_epoch = batch_size * num_batches, as the reference passes it (transfer_learning.py:89)
    history = fit(xfer, train_ds, val_ds, steps_per_epoch=batch_size * num_batches, epochs=num_epochs, lr=primary_lr,
                  csvlog_dest=csvlog_dest, verbose=verbose)
    if backprop_into_embedding:
        # Reference phase 2 (transfer_learning.py:97-112): "unfreeze the top 20 layers while leaving BatchNorm layers
        # frozen", recompile with Adam(embedding_lr) (fresh optimiser state), fit again.  The embedding's last 20 layers
        # are block7a, top_conv and the dense tower: finetune.TailTrainer (see its docstring for how the reference's
        # loop over the 3-layer Sequential differs from that stated intent).
        from ..finetune import TailTrainer
        xfer.head.reset_optimizer()
        trainer = TailTrainer(embedding, xfer.head)
        history = fit(xfer, train_ds, val_ds, steps_per_epoch=batch_size * num_batches, epochs=num_epochs, lr=embedding_lr,
                      csvlog_dest=csvlog_dest, verbose=verbose, trainer=trainer)
        tuned = EmbeddingModel(trainer.export_weights(), dtype=embedding.dtype)
        tuned.trainable = True
        xfer = FewShotModel(tuned, xfer.head)

    va = history["val_accuracy"][-1]
    name = f"xfer_epochs_{num_epochs}_bs_{batch_size}_nbs_{num_batches}_val_acc_{va:0.2f}_target_{target}"
    details = dict(num_epochs=num_epochs, batch_size=batch_size, num_batches=num_batches, val_accuracy=va, target=target)
    return name, xfer, details


# ---- batch evaluators (reference :177-273): one frontend launch + one predict per call
def _word_files(words_to_evaluate, data_dir, utterances_per_word):
    files = []
    for word in words_to_evaluate:
        wavs = glob.glob(data_dir + word + "/*.wav")
        if len(wavs) > utterances_per_word:
            fs = np.random.choice(wavs, utterances_per_word, replace=False)
        else:
            print("using all wavs for ", word)
            fs = wavs
        files.extend(fs)
    return files


def _split_confidences(preds, target_id):
    cols = np.argmax(preds, axis=1)
    conf = preds[np.arange(preds.shape[0]), cols]
    return dict(correct=conf[cols == target_id].tolist(), incorrect=conf[cols != target_id].tolist())


def evaluate_files_single_target(files_to_evaluate: List[os.PathLike], target_id: int, model, model_settings: Dict):
    specs = input_data.files2specs(model_settings, files_to_evaluate)
    preds = model.predict(np.expand_dims(specs, -1))
    return preds[:, target_id], preds


def evaluate_files_multiclass(files_to_evaluate: List[os.PathLike], target_id: int, model, model_settings: Dict):
    specs = input_data.files2specs(model_settings, files_to_evaluate)
    return _split_confidences(model.predict(np.expand_dims(specs, -1)), target_id)


def evaluate_fast_single_target(words_to_evaluate: List[str], target_id: int, data_dir: os.PathLike,
                                utterances_per_word: int, model, model_settings: Dict):
    return evaluate_files_single_target(_word_files(words_to_evaluate, data_dir, utterances_per_word), target_id, model,
                                        model_settings)


def evaluate_fast_multiclass(words_to_evaluate: List[str], target_id: int, data_dir: os.PathLike,
                             utterances_per_word: int, model, model_settings: Dict):
    return evaluate_files_multiclass(_word_files(words_to_evaluate, data_dir, utterances_per_word), target_id, model,
                                     model_settings)
