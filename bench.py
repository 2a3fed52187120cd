#!/usr/bin/env python
"""bench.py — throughput of the hot path (1 s / 16 kHz PCM -> log-mel frontend -> EfficientNet-B0 embedding) on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N ...            the reference algorithm on the host CPU cores
                                                           (oracle port: C frontend + torch-CPU fp32 network)

A "step" is one pass of the hot path over one batch of synthetic clips (BASELINE.json configs[1]: batch = 1024
clips per GPU; weak scaling, so N GPUs = config 4's 8192 at N = 8).  Rank 0 prints ONE JSON line.
  value     utterances/s, inputs (int16 PCM) resident in HBM when the timed region starts, whole job over N GPUs
  e2e       same metric through the public Python API with HOST buffers: pinned H2D of the PCM + frontend +
            embedding + D2H of the embeddings inside the timed region
  roofline  for the kernel with the largest share of the step (per-op CUDA events on the launch stream)
  cpu_baseline  the oracle restatement of the reference algorithm timed on this box's host cores (rank 0, N = 1)
  finetune  the 5-shot head fine-tune step (embedding forward + head fwd/bwd + one NCCL all-reduce + Adam)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "utterances/sec (1s@16kHz) embed-fwd"
UNIT = "utterances/s"


def workload(batch: int) -> str:
    """config.workload: the SAME string in both arms (the driver compares them)."""
    return f"configs[1]: log-mel frontend + EfficientNet-B0 embedding forward, batch {batch} x 1 s @ 16 kHz clips per GPU"
FRONTEND_BYTES_PER_CLIP = 39840          # 32 000 B int16 PCM in + 49*40*4 B features out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def host_threads() -> int:
    """Host threads this process may really use: the affinity mask, capped by the cgroup CPU quota (a container that
    sees 128 CPUs but owns a 16-CPU quota is slower with 128 threads than with 16)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = int(q) / int(per)
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    if quota:
        n = min(n, max(1, int(quota + 0.999)))
    return max(1, n)


def bounded_cpu_sample(step_fn, max_clips: int, budget_s: float, n_steps: int, probe: int = 32) -> int:
    """Clips per CPU step such that n_steps steps take about budget_s: times a small probe first."""
    probe = min(probe, max_clips)
    step_fn(probe)                                   # page-in / thread-pool start-up
    t0 = time.perf_counter()
    step_fn(probe)
    per_clip = (time.perf_counter() - t0) / probe
    n = int(budget_s / max(n_steps, 1) / max(per_clip, 1e-9))
    return int(max(probe, min(max_clips, n)))


def load_traffic():
    """DRAM bytes per launch of each kernel, measured by ncu (tools/traffic_from_ncu.py -> profiles/*_traffic.json).
    bench.py never runs under a profiler; it only quotes the committed capture, and says which."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        return json.load(f), os.path.relpath(files[-1], ROOT)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's algorithm on the host CPU: C frontend oracle (all cores) + torch-CPU fp32 EfficientNet.
    TensorFlow itself cannot be installed offline, so this is kind "port"."""
    if rank != 0:
        return
    import torch
    from multilingual_kws_b200 import weights as W
    from multilingual_kws_b200.synthetic import synthetic_pcm
    from oracle import effnet_oracle as EO
    from oracle.frontend_oracle import FrontendOracle
    cores = host_threads()
    torch.set_num_threads(cores)
    pcm_all = synthetic_pcm(min(args.batch, args.ref_sample), cfg_id=2)
    w = W.random_init(0, randomize_bn=True)
    orc = FrontendOracle()

    def run(n):
        feats = orc.features(pcm_all[:n], threads=cores)
        with torch.no_grad():
            return EO.forward(w, feats)

    # bounded sample: the whole --steps/--warmup run has to end within a few minutes on any host
    B = bounded_cpu_sample(run, pcm_all.shape[0], args.ref_budget_s, args.steps + args.warmup)
    for _ in range(args.warmup):
        run(B)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(B)
    dt = (time.perf_counter() - t0) / args.steps
    v = B / dt
    sample = f"{B} synthetic clips per step (of the {args.batch}-clip workload), frontend + embedding, {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32 (CPU)", "data": "synthetic",
        "config": {"workload": workload(args.batch), "sample_clips_per_step": B},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from multilingual_kws_b200 import build as kbuild
    if rank == 0:
        kbuild.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    from multilingual_kws_b200 import weights as W
    from multilingual_kws_b200.fewshot import FewShotModel, Head
    from multilingual_kws_b200.frontend import FEATURE_SCALE, MicroFrontend
    from multilingual_kws_b200.model import EmbeddingModel
    from multilingual_kws_b200.synthetic import synthetic_pcm
    from multilingual_kws_b200.embedding.transfer_learning import train_step

    dev = torch.device("cuda", local_rank)
    B = args.batch
    fe = MicroFrontend()
    weights = W.random_init(0, randomize_bn=True, residual_gamma_scale=0.3)
    emb_model = EmbeddingModel(weights, chunk=args.chunk, dtype=args.dtype)
    emb_model.set_chunk_late(args.chunk_late)
    if args.no_graph:
        emb_model.set_graph(False)      # plain launches (needed under ncu: kernels inside a graph capture cannot be profiled)
    base = synthetic_pcm(min(B, 256), cfg_id=2 + rank)
    pcm_host = np.tile(base, (-(-B // base.shape[0]), 1))[:B]
    pcm = torch.from_numpy(pcm_host).to(dev)
    feats = torch.empty((B, 49, 40), dtype=torch.float32, device=dev)
    emb = torch.empty((B, emb_model.output_dim), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def step():
        fe.forward(pcm, out=feats)
        emb_model.forward_device(feats, out=emb)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()                                                  # L2 flush between timed iterations
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            evs.append((s, e))
        barrier()
        wall = time.perf_counter() - t0
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps, wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, wall = timed(step, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms_step * 1e-3)

    # ---- the same device-resident work with consecutive steps on alternating streams (what the host pipeline does):
    # late layers are latency-bound and leave SMs idle, early layers are throughput-bound, neighbouring steps fill
    # each other's gaps.  Region timing: first launch -> last kernel, L2 flushes included, / K.
    n_str = args.pipe_streams
    ov_streams = [torch.cuda.Stream(device=dev) for _ in range(n_str)]
    ov_bufs = [(torch.empty_like(feats), torch.empty_like(emb),
                torch.empty(emb_model.workspace_bytes(B), dtype=torch.uint8, device=dev)) for _ in range(n_str)]

    from multilingual_kws_b200.pipeline import EmbedPipeline
    ov_budget = EmbedPipeline.SM_BUDGET if n_str > 1 else None        # the schedule the host pipeline uses
    if args.sm_budget:
        ov_budget = tuple(int(x) for x in args.sm_budget.split(","))

    def ov_step(k):
        f, o, w_ = ov_bufs[k % n_str]
        st = ov_streams[k % n_str]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            fe.forward(pcm, out=f)
            emb_model.forward_device(f, out=o, workspace=w_, sm_budget=ov_budget)

    def ov_join():
        for st in ov_streams:
            torch.cuda.current_stream().wait_stream(st)

    ms_overlap = None
    if not args.no_graph:
        for k in range(max(args.warmup, 3) * n_str):
            ov_step(k)
        ov_join()
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for k in range(args.steps):
            flush.zero_()
            ov_step(k)
        ov_join()
        e_ev.record()
        barrier()
        ms_overlap = s_ev.elapsed_time(e_ev) / args.steps
        if world > 1:
            t = torch.tensor([ms_overlap], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_overlap = float(t.item())

    # ---- e2e: public API, host buffers (pinned), H2D + D2H inside the timed region
    from multilingual_kws_b200.pipeline import EmbedPipeline
    sub_b = args.pipe_sub_batch or B
    pipe = EmbedPipeline(fe, emb_model, n_samples=16000, sub_batch=sub_b, depth=2 * args.pipe_streams * max(1, B // sub_b),
                         streams=args.pipe_streams, sm_budget=ov_budget)
    # host buffers: two input sets in write-combined pinned memory (kws_host_alloc), two pinned result buffers
    pcm_pinned2 = [pipe.alloc_input(B), pipe.alloc_input(B)]
    pcm_pinned2[0].copy_(torch.from_numpy(pcm_host))
    pcm_pinned2[1].copy_(torch.from_numpy(np.ascontiguousarray(pcm_host[::-1])))
    emb_pinned2 = [pipe.alloc_output(B), pipe.alloc_output(B)]
    pcm_pinned, emb_pinned = pcm_pinned2[0], emb_pinned2[0]

    def timed_e2e(steps, warmup):
        """K steps through EmbedPipeline.run_host, each with its own pinned H2D (B x 32 000 B) and D2H (B x 4 096 B).
        run_host only enqueues, so the upload of step i+1 and the download of step i-1 overlap the kernels of
        step i; the region is timed as a whole (first enqueue -> last download) and divided by K."""
        for i in range(warmup):
            pipe.run_host(pcm_pinned2[i & 1], emb_pinned2[i & 1])
        pipe.join()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            flush.zero_()                                                  # L2 flush, inside the timed region
            pipe.run_host(pcm_pinned2[i & 1], emb_pinned2[i & 1])
        pipe.join()
        e.record()
        barrier()
        total_ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps

    # the K-step region is short (tens of ms) and includes pipeline fill / drain and host enqueue jitter: it is timed
    # three times (K steps each, warm-up before each) and the median region is reported, all three are listed
    e2e_regions = [timed_e2e(args.steps, max(args.warmup, 3)) for _ in range(3)]
    ms_e2e = float(np.median(e2e_regions))

    # one blocking call at a time (host waits for the rows before it submits the next batch): wall clock per call
    lat = []
    for i in range(max(args.warmup, 3) + args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.run_host(pcm_pinned, emb_pinned)
        pipe.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    ms_e2e_sync = float(np.median(lat[max(args.warmup, 3):]))
    e2e_value = world * B / (ms_e2e * 1e-3)

    # ---- the e2e ceiling set by the host -> device fabric: the same pinned PCM buffers copied by every rank at the same
    # time, nothing else running (32 000 B per clip have to cross it; at N = 8 all ranks share one host memory system)
    h2d_dst = torch.empty((B, 16000), dtype=torch.int16, device=dev)
    for _ in range(3):
        h2d_dst.copy_(pcm_pinned2[0], non_blocking=True)
    barrier()
    s_h, e_h = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_h.record()
    for i in range(20):
        h2d_dst.copy_(pcm_pinned2[i & 1], non_blocking=True)
    e_h.record()
    barrier()
    ms_h2d = s_h.elapsed_time(e_h) / 20
    if world > 1:
        t = torch.tensor([ms_h2d], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_h2d = float(t.item())
    h2d_gbps = B * 32000 / (ms_h2d * 1e-3) / 1e9
    del h2d_dst

    # ---- fine-tune step (BASELINE config 3 shape): batch 512 / GPU, embedding fwd + head fwd/bwd + all-reduce + Adam
    ft_B = 512
    ft_model = FewShotModel(emb_model, Head.keras_init(emb_model.output_dim, 18, 3, seed=0))
    ft_specs = feats[:ft_B].clone()
    ft_labels = torch.randint(0, 3, (ft_B * world,), device=dev, dtype=torch.int32)

    def step_ft():
        # every rank owns its shard already (global batch = world * ft_B): emulate train_step's sharded path
        e = emb_model.forward_device(ft_specs)
        flat = ft_model.head.grad(e, ft_labels[rank * ft_B:(rank + 1) * ft_B])
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        ft_model.head.apply_adam(flat, 1e-3)

    ms_ft, _ = timed(step_ft, args.steps, max(args.warmup, 3))

    # the same steps the way `fit` runs them: the embedding is frozen, so G consecutive steps share ONE embedding
    # forward of G x 512 clips (transfer_learning.train_steps_grouped); timed per group, reported per step
    ft_G = 8
    ft_specs_g = feats[:ft_B].repeat(ft_G, 1, 1)[:ft_G * ft_B].contiguous() if B >= ft_B else None
    ft_emb_g = torch.empty((ft_G * ft_B, emb_model.output_dim), dtype=torch.float32, device=dev)

    def group_ft():
        emb_model.forward_device(ft_specs_g, out=ft_emb_g)
        for g in range(ft_G):
            flat = ft_model.head.grad(ft_emb_g[g * ft_B:(g + 1) * ft_B], ft_labels[rank * ft_B:(rank + 1) * ft_B])
            if world > 1:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            ft_model.head.apply_adam(flat, 1e-3)

    ms_ft_group = None
    if ft_specs_g is not None:
        ms_ft_group, _ = timed(group_ft, max(args.steps // ft_G, 3), 3)

    # ---- phase 2 of transfer_learn (BASELINE config 3: "head + last block"): forward + backward through block7a, the top
    # conv and the dense tower + head, ONE all-reduce of the flat fp32 gradient buffer (~10.1 M floats), Adam
    from multilingual_kws_b200.finetune import TailTrainer
    ft2_head = Head.keras_init(emb_model.output_dim, 18, 3, seed=1)
    trainer = TailTrainer(emb_model, ft2_head)
    ft2_labels = ft_labels[rank * ft_B:(rank + 1) * ft_B]

    def step_ft2_eager():
        trainer.forward_tail(trainer.tail_input(ft_specs), keep=True)
        trainer.backward(ft2_labels)
        if world > 1:
            dist.all_reduce(trainer.flat, op=dist.ReduceOp.SUM)
        trainer.apply_adam(1e-4)

    ms_ft2_eager, _ = timed(step_ft2_eager, max(args.steps // 2, 5), 3)
    # the same step the way TailTrainer.step runs it: captured once into a CUDA graph (all-reduce included), replayed;
    # the returned loss / accuracy are read back every step as Keras' fit loop does
    ms_ft2, _ = timed(lambda: trainer.step(ft_specs, ft2_labels, 1e-4), max(args.steps // 2, 5), 3)
    ms_ar = None
    if world > 1:
        ms_ar, _ = timed(lambda: dist.all_reduce(trainer.flat, op=dist.ReduceOp.SUM), 10, 3)

    # ---- N > 1: asserted parity of the multi-rank paths against the same work done by one rank (the driver's pytest box
    # has one GPU, so this is where the NCCL paths are checked under the driver); any mismatch fails the run
    parity = None
    if world > 1:
        from multilingual_kws_b200.embedding.batch_streaming_analysis import stream_inferences
        from multilingual_kws_b200.embedding.input_data import standard_microspeech_model_settings
        from multilingual_kws_b200.embedding.transfer_learning import train_step_embedding, train_steps_grouped
        from multilingual_kws_b200.synthetic import synthetic_stream
        g = torch.Generator(device="cpu").manual_seed(77)
        nb_ = 8 * world
        p_specs = fe.forward(torch.from_numpy(synthetic_pcm(nb_, cfg_id=9)).to(dev))          # same global batch on every rank
        p_labels = torch.randint(0, 3, (nb_,), generator=g, dtype=torch.int32).to(dev)
        # (1) head fine-tune, 10 grouped steps: sharded + all-reduce vs the whole batch on this rank alone
        m_dist = FewShotModel(emb_model, Head.keras_init(emb_model.output_dim, 18, 3, seed=5))
        m_solo = FewShotModel(emb_model, Head.keras_init(emb_model.output_dim, 18, 3, seed=5))
        train_steps_grouped(m_dist, [(p_specs, p_labels)] * 10, 1e-3)
        e_all = emb_model.forward_device(p_specs)
        for _ in range(10):
            m_solo.head.apply_adam(m_solo.head.grad(e_all, p_labels), 1e-3)
        a, b = m_dist.head.get_params(), m_solo.head.get_params()
        head_diff = float(np.abs(a - b).max())
        # Adam divides the gradient by its own magnitude, so equal parameters do not prove an equal gradient SCALE:
        # compare the all-reduced gradient sums themselves (shards + NCCL sum vs the whole batch on this rank)
        lo_, hi_ = nb_ * rank // world, nb_ * (rank + 1) // world
        g_solo = m_solo.head.grad(e_all, p_labels).clone()
        g_dist = m_solo.head.grad(e_all[lo_:hi_].contiguous(), p_labels[lo_:hi_].contiguous()).clone()
        dist.all_reduce(g_dist, op=dist.ReduceOp.SUM)
        head_grad_rel = float((g_dist - g_solo).norm() / g_solo.norm())
        # (2) phase-2 steps: sharded + 40 MB all-reduce vs the whole batch on this rank alone
        t_dist = TailTrainer(emb_model, Head.keras_init(emb_model.output_dim, 18, 3, seed=6))
        t_solo = TailTrainer(emb_model, Head.keras_init(emb_model.output_dim, 18, 3, seed=6))
        t_dist.forward_tail(t_dist.tail_input(p_specs[lo_:hi_].contiguous()), keep=True)
        t_dist.backward(p_labels[lo_:hi_].contiguous())
        dist.all_reduce(t_dist.flat, op=dist.ReduceOp.SUM)
        t_solo.forward_tail(t_solo.tail_input(p_specs), keep=True)
        t_solo.backward(p_labels)
        tail_grad_rel = float((t_dist.flat - t_solo.flat).norm() / t_solo.flat.norm())      # 10 058 110 gradient sums
        for _ in range(3):
            train_step_embedding(t_dist, p_specs, p_labels, 1e-4)
            t_solo.forward_tail(t_solo.tail_input(p_specs), keep=True)
            t_solo.backward(p_labels)
            t_solo.apply_adam(1e-4)
        tail_rel = 0.0
        for pd, ps in zip(t_dist.params, t_solo.params):
            tail_rel = max(tail_rel, float((pd.master - ps.master).norm()) / max(float(ps.master.norm()), 1e-12))
        # (3) streaming: window range sharded over the ranks + all-gather vs unsharded
        settings = standard_microspeech_model_settings(3)
        audio = synthetic_stream(16000 * 20, cfg_id=5).astype(np.float32) / np.float32(32768)
        rows_dist = stream_inferences(m_solo, settings, audio, 16000, 1000, 100)
        import multilingual_kws_b200.embedding.batch_streaming_analysis as bsa
        saved = bsa._dist
        bsa._dist = lambda: None                       # the same call, un-sharded, on this rank
        rows_solo = stream_inferences(m_solo, settings, audio, 16000, 1000, 100)
        bsa._dist = saved
        stream_equal = bool(np.array_equal(rows_dist, rows_solo))
        ok = head_diff < 1e-5 and tail_rel < 1e-3 and stream_equal and head_grad_rel < 1e-5 and tail_grad_rel < 1e-4
        flags = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        parity = {"finetune_parity_vs_single_rank": bool(head_diff < 1e-5), "head_param_max_abs_diff_after_10_steps": head_diff,
                  "head_gradient_sums_rel_diff_after_allreduce": head_grad_rel,
                  "phase2_parity_vs_single_rank": bool(tail_rel < 1e-3), "phase2_param_max_rel_diff_after_3_steps": tail_rel,
                  "phase2_gradient_sums_rel_diff_after_allreduce": tail_grad_rel,
                  "streaming_sharded_equals_unsharded": stream_equal, "all_ranks_ok": bool(flags.item() == 1.0),
                  "note": "sums over shards are added by NCCL in a different order than one rank adds them: equal to fp32 rounding"}
        if flags.item() != 1.0:
            if rank == 0:
                print(json.dumps({"error": "multi-rank parity check failed", "parity": parity}), file=sys.stderr)
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(3)

    # ---- the other BASELINE configurations as extra keys (driver-run): config 1 (frontend only, batch 32) and config 5
    # (30 min of audio, 1 s window, the reference's default 20 ms hop and BASELINE's 100 ms hop)
    extra = None
    if rank == 0:
        from multilingual_kws_b200.synthetic import synthetic_stream
        from multilingual_kws_b200.frontend import FEATURE_SCALE as FS
        extra = {}
        pcm32 = torch.from_numpy(synthetic_pcm(32, cfg_id=1)).to(dev)
        out32 = torch.empty((32, 49, 40), dtype=torch.float32, device=dev)
        ms32, _ = timed(lambda: fe.forward(pcm32, out=out32), 20, 3) if world == 1 else (None, None)
        if ms32:
            extra["config1_frontend_only_batch32"] = {"ms": ms32, "clips_per_s": 32 / ms32 * 1e3,
                                                      "note": "one launch, 32 CTAs: latency-bound (bit-exact vs the oracle: tests)"}
        if world == 1:
            n_s = 30 * 60 * 16000
            audio_d = torch.from_numpy(synthetic_stream(n_s, cfg_id=5)).to(dev)
            rowsx = {}
            for hop_ms, hop in ((100, 1600), (20, 320)):
                Wn = fe.stream_num_windows(n_s, 16000, hop)

                def offline():
                    st_ = fe.stream_prepare(audio_d)
                    return [ft_model.forward_device(st_.windows(16000, hop, w0, min(4096, Wn - w0), FS)) for w0 in range(0, Wn, 4096)]
                offline()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pr_ = torch.cat(offline()).cpu()
                wall_ = time.perf_counter() - t0
                rowsx[f"hop_{hop_ms}ms"] = {"windows": int(Wn), "wall_s": wall_, "windows_per_s": Wn / wall_,
                                            "realtime_factor": n_s / 16000 / wall_}
            st_ = fe.stream_prepare(audio_d)
            lat_ = []
            for i in range(1, 251):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ft_model.forward_device(st_.windows(16000, 1600, i, 1, FS)).cpu()
                lat_.append((time.perf_counter() - t0) * 1e3)
            rowsx["online_one_window_per_100ms_hop"] = {"p50_ms": float(np.percentile(lat_[50:], 50)), "p99_ms": float(np.percentile(lat_[50:], 99))}
            extra["config5_streaming_30min"] = rowsx
            del audio_d

    # ---- fine-tuning END TO END through the public API (dataset from WAV files, device augmentation, frontend, embedding,
    # head): transfer_learn() at the reference's defaults (run.py:219-223: 4 epochs x 64 steps x batch 64 = 16 384 augmented
    # clips) and fit() at BASELINE config 3's shape (batch 512 x 100 steps), wall clock
    ft_e2e = None
    if rank == 0 and world == 1 and not args.no_finetune_e2e:
        import shutil
        import tempfile
        from multilingual_kws_b200.embedding import input_data, transfer_learning
        tmp = tempfile.mkdtemp(prefix="kws_ft_")
        try:
            rng_ = np.random.default_rng(0)
            clips = synthetic_pcm(96, cfg_id=11).astype(np.float64) / 32768.0

            def wavs(sub, idx):
                os.makedirs(os.path.join(tmp, sub), exist_ok=True)
                out_ = []
                for i in idx:
                    p_ = os.path.join(tmp, sub, f"clip{i}.wav")
                    input_data.encode_wav(p_, clips[i])
                    out_.append(p_)
                return out_
            train_f, val_f, unk_f = wavs("kw", range(0, 5)), wavs("kw_val", range(5, 21)), wavs("other", range(21, 96))
            os.makedirs(os.path.join(tmp, "_background_noise_"))
            input_data.encode_wav(os.path.join(tmp, "_background_noise_", "noise.wav"), rng_.normal(0, 0.05, 16000 * 60))
            os.makedirs(os.path.join(tmp, "base"))
            W.save_npz(os.path.join(tmp, "base", "weights.npz"), weights)
            settings_ = input_data.standard_microspeech_model_settings(3)
            kw = dict(target="kw", train_files=train_f, val_files=val_f, unknown_files=unk_f, num_epochs=4, num_batches=1,
                      batch_size=64, primary_lr=1e-3, backprop_into_embedding=False, embedding_lr=0, model_settings=settings_,
                      base_model_path=os.path.join(tmp, "base"), base_model_output="dense_2", UNKNOWN_PERCENTAGE=50.0,
                      bg_datadir=os.path.join(tmp, "_background_noise_") + "/", verbose=0)
            transfer_learning.transfer_learn(**dict(kw, num_epochs=1))           # warm-up: library load, graphs, allocator
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, m_, det_ = transfer_learning.transfer_learn(**kw)
            torch.cuda.synchronize()
            wall_tl = time.perf_counter() - t0
            # config 3 shape through the same dataset pipeline: batch 512, 100 steps (fit on a prepared model)
            ds = input_data.AudioDataset(model_settings=settings_, commands=["kw"], background_data_dir=kw["bg_datadir"],
                                         unknown_files=unk_f, unknown_percentage=50.0,
                                         spec_aug_params=input_data.SpecAugParams(percentage=80), seed=1, device_augment="batched")
            tr_ds = ds.init_single_target(-1, train_f, is_training=True).shuffle(buffer_size=1000).repeat().batch(512)
            va_ds = ds.init_single_target(-1, val_f, is_training=False).batch(512)
            transfer_learning.fit(m_, tr_ds, va_ds, steps_per_epoch=8, epochs=1, lr=1e-3, verbose=0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            transfer_learning.fit(m_, tr_ds, va_ds, steps_per_epoch=100, epochs=1, lr=1e-3, verbose=0)
            torch.cuda.synchronize()
            wall_c3 = time.perf_counter() - t0
            # the dataset pipeline alone (element generation on the host + augmentation kernel + frontend), no model
            it_ = iter(tr_ds)
            next(it_)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                next(it_)
            torch.cuda.synchronize()
            wall_ds = (time.perf_counter() - t0) / 20
            ft_e2e = {
                "transfer_learn_reference_defaults": {"epochs": 4, "steps": 256, "batch": 64, "clips": 16384, "wall_s": wall_tl,
                                                       "utt_per_s": 16384 / wall_tl, "val_accuracy": det_["val_accuracy"],
                                                       "includes": "model load from disk, WAV decode, dataset, device augmentation, frontend, "
                                                                   "embedding, head steps, validation pass per epoch"},
                "config3_batch512_x100_steps": {"wall_s": wall_c3, "ms_per_step": wall_c3 * 10, "utt_per_s": 51200 / wall_c3,
                                                "dataset_pipeline_alone_ms_per_batch": wall_ds * 1e3,
                                                "device_only_ms_per_step": None if ms_ft_group is None else ms_ft_group / ft_G,
                                                "note": "fit() over AudioDataset batches (vectorised decision draws -> plan items -> kws_augment_pcm -> "
                                                        "frontend -> grouped embedding forward -> head steps); device_only = finetune.grouped"},
            }
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    # ---- per-kernel shares (CUDA events around every launch, on the launch stream) -> roofline of the dominant kernel
    peaks = load_peaks()
    roof, shares = None, None
    if rank == 0:
        op_ms = np.zeros(emb_model.n_ops, np.float64)
        fe_ms = 0.0
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fe.forward(pcm, out=feats); e.record()
            _, ms = emb_model.forward_timed(feats)
            op_ms += ms
            fe_ms += s.elapsed_time(e)
        op_ms /= reps
        fe_ms /= reps
        info = emb_model.op_info()
        kinds = {0: "stem_conv_kernel", 1: "gemm_tcgen05_kernel", 2: "dwse_kernel"}
        agg = {"frontend_clip_kernel": dict(ms=fe_ms, flops=0.0, bytes=FRONTEND_BYTES_PER_CLIP * B, launches=1)}
        per_op_launches = emb_model.launches(B) / emb_model.n_ops      # average (early ops run per small chunk)
        for (name, kind, fl, by, n, k, r), ms in zip(info, op_ms):
            a = agg.setdefault(kinds[kind], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            a["ms"] += float(ms); a["flops"] += fl * B; a["bytes"] += by * B; a["launches"] += per_op_launches
        total = sum(a["ms"] for a in agg.values())
        shares = {k: round(a["ms"] / total, 4) for k, a in agg.items()}
        top = max(agg, key=lambda k: agg[k]["ms"])
        a = agg[top]
        hbm_gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9
        roof = dict(kernel=top, bound="hbm", achieved=hbm_gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=hbm_gbs / peaks["hbm_gbs"],
                    traffic=None, launches_per_step=a["launches"], avg_launch_ms=a["ms"] / a["launches"],
                    peak_source=peaks["source"],
                    note="all launches of the kernel in one step: sum of algorithmic bytes (operands + results of every "
                         "launch, activations in 16 bit) / sum of launch durations")
        if top == "gemm_tcgen05_kernel":
            # the dense contractions have two roofs; the binding one is the roof with the larger minimum time.  With
            # K = 16 ... 1152 and M = batch x pixels the activation traffic, not the tensor pipe, is the tighter bound.
            tf = a["flops"] / (a["ms"] * 1e-3) / 1e12
            tpeak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
            tensor = dict(bound="tensor", achieved=tf, peak=tpeak, unit="TFLOP/s", frac=tf / tpeak)
            if tensor["frac"] > roof["frac"]:
                roof.update(tensor)
                roof["other_roof"] = dict(bound="hbm", achieved=hbm_gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=hbm_gbs / peaks["hbm_gbs"])
            else:
                roof["other_roof"] = tensor
        # per-op events are taken with plain launches (no graph, no overlap of a kernel's prologue with its predecessor),
        # so they add up to more than the graph-replayed step; the in-step figures scale them to the timed step
        roof["per_kernel_ms"] = {k: round(v["ms"], 4) for k, v in agg.items()}
        roof["per_kernel_ms_in_step"] = {k: round(v["ms"] / total * ms_step, 4) for k, v in agg.items()}
        roof["per_kernel_note"] = ("per_kernel_ms: per-op CUDA events, plain launches (sum > ms_per_step); per_kernel_ms_in_step: the "
                                   "same shares of the graph-replayed, timed step (sum = ms_per_step)")
        tr, tr_file = load_traffic()
        if tr and top in tr["kernels"] and tr.get("batch") == B:
            t_ = tr["kernels"][top]
            roof["traffic"] = t_["dram_bytes_per_launch"]
            roof["traffic_note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the {t_['launches']} launches of "
                                    f"one forward pass at batch {B}, from the committed ncu capture {tr_file}")
            roof["algorithmic_bytes_per_launch"] = a["bytes"] / a["launches"]

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import effnet_oracle as EO
        from oracle.frontend_oracle import FrontendOracle
        cores = host_threads()
        torch.set_num_threads(cores)
        orc = FrontendOracle()

        def cpu_run(n):
            f_ = orc.features(pcm_host[:n], threads=cores)
            with torch.no_grad():
                return f_, EO.forward(weights, f_)

        nb = bounded_cpu_sample(cpu_run, min(B, args.ref_sample), 15.0, 3)       # ~5 s per repetition at most
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 50):
            f, ref_emb = cpu_run(nb)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        # frontend alone: single-threaded (the TF op is single-threaded per call) and one clip per thread
        t0 = time.perf_counter(); orc.features_u16(pcm_host[:64]); fe_1 = 64 / (time.perf_counter() - t0)
        nfe = min(B, 512)
        t0 = time.perf_counter(); orc.features_u16(pcm_host[:nfe], threads=cores); fe_n = nfe / (time.perf_counter() - t0)
        cpu = {"value": nb / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "frontend_only_clips_per_s": {"single_thread": round(fe_1, 1), "all_threads": round(fe_n, 1)},
               "sample": f"{reps} x {nb} synthetic clips (frontend C oracle on {cores} threads + torch-CPU fp32 network)"}
        # parity spot-check against the oracle (checker only): same architecture/kernels, BN statistics calibrated
        # by the oracle so activations have trained-like scale (the timed model uses un-calibrated random weights)
        w2 = {k: v.copy() for k, v in weights.items()}
        fs = f[:min(nb, 128)]
        EO.forward(w2, fs, calibrate_bn=True)
        with torch.no_grad():
            want = EO.forward(w2, fs).numpy()
        got = EmbeddingModel(w2, chunk=args.chunk, dtype=args.dtype).predict(fs)
        cpu["parity_min_cosine_vs_oracle"] = float(EO.cosine(got, want).min())
        feats_dev = fe.forward(pcm[:fs.shape[0]]).cpu().numpy()
        cpu["parity_frontend_bit_exact"] = bool(np.array_equal(feats_dev, fs))

    if rank == 0:
        launches_per_step = 1 + emb_model.launches(B)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype + " storage / tensor-core operands, fp32 accumulate; frontend int16/int32/uint64 fixed point",
            "data": "synthetic",
            "value_note": "value / ms_per_step: one step at a time on one stream (CUDA events around each step, L2 flushed "
                          "before it); overlapped: K steps rotating over the compute streams, whole region / K, flushes included",
            "overlapped": None if ms_overlap is None else {"value": world * B / (ms_overlap * 1e-3), "unit": UNIT,
                                                           "ms_per_step": ms_overlap, "streams": n_str},
            "config": {"workload": workload(B),
                       "global_batch": B * world, "clip_samples": 16000, "parallelism": f"dp{world} (clips sharded, no collective)",
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)", "chunk": args.chunk, "chunk_late": args.chunk_late,
                       "weights": "random init (Keras initialisers, randomised BN), no checkpoint available offline"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 32000, "d2h_bytes_per_step": B * emb_model.output_dim * 4,
                    "ms_per_step": ms_e2e, "regions_ms_per_step": [round(x, 4) for x in e2e_regions],
                    "timing": "median of three K-step regions; each: whole region (first enqueue -> last download, L2 flushes included) / K; run_host "
                              "enqueues only: upload i+1 / kernels i / download i-1 overlap, and consecutive steps rotate over the compute streams (2 device slots per stream)",
                    "ms_per_blocking_call_wall": ms_e2e_sync,
                    "h2d_alone": {"per_rank_GBps": h2d_gbps, "aggregate_GBps": h2d_gbps * world, "ms_per_step": ms_h2d,
                                  "ceiling_utt_per_s": world * B / (ms_h2d * 1e-3),
                                  "note": "all ranks copying their step's PCM (write-combined pinned -> HBM) at the same time, no kernels: "
                                          "the e2e figure cannot exceed this; on the 8-GPU lease every GPU hangs off one (virtual) NUMA "
                                          "node, so the ranks share one host memory system"},
                    "host_buffers": "PCM in write-combined pinned memory (kws_host_alloc), results in pinned memory"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": roof,
            "kernel_time_shares": shares,
            "cpu_baseline": cpu,
            "finetune": {"metric": "utterances/sec, 5-shot 3-way head fine-tune step (embedding fwd + head fwd/bwd + Adam)",
                         "value": world * ft_B / (ms_ft * 1e-3), "unit": UNIT, "ms_per_step": ms_ft, "batch_per_gpu": ft_B,
                         "grouped": None if ms_ft_group is None else {
                             "steps_per_embedding_forward": ft_G, "ms_per_step": ms_ft_group / ft_G,
                             "value": world * ft_B * ft_G / (ms_ft_group * 1e-3), "unit": UNIT,
                             "note": "what transfer_learning.fit does: the embedding is frozen, G steps share one forward"},
                         "collective": "one NCCL all-reduce(sum) of 18 510 fp32 per step" if world > 1 else "none (1 GPU)",
                         "phase2": {"what": "transfer_learn phase 2 (reference transfer_learning.py:97-112): forward + backward through "
                                            "block7a + top conv + dense tower + head, Adam; BASELINE config 3 'head + last block'",
                                    "ms_per_step": ms_ft2, "ms_per_step_without_cuda_graph": ms_ft2_eager, "batch_per_gpu": ft_B,
                                    "value": world * ft_B / (ms_ft2 * 1e-3), "unit": UNIT,
                                    "gradient_floats": int(trainer.flat.numel()),
                                    "collective": (f"one NCCL all-reduce(sum) of {trainer.flat.numel()} fp32 ({trainer.flat.numel() * 4 / 1e6:.1f} MB) per step"
                                                   if world > 1 else "none (1 GPU)"),
                                    "allreduce_alone_ms": ms_ar,
                                    "allreduce_bus_GBps": (None if not ms_ar else
                                                           2 * (world - 1) / world * trainer.flat.numel() * 4 / (ms_ar * 1e-3) / 1e9)}},
            "finetune_e2e": ft_e2e,
            "multi_rank_parity": parity,
            "other_configs": extra,
            "wall_s_timed_region": wall,
        }
        print(json.dumps(out))
    if world > 1:
        # Leaving: every measurement is printed.  Tearing the NCCL communicator down (barrier + destroy_process_group) was
        # seen to HANG at N = 8 once TailTrainer.step had captured its all-reduce into a CUDA graph (fine at N = 2), so the
        # ranks leave without touching NCCL again: the others wait on the rendezvous store (plain TCP) until rank 0 has
        # printed its line, then every process exits.
        import datetime
        sys.stdout.flush()
        sys.stderr.flush()
        try:
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                store.set("kws_bench_done", "1")
                time.sleep(1.0)                      # let the others read the key while the store (hosted here) is alive
            else:
                store.wait(["kws_bench_done"], datetime.timedelta(seconds=300))
        except Exception:
            pass
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="clips per GPU per step")
    ap.add_argument("--chunk", type=int, default=1024, help="clips per pass for the early (large-activation) layers")
    ap.add_argument("--chunk-late", type=int, default=4096, help="clips per pass for the late (small-activation) layers")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--ref-sample", type=int, default=1024, help="clips per reference / cpu_baseline step")
    ap.add_argument("--ref-budget-s", type=float, default=120.0,
                    help="wall-clock budget of the whole --impl reference run (the per-step sample is sized to fit)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-finetune-e2e", action="store_true", help="skip the transfer_learn() end-to-end timing")
    ap.add_argument("--pipe-streams", type=int, default=3, help="compute streams the host pipeline / overlapped figure rotate over")
    ap.add_argument("--pipe-sub-batch", type=int, default=0, help="clips per pipeline job of the e2e path (0 = the whole step)")
    ap.add_argument("--sm-budget", default="", help="head,tail SM budget of the throughput schedule (default: the pipeline's)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels directly instead of replaying the CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
