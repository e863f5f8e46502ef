#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its quoted config: audio-seconds per second of log-mel + Whisper encoder,
distil-large-v3 shape (128 mel bins, 32 x d=1280 encoder), bf16, synthetic 30 s PCM windows, random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--windows B] [--impl ours|reference] [--quick] [--single-process]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path (PCM -> log-mel -> encoder) over a batch of B windows per GPU.
  value     whole-job audio-s/s with the PCM batch already resident in HBM (CUDA events on the ctx stream, max over ranks)
  e2e       the same through the reference-facing C-ABI calls nb200_transcode_submit/_collect with HOST (pinned) PCM in and HOST
            encoder features out: H2D + both stages + D2H inside the timed region
  roofline  dominant kernel = the tcgen05 bf16 GEMM: algorithmic FLOPs (2MNK) / its live CUDA-event time, against the
            measured cuBLAS bf16 peak in MEASURED_PEAKS.json; `kernels` lists every kernel class of the step the same way
  cpu_baseline  the CPU oracle (a port of the reference's candle path; the Rust reference cannot be built here) timed on
            a bounded sample (1 window) on all of this host's cores
The rest of the north-star path and the other BASELINE configs ride along as extra records of the same JSON line, each
measured live in this run (clocks are sampled over the whole run):
  single_window  BASELINE config 2 as written: ONE 30 s window per step (latency-shaped)
  decode         the KV-cached greedy decode step (us/step, tokens/s, fraction of the HBM weight-streaming floor) at B = 1 and 8
  stream         BASELINE config 4: 10 ms appends, incremental mel, encoder + decode every R ms: p50 / p99
  config3        BASELINE config 3: base.en, 120 x 30 s windows sharded over the ranks of this run
  config5        BASELINE config 5: large-v3-shaped encoder, batch 1..256 windows per GPU on the GPUs of this run
  gpu_baseline_standin  SURVEY §2b's bar ("the existing GPU path"): the same graph in torch eager fp32 with TF32 off, unfused
            (cuBLAS SGEMM + separate bias / GELU / softmax kernels, materialised 1500 x 1500 scores) — NOT candle, same class
`--impl reference` times the CPU port alone (the reference's own CPU implementation of the path is Rust/candle: no toolchain in
the image) on all host cores, one window per step.  `--quick` skips the extra records.
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
sys.path.insert(0, ROOT)

from norma_b200 import workload as wl  # noqa: E402  (pure Python)

METRIC = "audio_seconds_per_second_logmel_plus_encoder"
UNIT = "audio-s/s"
MODEL = "distil-large-v3"
WINDOW_S = wl.WINDOW_S


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.marks = index, [], None, {}

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self, name: str):
        """Start of a named section: summary() reports the samples of [this mark, the next mark)."""
        self.marks[name] = time.perf_counter()

    @staticmethod
    def _summarise(rows):
        sm, mx, pw, reasons = [], [], [], set()
        for _, r in rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}

    def section(self, name: str, end: float = None):
        t0 = self.marks.get(name)
        if t0 is None or not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        t1 = end if end is not None else time.perf_counter()
        return self._summarise([r for r in self.rows if t0 <= r[0] <= t1])

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get("bf16_tflops_sustained", 1416.7), j.get("hbm_gbs", 6536.4), "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16 / copy bandwidth)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """Every core this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_port_window(c, weights, n_windows=1, seed0=0):
    """One bounded sample of the workload on ALL host cores with the CPU oracle (port of the candle path)."""
    import torch
    from norma_b200 import filters, synth
    from oracle import mel_c
    from oracle.whisper_oracle import Config, WhisperOracle

    n_thr = host_threads()
    torch.set_num_threads(n_thr)
    f = filters.mel_filters(c["num_mel_bins"])
    orc = WhisperOracle(Config(**c), weights)
    pcm = [synth.synth_pcm_window(seed0 + i) for i in range(n_windows)]

    def step():
        t = time.perf_counter()
        mel = np.stack([mel_c.pcm_to_mel(p, f, n_threads=n_thr)[:, :3000] for p in pcm])
        y = orc.encoder_forward(torch.from_numpy(mel))
        return time.perf_counter() - t, y

    return step, n_thr


def workload_config(args, c, world, B):
    """The `config` object BOTH arms print, key for key."""
    return {"workload": f"{args.model}-shaped log-mel + encoder ({c['num_mel_bins']} mel, {c['encoder_layers']} x d{c['d_model']}), "
                        + (f"{args.total_windows} x 30 s windows per step sharded over {world} GPU(s)" if args.total_windows
                           else f"{B} x 30 s windows per step per GPU") + ", random-init weights",
            "windows_per_step_per_gpu": B, "parallelism": f"window-sharded x{world}, no collective",
            "l2": "working set per step (1.27 GB bf16 weights + activations) exceeds the 126 MB L2; no explicit flush"}


def run_reference(args, rank, world):
    """`--impl reference`: the CPU port of the reference path on all host cores, one window per step (rank 0 only)."""
    if rank != 0:
        return
    from norma_b200 import synth

    c = synth.model_config(args.model)
    w = synth.synth_weights(c, seed=1, decoder=False)
    step, cores = cpu_port_window(c, w, 1)
    for _ in range(args.warmup):
        step()
    times = [step()[0] for _ in range(args.steps)]
    total = sum(times)
    val = WINDOW_S * args.steps / total
    sample = "each step = 1 of the workload's 30 s windows: C restatement of candle pcm_to_mel + torch-CPU fp32 encoder (Rust reference unbuildable here)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, c, max(1, world if world > 1 else args.gpus), args.windows),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------ extra records
def rec_single_window(ctx, c, steps, peak_tf):
    """BASELINE config 2 as written: one 30 s window per step, inputs resident (window 0 of the staged batch)."""
    for _ in range(5):
        ctx.run_resident(1)
    ctx.sync()
    n = max(steps, 20)
    per = []
    for _ in range(n):  # one window = one graph launch = one timed step; median over the steps (see rec_decode)
        ctx.timer_start()
        ctx.run_resident(1)
        per.append(ctx.timer_stop())
    ms = float(np.median(per))
    tf = wl.encoder_flops(c) / (ms / 1e3) / 1e12
    return {"value": WINDOW_S / (ms / 1e3), "unit": UNIT, "ms_per_window": ms, "steps": n, "ms_per_window_mean": float(np.mean(per)), "tflops": tf, "frac_of_peak": tf / peak_tf,
            "how": "nb200_run_resident(1): log-mel + encoder of ONE window replayed as one CUDA graph; CUDA events on the ctx stream around every window, median"}


def rec_decode(ctx, c, xa_windows, peak_hbm, n_launches=12):
    """The KV-cached greedy decode step on the resident features of `xa_windows` windows (fused cooperative kernel, 16 positions per launch)."""
    out = {}
    for B in (1, 8):
        if B > xa_windows:
            continue
        ctx.decode_begin(B, max_new_tokens=0)
        ctx.decode_advance(16)  # warm-up launch
        ctx.sync()
        per_launch, t0 = [], time.perf_counter()
        for _ in range(n_launches):  # every launch of 16 positions is timed on its own: the median is immune to an nvidia-smi query landing in one
            ctx.timer_start()
            stop = ctx.decode_advance(16)
            ms = ctx.timer_stop()
            if stop:
                break
            per_launch.append(ms)
        wall = (time.perf_counter() - t0) * 1e3
        res = ctx.decode_end()
        steps = 16 * len(per_launch)
        if steps == 0:
            continue
        us = 1e3 * float(np.median(per_launch)) / 16
        floor_bytes = wl.decode_bytes_per_step(c, B)
        out[f"B{B}"] = {"us_per_step": us, "tokens_per_s": B * 1e6 / us, "steps_timed": steps, "us_per_step_mean": 1e3 * float(np.mean(per_launch)) / 16,
                        "hbm_floor_bytes_per_step": floor_bytes,
                        "hbm_floor_us": floor_bytes / (peak_hbm * 1e9) * 1e6, "frac_of_hbm_floor": floor_bytes / (peak_hbm * 1e9) * 1e6 / us,
                        "wall_us_per_step_incl_host_polls": 1e3 * wall / steps, "tokens_decoded": len(res[0]["tokens"])}
    out["how"] = ("nb200_decode_begin / _advance(16) / _end on resident encoder features, random-init decoder (embed_tokens x 8); CUDA events on the ctx "
                  "stream around EVERY advance call, median over the launches (the clock sampler's nvidia-smi queries stall a launch now and then: 242 vs 186 us "
                  "per step as a mean over eight launches); floor = 2 (L_dec 14 d^2 + V d) B of bf16 weights + the cross K/V of B windows at the measured HBM bandwidth")
    return out


def rec_stream(ctx, c, pcm):
    """BASELINE config 4: cpal-shaped 10 ms (160-sample) appends, incremental mel; encoder + greedy decode (7 sampled tokens) every R ms."""
    out = {"results": []}
    for R_ms, seconds in ((1000, 12.0), (100, 6.0)):
        ctx.stream_reset()
        push, trig, enc_only = [], [], []
        every = R_ms * 16
        for lo in range(0, int(seconds * 16000), 160):
            t0 = time.perf_counter(); ctx.stream_push(pcm[lo:lo + 160]); push.append(time.perf_counter() - t0)
            if (lo + 160) % every == 0:
                t0 = time.perf_counter()
                ctx.stream_features(run_encoder=True)
                t1 = time.perf_counter()
                ctx.decode_greedy(1, max_new_tokens=7)
                t2 = time.perf_counter()
                trig.append(t2 - t0); enc_only.append(t1 - t0)
        q = lambda a, p: float(np.percentile(np.asarray(a) * 1e3, p))
        skip = 2 if len(trig) > 4 else 0
        out["results"].append({"trigger_every_ms": R_ms, "audio_seconds": seconds, "n_triggers": len(trig),
                               "chunk_push_ms": {"p50": q(push, 50), "p99": q(push, 99)},
                               "mel_norm+encoder_ms": {"p50": q(enc_only[skip:], 50), "p99": q(enc_only[skip:], 99)},
                               "chunk_to_tokens_ms": {"p50": q(trig[skip:], 50), "p99": q(trig[skip:], 99)}})
    out["how"] = ("nb200_stream_push(160 samples) per chunk; every R ms nb200_stream_features (normalise + encoder of the buffered audio) then nb200_decode_greedy "
                  "with a budget of 7 sampled tokens (random-init decoder); host wall clock per call, chunk_to_tokens = features + decode")
    return out


def rec_config3(args, rank, local_rank, world, barrier, max_over_ranks, steps):
    """BASELINE config 3: base.en, 1 h = 120 x 30 s windows (seeds 0..119: gauss / uniform / bursts mix) sharded round-robin over the ranks."""
    from norma_b200 import ffi, filters, synth

    c = synth.model_config("base.en")
    total = 120
    ids = wl.windows_of_rank(rank, world, total)
    B = len(ids)
    ms = 0.0
    if B > 0:
        ctx = ffi.Context(c, ordinal=local_rank, compute="bf16", max_batch=B)
        ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"]))
        ctx.load_weights(synth.synth_weights(c, seed=1, decoder=False))
        kinds = ("gauss", "uniform", "bursts")
        pcm = np.stack([synth.synth_pcm(kinds[w % 3], w) for w in ids])
        ctx.stage_pcm(pcm)
        for _ in range(3):
            ctx.run_resident(B)
    barrier()
    if B > 0:
        ctx.timer_start()
        for _ in range(steps):
            ctx.run_resident(B)
        ms = ctx.timer_stop()
    barrier()
    ms = max_over_ranks(ms)
    if B > 0:
        ctx.close()
    return {"value": total * WINDOW_S * steps / (ms / 1e3), "unit": UNIT, "model": "base.en", "total_windows": total, "n_gpus": world, "steps": steps,
            "ms_per_step": ms / steps, "scaling": "strong", "tflops": wl.encoder_flops(c) * total * steps / (ms / 1e3) / 1e12,
            "how": "120 windows per step, window w -> rank w mod world, inputs resident, CUDA events, max over ranks"}


def rec_config5(args, c, weights, local_rank, world, barrier, max_over_ranks, pcm25):
    """BASELINE config 5: large-v3-shaped encoder (= this encoder: 128 mel, 32 x d1280), batch 1..256 windows per GPU on every GPU of the run."""
    from norma_b200 import ffi, filters

    sizes = [1, 2, 4, 8, 16, 32, 64, 128, 256]
    ctx = ffi.Context(c, ordinal=local_rank, compute="bf16", max_batch=max(sizes))
    ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"]))
    ctx.load_weights(weights)
    reps = -(-max(sizes) // pcm25.shape[0])
    pcm = np.concatenate([pcm25] * reps)[: max(sizes)]
    ctx.stage_pcm(pcm)
    rows = []
    for B in sizes:
        n = 3 if B >= 64 else 6
        for _ in range(2):
            ctx.run_resident(B)
        barrier()
        ctx.timer_start()
        for _ in range(n):
            ctx.run_resident(B)
        ms = max_over_ranks(ctx.timer_stop()) / n
        barrier()
        rows.append({"windows_per_gpu": B, "ms_per_step": ms, "value": world * B * WINDOW_S / (ms / 1e3),
                     "per_gpu": B * WINDOW_S / (ms / 1e3), "tflops_per_gpu": wl.encoder_flops(c) * B / (ms / 1e3) / 1e12})
    ctx.close()
    return {"unit": UNIT, "model": "large-v3-shaped encoder (128 mel, 32 x d1280)", "n_gpus": world, "rows": rows,
            "how": "log-mel + encoder, inputs resident, CUDA events, max over ranks; value = all GPUs of this run"}


def rec_gpu_standin(c, weights, pcm_window):
    """SURVEY §2b's bar — 'the existing GPU path' = candle(CUDA): cuBLAS SGEMM (TF32 off) + unfused SIMT kernels with a materialised
    [h, 1500, 1500] f32 score tensor, and the log-mel on the CPU.  candle cannot be built here; this is the SAME GRAPH in torch eager fp32
    on this GPU, op by op — same class of execution, clearly not candle itself.  Context only, never the product path."""
    import torch
    import torch.nn.functional as F

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", torch.cuda.current_device())
    w = {k: v.to(dev) for k, v in weights.items() if k.startswith("model.encoder.")}
    d, H, L = c["d_model"], c["encoder_attention_heads"], c["encoder_layers"]
    half = d // 2
    inv = torch.exp(-torch.arange(half, dtype=torch.float32) * (np.log(10000.0) / (half - 1)))
    t = torch.arange(1500, dtype=torch.float32)[:, None] * inv[None, :]
    pos = torch.cat([t.sin(), t.cos()], 1).to(dev)
    win = torch.hann_window(400, periodic=True, device=dev)
    from norma_b200 import filters as flt

    fb = torch.from_numpy(flt.mel_filters(c["num_mel_bins"])).to(dev)
    gelu = lambda x: F.gelu(x, approximate="tanh")
    scale = 64.0 ** -0.25

    def step(pcm_dev):
        x = torch.zeros(4500 * 160 + 400, device=dev)
        x[: pcm_dev.numel()] = pcm_dev
        st = torch.stft(x, 400, 160, window=win, center=False, return_complex=True)[:, :3000]
        p = st.real ** 2 + st.imag ** 2
        p[1:200] *= 2.0
        mel = torch.log10(torch.clamp(fb @ p, min=1e-10))
        mel = torch.maximum(mel, torch.clamp(mel.max(), min=-10.0) - 8.0) / 4.0 + 1.0
        e = "model.encoder."
        x = gelu(F.conv1d(mel[None], w[e + "conv1.weight"], w[e + "conv1.bias"], padding=1))
        x = gelu(F.conv1d(x, w[e + "conv2.weight"], w[e + "conv2.bias"], stride=2, padding=1))
        x = x.transpose(1, 2) + pos
        for i in range(L):
            lp = f"{e}layers.{i}."
            h = F.layer_norm(x, (d,), w[lp + "self_attn_layer_norm.weight"], w[lp + "self_attn_layer_norm.bias"], 1e-5)
            q = F.linear(h, w[lp + "self_attn.q_proj.weight"], w[lp + "self_attn.q_proj.bias"])
            k = F.linear(h, w[lp + "self_attn.k_proj.weight"])
            v = F.linear(h, w[lp + "self_attn.v_proj.weight"], w[lp + "self_attn.v_proj.bias"])
            q = q.view(1, 1500, H, 64).transpose(1, 2) * scale
            k = k.view(1, 1500, H, 64).transpose(1, 2).transpose(2, 3) * scale
            v = v.view(1, 1500, H, 64).transpose(1, 2)
            a = (torch.softmax(q @ k, -1) @ v).transpose(1, 2).flatten(2)
            x = x + F.linear(a, w[lp + "self_attn.out_proj.weight"], w[lp + "self_attn.out_proj.bias"])
            h = F.layer_norm(x, (d,), w[lp + "final_layer_norm.weight"], w[lp + "final_layer_norm.bias"], 1e-5)
            x = x + F.linear(gelu(F.linear(h, w[lp + "fc1.weight"], w[lp + "fc1.bias"])), w[lp + "fc2.weight"], w[lp + "fc2.bias"])
        return F.layer_norm(x, (d,), w[e + "layer_norm.weight"], w[e + "layer_norm.bias"], 1e-5)

    with torch.no_grad():
        pcm_dev = torch.from_numpy(pcm_window).to(dev)
        for _ in range(2):
            y = step(pcm_dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 4
        a.record()
        for _ in range(n):
            y = step(pcm_dev)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        chk = float(y.abs().mean())
    del w
    torch.cuda.empty_cache()
    return {"value": WINDOW_S / (ms / 1e3), "unit": UNIT, "ms_per_window": ms, "windows_per_step": 1, "dtype": "f32 (TF32 off)", "feature_checksum": chk,
            "what": "NOT candle: the same graph in torch eager fp32 on this GPU, unfused (cuBLAS SGEMM, separate bias / GELU / softmax kernels, materialised "
                    "scores, torch.stft log-mel on the GPU) — the class of GPU path SURVEY §2b names as the bar; context only"}


# ------------------------------------------------------------------------------------------------------------ single-process N-context mode
def run_single_process(args):
    """One process, N contexts (ordinals 0..N-1), N host threads through the C ABI — the mode one `Transcriber` per `SelectedDevice::Cuda(n)`
    gives (src/models/mod.rs:38-55).  Same timing rules: per-context CUDA events, max over contexts, threads released together."""
    import torch
    from norma_b200 import ffi, filters, synth

    N, B = args.gpus, args.windows
    c = synth.model_config(args.model)
    weights = synth.synth_weights(c, seed=1, decoder=False)
    f = filters.mel_filters(c["num_mel_bins"])
    ctxs, pcms = [], []
    for g in range(N):
        ctx = ffi.Context(c, ordinal=g, compute=args.compute, max_batch=B)
        ctx.set_mel_filters(f)
        ctx.load_weights(weights)
        pcm = np.stack([synth.synth_pcm_window(w) for w in wl.window_ids(g, N, B)])
        ctx.stage_pcm(pcm)
        ctxs.append(ctx); pcms.append(pcm)
    bar = threading.Barrier(N)
    ms = [0.0] * N
    sampler = ClockSampler(0)
    sampler.start()
    sampler.mark("timed")

    def worker(g):
        ctx = ctxs[g]
        for _ in range(args.warmup):
            ctx.run_resident(B)
        ctx.sync()
        bar.wait()
        ctx.timer_start()
        for _ in range(args.steps):
            ctx.run_resident(B)
        ms[g] = ctx.timer_stop()

    th = [threading.Thread(target=worker, args=(g,)) for g in range(N)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    wall = (time.perf_counter() - t0) * 1e3
    clocks = sampler.section("timed")
    sampler.stop()
    worst = max(ms)
    line = {"metric": METRIC, "value": N * B * WINDOW_S * args.steps / (worst / 1e3), "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": worst / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.compute, "data": "synthetic",
            "config": dict(workload_config(args, c, N, B), mode="ONE process, one nb200_ctx + one host thread per GPU (no torchrun, no NCCL)"),
            "clocks": clocks, "per_context_ms": ms, "wall_ms_incl_warmup": wall,
            "gpu_launches": int(sum(ctx.query("kernel_launches") for ctx in ctxs))}
    print(json.dumps(line), flush=True)
    for ctx in ctxs:
        ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=int(os.environ.get("NB200_BENCH_WINDOWS", "25")), help="30 s windows per step per GPU")
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline only: skip single_window / decode / stream / config3 / config5 / gpu_baseline_standin")
    ap.add_argument("--single-process", action="store_true", help="N contexts driven by N threads of ONE process instead of one process per GPU")
    ap.add_argument("--compute", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--total-windows", type=int, default=0,
                    help="strong-scaling mode (BASELINE config 3): this many windows in total, sharded round-robin over the ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from norma_b200 import ffi, filters, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    if args.single_process:
        if world > 1:
            raise SystemExit("--single-process is launched with plain python, not torchrun")
        run_single_process(args)
        return
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    c = synth.model_config(args.model)
    ids = wl.window_ids(rank, world, args.windows, args.total_windows)
    B = len(ids)
    if B == 0:
        raise SystemExit(f"bench.py: rank {rank} of {world} has no window to process (--total-windows {args.total_windows} < world size)")
    full = (not args.quick) and args.compute == "bf16" and args.model == MODEL and not args.total_windows
    ctx = ffi.Context(c, ordinal=local_rank, compute=args.compute, max_batch=B)
    ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"]))
    weights = synth.synth_weights(c, seed=1, decoder=full)  # the decoder only rides along for the decode / stream records
    ctx.load_weights(weights)
    if full:
        ctx.set_tokens(**synth.special_tokens(c["vocab_size"]))

    # synthetic PCM: one independent window per global id (sharded with no collective)
    pcm_pinned = torch.empty((B, 480_000), dtype=torch.float32, pin_memory=True)
    pcm = pcm_pinned.numpy()
    for i, w in enumerate(ids):
        pcm[i] = synth.synth_pcm_window(w)
    out_pinned = torch.empty((B, 1500, c["d_model"]), dtype=torch.float32, pin_memory=True)
    out = out_pinned.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if ctx.h:  # the headline context is closed before the config 3 / 5 records open their own
            ctx.sync()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---------------- value: inputs resident in HBM ----------------
    ctx.stage_pcm(pcm)
    for _ in range(args.warmup):
        ctx.run_resident(B)
    barrier()
    l0 = ctx.query("kernel_launches")
    barrier()
    sampler.mark("headline")
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.run_resident(B)
    ms = ctx.timer_stop()
    barrier()
    t_headline_end = time.perf_counter()
    launches = ctx.query("kernel_launches") - l0
    ms = max_over_ranks(ms)
    n_total = wl.windows_per_step(world, B, args.total_windows)  # windows per step over all ranks
    value = n_total * WINDOW_S * args.steps / (ms / 1e3)

    # ---------------- roofline: same steps with per-kernel CUDA events on the launching stream ----------------
    ctx.profile_reset()
    ctx.profile_enable(True)
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.run_resident(B)
    ms_prof = ctx.timer_stop()
    prof, gflops = ctx.profile_read()
    ctx.profile_enable(False)
    peak_tf, peak_hbm, peak_src = measured_peaks()
    es = 2 if args.compute == "bf16" else 4

    def per_launch(k):
        return prof[k]["ms"] / max(prof[k]["launches"], 1)

    gemm_ms = prof["gemm"]["ms"]
    gemm_launches = max(prof["gemm"]["launches"], 1)
    achieved_tf = gflops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    attn_tf = wl.attention_flops(c) * B * args.steps / (max(prof["attn"]["ms"], 1e-9) / 1e3) / 1e12
    mel_gbs = wl.mel_bytes(c) * B * args.steps / 1e9 / (max(prof["mel"]["ms"], 1e-9) / 1e3)
    ln_rows = prof["layernorm"]["launches"] * B * 1500
    ln_gbs = ln_rows * wl.layernorm_bytes_per_row(c["d_model"], es) / 1e9 / (max(prof["layernorm"]["ms"], 1e-9) / 1e3)
    share = {k: (v["ms"] / ms_prof if ms_prof > 0 else 0.0) for k, v in prof.items() if v["ms"] > 0}
    kernels = [
        {"kernel": "gemm_tc_kernel", "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
         "share_of_step": share.get("gemm", 0.0), "launches_per_step": prof["gemm"]["launches"] / args.steps, "avg_launch_us": 1e3 * per_launch("gemm")},
        {"kernel": "attn_tc_kernel", "bound": "tensor (XU-limited at head_dim 64, DESIGN.md §4)", "achieved": attn_tf, "peak": peak_tf, "unit": "TFLOP/s",
         "frac": attn_tf / peak_tf, "share_of_step": share.get("attn", 0.0), "launches_per_step": prof["attn"]["launches"] / args.steps,
         "avg_launch_us": 1e3 * per_launch("attn")},
        {"kernel": "layernorm_kernel", "bound": "hbm", "achieved": ln_gbs, "peak": peak_hbm, "unit": "GB/s", "frac": ln_gbs / peak_hbm,
         "share_of_step": share.get("layernorm", 0.0), "launches_per_step": prof["layernorm"]["launches"] / args.steps, "avg_launch_us": 1e3 * per_launch("layernorm")},
        {"kernel": "mel_kernel", "bound": "hbm", "achieved": mel_gbs, "peak": peak_hbm, "unit": "GB/s", "frac": mel_gbs / peak_hbm,
         "share_of_step": share.get("mel", 0.0), "launches_per_step": prof["mel"]["launches"] / args.steps, "avg_launch_us": 1e3 * per_launch("mel")},
    ]
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16, fused epilogue)" if args.compute == "bf16" else "sgemm_kernel (fp32 SIMT)",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        # `traffic` is NOT measured by this run: it is the ALGORITHMIC HBM bytes per launch computed from the shape (operands once + f32 residual in /
        # out, mean of a layer's four GEMMs).  The measured DRAM bytes of the same launches are in the ncu capture named below.
        "traffic": wl.gemm_bytes(c, B, es, ln_folded=(args.compute == "bf16" and os.environ.get("NB200_LN_FUSED", "1") != "0")),
        "traffic_kind": "algorithmic bytes per launch computed from the shape, LayerNorm folded (not measured here)",
        "traffic_measured_in": "profiles/r2h_gemm_full_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
        "peak_source": peak_src,
        "flops_per_launch": gflops / gemm_launches, "avg_launch_ms": gemm_ms / gemm_launches, "launches": gemm_launches,
        "profiled_ms_per_step": ms_prof / args.steps,
        "share_of_step": share,
        "kernels": kernels,
        "mel_stage": {"achieved_gbs": mel_gbs, "peak_gbs": peak_hbm, "bound": "hbm"},
        "whole_step_tflops": wl.encoder_flops(c) * B * args.steps / (ms / 1e3) / 1e12,
        "whole_step_frac_of_peak": wl.encoder_flops(c) * B * args.steps / (ms / 1e3) / 1e12 / peak_tf,
    }

    # ---------------- e2e: host PCM in, host features out, through the C-ABI calls a norma binding makes ----------------
    # nb200_transcode_submit / _collect: every step uploads its PCM from pinned host memory and downloads its features into
    # pinned host memory inside the timed region; batch k+1's upload and batch k-1's download overlap batch k's compute.
    out2_pinned = torch.empty((B, 1500, c["d_model"]), dtype=torch.float32, pin_memory=True)
    outs = [out, out2_pinned.numpy()]
    ctx.transcode_batch(pcm, out=out)  # blocking form once (also the correctness reference for the pipelined form below)
    ref_sum = float(np.abs(out).sum())
    blocking = out.copy()
    for i in range(2):
        ctx.transcode_submit(pcm, outs[i & 1])
    ctx.transcode_collect(); ctx.transcode_collect()
    assert abs(float(np.abs(outs[1]).sum()) - ref_sum) <= 1e-6 * ref_sum, "pipelined and blocking transcode disagree"
    bit_identical = bool(np.array_equal(outs[1], blocking))  # the same kernels on the same inputs: every bit (reported, not asserted)
    del blocking
    barrier()
    sampler.mark("e2e")
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctx.transcode_submit(pcm, outs[i & 1])
        if i >= 1:
            ctx.transcode_collect()
    ctx.transcode_collect()
    e2e_ms_wall = (time.perf_counter() - t0) * 1e3
    t_e2e_end = time.perf_counter()
    t0 = time.perf_counter()
    for _ in range(2):
        ctx.transcode_batch(pcm, out=out)
    e2e_blocking_ms = (time.perf_counter() - t0) * 1e3 / 2
    e2e_ms = max_over_ranks(e2e_ms_wall)
    e2e = {"value": n_total * WINDOW_S * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(n_total * 480_000 * 4),
           "d2h_bytes_per_step": int(n_total * 1500 * c["d_model"] * 4), "ms_per_step": e2e_ms / args.steps,
           "api": "nb200_transcode_submit/_collect(host pinned PCM) -> host pinned f32 encoder features, copies overlapped with compute",
           "blocking_call_ms_per_step": e2e_blocking_ms, "pipelined_bit_identical_to_blocking": bit_identical, "timing": "host wall clock around the submit/collect loop (max over ranks)",
           "clocks": sampler.section("e2e", t_e2e_end)}
    checksum = float(np.abs(out[0]).mean())

    extra = {}

    def record(name, fn, cooldown_s=0.0):
        """An extra record never takes the headline line down with it: a failure is reported in place of the record.  `cooldown_s`: idle
        time in front of a latency-shaped record — the headline leaves the GPU at its power cap (~1.45 GHz) and the cap's hysteresis outlives
        a 30 ms measurement: the decode step read 242 us right behind the 25-window steps and 186 us on an idle GPU (it is latency-bound,
        so it scales with the SM clock)."""
        if cooldown_s > 0:
            time.sleep(cooldown_s)
        sampler.mark(name)
        try:
            extra[name] = fn()
            extra[name]["clocks"] = sampler.section(name)
            if cooldown_s > 0:
                extra[name]["cooldown_s"] = cooldown_s
        except Exception as e:
            extra[name] = {"unavailable": f"{type(e).__name__}: {e}"}

    if full:
        ctx.stage_pcm(pcm)
        ctx.run_resident(B)  # encoder features of B windows resident: the decode record's input
        ctx.sync()
        record("decode", lambda: rec_decode(ctx, c, B, peak_hbm), cooldown_s=3.0)
        record("stream", lambda: rec_stream(ctx, c, pcm[0]))
        ctx.stage_pcm(pcm)  # window 0 resident again (the streaming record used the buffer)
        record("single_window", lambda: rec_single_window(ctx, c, args.steps, peak_tf))
    ctx.close()
    if full:
        record("config3", lambda: rec_config3(args, rank, local_rank, world, barrier, max_over_ranks, max(args.steps, 10)))
        enc_w = {k: v for k, v in weights.items() if k.startswith("model.encoder.")}
        record("config5", lambda: rec_config5(args, c, enc_w, local_rank, world, barrier, max_over_ranks, pcm))

    cpu_baseline = None
    if rank == 0 and world == 1 and full:
        record("gpu_baseline_standin", lambda: rec_gpu_standin(c, weights, pcm[0]))  # context only: never fails the bench
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        enc_w = {k: v for k, v in weights.items() if k.startswith("model.encoder.")}
        step, cores = cpu_port_window(c, enc_w, 1)
        dt, _ = step()  # parity is the tests' job; this leg only times the port
        cpu_baseline = {"value": WINDOW_S / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "1 x 30 s window (seed 0): C restatement of candle pcm_to_mel + torch-CPU fp32 encoder of the same shape, all host cores",
                        "seconds": dt}
    clocks = sampler.section("headline", t_headline_end)
    sampler.stop()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.total_windows else "weak", "vs_baseline": None, "dtype": args.compute, "data": "synthetic",
            "config": workload_config(args, c, world, B),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "feature_checksum": checksum,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
