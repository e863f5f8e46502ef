#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its quoted config: audio-seconds per second of log-mel + Whisper encoder,
distil-large-v3 shape (128 mel bins, 32 x d=1280 encoder), bf16, synthetic 30 s PCM windows, random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--windows B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path (PCM -> log-mel -> encoder) over a batch of B windows per GPU.
  value     whole-job audio-s/s with the PCM batch already resident in HBM (CUDA events on the ctx stream, max over ranks)
  e2e       the same through the reference-facing C-ABI call nb200_transcode_batch with HOST (pinned) PCM in and HOST
            encoder features out: H2D + both stages + D2H inside the timed region
  roofline  dominant kernel = the tcgen05 bf16 GEMM: algorithmic FLOPs (2MNK) / its live CUDA-event time, against the
            measured cuBLAS bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a port of the reference's candle path; the Rust reference cannot be built here) timed on
            a bounded sample (1 window) on this host's cores
`--impl reference` times that CPU port alone (the reference's own CPU implementation of the path is Rust/candle:
no toolchain in the image), one window per step.
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

METRIC = "audio_seconds_per_second_logmel_plus_encoder"
UNIT = "audio-s/s"
MODEL = "distil-large-v3"
WINDOW_S = 30.0


def encoder_flops(c) -> float:
    """2MNK FLOPs of one window (SURVEY §8 d): conv1 + conv2 + L x (qkv/out proj + attention + MLP)."""
    d, T, L, nm = c["d_model"], 1500, c["encoder_layers"], c["num_mel_bins"]
    return 2.0 * 3000 * d * 3 * nm + 2.0 * T * d * 3 * d + L * (8.0 * T * d * d + 4.0 * T * T * d + 16.0 * T * d * d)


def gemm_flops(c) -> float:
    d, T, L, nm = c["d_model"], 1500, c["encoder_layers"], c["num_mel_bins"]
    return 2.0 * 3000 * d * 3 * nm + 2.0 * T * d * 3 * d + L * (8.0 * T * d * d + 16.0 * T * d * d)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

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
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
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
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get("bf16_tflops_sustained", 1416.7), j.get("hbm_gbs", 6536.4), "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def cpu_port_window(c, weights, n_windows=1, seed0=0):
    """One bounded sample of the workload on the host cores with the CPU oracle (port of the candle path)."""
    import torch
    from norma_b200 import filters, synth
    from oracle import mel_c
    from oracle.whisper_oracle import Config, WhisperOracle

    f = filters.mel_filters(c["num_mel_bins"])
    orc = WhisperOracle(Config(**c), weights)
    pcm = [synth.synth_pcm_window(seed0 + i) for i in range(n_windows)]

    def step():
        t = time.perf_counter()
        mel = np.stack([mel_c.pcm_to_mel(p, f)[:, :3000] for p in pcm])
        y = orc.encoder_forward(torch.from_numpy(mel))
        return time.perf_counter() - t, y

    return step


def workload_config(args, c, world, B):
    """The `config` object both arms print (the reference arm adds `sample`)."""
    return {"workload": f"{args.model}-shaped log-mel + encoder ({c['num_mel_bins']} mel, {c['encoder_layers']} x d{c['d_model']}), "
                        + (f"{args.total_windows} x 30 s windows per step sharded over {world} GPU(s)" if args.total_windows
                           else f"{B} x 30 s windows per step per GPU") + ", random-init weights",
            "windows_per_step_per_gpu": B, "parallelism": f"window-sharded x{world}, no collective",
            "l2": "working set per step (1.27 GB bf16 weights + activations) exceeds the 126 MB L2; no explicit flush"}


def run_reference(args, rank):
    """`--impl reference`: the CPU port of the reference path, all host threads, one window per step."""
    if rank != 0:
        return
    import torch
    from norma_b200 import synth

    c = synth.model_config(args.model)
    w = synth.synth_weights(c, seed=1, decoder=False)
    step = cpu_port_window(c, w, 1)
    for _ in range(args.warmup):
        step()
    times = [step()[0] for _ in range(args.steps)]
    total = sum(times)
    val = WINDOW_S * args.steps / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, c, max(1, args.gpus), args.windows),
                       sample="each step = 1 of the workload's 30 s windows through the CPU port of the candle path (Rust reference unbuildable here)"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": "1 x 30 s window per step: C restatement of pcm_to_mel + torch-CPU fp32 encoder"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=int(os.environ.get("NB200_BENCH_WINDOWS", "25")), help="30 s windows per step per GPU")
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compute", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--total-windows", type=int, default=0,
                    help="strong-scaling mode (BASELINE config 3): this many windows in total, sharded round-robin over the ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from norma_b200 import ffi, filters, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    c = synth.model_config(args.model)
    B = args.windows
    if args.total_windows:  # window w -> rank w mod world (SURVEY §8 e); every rank gets ceil or floor of the share
        B = len(range(rank, args.total_windows, world))
    ctx = ffi.Context(c, ordinal=local_rank, compute=args.compute, max_batch=B)
    ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"]))
    weights = synth.synth_weights(c, seed=1, decoder=False)
    ctx.load_weights(weights)

    # synthetic PCM: window w of rank r has seed r*B + w (independent windows, sharded with no collective)
    pcm_pinned = torch.empty((B, 480_000), dtype=torch.float32, pin_memory=True)
    pcm = pcm_pinned.numpy()
    for w in range(B):
        pcm[w] = synth.synth_pcm_window(rank + w * world if args.total_windows else rank * B + w)
    out_pinned = torch.empty((B, 1500, c["d_model"]), dtype=torch.float32, pin_memory=True)
    out = out_pinned.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: inputs resident in HBM ----------------
    ctx.stage_pcm(pcm)
    for _ in range(args.warmup):
        ctx.run_resident(B)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.query("kernel_launches")
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.run_resident(B)
    ms = ctx.timer_stop()
    barrier()
    launches = ctx.query("kernel_launches") - l0
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    n_total = args.total_windows if args.total_windows else world * B  # windows per step over all ranks
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
    gemm_ms = prof["gemm"]["ms"]
    gemm_launches = max(prof["gemm"]["launches"], 1)
    achieved_tf = gflops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM bytes per GEMM launch from ONE `ncu --set full` capture of a layer's four GEMMs (qkv 346 MB, out 436, fc1 444, fc2 850; profiles/
    # r1e_gemm_full_summary.md) — only valid for the shape and batch it was captured on; the algorithmic bytes of the same four launches
    # average 538 MB (operands + f32 residual in / out), so nothing is re-read from HBM
    traffic = 5.19e8 if (args.model == "distil-large-v3" and B == 25 and args.compute == "bf16") else None
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16, fused epilogue)" if args.compute == "bf16" else "sgemm_kernel (fp32 SIMT)",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of the 4 GEMMs of a layer)",
        "peak_source": peak_src,
        "flops_per_launch": gflops / gemm_launches, "avg_launch_ms": gemm_ms / gemm_launches, "launches": gemm_launches,
        "profiled_ms_per_step": ms_prof / args.steps,
        "share_of_step": {k: (v["ms"] / ms_prof if ms_prof > 0 else 0.0) for k, v in prof.items() if v["ms"] > 0},
        "mel_stage": {"achieved_gbs": (4.0 * 480_000 + 4.0 * c["num_mel_bins"] * 3000) * B * args.steps / 1e9 / (max(prof["mel"]["ms"], 1e-9) / 1e3),
                      "peak_gbs": peak_hbm, "bound": "hbm"},
        "whole_step_tflops": encoder_flops(c) * B * args.steps / (ms / 1e3) / 1e12 / 1.0,
        "whole_step_frac_of_peak": encoder_flops(c) * B * args.steps / (ms / 1e3) / 1e12 / peak_tf,
    }

    # ---------------- e2e: host PCM in, host features out, through the C-ABI calls a norma binding makes ----------------
    # nb200_transcode_submit / _collect: every step uploads its PCM from pinned host memory and downloads its features into
    # pinned host memory inside the timed region; batch k+1's upload and batch k-1's download overlap batch k's compute.
    out2_pinned = torch.empty((B, 1500, c["d_model"]), dtype=torch.float32, pin_memory=True)
    outs = [out, out2_pinned.numpy()]
    ctx.transcode_batch(pcm, out=out)  # blocking form once (also the correctness reference for the pipelined form below)
    ref_sum = float(np.abs(out).sum())
    for i in range(2):
        ctx.transcode_submit(pcm, outs[i & 1])
    ctx.transcode_collect(); ctx.transcode_collect()
    assert abs(float(np.abs(outs[1]).sum()) - ref_sum) <= 1e-6 * ref_sum, "pipelined and blocking transcode disagree"
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctx.transcode_submit(pcm, outs[i & 1])
        if i >= 1:
            ctx.transcode_collect()
    ctx.transcode_collect()
    e2e_ms_wall = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    for _ in range(2):
        ctx.transcode_batch(pcm, out=out)
    e2e_blocking_ms = (time.perf_counter() - t0) * 1e3 / 2
    e2e_ms = max_over_ranks(e2e_ms_wall)
    e2e = {"value": n_total * WINDOW_S * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(n_total * 480_000 * 4),
           "d2h_bytes_per_step": int(n_total * 1500 * c["d_model"] * 4), "ms_per_step": e2e_ms / args.steps,
           "api": "nb200_transcode_submit/_collect(host pinned PCM) -> host pinned f32 encoder features, copies overlapped with compute",
           "blocking_call_ms_per_step": e2e_blocking_ms, "timing": "host wall clock around the submit/collect loop (max over ranks)"}
    checksum = float(np.abs(out[0]).mean())

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step = cpu_port_window(c, weights, 1)
        import torch as _t

        dt, _ = step()  # parity is the tests' job; this leg only times the port
        cpu_baseline = {"value": WINDOW_S / dt, "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
                        "sample": "1 x 30 s window (gauss seed 0): C restatement of candle pcm_to_mel + torch-CPU fp32 encoder of the same shape",
                        "seconds": dt}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.total_windows else "weak", "vs_baseline": None, "dtype": args.compute, "data": "synthetic",
            "config": workload_config(args, c, world, B),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "feature_checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
