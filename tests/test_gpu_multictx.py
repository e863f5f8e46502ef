"""The one-process multi-context mode include/norma_b200.h promises ("different ctxs ... may be driven concurrently from different
threads"), which is what one `Transcriber` per `SelectedDevice::Cuda(n)` amounts to (/root/reference/src/models/mod.rs:38-55; every
model is driven by its own thread, /root/reference/src/lib.rs:377).  Several contexts are created in ONE process and driven from as many
host threads through the C ABI (ctypes releases the GIL during a call); every result must equal the single-context, single-thread one."""
import threading

import numpy as np
import pytest
import torch

from norma_b200 import ffi, filters, synth
from oracle.whisper_oracle import special_tokens_for_vocab

pytestmark = pytest.mark.gpu


def planted(name="tiny.en"):
    c = synth.model_config(name)
    st = special_tokens_for_vocab(c["vocab_size"])
    ts = lambda s: st.no_timestamps + 1 + int(round(s / 0.02))
    plan = {0: 7, 1: 8, 2: ts(0.0), 3: 100, 4: 200, 5: ts(2.0), 6: ts(2.02), 7: 300, 8: st.eot}
    return c, st, synth.plant_decoder_plan(synth.synth_weights(c, seed=1, embed_scale=1.0), c, plan), plan


def make(c, st, w, ordinal, compute="bf16", B=2):
    ctx = ffi.Context(c, ordinal=ordinal, compute=compute, max_batch=B)
    ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"]))
    ctx.load_weights(w)
    ctx.set_tokens(st.sot, st.eot, st.task, st.lang, st.no_speech, st.no_timestamps, st.ts_zero, st.ts_one)
    return ctx


def drive(ctxs, fn):
    """Run fn(i, ctx) on one thread per context, released together; re-raise the first failure."""
    bar = threading.Barrier(len(ctxs))
    out, err = [None] * len(ctxs), []

    def work(i):
        try:
            bar.wait()
            out[i] = fn(i, ctxs[i])
        except BaseException as e:  # noqa: BLE001
            err.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(ctxs))]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    return out


def test_contexts_on_one_gpu_driven_from_threads(lib):
    """Four contexts on GPU 0, four threads, different inputs and even different compute modes at the same time: nothing on the launch path
    is process-wide state (tile configuration, epilogue mode, attention choice and the tensor-map encoder used to be)."""
    c, st, w, _ = planted()
    pcm = [np.stack([synth.synth_pcm(k, s) for k, s in pair]) for pair in
           ((("gauss", 0), ("uniform", 1)), (("bursts", 2), ("gauss", 3)), (("uniform", 4), ("bursts", 5)), (("gauss", 6), ("gauss", 7)))]
    modes = ["bf16", "f32", "bf16", "f32"]
    solo = []
    for i in range(4):
        ctx = make(c, st, w, 0, modes[i])
        solo.append(ctx.transcode_batch(pcm[i]))
        ctx.close()
    ctxs = [make(c, st, w, 0, modes[i]) for i in range(4)]

    def fn(i, ctx):
        res = None
        for _ in range(6):  # several passes: graph capture on the first, replays afterwards, all interleaved with the other threads
            res = ctx.transcode_batch(pcm[i])
        return res

    got = drive(ctxs, fn)
    for i in range(4):
        assert np.array_equal(got[i], solo[i]), f"context {i} ({modes[i]}) differs when driven concurrently"
    for ctx in ctxs:
        ctx.close()


def test_one_context_per_gpu_driven_from_threads(lib):
    """N = all visible GPUs (skipped below 2): one context per ordinal, one thread each — encoder features bit-equal to ordinal 0 alone,
    greedy tokens of the planted model equal to the plan on every GPU."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    c, st, w, plan = planted()
    pcm = np.stack([synth.synth_pcm("gauss", 0), synth.synth_pcm("uniform", 1)])
    ref_ctx = make(c, st, w, 0)
    ref = ref_ctx.transcode_batch(pcm)
    ref_ctx.close()
    want = [st.sot, st.lang, st.task] + [plan[p] for p in range(2, 9)]
    ctxs = [make(c, st, w, g) for g in range(n)]

    def fn(g, ctx):
        feats = None
        for _ in range(4):
            feats = ctx.transcode_batch(pcm)
        return feats, ctx.decode(2, 0.0)

    got = drive(ctxs, fn)
    for g in range(n):
        assert np.array_equal(got[g][0], ref), f"GPU {g}"
        for b in range(2):
            assert got[g][1][b]["tokens"] == want, (g, b)
    for ctx in ctxs:
        ctx.close()
