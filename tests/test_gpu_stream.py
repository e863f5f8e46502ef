"""Streaming front half (BASELINE config 4 / SURVEY f-2): appending small PCM chunks and recomputing only the touched
mel frames must give exactly what `pcm_to_mel` gives on the whole buffer, before and after norma-style seeks."""
import numpy as np
import pytest

from norma_b200 import ffi, filters, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = synth.model_config("test-micro")
    x = ffi.Context(c, compute="f32", max_batch=1)
    x.set_mel_filters(filters.mel_filters(80))
    x.load_weights(synth.synth_weights(c, seed=1, decoder=False))
    yield x
    x.close()


def batch_mel(ctx, pcm):
    return ctx.pcm_to_mel_batch(pcm[None, :].copy() if pcm.size else np.zeros((1, 1), np.float32), lens=[pcm.size])[0]


def test_incremental_mel_equals_full_recompute(lib, ctx):
    pcm = synth.synth_pcm("gauss", 21, 100_000)
    rng = np.random.default_rng(0)
    ctx.stream_reset()
    pos = 0
    checks = 0
    while pos < pcm.size:
        n = int(rng.choice([160, 160, 160, 37, 400, 1600, 5000]))
        ctx.stream_push(pcm[pos:pos + n])
        pos = min(pcm.size, pos + n)
        if checks < 12 and rng.random() < 0.15:
            mel, _ = ctx.stream_features(run_encoder=False, want_mel=True)
            assert np.array_equal(mel, batch_mel(ctx, pcm[:pos]))  # bit-identical: same kernel, same arithmetic per frame
            checks += 1
    mel, feat = ctx.stream_features(run_encoder=True, want_mel=True, want_features=True)
    assert np.array_equal(mel, batch_mel(ctx, pcm))
    ref = ctx.transcode_batch(pcm[None, :].copy())[0]
    assert np.array_equal(feat, ref)


@pytest.mark.parametrize("drain", [320 * 50, 320 * 1 + 7, 99_999])
def test_seek_then_continue(lib, ctx, drain):
    pcm = synth.synth_pcm("uniform", 22, 140_000)
    ctx.stream_reset()
    for lo in range(0, 100_000, 4000):
        ctx.stream_push(pcm[lo:lo + 4000])
    ctx.stream_drain(drain)  # norma: buf.drain(..s_timestamp * 320) (model.rs:126-127)
    mel, _ = ctx.stream_features(run_encoder=False, want_mel=True)
    assert np.array_equal(mel, batch_mel(ctx, pcm[drain:100_000]))
    for lo in range(100_000, 140_000, 160):
        ctx.stream_push(pcm[lo:lo + 160])
    mel, _ = ctx.stream_features(run_encoder=False, want_mel=True)
    assert np.array_equal(mel, batch_mel(ctx, pcm[drain:]))


def test_stream_bounds(lib, ctx):
    ctx.stream_reset()
    ctx.stream_push(np.zeros(480_000, np.float32))
    with pytest.raises(ffi.Nb200Error):
        ctx.stream_push(np.zeros(1, np.float32))  # more than one 30 s window
    with pytest.raises(ffi.Nb200Error):
        ctx.stream_drain(480_001)
    ctx.stream_drain(480_000)
    mel, _ = ctx.stream_features(run_encoder=False, want_mel=True)
    assert np.all(mel == np.float32(-1.5))
