"""Decode step time through the two APIs on the same context: nb200_decode (one call) against nb200_decode_begin / _advance(16) / _end
(what bench.py's `decode` record uses), CUDA events on the context's stream and host wall clock."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, filters, synth
c = synth.model_config("distil-large-v3")
w = synth.synth_weights(c, seed=1)
ctx = ffi.Context(c, compute="bf16", max_batch=int(os.environ.get("MAXB", "1")))
ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"])); ctx.load_weights(w); ctx.set_tokens(**synth.special_tokens(c["vocab_size"]))
ctx.transcode_batch(np.stack([synth.synth_pcm_window(0)]), want_output=False)
ctx.decode_greedy(1, max_new_tokens=8); ctx.sync()
t = time.perf_counter(); r = ctx.decode_greedy(1, max_new_tokens=144); dt = time.perf_counter() - t
print(f"nb200_decode: {len(r[0]['tokens'])} tokens, {dt / 144 * 1e6:.1f} us/step wall")
for k in (16, 1):
    ctx.decode_begin(1, max_new_tokens=0); ctx.decode_advance(16); ctx.sync()
    steps = 0; ctx.timer_start(); t = time.perf_counter()
    for _ in range(128 // k):
        if ctx.decode_advance(k): break
        steps += k
    ms = ctx.timer_stop(); wall = time.perf_counter() - t
    ctx.decode_end()
    print(f"begin / advance({k}) x {128 // k} / end: {steps} steps, {1e3 * ms / steps:.1f} us/step by events, {1e6 * wall / steps:.1f} us/step wall")
