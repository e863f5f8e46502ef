"""Bitwise repeatability of the whole log-mel + encoder pass at B windows (distil-large-v3 shape, bf16): nb200_transcode_batch REPS times."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, filters, synth
B, reps = int(os.environ.get("B", "25")), int(os.environ.get("REPS", "4"))
c = synth.model_config(os.environ.get("MODEL", "distil-large-v3"))
if os.environ.get("LAYERS"):
    c = dict(c, encoder_layers=int(os.environ["LAYERS"]))
ctx = ffi.Context(c, compute="bf16", max_batch=B)
ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"]))
ctx.load_weights(synth.synth_weights(c, seed=1, decoder=False))
pcm = np.stack([synth.synth_pcm_window(i) for i in range(B)])
ref = ctx.transcode_batch(pcm).copy()
bad = 0
for i in range(reps):
    out = ctx.transcode_batch(pcm)
    if not np.array_equal(out, ref):
        d = np.argwhere(out != ref)
        wins = sorted(set(int(x) for x in d[:, 0]))
        rows = len({(int(a), int(b)) for a, b, _ in d})
        bad += 1
        print(f"rep {i}: {len(d)} elements differ in {rows} rows of windows {wins}, max |d| {np.abs(out - ref).max():.3e}, rel fro {np.linalg.norm(out - ref) / np.linalg.norm(ref):.3e}")
print(f"B={B}: {bad} of {reps} repetitions differ from the first run")
