import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, filters, synth
c = synth.model_config("test-micro")
st = synth.special_tokens(c["vocab_size"])
TS = lambda s: st["no_timestamps"] + 1 + int(round(s / 0.02))
plan = {0: 7, 1: 8, 2: TS(0.0), 3: 100, 4: 200, 5: TS(2.0), 6: TS(2.02), 7: 300, 8: st["eot"]}
w = synth.plant_decoder_plan(synth.synth_weights(c, seed=1, embed_scale=1.0), c, plan)
ctx = ffi.Context(c, compute="f32", max_batch=2)
ctx.set_mel_filters(filters.mel_filters(80)); ctx.load_weights(w); ctx.set_tokens(**st)
pcm = np.stack([synth.synth_pcm("gauss", 0), synth.synth_pcm("uniform", 1)])
ctx.transcode_batch(pcm, want_output=False)
bad = 0
for seed in range(300):
    r = ctx.decode(2, 1.0, seed=seed, max_new_tokens=1)
    for b in range(2):
        t = r[b]["tokens"]
        if not (st["ts_zero"] <= t[3] <= st["ts_one"]):
            bad += 1
            if bad < 8: print("seed", seed, "b", b, t, r[b]["avg_logprob"], r[b]["no_speech_prob"])
print("bad", bad, "of 600")
