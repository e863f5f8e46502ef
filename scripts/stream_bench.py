"""BASELINE config 4: distil-large-v3-shaped streaming — cpal-shaped 10 ms (160-sample) PCM chunks, incremental mel on the device,
encoder + KV-cached greedy decode every R ms of audio; reports p50 / p99 latencies.  Random-init weights with a planted,
confident decoder plan (synth.plant_decoder_plan) so that every decode emits the same 7 tokens and stops at eot."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, filters, synth

name = os.environ.get("MODEL", "distil-large-v3")
c = synth.model_config(name)
st = synth.special_tokens(c["vocab_size"])
TS = lambda s: st["no_timestamps"] + 1 + int(round(s / 0.02))
plan = {0: 7, 1: 8, 2: TS(0.0), 3: 100, 4: 200, 5: TS(2.0), 6: TS(2.02), 7: 300, 8: st["eot"]}
w = synth.plant_decoder_plan(synth.synth_weights(c, seed=1, embed_scale=1.0), c, plan)
ctx = ffi.Context(c, compute="bf16", max_batch=1)
ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"])); ctx.load_weights(w); ctx.set_tokens(**st)
pcm = synth.synth_pcm("gauss", 0, 480_000)
out = {"config": f"{name}-shaped streaming, 160-sample chunks, 1 x B200", "results": []}
for R_ms in (1000, 100):
    ctx.stream_reset()
    push, trig, enc_only = [], [], []
    every = R_ms * 16
    seconds = float(os.environ.get("SECONDS", "20"))
    n_total = int(seconds * 16000)
    toks = None
    for lo in range(0, n_total, 160):
        t0 = time.perf_counter(); ctx.stream_push(pcm[lo:lo + 160]); push.append(time.perf_counter() - t0)
        if (lo + 160) % every == 0:
            t0 = time.perf_counter()
            ctx.stream_features(run_encoder=True)
            t1 = time.perf_counter()
            r = ctx.decode_greedy(1)
            t2 = time.perf_counter()
            trig.append(t2 - t0); enc_only.append(t1 - t0)
            toks = r[0]["tokens"]
    q = lambda a, p: float(np.percentile(np.asarray(a) * 1e3, p))
    res = {"trigger_every_ms": R_ms, "audio_seconds": seconds, "n_triggers": len(trig), "decoded_tokens": len(toks),
           "chunk_push_ms": {"p50": q(push, 50), "p99": q(push, 99)},
           "mel_norm+encoder_ms": {"p50": q(enc_only[2:], 50), "p99": q(enc_only[2:], 99)},
           "chunk_to_tokens_ms": {"p50": q(trig[2:], 50), "p99": q(trig[2:], 99)}}
    out["results"].append(res)
    print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/stream_bench.json", "w"), indent=1)
