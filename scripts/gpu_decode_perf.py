import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, filters, synth
name = os.environ.get("MODEL", "distil-large-v3")
c = synth.model_config(name)
w = synth.synth_weights(c, seed=1)
st = synth.special_tokens(c["vocab_size"])
d, V, L = c["d_model"], c["vocab_size"], c["decoder_layers"]
bytes_tok = 2.0 * (L * 14 * d * d + V * d)
for B in [int(x) for x in os.environ.get("BS", "1,8").split(",")]:
    ctx = ffi.Context(c, compute="bf16", max_batch=max(B, int(os.environ.get("MAXB", "0"))))
    ctx.set_mel_filters(filters.mel_filters(c["num_mel_bins"])); ctx.load_weights(w); ctx.set_tokens(**st)
    pcm = np.stack([synth.synth_pcm_window(i) for i in range(B)])
    ctx.transcode_batch(pcm, want_output=False)
    ctx.decode_greedy(B, max_new_tokens=8)
    ctx.sync()
    t = time.perf_counter(); r = ctx.decode_greedy(B, max_new_tokens=int(os.environ.get("NTOK", "200"))); dt = time.perf_counter() - t
    ctx.profile_reset(); ctx.profile_enable(True)   # per-class device times come from a second, un-graphed run
    ctx.decode_greedy(B, max_new_tokens=int(os.environ.get("NTOK", "200")))
    prof, _ = ctx.profile_read(); ctx.profile_enable(False)
    n = len(r[0]["tokens"]) - 4
    print(f"{name} B={B}: {n} tokens/window in {dt*1e3:.1f} ms -> {dt/n*1e6:.1f} us/step ({B*n/dt:.0f} tok/s); HBM floor {bytes_tok/6536.4e9*1e6:.1f} us/step; "
          f"device ms: gemv {prof['decode_gemv']['ms']:.1f} attn {prof['decode_attn']['ms']:.1f} select {prof['decode_select']['ms']:.1f} ln {prof['layernorm']['ms']:.1f} gemm {prof['gemm']['ms']:.1f} misc {prof['misc']['ms']:.1f}; launches {sum(v['launches'] for v in prof.values())}", flush=True)
    ctx.close()
