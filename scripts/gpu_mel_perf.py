import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, filters, synth
B = int(os.environ.get("B", "32")); n_mel = int(os.environ.get("NMEL", "128")); iters = int(os.environ.get("ITERS", "20"))
c = dict(synth.model_config("test-micro"), num_mel_bins=n_mel)
ctx = ffi.Context(c, compute="bf16", max_batch=B)
ctx.set_mel_filters(filters.mel_filters(n_mel))
pcm = np.stack([synth.synth_pcm_window(i) for i in range(B)])
ctx.stage_pcm(pcm)
for _ in range(3): ctx.run_resident(B, True, False)
ctx.sync(); ctx.profile_reset(); ctx.profile_enable(True)
ctx.timer_start()
for _ in range(iters): ctx.run_resident(B, True, False)
ms = ctx.timer_stop() / iters
prof, _ = ctx.profile_read()
byt = (4.0 * 480000 + 4.0 * n_mel * 3000) * B
print(f"B={B} n_mel={n_mel}: mel stage {ms*1e3:.1f} us/step; mel_kernel {prof['mel']['ms']/iters*1e3:.1f} us -> {byt/ (prof['mel']['ms']/iters/1e3)/1e9:.0f} GB/s algorithmic; norm {prof['mel_norm']['ms']/iters*1e3:.1f} us")
