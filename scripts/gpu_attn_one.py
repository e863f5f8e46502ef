import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
rng = np.random.default_rng(1)
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
B, T, H = 8, 1500, 20
qkv = (rng.standard_normal((B * T, 3 * H * 64)) * 1.0).astype(np.float32)
for _ in range(3):
    got = ctx.test_attention(qkv, B, T, H)
print("ok", float(np.abs(got).mean()))
