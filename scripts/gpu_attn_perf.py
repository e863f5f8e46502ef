import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
rng = np.random.default_rng(1)
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
T, H = 1500, 20
for B in (8, 25):
    qkv = (rng.standard_normal((B * T, 3 * H * 64)) * 1.0).astype(np.float32)
    ms = ctx.test_attention_perf(qkv, B, T, H, 20)
    print(f"[{os.environ.get('TAG','')}] attention B={B}: {ms*1e3:.1f} us  {4.0*T*T*H*64*B/ms/1e9:.1f} TF", flush=True)
