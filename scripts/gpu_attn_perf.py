"""Attention microbenchmark: `ATTN_CASES="B:T,B:T"` (default 8:1500,25:1500), 20 launches each, CUDA events on the ctx stream.
`ATTN_QK_STD` (default 1.0): std of q and k; 1.0 gives scores of std 8 (a stress case for the online-softmax rescale), 0.3536 = 64^-0.25
gives unit-variance scores like the encoder's pre-scaled q, k."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
rng = np.random.default_rng(1)
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
H = 20
for case in os.environ.get("ATTN_CASES", "8:1500,25:1500").split(","):
    B, T = (int(x) for x in case.split(":"))
    qkv = rng.standard_normal((B * T, 3 * H * 64)).astype(np.float32)
    qkv[:, : 2 * H * 64] *= float(os.environ.get("ATTN_QK_STD", "1.0"))
    ms = ctx.test_attention_perf(qkv, B, T, H, 20)
    tiles = B * H * ((T + 127) // 128) ** 2 / 148.0
    print(f"[{os.environ.get('TAG','')}] attention B={B} T={T}: {ms*1e3:.1f} us  {4.0*T*T*H*64*B/ms/1e9:.1f} TF  {ms*1e6/tiles:.0f} ns per 128x128 tile per SM", flush=True)
