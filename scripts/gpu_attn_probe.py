import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
rng = np.random.default_rng(1)
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
for (B, T, H, sc) in ((1, 128, 1, 0.5), (1, 256, 1, 0.5), (1, 64, 2, 0.5), (2, 200, 2, 0.5), (1, 1500, 3, 0.5), (2, 1500, 20, 1.0), (1, 1500, 2, 3.0)):
    d = H * 64
    qkv = ffi.bf16_round((rng.standard_normal((B * T, 3 * d)) * sc).astype(np.float32))
    q, k, v = [torch.from_numpy(qkv[:, i*d:(i+1)*d]).double().view(B, T, H, 64).transpose(1, 2) for i in range(3)]
    ref = (torch.softmax(q @ k.transpose(2, 3), -1) @ v).transpose(1, 2).reshape(B * T, d).numpy()
    for impl in ("simt", "tc"):
        os.environ["NB200_ATTN"] = impl
        got = ctx.test_attention(qkv, B, T, H)
        err = np.abs(got - ref)
        print(f"{impl:4s} B{B} T{T} H{H} scale {sc}: maxabs {err.max():.3e} nan={np.isnan(got).sum()} worst row {np.unravel_index(err.argmax(), err.shape)}", flush=True)
