import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
rng = np.random.default_rng(0)
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
for (M, N, K) in ((256, 256, 64), (128, 256, 64), (256, 256, 128), (300, 512, 256), (1500, 384, 384), (1500, 1152, 384), (3000, 1280, 240), (1000, 3840, 1280), (4500, 1280, 5120), (12000, 5120, 1280)):
    a = ffi.bf16_round(rng.standard_normal((M, K)).astype(np.float32)); w = ffi.bf16_round((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
    bias = rng.standard_normal(N).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T + bias
    got = ctx.test_gemm(a, w, bias)
    err = np.abs(got - ref)
    print(f"gemm[{os.environ.get('NB200_GEMM','2cta')}] {M}x{N}x{K}: maxabs {err.max():.3e} nan={np.isnan(got).sum()}", flush=True)
    if err.max() > 1e-2 or np.isnan(got).any():
        bad = np.argwhere(~(err < 1e-2))
        print("   n_bad", len(bad), "rows", np.unique(bad[:, 0])[:12], "cols", np.unique(bad[:, 1])[:12], flush=True)
