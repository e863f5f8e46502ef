import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
M = int(os.environ.get("M", "37500"))
shapes = [("qkv", 3840, 1280, 0), ("out", 1280, 1280, 2), ("fc1", 5120, 1280, 1), ("fc2", 1280, 5120, 2), ("out_bf16", 1280, 1280, 0), ("qkv_res", 3840, 1280, 2)]
for name, N, K, ek in shapes:
    ms = ctx.test_gemm_perf(M, N, K, ek, 20)
    print(f"[{os.environ.get('NB200_GEMM','2cta')} dbg={os.environ.get('NB200_GEMM_DEBUG','0')}] {name:8s} M={M} N={N} K={K} epi={ek}: {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:8.1f} TF", flush=True)
