import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
N, K, ek = [int(x) for x in os.environ.get("SHAPE", "3840,1280,0").split(",")]
ms = ctx.test_gemm_perf(37500, N, K, ek, 3)
print(f"N={N} K={K} epi={ek}: {ms*1e3:.1f} us {2.0*37500*N*K/ms/1e9:.1f} TF")
