"""Where an epilogue warp of gemm_tc_kernel's f32 + residual path spends its cycles at the bench shape (NB200_GEMM_DEBUG=256)."""
import os, sys
os.environ["NB200_GEMM_DEBUG"] = "256"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
M = int(os.environ.get("M", "37500"))
for (N, K) in ((1280, 1280), (1280, 5120)):
    ms = ctx.test_gemm_perf(M, N, K, 2, 2)
