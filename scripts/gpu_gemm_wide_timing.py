"""Where a CTA of gemm_wide_kernel spends its cycles (NB200_GEMM_DEBUG=256 prints clock64 phase deltas per launch)."""
import os, sys
os.environ["NB200_GEMM_DEBUG"] = "256"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
for (N, K, epi) in ((3840, 1280, 0), (5120, 1280, 1)):
    ms = ctx.test_gemm_perf(1500, N, K, epi, 3)
    print(f"N={N} K={K}: {ms*1e3:.1f} us per launch (with the host sync of the timing build)", flush=True)
