"""Per-shape GEMM time for the three tile configurations at small M (NB200_GEMM=… NB200_GEMM_NOFIT=1 select them per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from norma_b200 import ffi, synth
ctx = ffi.Context(synth.model_config("test-micro"), compute="bf16", max_batch=1)
tag = os.environ.get("TAG", "")
for M in (1500, 3000, 6000, 12000):
    row = []
    for (N, K, epi) in ((3840, 1280, 0), (1280, 1280, 2), (5120, 1280, 1), (1280, 5120, 2)):
        ms = ctx.test_gemm_perf(M, N, K, epi, 30)
        row.append(f"{ms*1e3:7.1f}")
    print(f"[{tag}] M={M:6d} us: qkv {row[0]} out {row[1]} fc1 {row[2]} fc2 {row[3]}", flush=True)
