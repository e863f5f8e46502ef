"""norma_b200 — B200-native (sm_100a) implementation of norma's hot path behind a C ABI (include/norma_b200.h).

The arithmetic lives in `libnorma_b200.so` (hand-written CUDA, built in-tree by `norma_b200.build`).  Python here is
only the host-side mirror used by tests and bench.py; it never computes the path itself and has no fallback."""
from . import ffi, filters, synth  # noqa: F401

__all__ = ["ffi", "filters", "synth"]
