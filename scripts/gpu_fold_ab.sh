#!/bin/bash
# A/B inside the bench: per-shape GEMM times (CUDA events) and the step, 25 windows and 1 window
run() { echo "== $*"; env "$@" NB200_PROF_DUMP=1 python bench.py --quick --no-cpu-baseline 2>&1 | grep -E "class 2 tag (128001280|128005120|384001280|512001280)|^\{" | sed -e 's/^{.*"ms_per_step": \([0-9.]*\).*/ms_per_step \1/'; }
run NB200_GEMM_DIRECT_F32=1
run NB200_GEMM_DIRECT_F32=0
run NB200_GEMM_DIRECT_F32=1 NB200_BENCH_WINDOWS=1
run NB200_GEMM_DIRECT_F32=0 NB200_BENCH_WINDOWS=1
