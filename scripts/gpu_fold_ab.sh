#!/bin/bash
# A/B inside the bench: per-shape GEMM times (CUDA events) and the step, 25 windows
run() { echo "== $*"; env "$@" NB200_PROF_DUMP=1 python bench.py --quick --no-cpu-baseline 2>&1 | grep -E "class 2 tag (128001280|128005120|384001280|512001280)|^\{" | sed -e 's/^{.*"ms_per_step": \([0-9.]*\).*/ms_per_step \1/'; }
run NB200_GEMM_NP3=1
run NB200_GEMM_NP3=0
