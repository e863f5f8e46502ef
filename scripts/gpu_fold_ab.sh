#!/bin/bash
# A/B of the fused LayerNorm against standalone LayerNorm kernels, 25 windows and 1 window per step
run() { echo "== $*"; env "$@" NB200_PROF_DUMP=1 python bench.py --quick --no-cpu-baseline 2>&1 | grep -E "class 2 tag (128001280|128005120|384001280|512001280)|^\{" | sed -e 's/^{.*"ms_per_step": \([0-9.]*\).*/ms_per_step \1/'; }
run NB200_LN_FUSED=1
run NB200_LN_FUSED=0
run NB200_LN_FUSED=1 NB200_BENCH_WINDOWS=1
run NB200_LN_FUSED=0 NB200_BENCH_WINDOWS=1
