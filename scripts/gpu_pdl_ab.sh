#!/bin/bash
# A/B of programmatic dependent launch on the encoder chain, 25 windows and 1 window per step
run() { echo "== $*"; env "$@" python bench.py --quick --no-cpu-baseline 2>&1 | grep -E "^\{" | sed -e 's/^{.*"ms_per_step": \([0-9.]*\).*/ms_per_step \1/'; }
run NB200_PDL=1
run NB200_PDL=0
run NB200_PDL=1 NB200_BENCH_WINDOWS=1
run NB200_PDL=0 NB200_BENCH_WINDOWS=1
