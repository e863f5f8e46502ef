#!/bin/bash
# Instrumented build of the attention kernel (-DNB200_ATTN_TIMING: clock64 brackets per softmax phase, printed per launch on stderr).
#   bash scripts/probes/build_attn_timing.sh && NB200_LIB_PATH=$PWD/scripts/probes/_build/libnorma_b200_timing.so python scripts/gpu_attn_perf.py
set -eu
cd "$(dirname "$0")/../.."
python -m norma_b200.build > /dev/null
mkdir -p scripts/probes/_build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr \
     -DNB200_ATTN_TIMING -c norma_b200/csrc/attn_tcgen05.cu -o scripts/probes/_build/attn_timing.o
nvcc -shared -o scripts/probes/_build/libnorma_b200_timing.so $(ls norma_b200/_obj/*.o | grep -v attn_tcgen05) scripts/probes/_build/attn_timing.o \
     -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo scripts/probes/_build/libnorma_b200_timing.so
