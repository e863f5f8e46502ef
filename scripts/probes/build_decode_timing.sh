#!/bin/bash
# Instrumented build of the fused decoder step (-DNB200_DECODE_TIMING: %globaltimer of CTA 0 after every phase, printed every 64th step).
#   bash scripts/probes/build_decode_timing.sh && NB200_LIB_PATH=$PWD/scripts/probes/_build/libnorma_b200_dectiming.so python scripts/gpu_decode_perf.py
set -eu
cd "$(dirname "$0")/../.."
python -m norma_b200.build > /dev/null
mkdir -p scripts/probes/_build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr \
     -DNB200_DECODE_TIMING -c norma_b200/csrc/decoder.cu -o scripts/probes/_build/decoder_timing.o
nvcc -shared -o scripts/probes/_build/libnorma_b200_dectiming.so $(ls norma_b200/_obj/*.o | grep -v "/decoder.o") scripts/probes/_build/decoder_timing.o \
     -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo scripts/probes/_build/libnorma_b200_dectiming.so
