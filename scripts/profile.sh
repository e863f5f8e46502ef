#!/bin/bash
# ncu evidence for profiles/: launch list (shares) + one full capture each of the GEMM and attention kernels.
# Usage (under gpurun): bash scripts/profile.sh <tag>
set -u
TAG=${1:-r1}
CMD="python bench.py --windows 8 --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 460 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 330 -c 4 -o gpurun_out/gemm_$TAG $CMD > gpurun_out/ncu_gemm_$TAG.log 2>&1
$CMD > gpurun_out/plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 100 -c 2 -o gpurun_out/attn_$TAG $CMD > gpurun_out/ncu_attn_$TAG.log 2>&1
$CMD > gpurun_out/plain4_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mel_kernel|layernorm_kernel" -s 8 -c 3 -o gpurun_out/melln_$TAG $CMD > gpurun_out/ncu_melln_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log
ls -la gpurun_out/
