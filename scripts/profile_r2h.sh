#!/bin/bash
# Round-2 final-tree ncu evidence (under gpurun, one GPU): launch list of the headline step, `ncu --set full` of one encoder layer's four
# GEMMs (LayerNorm folded), the attention kernel, the single-round wide GEMMs of the one-window path, and the fused decoder step.
# Usage: bash scripts/profile_r2h.sh <tag>
set -u
TAG=${1:-r2h}
mkdir -p gpurun_out
CMD="python bench.py --quick --windows 8 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 850 -c 340 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
# one layer = qkv, attention, out-proj, fc1, fc2: skip the warm-up steps (5 x 170 launches) and a few layers of the first timed step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 670 -c 4 -o gpurun_out/gemm_$TAG -f $CMD > gpurun_out/ncu_gemm_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 165 -c 1 -o gpurun_out/attn_$TAG -f $CMD > gpurun_out/ncu_attn_$TAG.log 2>&1
CMD1="python bench.py --quick --windows 1 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_wide_kernel -s 320 -c 2 -o gpurun_out/wide_$TAG -f $CMD1 > gpurun_out/ncu_wide_$TAG.log 2>&1
BS=1 NTOK=48 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_step_fused -s 1 -c 1 -o gpurun_out/decfused_$TAG -f python scripts/gpu_decode_perf.py > gpurun_out/ncu_decfused_$TAG.log 2>&1
ls -la gpurun_out/ | grep $TAG
