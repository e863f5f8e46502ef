#!/bin/bash
# Round-2 evidence run (under gpurun): GPU tests, the default bench line, the ncu launch list of the headline step and
# `ncu --set full` captures of mel_kernel and decoder_step_fused_kernel (VERDICT r1 "missing" item 3).
# Usage: bash scripts/profile_r2.sh <tag> [tests|notests]
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
if [ "${2:-tests}" = "tests" ]; then
  ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$TAG.log 2>&1
  tail -5 gpurun_out/pytest_$TAG.log
fi
( time timeout 900 python bench.py ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
CMD="python bench.py --quick --windows 8 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 460 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
B=25 NMEL=128 ITERS=2 timeout 300 python scripts/gpu_mel_perf.py > gpurun_out/mel_plain_$TAG.log 2>&1 &&
B=25 NMEL=128 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^mel_kernel" -s 3 -c 2 -o gpurun_out/mel_$TAG -f python scripts/gpu_mel_perf.py > gpurun_out/ncu_mel_$TAG.log 2>&1
BS=1,8 NTOK=48 timeout 300 python scripts/gpu_decode_perf.py > gpurun_out/decode_plain_$TAG.log 2>&1 &&
BS=1 NTOK=48 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_step_fused -s 1 -c 2 -o gpurun_out/decfused_$TAG -f python scripts/gpu_decode_perf.py > gpurun_out/ncu_decfused_$TAG.log 2>&1
cat gpurun_out/mel_plain_$TAG.log gpurun_out/decode_plain_$TAG.log
ls -la gpurun_out/
