#!/bin/bash
# BASELINE config 5 (large-v3-shaped encoder, batch sweep) and config 3 (base.en, 120 windows = 1 h, strong scaling) on this box
set -u
mkdir -p gpurun_out
NG=${1:-1}
for B in 1 2 4 8 16 32 64 128 256; do
  S=$(( B >= 64 ? 3 : (B >= 8 ? 6 : 12) ))
  timeout 600 python bench.py --model large-v3 --windows $B --steps $S --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/config5_B$B.json
done
timeout 600 python bench.py --model base.en --total-windows 120 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/config3_1gpu.json
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --model base.en --total-windows 120 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/config3_2gpu.json
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/config5_B*.json"), key=lambda s: int(s.split("_B")[1].split(".")[0])) + sorted(glob.glob("gpurun_out/config3_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j["roofline"]
        print(f"{f}: value {j['value']:.0f} audio-s/s  {j['ms_per_step']:.2f} ms/step  e2e {j['e2e']['value']:.0f}  gemm {r['achieved']:.0f} TF  whole {r['whole_step_frac_of_peak']:.3f}  n_gpus {j['n_gpus']}")
    except Exception as e:
        print(f, "failed", e)
PY
