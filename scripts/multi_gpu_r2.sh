#!/bin/bash
# Round-2 multi-GPU evidence: usage  bash scripts/multi_gpu_r2.sh N TAG   (under gpurun --gpus N)
# 1. the one-process N-context test (one nb200_ctx + one host thread per GPU), 2. the bench under torchrun (headline + config 3 strong
# scaling + config 5 sweep on N GPUs in the same JSON line), 3. the bench in single-process mode.
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "multictx" 2>&1 | tail -3) | tee gpurun_out/multictx_${TAG}_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline \
  > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
tail -2 gpurun_out/bench_${TAG}_${N}gpu.err
timeout 600 python bench.py --gpus $N --single-process > gpurun_out/bench_${TAG}_${N}gpu_single_process.json 2> gpurun_out/bench_${TAG}_${N}gpu_single_process.err
tail -2 gpurun_out/bench_${TAG}_${N}gpu_single_process.err
python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_${N}gpu.json", "gpurun_out/bench_${TAG}_${N}gpu_single_process.json"):
    try:
        j = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value", round(j["value"]), "ms", round(j["ms_per_step"], 2), "n_gpus", j["n_gpus"], j.get("clocks", {}).get("sm_mhz"))
        if "config3" in j: print("  config3", round(j["config3"]["value"]), j["config3"]["ms_per_step"])
        if "config5" in j: print("  config5", [(r["windows_per_gpu"], round(r["value"])) for r in j["config5"]["rows"]])
    except Exception as e:
        print(f, "unreadable:", e)
PY
