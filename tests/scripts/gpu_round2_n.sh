#!/bin/bash
# strong scaling of config 4 (TDO, S = 20, global 64 episodes = 1280 frames / step) on N GPUs
mkdir -p gpurun_out
N=${1:-8}
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --model tdo --global-batch 64 --only-main --no-cpu-baseline --steps 10 --warmup 3 2> gpurun_out/r2n_strong_${N}gpu.err | grep '^{' > gpurun_out/r2n_strong_${N}gpu.json
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --model tdo --global-batch 64 --only-main --no-cpu-baseline --steps 10 --warmup 3 2> gpurun_out/r2n_strong_${N}gpu.err | grep '^{' > gpurun_out/r2n_strong_${N}gpu.json
fi
python - <<PY
import json
d = json.load(open("gpurun_out/r2n_strong_${N}gpu.json"))
print("strong N=${N} tdo", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["config"].get("per_gpu_batch"), d.get("scaling"))
PY
